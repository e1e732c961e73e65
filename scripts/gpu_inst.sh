#!/bin/bash
# dynamic instruction count, fp64-pipe share and duration of the flux kernels (256^3)
mkdir -p gpurun_out
TAG=${TAG:-inst}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu_$TAG.log
for lib in ${VARIANTS:-enzo-e_b200/csrc/libvlct_b200.so}; do
name=$(basename $lib .so)
VLCT_B200_LIB=$PWD/$lib timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_fp64.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:'k_flux|k_update|k_edge|k_face|k_ct' -s 18 -c 9 --csv --log-file gpurun_out/inst_${TAG}_$name.csv python bench.py --size ${SIZE:-256} --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/inst_${TAG}_$name.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/inst_${TAG}_$name.csv")) if len(r)>10]
hdr=rows[0]; ix={h:i for i,h in enumerate(hdr)}
out={}
for r in rows[1:]:
    key=(r[ix["ID"]], r[ix["Kernel Name"]].split("(")[0][-40:])
    out.setdefault(key,{})[r[ix["Metric Name"]]]=(r[ix["Metric Value"]], r[ix["Metric Unit"]])
print("$name")
for k,v in out.items():
    print("  %-44s" % k[1], " ".join("%s=%s%s" % (m.split("__")[-1][:28], x[0], x[1]) for m,x in v.items()))
PY
done
