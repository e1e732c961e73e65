"""CPU: the C-ABI library loads, exports every symbol include/vlct.h declares,
parses/validates parameters like the reference's constructors, and refuses to
work without a GPU (no CPU fallback). No compute call is made here."""
import ctypes as C
import os
import re

import pytest

from enzo_e_b200 import abi, lib as libmod
from enzo_e_b200.method import config_from_parameters, EnzoMethodMHDVlct
from enzo_e_b200.lib import VlctError

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "vlct.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vlct_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = libmod.load()
    declared = header_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared but not exported"
    assert set(declared) == set(libmod.EXPORTED_SYMBOLS)


def test_struct_layouts_match_the_header(tmp_path):
    """ctypes mirror (abi.py) vs the C compiler's layout of include/vlct.h."""
    import subprocess
    src = tmp_path / "probe.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "vlct.h"\n'
        'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu\\n",'
        'sizeof(vlct_config), sizeof(vlct_block),'
        'offsetof(vlct_config, pressure_floor), offsetof(vlct_block, dx),'
        'offsetof(vlct_block, pressure), offsetof(vlct_block, passive),'
        'offsetof(vlct_block, stream)); return 0;}\n')
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o",
                    str(exe)], check=True)
    got = [int(x) for x in subprocess.run(
        [str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    want = [C.sizeof(abi.VlctConfig), C.sizeof(abi.VlctBlock),
            abi.VlctConfig.pressure_floor.offset, abi.VlctBlock.dx.offset,
            abi.VlctBlock.pressure.offset, abi.VlctBlock.passive.offset,
            abi.VlctBlock.stream.offset]
    assert got == want


def test_name_and_defaults():
    lib = libmod.load()
    assert lib.vlct_name() == b"mhd_vlct"
    cfg = abi.VlctConfig()
    assert lib.vlct_config_init(C.byref(cfg)) == abi.VLCT_OK
    ref = abi.default_config()
    for name, _ in abi.VlctConfig._fields_:
        assert getattr(cfg, name) == getattr(ref, name), name


def test_parameter_keys_roundtrip():
    cfg = config_from_parameters({
        "Method:mhd_vlct:mhd_choice": "constrained_transport",
        "riemann_solver": "hlle", "reconstruct_method": "plm_athena",
        "theta_limiter": 1.25, "time_scheme": "vl", "courant": 0.4,
        "Physics:fluid_props:eos:gamma": 1.4,
        "Physics:fluid_props:dual_energy:type": "modern",
        "Physics:fluid_props:dual_energy:eta": 0.00769,
        "Physics:fluid_props:floors:density": 1e-200,
        "Physics:fluid_props:floors:pressure": 1e-100}, n_passive=2)
    assert cfg.riemann_solver == abi.RIEMANN["hlle"]
    assert cfg.reconstruct_method == abi.RECON["plm_athena"]
    assert cfg.theta_limiter == 1.25 and cfg.courant == 0.4
    assert cfg.gamma == 1.4 and cfg.dual_energy == 1
    assert cfg.dual_energy_eta == 0.00769
    assert cfg.density_floor == 1e-200 and cfg.pressure_floor == 1e-100
    assert cfg.n_passive == 2


def test_legacy_aliases():
    cfg = config_from_parameters({
        "mhd_choice": "no_bfield", "Method:mhd_vlct:dual_energy": True,
        "Method:mhd_vlct:dual_energy_eta": 0.01,
        "Method:mhd_vlct:density_floor": 1e-10,
        "Method:mhd_vlct:pressure_floor": 1e-11, "Field:gamma": 1.5})
    assert cfg.dual_energy == 1 and cfg.dual_energy_eta == 0.01
    assert cfg.density_floor == 1e-10 and cfg.pressure_floor == 1e-11
    assert cfg.gamma == 1.5


@pytest.mark.parametrize("key", ["half_dt_reconstruct_method",
                                 "full_dt_reconstruct_method"])
def test_removed_keys_are_rejected(key):
    with pytest.raises(VlctError) as e:
        config_from_parameters({key: "nn"})
    assert "have been removed" in e.value.message


def test_unknown_key():
    with pytest.raises(VlctError) as e:
        config_from_parameters({"Method:mhd_vlct:nonsense": "1"})
    assert e.value.status == abi.VLCT_ERR_UNKNOWN_KEY


def _validate(**kw):
    from helpers import make_config
    lib = libmod.load()
    cfg = make_config(**kw)
    err = C.create_string_buffer(512)
    rc = lib.vlct_config_validate(C.byref(cfg), err, len(err))
    return rc, err.value.decode()


@pytest.mark.parametrize("kw,needle", [
    (dict(riemann="hllc", mhd=True), "can't support mhd"),
    (dict(riemann="hlld", mhd=False), "requires magnetic fields"),
    (dict(riemann="hlle", mhd=False), "untested"),
    (dict(riemann="hll", mhd=True), "hasn't been tested"),
    (dict(theta=2.5), "theta_limiter"),
    (dict(theta=0.5), "theta_limiter"),
    (dict(dfloor=0.0), "floors must be defined"),
    (dict(pfloor=0.0), "floors must be defined"),
    (dict(gamma=1.0), "gamma"),
    (dict(time_scheme="euler", mhd=True), "num_partial_timesteps"),
    (dict(n_passive=17), "n_passive"),
])
def test_validation_mirrors_reference_errors(kw, needle):
    rc, msg = _validate(**kw)
    assert rc == abi.VLCT_ERR_INVALID_CONFIG
    assert needle in msg


def test_mhd_choice_is_required():
    lib = libmod.load()
    cfg = abi.default_config()
    cfg.density_floor = cfg.pressure_floor = 1e-200
    err = C.create_string_buffer(512)
    assert lib.vlct_config_validate(C.byref(cfg), err, len(err)) \
        == abi.VLCT_ERR_INVALID_CONFIG
    assert b"mhd_choice" in err.value


def test_valid_combinations():
    for kw in (dict(riemann="hlld", mhd=True), dict(riemann="hlle", mhd=True),
               dict(riemann="hllc", mhd=False),
               dict(riemann="hllc", mhd=False, time_scheme="euler"),
               dict(riemann="hlld", recon="nn", dual_energy=True, eta=0.0)):
        rc, msg = _validate(**kw)
        assert rc == abi.VLCT_OK, msg


def test_no_cpu_fallback():
    """Without a CUDA device the Method cannot be constructed."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from helpers import make_config
    with pytest.raises(VlctError) as e:
        EnzoMethodMHDVlct(config=make_config())
    assert e.value.status == abi.VLCT_ERR_NO_DEVICE
    assert "no CPU fallback" in e.value.message


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "enzo-e_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert "import oracle" not in text, fn
                assert "vlct_oracle" not in text, fn
                assert "libvlct_ref" not in text, fn
