// dp_pipe.cu -- fp64 pipe latency / throughput on one SM (design input for the
// flux kernels). Prints cycles per warp-instruction for dependent chains with
// ILP = 1,2,4 and 1..16 warps on one SM (4 SMSPs), for DFMA, DADD, DMUL and a
// division / sqrt / rcp mix.
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP, int OP>
__global__ void k(double* out, long long* cyc, int iters, double seed)
{
  double a[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) a[i] = seed + i + threadIdx.x * 1e-3;
  const double b = 1.0000001, c = 1e-9;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
#pragma unroll
      for (int i = 0; i < ILP; i++) {
        if (OP == 0) a[i] = __fma_rn(a[i], b, c);
        if (OP == 1) a[i] = __dadd_rn(a[i], c);
        if (OP == 2) a[i] = __dmul_rn(a[i], b);
        if (OP == 3) a[i] = b / a[i] + 1.0;         // IEEE division
        if (OP == 4) a[i] = sqrt(a[i]) + 1.0;
        if (OP == 5) a[i] = 1.0 / a[i] + 1.0;
        if (OP == 6) a[i] = (a[i] < b) ? a[i] + c : a[i] * b;   // DSETP + select
      }
    }
  }
  long long t1 = clock64();
  __syncthreads();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int ILP, int OP>
void run(const char* name, double* out, long long* cyc)
{
  const int iters = 2000;
  printf("%-8s ILP=%d :", name, ILP);
  for (int warps : {1, 4, 8, 12, 16, 24, 32}) {
    k<ILP, OP><<<1, warps * 32>>>(out, cyc, iters, 1.5);
    cudaDeviceSynchronize();
    k<ILP, OP><<<1, warps * 32>>>(out, cyc, iters, 1.5);
    cudaDeviceSynchronize();
    long long h;
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    // cycles per (warp-level op) per SMSP: warps/4 warps share one SMSP
    double per_op_chain = (double) h / (iters * 8.0);            // cycles per chain step (all ILP)
    double smsp_ops = (double) iters * 8.0 * ILP * ((warps + 3) / 4);
    printf("  w%-2d %6.2f c/step %5.2f c/op", warps, per_op_chain, (double) h / smsp_ops);
  }
  printf("\n");
}

int main()
{
  double* out; long long* cyc;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
  run<1, 0>("DFMA", out, cyc); run<2, 0>("DFMA", out, cyc); run<4, 0>("DFMA", out, cyc);
  run<1, 1>("DADD", out, cyc); run<2, 1>("DADD", out, cyc);
  run<1, 2>("DMUL", out, cyc); run<2, 2>("DMUL", out, cyc);
  run<1, 6>("DSETSEL", out, cyc); run<2, 6>("DSETSEL", out, cyc);
  run<1, 3>("DIV", out, cyc); run<2, 3>("DIV", out, cyc); run<4, 3>("DIV", out, cyc);
  run<1, 4>("SQRT", out, cyc); run<2, 4>("SQRT", out, cyc);
  run<1, 5>("RCP", out, cyc); run<2, 5>("RCP", out, cyc);
  return 0;
}
