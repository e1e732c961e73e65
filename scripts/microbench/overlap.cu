// overlap.cu -- can an HBM-bound persistent kernel (few warps per SM, high
// priority stream) run under an fp64-bound kernel that fills the register file?
// Prints: A alone, B alone, A+B concurrent (two streams), for several B grids.
#include <cstdio>
#include <cuda_runtime.h>

// fp64-bound: ~600 dependent-ish DP ops per element, 128 registers per thread
__global__ void __launch_bounds__(128, 4)
kA(const double* __restrict__ in, double* __restrict__ out, size_t n, int reps)
{
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  double a[8];
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = in[t] + i;
  for (int r = 0; r < reps; r++) {
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = __fma_rn(a[i], 1.0000001, a[(i + 1) & 7] * 1e-9);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += a[i];
  out[t] = s;
}

// HBM-bound persistent copy-add: grid-stride, 4 independent loads per thread
__global__ void __launch_bounds__(256)
kB(const double2* __restrict__ x, const double2* __restrict__ y, double2* __restrict__ z, size_t n2)
{
  const size_t stride = (size_t) gridDim.x * blockDim.x;
  for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += 4 * stride) {
    double2 a[4], b[4];
#pragma unroll
    for (int u = 0; u < 4; u++) if (i + u * stride < n2) { a[u] = x[i + u * stride]; b[u] = y[i + u * stride]; }
#pragma unroll
    for (int u = 0; u < 4; u++) if (i + u * stride < n2) z[i + u * stride] = make_double2(a[u].x + b[u].x, a[u].y + b[u].y);
  }
}

int main()
{
  const size_t nA = (size_t) 64 << 20;      // elements for A
  const size_t nB = (size_t) 512 << 20;     // doubles per array for B (4 GiB each)
  double *ain, *aout, *x, *y, *z;
  cudaMalloc(&ain, nA * 8); cudaMalloc(&aout, nA * 8);
  cudaMalloc(&x, nB * 8); cudaMalloc(&y, nB * 8); cudaMalloc(&z, nB * 8);
  cudaMemset(ain, 0, nA * 8); cudaMemset(x, 0, nB * 8); cudaMemset(y, 0, nB * 8);
  int lo, hi; cudaDeviceGetStreamPriorityRange(&lo, &hi);
  cudaStream_t sa, sb;
  cudaStreamCreateWithPriority(&sa, cudaStreamNonBlocking, lo);
  cudaStreamCreateWithPriority(&sb, cudaStreamNonBlocking, hi);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int reps = 75;   // 8*75 = 600 DFMA + 600 DMUL per element
  auto timeit = [&](bool runA, bool runB, int gridB, int blockB) {
    float best = 1e9;
    for (int it = 0; it < 3; it++) {
      cudaDeviceSynchronize();
      cudaEventRecord(e0, sa);
      cudaStreamWaitEvent(sb, e0, 0);
      if (runA) kA<<<(unsigned) ((nA + 127) / 128), 128, 0, sa>>>(ain, aout, nA, reps);
      if (runB) kB<<<gridB, blockB, 0, sb>>>((double2*) x, (double2*) y, (double2*) z, nB / 2);
      cudaEventRecord(e1, sb);
      cudaStreamWaitEvent(sa, e1, 0);
      cudaEventRecord(e1, sa);
      cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    return best;
  };
  const float tA = timeit(true, false, 0, 0);
  printf("A alone: %.2f ms (%.1f G DP-inst/s-per-thread-equivalent)\n", tA, nA * 1200.0 / tA / 1e6);
  for (int blockB : {256, 512}) for (int perSM : {1, 2, 4}) {
    const int gridB = 148 * perSM;
    const float tB = timeit(false, true, gridB, blockB);
    const float tAB = timeit(true, true, gridB, blockB);
    printf("B %d blocks/SM x %d thr: alone %.2f ms (%.0f GB/s)  A+B concurrent %.2f ms  (sum %.2f, overlap gain %.2f ms)\n",
           perSM, blockB, tB, 3.0 * nB * 8 / tB / 1e6, tAB, tA + tB, tA + tB - tAB);
  }
  return 0;
}
