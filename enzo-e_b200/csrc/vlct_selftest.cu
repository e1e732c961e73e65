// vlct_selftest.cu -- device self-test of vlct_fpops.cuh: the straight-line
// division / reciprocal / square root must equal the built-in IEEE operators
// bit for bit wherever their range guard passes. Exported through the C ABI so
// that tests/test_gpu_fpops.py can run it on the GPU box.
#include "vlct_fpops.cuh"
#include "../../include/vlct.h"

namespace vlct {
namespace {

__device__ __forceinline__ unsigned long long mix64(unsigned long long z)
{
  z += 0x9e3779b97f4a7c15ULL;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  return z ^ (z >> 31);
}

/// operand generator: mode 0 = O(1) magnitudes (what the solver sees), mode 1 =
/// any bit pattern (all exponents, denormals, infinities, NaNs), mode 2 =
/// specials and near-specials
__device__ double operand(unsigned long long h, int mode)
{
  if (mode == 0) {
    // sign, exponent in [-40, 40], random mantissa
    const unsigned long long mant = h & 0x000fffffffffffffULL;
    const long long e = 1023 + (long long) ((h >> 52) % 81) - 40;
    const unsigned long long s = (h >> 63) << 63;
    return __longlong_as_double((long long) (s | ((unsigned long long) e << 52) | mant));
  }
  if (mode == 1) return __longlong_as_double((long long) h);
  const double specials[16] = {
    0.0, -0.0, 1.0, -1.0, 4.9406564584124654e-324, 2.2250738585072014e-308,
    2.2250738585072009e-308, 1.7976931348623157e308, __longlong_as_double(0x7ff0000000000000LL),
    __longlong_as_double(0xfff0000000000000LL), __longlong_as_double(0x7ff8000000000000LL),
    1e-300, 1e300, 3.0, 1e-8, 0.5 };
  double v = specials[h & 15];
  if (h & 16) v = __longlong_as_double(__double_as_longlong(v) ^ (long long) ((h >> 8) & 3));
  return v;
}

__device__ __forceinline__ bool same_bits(double x, double y)
{
  const long long a = __double_as_longlong(x), b = __double_as_longlong(y);
  // all NaNs are equivalent (the built-in's NaN payload is not contractual)
  if (x != x && y != y) return true;
  return a == b;
}

// ops: div, rcp, sqrt, shared-reciprocal pair (quotz + quot), divz
// counters: [op][0] = guarded-ok results that differ from the built-in (must be
// 0), [op][1] = results flagged for the slow path
__global__ void k_selftest(long long n, unsigned long long seed, int mode,
                           unsigned long long* counters)
{
  unsigned long long wrong[5] = { 0, 0, 0, 0, 0 }, slow[5] = { 0, 0, 0, 0, 0 };
  for (long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x; t < n;
       t += (long long) gridDim.x * blockDim.x) {
    const unsigned long long h0 = mix64(seed + 3ULL * (unsigned long long) t);
    const double a = operand(h0, mode);
    const double b = operand(mix64(h0), mode);
    const double c = operand(mix64(h0 ^ 0x5555555555555555ULL), mode);
    { FastOps op; const double q = op.div(a, b);
      if (op.bad) slow[0]++; else if (!same_bits(q, a / b)) wrong[0]++; }
    { FastOps op; const double q = op.rcp(b);
      if (op.bad) slow[1]++; else if (!same_bits(q, 1.0 / b)) wrong[1]++; }
    { FastOps op; const double q = op.sqrt(a);
      if (op.bad) slow[2]++; else if (!same_bits(q, ::sqrt(a))) wrong[2]++; }
    { FastOps op; const double r = op.prep(b);
      const double q1 = op.quotz(a, b, r), q2 = op.quot(c, b, r);
      if (op.bad) slow[3]++;
      else if (!same_bits(q1, a / b) || !same_bits(q2, c / b)) wrong[3]++; }
    { FastOps op; const double q = op.divz(a, b);
      if (op.bad) slow[4]++; else if (!same_bits(q, a / b)) wrong[4]++; }
  }
  for (int i = 0; i < 5; i++) {
    if (wrong[i]) atomicAdd(counters + 2 * i, wrong[i]);
    if (slow[i]) atomicAdd(counters + 2 * i + 1, slow[i]);
  }
}

}  // namespace
}  // namespace vlct

extern "C" int vlct_selftest_fpops(long long n, unsigned long long seed, int mode,
                                   long long* counters_out)
{
  if (n <= 0 || counters_out == nullptr || mode < 0 || mode > 2)
    return VLCT_ERR_INVALID_CONFIG;
  unsigned long long* d = nullptr;
  if (cudaMalloc(&d, 10 * sizeof(unsigned long long)) != cudaSuccess) return VLCT_ERR_CUDA;
  cudaMemset(d, 0, 10 * sizeof(unsigned long long));
  vlct::k_selftest<<<148 * 8, 256>>>(n, seed, mode, d);
  unsigned long long h[10];
  const cudaError_t err = cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (err != cudaSuccess) return VLCT_ERR_CUDA;
  for (int i = 0; i < 10; i++) counters_out[i] = (long long) h[i];
  return VLCT_OK;
}
