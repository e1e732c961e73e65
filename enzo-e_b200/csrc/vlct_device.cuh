// vlct_device.cuh -- helpers shared by the kernel translation units
// (vlct_flux.cu, vlct_kernels.cu): index boxes, array indexing, launch
// bookkeeping. Internal to csrc/.
#pragma once

#include "vlct_kernels.cuh"

namespace vlct {

struct Box { int lo[3], hi[3]; };  // [lo,hi) along x,y,z

/// counts the launch and, when profiling is on, brackets it with CUDA events
struct ScopedLaunch {
  const LaunchCtx& c;
  ScopedLaunch(const LaunchCtx& ctx, const char* name) : c(ctx)
  { if (c.prof && c.prof->enabled) c.prof->begin(c.st, name); }
  ~ScopedLaunch()
  { if (c.prof && c.prof->enabled) c.prof->end(c.st); ++*c.launches; }
};

inline bool empty(const Box& b)
{ return b.hi[0] <= b.lo[0] || b.hi[1] <= b.lo[1] || b.hi[2] <= b.lo[2]; }

inline Box full_box(const Geom& G, int s)
{
  Box b;
  b.lo[0] = b.lo[1] = b.lo[2] = s;
  b.hi[0] = G.mx - s; b.hi[1] = G.my - s; b.hi[2] = G.mz - s;
  return b;
}

/// clip a box along z; false if nothing is left
inline bool clip_z(Box& b, const ZClip& zc)
{
  if (b.lo[2] < zc.lo) b.lo[2] = zc.lo;
  if (b.hi[2] > zc.hi) b.hi[2] = zc.hi;
  return !empty(b);
}

/// z levels a launch covers over all stacked blocks
inline unsigned stacked_nz(const Geom& G, const Box& b)
{ return (unsigned) (b.hi[2] - b.lo[2]) * (unsigned) G.nrep; }

/// stacked level number kk in [0, nz * nrep) -> the level inside its block
/// (for index-box tests) and the level in the stacked arrays (for indexing)
__device__ __forceinline__ void unstack(const Geom& G, const Box& b, unsigned kk,
                                        int& k_local, int& k_global)
{
  const unsigned nz = (unsigned) (b.hi[2] - b.lo[2]);
  const unsigned r = kk / nz;
  k_local = b.lo[2] + (int) (kk - r * nz);
  k_global = k_local + (int) r * G.zper;
}

/// What a single-block cell kernel needs of the geometry. Kernels for one
/// block take this instead of Geom: the two extra ints of Geom shift the
/// kernel-parameter layout and cost k_edge_efield 4 registers = one resident
/// block per SM (4.4 -> 5.1 ms at 512^3).
struct GeomLite {
  int mx, my, mz;
};
template <bool STACKED> struct GeomFor { typedef GeomLite type; };
template <> struct GeomFor<true> { typedef Geom type; };
inline GeomLite lite(const Geom& G) { return GeomLite{ G.mx, G.my, G.mz }; }

template <class GEOM>
__device__ __forceinline__ size_t cidx(const GEOM& G, int k, int j, int i)
{ return ((size_t) k * (size_t) G.my + (size_t) j) * (size_t) G.mx + (size_t) i; }

/// index into the face-centred array of component d
template <class GEOM>
__device__ __forceinline__ size_t fidx(const GEOM& G, int d, int k, int j, int i)
{
  const size_t n2 = (size_t) G.mx + (d == 0), n1 = (size_t) G.my + (d == 1);
  return ((size_t) k * n1 + (size_t) j) * n2 + (size_t) i;
}

struct ScalarPtrs { double* p[kMaxPassive]; };

inline ScalarPtrs scalar_ptrs(double* const* p, int n)
{
  ScalarPtrs s;
  for (int i = 0; i < kMaxPassive; i++) s.p[i] = (i < n) ? p[i] : nullptr;
  return s;
}

}  // namespace vlct
