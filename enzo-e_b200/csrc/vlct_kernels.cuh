// vlct_kernels.cuh -- launch interface between the C ABI (vlct_api.cu) and the
// CUDA kernels (vlct_kernels.cu). Plain structs of device pointers; no ownership.
#pragma once

#include <cuda_runtime.h>
#include <cstddef>
#include <string>
#include <vector>

#include "../../include/vlct.h"

namespace vlct {

constexpr int kMaxPassive = VLCT_MAX_PASSIVE;

/// extents of a cell-centred array including ghost zones. A batch of `nrep`
/// equally shaped blocks may be stacked along z (vlct_compute_batch): block r
/// occupies the levels [r*zper, r*zper + mz) of every array (zper = mz + 1, so
/// that the z-face-centred arrays with their mz + 1 levels stack at the same
/// period). Index boxes are always those of ONE block; a launch repeats them
/// nrep times. Stencils never leave a block's own levels, so stacked blocks
/// cannot see each other.
struct Geom {
  int mx, my, mz;
  int nrep = 1, zper = 0;
  __host__ __device__ size_t cells() const
  { return (size_t) mx * (size_t) my * (size_t) mz; }
  /// levels of a stacked cell-centred array
  __host__ __device__ size_t levels() const
  { return nrep > 1 ? (size_t) nrep * (size_t) zper : (size_t) mz; }
};

/// the integration quantities of one state (all cell-centred, shape mz,my,mx)
struct State {
  double *rho, *vx, *vy, *vz, *etot, *eint;
  double *bx, *by, *bz;
  double *sc[kMaxPassive];
};

/// face-centred B: bi[0] (mz,my,mx+1), bi[1] (mz,my+1,mx), bi[2] (mz+1,my,mx)
struct FaceB { double *bi[3]; };

/// fluxes through the faces along one dimension. Every array uses the
/// cell-centred strides (mz,my,mx); entry (k,j,i) is the face between cell
/// (k,j,i) and its +1 neighbour along the sweep dimension.
struct FluxSet {
  double *rho, *mx_, *my_, *mz_, *e;
  double *bx, *by, *bz;      // the component along the sweep is unused (NULL)
  double *eint, *vbar;       // dual energy only
  double *sc[kMaxPassive];
};

/// run-time constants of a handle
struct Params {
  double gamma, theta;
  double density_floor, pressure_floor;
  double de_eta;
  double ggm1;          // (double)(float)(gamma*(gamma-1)), FluidProps.cpp:234
  double igm1;          // 1. / (gamma - 1.), HLLD.hpp:66 (IEEE, host-side)
  int nsc;
  // Passive-scalar fluxes: by default they are never materialised -- the
  // update kernel forms the two face fluxes of a cell per direction itself
  // from the specific scalars and the density fluxes (same expressions, same
  // bits). nsc_flux = nsc brings back the flux arrays written by the sweeps
  // (handle option "scalar_flux_arrays"), which vlct_save_face_fluxes needs.
  int nsc_flux;
  int mhd, de;
  int riemann, recon;
};

struct Scratch {
  State temp;           // temp_integration_map
  FaceB tbi;            // temp_bfieldi_l_
  FluxSet flux[3];
  // specific passive scalars of each stage's input state (one set per stage:
  // the update of a stage reads them two levels behind the sweeps)
  double *prim_sc[2][kMaxPassive];
  double *edge[3];      // edge-centred E (cell strides)
};

/// Optional per-kernel timing (CUDA events on the launching stream). Used by
/// bench.py to measure the dominant kernel's launch duration live.
struct Profiler {
  struct Entry { const char* name; cudaEvent_t beg, end; };
  bool enabled = false;
  std::vector<Entry> pending;
  std::vector<std::string> names;
  std::vector<double> total_ms;
  std::vector<long long> calls;
  void begin(cudaStream_t st, const char* name);
  void end(cudaStream_t st);
  void collect();          // synchronises the recorded events
  void reset();
};

/// what every launcher needs: the stream, the launch counter, the profiler
struct LaunchCtx {
  cudaStream_t st;
  long long* launches;
  Profiler* prof;
  // which cell kernels run as pair kernels (two x-cells per thread, 128-bit
  // accesses): bit 0 edge E, bit 1 face B, bit 2 update; bit 3: edge E of a
  // single block with TMA-staged inputs; bit 4: edge E + face B of a single
  // block in one TMA-staged kernel, which then replaces both; bit 5: the
  // TMA-staged kernels also for blocks too small to fill the chip with their
  // tiles (tests) (option "pair_kernels")
  int pair_mask = 30;
};

/// default of the option "pair_kernels": the fused TMA-staged CT kernel, and
/// where it does not apply the TMA-staged edge E, face B and update as pair
/// kernels (measured; the edge-E pair kernel executes 26 % fewer
/// instructions but needs 128 registers, holds a third of the warps and is
/// 3 % slower);
/// VLCT_PAIR_MASK in the environment overrides it for A/B runs
int default_pair_kernels();

/// A half-open range [lo, hi) along z that a launch is clipped to, in the
/// kernel's own z index space (cells for the cell kernels and the x / y
/// sweeps, faces for the z sweep). Kernels are pure functions of their inputs
/// at a given index, so any partition of a launch's index box into clipped
/// launches gives bit-identical results; vlct_api.cu uses this to run a step
/// as a z-skewed pipeline (HOST staging overlapped with the kernels, ghost
/// exchange overlapped with the interior).
struct ZClip {
  int lo, hi;
};
constexpr ZClip kNoClip{ -(1 << 30), 1 << 30 };

/// specific passive scalars over [s, m-s)^3 (no-op without scalars; the
/// primitive pressure is computed on the fly by the flux kernels)
void launch_primitives(const LaunchCtx& ctx, const Params& P, const Geom& G,
                       const State& cur, const Scratch& S, int stage, int stale,
                       ZClip zc = kNoClip);

/// reconstruct -> fix longitudinal B -> Riemann -> passive fluxes along dim
void launch_flux(const LaunchCtx& ctx, const Params& P, const Geom& G, int dim,
                 int recon, const State& cur, const Scratch& S, int stage,
                 const FaceB& bi_cur, int cur_stale, ZClip zc = kNoClip);

/// constrained transport: edge E, face-B update
/// (step_params: dt/dx, dt/dy, dt/dz, dt of the stage, in device memory)
void launch_ct(const LaunchCtx& ctx, const Params& P, const Geom& G,
               const State& cur, const Scratch& S, const FaceB& bi0,
               const FaceB& bi_out, const double* step_params, int stale,
               ZClip z_edge = kNoClip, ZClip z_face = kNoClip);

/// CFL fold: the update of the LAST stage also evaluates the timestep() of the
/// next cycle -- dual-energy sync, "pressure", CFL minimum into *dt_bits (see
/// launch_timestep) -- on the cells it updates, from registers; the other cells
/// of the clipped levels (the ghost shell, which compute() never updates) go
/// through a small second launch. Single blocks only (Geom::nrep == 1).
struct CflFold {
  double* pressure;
  unsigned long long* dt_bits;
  double width[3];
};

/// centred B + flux divergence + sources + conserved update + floors/sync
void launch_update(const LaunchCtx& ctx, const Params& P, const Geom& G,
                   const State& u0, const State& cur, const State& out,
                   const Scratch& S, const FaceB& bi_out,
                   const double* accel[3], bool gravity,
                   const double* step_params, int stage, int recon, int stale,
                   ZClip zc = kNoClip, const CflFold* cfl = nullptr);

/// the per-stage constants of a step from a host or device dt (see k_step_params)
void launch_step_params(const LaunchCtx& ctx, const double* dt_dev, double dt_host,
                        int nstages, const double* width, double* out);

/// *dt_out = courant * minimum of k_timestep, on the device
void launch_finish_dt(const LaunchCtx& ctx, const unsigned long long* bits,
                      double courant, double* dt_out);

/// DE sync + pressure field + CFL minimum over all cells; *dt_bits receives
/// the bit pattern of the minimum local dt (not yet multiplied by courant).
/// launch_timestep_reset sets *dt_bits to DBL_MAX; launch_timestep folds the
/// cells of z levels [zc.lo, zc.hi) into it (atomicMin), so a block may be
/// processed in several launches.
void launch_timestep_reset(const LaunchCtx& ctx, unsigned long long* dt_bits);
void launch_timestep(const LaunchCtx& ctx, const Params& P, const Geom& G,
                     const State& u, double* pressure, const double* width,
                     unsigned long long* dt_bits, ZClip zc = kNoClip);

/// periodic self-refresh of one field along one axis
/// every field of a block, for the one-launch periodic wrap along an axis
constexpr int kMaxWrapFields = 16 + kMaxPassive;
struct WrapTable {
  double* p[kMaxWrapFields];
  int face[kMaxWrapFields];   // -1: cell-centred; 0/1/2: face-centred along x/y/z
  int count;
};
void launch_wrap_axis_all(const LaunchCtx& ctx, const WrapTable& T, int mz, int my,
                          int mx, int axis, int n, int g);

/// outflow / reflecting boundary of one field on one face of the domain
struct BoundaryTable {
  double* p[kMaxWrapFields];
  int face[kMaxWrapFields];
  double sign[kMaxWrapFields];      // reflecting: +-1; inflow: the value
  int count;
};
void launch_boundary_axis(const LaunchCtx& ctx, const BoundaryTable& T, int mz, int my,
                          int mx, int axis, int n, int g, int side, int type);

/// out[(i1, i0)] = *dtdx * flux[.. at ..]: one face of the block for the
/// flux-correction output (dim: normal axis; n0/g0, n1/g1: active size and
/// ghost depth of the faster / slower transverse axis)
void launch_face_flux(const LaunchCtx& ctx, const Geom& G, const double* flux,
                      const double* dtdx, double* out, int dim, int at, int n0,
                      int n1, int g0, int g1);

/// batch of blocks <-> their stacked array (ptrs: device table of nblocks
/// device pointers, `count` elements each, stacked `stride` elements apart)
void launch_batch_copy(const LaunchCtx& ctx, double* stacked, double* const* ptrs,
                       int nblocks, size_t count, size_t stride, bool to_stacked,
                       bool over_pcie = false);

/// halo slab pack / unpack of one field along one axis
/// lo..lo+g: range along the axis; the slab spans the full other extents
struct SlabTable {
  double* p[kMaxWrapFields];
  int face[kMaxWrapFields];
  int lo[kMaxWrapFields];           // first layer of the slab along the axis
  long long off[kMaxWrapFields];    // offset of the field's slab in the buffer
  int count;
};
void launch_slab_copy_all(const LaunchCtx& ctx, const SlabTable& T, int mz, int my,
                          int mx, int axis, int width, double* buffer, bool pack);

}  // namespace vlct
