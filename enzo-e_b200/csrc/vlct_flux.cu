// vlct_flux.cu -- the flux kernels of the VL+CT update (sm_100a, fp64).
//
// One launch per sweep direction fuses, per face,
//   primitives   fluid-props/EnzoPhysicsFluidProps.cpp:64-138,
//                fluid-props/EnzoComputePressure.cpp:82-198
//   reconstruct  toolkit/EnzoReconstructorNN.cpp:14-48,
//                toolkit/EnzoReconstructorPLM.hpp:166-248
//   B fix        toolkit/EnzoBfieldMethodCT.cpp:122-166
//   Riemann      riemann/EnzoRiemannImpl.hpp:266-338 (HLLD / HLLE / HLLC)
//   passive flux riemann/EnzoRiemannUtils.hpp:267-314
// so that the reference's primitive / priml / primr arrays never exist. These
// kernels are bound by the FP64 pipe (~740 DP instructions per HLLD face), not
// by HBM, so the design goal is: every cell's primitives and every limited
// slope are evaluated ONCE per sweep, and enough warps stay resident to keep
// the DP pipe busy.
//
//   k_flux_x      sweep along x (the contiguous axis). A warp owns 32
//                 consecutive faces of one row; the cells' primitives and
//                 slopes are exchanged between lanes through a warp-private
//                 shared-memory strip (no block barrier).
//   k_flux_march  sweeps along y and z. A thread owns one (x, other) column
//                 and marches along the sweep axis with a rolling window of
//                 primitives in registers; a warp reads 32 consecutive x, so
//                 every load and store is coalesced and each array is read
//                 exactly once. The cells of the next faces arrive through a
//                 cp.async ring of thread-private shared-memory slots.
//
// Results are bit-identical to the reference's value-safe CPU build: the same
// expressions in the same order (vlct_physics.cuh), compiled with -fmad=false.
#include "vlct_device.cuh"
#include "vlct_physics.cuh"

namespace vlct {

namespace {

// slots of a primitive state in the permuted (i,j,k) frame of the sweep
enum { W_RHO = 0, W_VI = 1, W_VJ = 2, W_VK = 3, W_P = 4, W_BJ = 5, W_BK = 6 };

template <bool MHD> struct NVars { static constexpr int n = MHD ? 7 : 5; };

/// Primitives of a cell in the frame of sweep DIM from its field values, the
/// pressure computed on the fly (EnzoComputePressure.cpp:82-198; same operand
/// order as the reference). e = internal_energy with DE, else total_energy.
template <int DIM, bool MHD, bool DE>
__device__ __forceinline__ void
cell_primitives(const Params& P, double rho, const double (&v)[3],
                const double (&b)[3], double e, double (&w)[NVars<MHD>::n])
{
  constexpr int JD = (DIM + 1) % 3, KD = (DIM + 2) % 3;
  const double gm1 = P.gamma - 1.0;
  double p;
  if (DE) {
    p = gm1 * rho * e;
  } else {
    const double ke = 0.5 * (v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    double me_den = 0.;
    if (MHD) me_den = 0.5 * (b[0] * b[0] + b[1] * b[1] + b[2] * b[2]);
    p = gm1 * (rho * (e - ke) - me_den);
  }
  w[W_RHO] = rho; w[W_VI] = v[DIM]; w[W_VJ] = v[JD]; w[W_VK] = v[KD];
  w[W_P] = p;
  if (MHD) { w[W_BJ] = b[JD]; w[W_BK] = b[KD]; }
}

/// Primitives of cell c in the frame of sweep DIM
template <int DIM, bool MHD, bool DE>
__device__ __forceinline__ void load_cell(const Params& P, const State& u,
                                          size_t c, double (&w)[NVars<MHD>::n])
{
  double v[3], b[3] = { 0., 0., 0. };
  const double rho = __ldg(u.rho + c);
  v[0] = __ldg(u.vx + c); v[1] = __ldg(u.vy + c); v[2] = __ldg(u.vz + c);
  if (MHD) { b[0] = __ldg(u.bx + c); b[1] = __ldg(u.by + c); b[2] = __ldg(u.bz + c); }
  const double e = DE ? __ldg(u.eint + c) : __ldg(u.etot + c);
  cell_primitives<DIM, MHD, DE>(P, rho, v, b, e, w);
}

__device__ __forceinline__ void prefetch_l1(const double* p)
{
#ifndef VLCT_NO_PREFETCH
  asm volatile("prefetch.global.L1 [%0];" :: "l"(p));
#endif
}

/// pull the lines of cell c into L1 ahead of the load_cell that will need them
template <bool MHD, bool DE>
__device__ __forceinline__ void prefetch_cell(const State& u, size_t c)
{
  prefetch_l1(u.rho + c);
  prefetch_l1(u.vx + c); prefetch_l1(u.vy + c); prefetch_l1(u.vz + c);
  if (MHD) { prefetch_l1(u.bx + c); prefetch_l1(u.by + c); prefetch_l1(u.bz + c); }
  if (DE) prefetch_l1(u.eint + c); else prefetch_l1(u.etot + c);
}

template <bool MHD>
__device__ __forceinline__ void apply_floors(const Params& P,
                                             double (&w)[NVars<MHD>::n])
{
  w[W_RHO] = apply_floor(w[W_RHO], P.density_floor);
  w[W_P] = apply_floor(w[W_P], P.pressure_floor);
}

/// L/R states of a passive scalar straight from the specific-scalar array
template <int RECON>
__device__ __forceinline__ void
recon_pair(const double* __restrict__ a, size_t c, ptrdiff_t sd, double theta,
           double& wl, double& wr)
{
  if (RECON == RECON_NN) {
    wl = __ldg(a + c);
    wr = __ldg(a + c + sd);
  } else {
    const double w0 = __ldg(a + c - sd), w1 = __ldg(a + c);
    const double w2 = __ldg(a + c + sd), w3 = __ldg(a + c + 2 * sd);
    const double dvl = limited_slope<RECON>(w0, w1, w2, theta);
    const double dvr = limited_slope<RECON>(w1, w2, w3, theta);
    wl = w1 + dvl * 0.5;
    wr = w2 - dvr * 0.5;
  }
}

/// Riemann solve of one face and store of its fluxes at cell index c (the
/// face's left cell).
template <int DIM, int RECON, int SOLVER, bool DE>
__device__ __forceinline__ void
solve_and_store(const Params& P, const double* wl_, const double* wr_,
                double blong, const FluxSet& F, const ScalarPtrs& spec,
                size_t c, ptrdiff_t sd)
{
  constexpr bool MHD = (SOLVER != SOLVER_HLLC);
  constexpr int JD = (DIM + 1) % 3, KD = (DIM + 2) % 3;
  Prim wl, wr;
  wl.rho = wl_[W_RHO]; wl.vi = wl_[W_VI]; wl.vj = wl_[W_VJ]; wl.vk = wl_[W_VK];
  wl.p = wl_[W_P];
  wr.rho = wr_[W_RHO]; wr.vi = wr_[W_VI]; wr.vj = wr_[W_VJ]; wr.vk = wr_[W_VK];
  wr.p = wr_[W_P];
  if (MHD) {
    wl.bi = blong; wl.bj = wl_[W_BJ]; wl.bk = wl_[W_BK];
    wr.bi = blong; wr.bj = wr_[W_BJ]; wr.bk = wr_[W_BK];
  } else {
    wl.bi = wl.bj = wl.bk = 0.;
    wr.bi = wr.bj = wr.bk = 0.;
  }
  Flux f;
  riemann_solve<SOLVER, DE>(P.gamma, P.igm1, wl, wr, f);

  double* fm[3] = { F.mx_, F.my_, F.mz_ };
  F.rho[c] = f.rho;
  fm[DIM][c] = f.mi;
  fm[JD][c] = f.mj;
  fm[KD][c] = f.mk;
  F.e[c] = f.e;
  if (MHD) {
    double* fb[3] = { F.bx, F.by, F.bz };
    fb[JD][c] = f.bj;
    fb[KD][c] = f.bk;
  }
  if (DE) {
    F.eint[c] = f.eint;
    F.vbar[c] = f.vbar;
  }
  for (int s = 0; s < P.nsc_flux; s++) {
    double sl, sr;
    recon_pair<RECON>(spec.p[s], c, sd, P.theta, sl, sr);
    F.sc[s][c] = passive_flux(sl, sr, f.rho);
  }
}

// The cells a column needs next are fetched kRingAhead faces ahead with
// cp.async into a ring of thread-private shared-memory slots: no registers are
// held across the Riemann solve for them, and the loads that feed the slopes
// hit shared memory instead of waiting for L2 / HBM (the marching kernels spent
// a quarter of their stall cycles on that scoreboard).
constexpr int kRingAhead = 3, kRingDepth = 4;    // depth: a power of two > ahead
template <bool MHD> struct RingVars { static constexpr int n = MHD ? 9 : 5; };

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* src)
{
  const unsigned d = (unsigned) __cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit()
{ asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait()
{ asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

/// start the copy of cell c (and of the longitudinal face field at fb) into
/// one ring slot; slot points at this thread's first entry, T = entries per
/// variable
template <bool MHD, bool DE, int T>
__device__ __forceinline__ void ring_issue(double* slot, const State& u, size_t c,
                                           const double* bi, size_t fb)
{
  cp_async8(slot, u.rho + c);
  cp_async8(slot + T, u.vx + c);
  cp_async8(slot + 2 * T, u.vy + c);
  cp_async8(slot + 3 * T, u.vz + c);
  cp_async8(slot + 4 * T, (DE ? u.eint : u.etot) + c);
  if (MHD) {
    cp_async8(slot + 5 * T, u.bx + c);
    cp_async8(slot + 6 * T, u.by + c);
    cp_async8(slot + 7 * T, u.bz + c);
    cp_async8(slot + 8 * T, bi + fb);
  }
}

/// primitives of the cell held by a ring slot (+ the face field that came along)
template <int DIM, bool MHD, bool DE, int T>
__device__ __forceinline__ void ring_cell(const Params& P, const double* slot,
                                          double (&w)[NVars<MHD>::n], double& blong)
{
  double v[3], b[3] = { 0., 0., 0. };
  const double rho = slot[0];
  v[0] = slot[T]; v[1] = slot[2 * T]; v[2] = slot[3 * T];
  const double e = slot[4 * T];
  if (MHD) { b[0] = slot[5 * T]; b[1] = slot[6 * T]; b[2] = slot[7 * T]; blong = slot[8 * T]; }
  cell_primitives<DIM, MHD, DE>(P, rho, v, b, e, w);
}

// ---------------------------------------------------------------------------
// sweep along x: one warp = 32 consecutive faces of one row
// ---------------------------------------------------------------------------
#ifndef VLCT_FLUX_MINBLOCKS
#define VLCT_FLUX_MINBLOCKS 4     // 128-thread blocks per SM => <=128 registers
#endif
#ifndef VLCT_FLUX_X_MINBLOCKS     // (separate knobs for A/B builds)
#define VLCT_FLUX_X_MINBLOCKS VLCT_FLUX_MINBLOCKS
#endif
#ifndef VLCT_FLUX_MARCH_MINBLOCKS
#define VLCT_FLUX_MARCH_MINBLOCKS VLCT_FLUX_MINBLOCKS
#endif
// The PLM marches carry two more cells of state through the Riemann solve: at
// 3 blocks per SM (168 registers) they do not spill and run 2-5 % faster than
// at 4 (measured at 512^3: y 7.19 -> 7.03 ms, z 7.48 -> 7.13 ms); every other
// flux kernel is faster at 4 blocks.
#ifndef VLCT_FLUX_MARCH_PLM_MINBLOCKS
#define VLCT_FLUX_MARCH_PLM_MINBLOCKS 3
#endif
constexpr int kXWarps = 4;
#ifndef VLCT_XROWS
#define VLCT_XROWS 8
#endif
constexpr int kXRows = VLCT_XROWS;   // rows a warp walks through, one after another

template <int RECON, int SOLVER, bool DE>
__global__ void __launch_bounds__(kXWarps * 32, VLCT_FLUX_X_MINBLOCKS)
k_flux_x(const __grid_constant__ Params P, const __grid_constant__ Geom G,
         const __grid_constant__ State u, const __grid_constant__ ScalarPtrs spec,
         const double* __restrict__ bi, const __grid_constant__ FluxSet F,
         const __grid_constant__ Box box)
{
  constexpr bool MHD = (SOLVER != SOLVER_HLLC);
  constexpr int NV = NVars<MHD>::n;
  constexpr bool PLM = (RECON != RECON_NN);
  constexpr int H = PLM ? 1 : 0;          // cells needed left of the first face
  // Faces per warp. With PLM the 32 lanes own 32 cells (one limited slope
  // each) and solve the 31 faces between them: a 32nd face would need a 33rd
  // slope, evaluated by one lane while 31 wait (+20 % instructions).
  constexpr int FPW = PLM ? 31 : 32;
  // strip position p <-> cell f0 - H + p; PLM needs cells f0-1 .. f0+32
  __shared__ double sW[kXWarps][NV][36];
  __shared__ double sD[kXWarps][PLM ? NV : 1][PLM ? 32 : 1];

  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int nfx = box.hi[0] - box.lo[0];
  const int wpr = (nfx + FPW - 1) / FPW;  // warps per row
  const int nyb = box.hi[1] - box.lo[1];
  // rows of all stacked blocks (G.nrep = 1 for a single block)
  const unsigned nzs = (unsigned) (box.hi[2] - box.lo[2]) * (unsigned) G.nrep;
  // (32-bit index arithmetic: check_block / check_batch in vlct_api.cu refuse
  // blocks and batches of 2^31 or more cells, so the counts fit)
  const unsigned nrows = (unsigned) nyb * nzs;
  const unsigned ngroups = (nrows + kXRows - 1) / kXRows;
  const unsigned gw = blockIdx.x * kXWarps + w;
  if (gw >= (unsigned) wpr * ngroups) return;
  // consecutive warps take consecutive segments of the same rows
  const int seg = (int) (gw % (unsigned) wpr);
  const unsigned row0 = (gw / (unsigned) wpr) * kXRows;
  const unsigned row1 = (row0 + kXRows < nrows) ? row0 + kXRows : nrows;
  const int f0 = box.lo[0] + seg * FPW;
  const int nf = min(FPW, box.hi[0] - f0);
  const int i = f0 + lane;                // own cell = left cell of own face

  // the cells beyond the warp's 32: one or two lanes fetch an extra cell each
  int ecell = -1, epos = 0;
  if (PLM) {
    if (lane == 0)      { ecell = f0 - 1;  epos = 0; }
    else if (lane == 1) { ecell = f0 + 32; epos = 33; }
  } else if (lane == 0) { ecell = f0 + 32; epos = 32; }
  if (ecell >= G.mx) ecell = -1;

  // (j,k) of the current and of the next row (k: level in the stacked arrays)
  int j = box.lo[1] + (int) (row0 % (unsigned) nyb);
  int kl, k;
  unstack(G, box, row0 / (unsigned) nyb, kl, k);
#pragma unroll 1
  for (unsigned row = row0; row < row1; row++) {
    const size_t rowbase = cidx(G, k, j, 0);
    const int jc = j, kc = k;
    if (++j == box.hi[1]) {
      j = box.lo[1];
      unstack(G, box, (row + 1) / (unsigned) nyb, kl, k);
    }
    double W[NV];
#pragma unroll
    for (int v = 0; v < NV; v++) W[v] = 0.;
    double blong = 0.;
    if (row + 1 < row1 && i < G.mx) {
      prefetch_cell<MHD, DE>(u, cidx(G, k, j, i));
      if (MHD) prefetch_l1(bi + fidx(G, 0, k, j, i + 1));
    }
    // face f of the sweep <-> index f+1 of the face-centred array; loaded here,
    // with the cell, ahead of the slopes, so that its latency is hidden (x sweeps -6 %)
    if (MHD && lane < nf) blong = __ldg(bi + fidx(G, 0, kc, jc, i + 1));
    if (i < G.mx) {
      load_cell<0, MHD, DE>(P, u, rowbase + i, W);
#pragma unroll
      for (int v = 0; v < NV; v++) sW[w][v][lane + H] = W[v];
    }
    if (ecell >= 0) {
      double E[NV];
      load_cell<0, MHD, DE>(P, u, rowbase + ecell, E);
#pragma unroll
      for (int v = 0; v < NV; v++) sW[w][v][epos] = E[v];
    }
    __syncwarp();

    double wl[NV], wr[NV];
    if (PLM) {
      // limited slope of the own cell, once; the neighbour's comes from its lane
      double dv[NV];
#pragma unroll
      for (int v = 0; v < NV; v++) {
        dv[v] = limited_slope<RECON>(sW[w][v][lane], W[v], sW[w][v][lane + 2], P.theta);
        sD[w][v][lane] = dv[v];
      }
      __syncwarp();
#pragma unroll
      for (int v = 0; v < NV; v++) {
        wl[v] = W[v] + dv[v] * 0.5;
        wr[v] = sW[w][v][lane + 2] - sD[w][v][(lane + 1) & 31] * 0.5;
      }
      apply_floors<MHD>(P, wl);
      apply_floors<MHD>(P, wr);
    } else {
#pragma unroll
      for (int v = 0; v < NV; v++) { wl[v] = W[v]; wr[v] = sW[w][v][lane + 1]; }
    }
    __syncwarp();   // strip is rewritten by the next row

    if (lane < nf) {
      const size_t c = rowbase + i;
      solve_and_store<0, RECON, SOLVER, DE>(P, wl, wr, blong, F, spec, c, 1);
    }
  }
}

// ---------------------------------------------------------------------------
// sweeps along y and z: one thread = one column, marching along the sweep
// ---------------------------------------------------------------------------
constexpr int kMarchThreads = 128;


template <int DIM, int RECON, int SOLVER, bool DE>
__global__ void __launch_bounds__(kMarchThreads,
                                  (RECON != RECON_NN && SOLVER == SOLVER_HLLD)
                                      ? VLCT_FLUX_MARCH_PLM_MINBLOCKS
                                      : VLCT_FLUX_MARCH_MINBLOCKS)
k_flux_march(const __grid_constant__ Params P, const __grid_constant__ Geom G,
             const __grid_constant__ State u,
             const __grid_constant__ ScalarPtrs spec,
             const double* __restrict__ bi, const __grid_constant__ FluxSet F,
             const __grid_constant__ Box box, const int chunk)
{
  static_assert(DIM == 1 || DIM == 2, "marching sweeps are y and z");
  constexpr bool MHD = (SOLVER != SOLVER_HLLC);
  constexpr int NV = NVars<MHD>::n;
  constexpr bool PLM = (RECON != RECON_NN);
  constexpr int OD = (DIM == 1) ? 2 : 1;   // the non-x, non-sweep axis

  // Stacked blocks (G.nrep > 1) repeat the box along z: for the y sweep z is
  // the column axis `o`, for the z sweep it is the march axis.
  const int nxb = box.hi[0] - box.lo[0];
  const unsigned nob = (unsigned) (box.hi[OD] - box.lo[OD]) *
                       (DIM == 1 ? (unsigned) G.nrep : 1u);
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (unsigned) nxb * nob) return;
  const int i = box.lo[0] + (int) (t % (unsigned) nxb);
  int o, zshift = 0;
  if (DIM == 1) {
    int ol;
    unstack(G, box, t / (unsigned) nxb, ol, o);
  } else {
    o = box.lo[OD] + (int) (t / (unsigned) nxb);
    // blockIdx.y = (block of the batch) * chunks_per_block + chunk
    const unsigned cpb = gridDim.y / (unsigned) G.nrep;
    zshift = (int) (blockIdx.y / cpb) * G.zper;
  }
  const int cy = (DIM == 1) ? (int) blockIdx.y
                            : (int) (blockIdx.y % (gridDim.y / (unsigned) G.nrep));
  const int fl0 = box.lo[DIM] + cy * chunk;          // inside the block
  const int f0 = fl0 + zshift;                       // in the stacked arrays
  const int f1 = min(fl0 + chunk, box.hi[DIM]) + zshift;
  const int j = (DIM == 1) ? f0 : o, k = (DIM == 1) ? o : f0;
  // cell stride along the sweep; the face-centred array of component DIM has
  // the same stride along DIM, and face f sits at index f+1
  const ptrdiff_t sd = (DIM == 1) ? (ptrdiff_t) G.mx
                                  : (ptrdiff_t) G.mx * (ptrdiff_t) G.my;
  size_t c = cidx(G, k, j, i);
  size_t fb = 0;
  if (MHD) fb = fidx(G, DIM, k + (DIM == 2), j + (DIM == 1), i);

  double Wm[NV], Wc[NV], wl[NV];
  if (PLM) {
    double A[NV];
    load_cell<DIM, MHD, DE>(P, u, c - sd, A);
    load_cell<DIM, MHD, DE>(P, u, c, Wm);
    load_cell<DIM, MHD, DE>(P, u, c + sd, Wc);
#pragma unroll
    for (int v = 0; v < NV; v++)
      wl[v] = Wm[v] + limited_slope<RECON>(A[v], Wm[v], Wc[v], P.theta) * 0.5;
    apply_floors<MHD>(P, wl);
  } else {
    load_cell<DIM, MHD, DE>(P, u, c, Wc);
  }

  const int mdim = (DIM == 1) ? G.my : (int) G.levels();
  constexpr int kFirst = PLM ? 2 : 1;       // the cell an iteration adds: c + kFirst sd
  constexpr int kSlot = RingVars<MHD>::n * kMarchThreads;      // doubles per slot
  __shared__ double ring[kRingDepth * kSlot];
  double* const ring0 = ring + threadIdx.x;
  // cells this chunk really needs: up to f1 - 1 + kFirst (prefetching further,
  // into the next chunk's cells, cost 3 rows of DRAM traffic per 64-face chunk)
  const int cell_end = min(mdim, f1 + kFirst);
#pragma unroll
  for (int a = 0; a < kRingAhead; a++) {
    if (f0 + kFirst + a < cell_end)
      ring_issue<MHD, DE, kMarchThreads>(ring0 + a * kSlot, u, c + (kFirst + a) * sd, bi,
                                         fb + a * sd);
    cp_async_commit();
  }
  int slot = 0;
#pragma unroll 1
  for (int f = f0; f < f1; f++, c += sd, fb += sd) {
    double Wn[NV], wr[NV], wl_next[NV];
    double blong = 0.;
    if (f + kFirst + kRingAhead < cell_end)
      ring_issue<MHD, DE, kMarchThreads>(
          ring0 + ((slot + kRingAhead) & (kRingDepth - 1)) * kSlot, u,
          c + (kFirst + kRingAhead) * sd, bi, fb + kRingAhead * sd);
    cp_async_commit();
    cp_async_wait<kRingAhead>();            // the group of this face has landed
    ring_cell<DIM, MHD, DE, kMarchThreads>(P, ring0 + slot * kSlot, Wn, blong);
    slot = (slot + 1) & (kRingDepth - 1);
    if (PLM) {
#pragma unroll
      for (int v = 0; v < NV; v++) {
        const double h = limited_slope<RECON>(Wm[v], Wc[v], Wn[v], P.theta) * 0.5;
        wr[v] = Wc[v] - h;
        wl_next[v] = Wc[v] + h;
      }
      apply_floors<MHD>(P, wr);
      apply_floors<MHD>(P, wl_next);
    } else {
#pragma unroll
      for (int v = 0; v < NV; v++) { wl[v] = Wc[v]; wr[v] = Wn[v]; }
    }
    solve_and_store<DIM, RECON, SOLVER, DE>(P, wl, wr, blong, F, spec, c, sd);
#pragma unroll
    for (int v = 0; v < NV; v++) {
      if (PLM) { wl[v] = wl_next[v]; Wm[v] = Wc[v]; }
      Wc[v] = Wn[v];
    }
  }
}

// ---------------------------------------------------------------------------
// dispatch
// ---------------------------------------------------------------------------
struct FluxLaunch {
  cudaStream_t st;
  Params P; Geom G; State cur; ScalarPtrs spec; const double* bi; FluxSet F;
  Box box;
};

template <int DIM, int RECON, int SOLVER, bool DE>
void flux_go(const FluxLaunch& L)
{
  const Box& b = L.box;
  if constexpr (DIM == 0) {
    constexpr int fpw = (RECON != RECON_NN) ? 31 : 32;   // k_flux_x: FPW
    const long long wpr = (b.hi[0] - b.lo[0] + fpw - 1) / fpw;
    const long long rows = (long long) (b.hi[1] - b.lo[1]) * (b.hi[2] - b.lo[2]) *
                           L.G.nrep;
    const long long warps = wpr * ((rows + kXRows - 1) / kXRows);
    const unsigned grid = (unsigned) ((warps + kXWarps - 1) / kXWarps);
    k_flux_x<RECON, SOLVER, DE><<<grid, kXWarps * 32, 0, L.st>>>(
        L.P, L.G, L.cur, L.spec, L.bi, L.F, b);
  } else {
    constexpr int D = DIM;
    const int od = (D == 1) ? 2 : 1;
    const long long cols = (long long) (b.hi[0] - b.lo[0]) * (b.hi[od] - b.lo[od]) *
                           (D == 1 ? L.G.nrep : 1);
    const int nf = b.hi[D] - b.lo[D];
    // chunks along the march: enough blocks for many waves, long enough that
    // the 2-3 warm-up cells per chunk stay a small overhead
    int chunk = 64;
    if (nf < 2 * chunk) chunk = nf;
    const unsigned gx = (unsigned) ((cols + kMarchThreads - 1) / kMarchThreads);
    const unsigned gy = (unsigned) ((nf + chunk - 1) / chunk) *
                        (D == 2 ? (unsigned) L.G.nrep : 1u);
    // 4 resident blocks need 4 x (36 KB ring + 1 KB) of the SM's shared memory
    // (a per-device attribute: remembered per device, one handle = one device
    // = one host thread at a time)
    static bool carveout_set[64] = {};
    int device = 0;
    cudaGetDevice(&device);
    if (device < 0 || device >= 64 || !carveout_set[device]) {
      cudaFuncSetAttribute(k_flux_march<D, RECON, SOLVER, DE>,
                           cudaFuncAttributePreferredSharedMemoryCarveout, 75);
      if (device >= 0 && device < 64) carveout_set[device] = true;
    }
    k_flux_march<D, RECON, SOLVER, DE><<<dim3(gx, gy), kMarchThreads, 0, L.st>>>(
        L.P, L.G, L.cur, L.spec, L.bi, L.F, b, chunk);
  }
}

template <int DIM, int RECON, int SOLVER>
void flux_de(const FluxLaunch& L, bool de)
{
  if (de) flux_go<DIM, RECON, SOLVER, true>(L);
  else    flux_go<DIM, RECON, SOLVER, false>(L);
}

template <int DIM, int RECON>
void flux_solver(const FluxLaunch& L, int solver, bool de)
{
  switch (solver) {
  case VLCT_RIEMANN_HLLD: flux_de<DIM, RECON, SOLVER_HLLD>(L, de); break;
  case VLCT_RIEMANN_HLLE: flux_de<DIM, RECON, SOLVER_HLLE>(L, de); break;
  default:                flux_de<DIM, RECON, SOLVER_HLLC>(L, de); break;
  }
}

template <int DIM>
void flux_recon(const FluxLaunch& L, int recon, int solver, bool de)
{
  switch (recon) {
  case VLCT_RECON_NN:         flux_solver<DIM, RECON_NN>(L, solver, de); break;
  case VLCT_RECON_PLM_ATHENA: flux_solver<DIM, RECON_PLM_ATHENA>(L, solver, de); break;
  default:                    flux_solver<DIM, RECON_PLM_ENZO>(L, solver, de); break;
  }
}

}  // namespace

void launch_flux(const LaunchCtx& ctx, const Params& P, const Geom& G, int dim,
                 int recon, const State& cur, const Scratch& S, int stage,
                 const FaceB& bi_cur, int cs, ZClip zc)
{
  // non-stale faces: [cs, f-cs) on every axis of the face-shaped array
  Box box = full_box(G, cs);
  box.hi[dim] -= 1;
  if (!clip_z(box, zc)) return;
  FluxLaunch L{ ctx.st, P, G, cur, scalar_ptrs(S.prim_sc[stage], P.nsc_flux),
                P.mhd ? bi_cur.bi[dim] : nullptr, S.flux[dim], box };
  const bool de = P.de != 0;
  static const char* const names[2][3] = {
    { "k_flux_x_nn", "k_flux_y_nn", "k_flux_z_nn" },
    { "k_flux_x_plm", "k_flux_y_plm", "k_flux_z_plm" } };
  ScopedLaunch sl(ctx, names[recon == VLCT_RECON_NN ? 0 : 1][dim]);
  switch (dim) {
  case 0:  flux_recon<0>(L, recon, P.riemann, de); break;
  case 1:  flux_recon<1>(L, recon, P.riemann, de); break;
  default: flux_recon<2>(L, recon, P.riemann, de); break;
  }
}

}  // namespace vlct
