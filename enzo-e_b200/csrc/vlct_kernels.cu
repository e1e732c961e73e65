// vlct_kernels.cu -- sm_100a kernels of the VL+CT block update.
//
// One stage of the integrator (EnzoMHDIntegratorStageCommands::
// compute_update_stage, hydro-mhd/EnzoMHDIntegratorStageCommands.cpp:102-203)
// is executed as
//
//   k_specific_scalars   passive scalars / density (only when scalars exist)
//   k_flux_x, k_flux_march<y|z>  (vlct_flux.cu) fused  primitives ->
//                  reconstruct -> longitudinal-B fix -> Riemann -> passive
//                  fluxes; nothing but the fluxes reaches HBM
//   k_edge_efield  fused  cell-centred E -> upwind weights -> edge E
//   k_face_bfield  CT update of the three face-centred components
//   k_update       fused  centred B -> flux divergence (+ dual-energy source,
//                  gravity) -> conserved update -> floors / dual-energy sync
//
// where the reference makes ~25 separate full-array passes. The arrays
// priml/primr/dUcons/weights/centre-E of the reference never exist here.
//
// Regions: every kernel works on the reference's stale-depth-trimmed index
// boxes so that even the ghost-zone content left behind is identical.
// All arithmetic is fp64; compile with -fmad=false for bit parity.
#include "vlct_device.cuh"
#include "vlct_physics.cuh"

#include <cuda.h>
#include <cfloat>
#include <map>
#include <mutex>
#include <tuple>
#include <cstdint>
#include <cstdlib>

namespace vlct {

namespace {

// Cell-parallel kernels: a block owns kBlock consecutive entries of the
// flattened (y,x) plane of the box (no partially filled rows: 514 of 518 wide
// boxes would otherwise waste 10-20 % of the threads), blockIdx.y is z.
constexpr int kBlock = 256;

inline dim3 grid_for(const Geom& G, const Box& b)
{
  const unsigned nx = b.hi[0] - b.lo[0], ny = b.hi[1] - b.lo[1];
  return dim3((nx * ny + kBlock - 1) / kBlock, stacked_nz(G, b), 1);
}

// i, j: x / y index; kl: z level inside the block (what index boxes are tested
// against); k: z level in the (possibly stacked) arrays
// (kernels are instantiated for a single block, STACKED = false: kl == k, one
// register and one division less -- k_edge_efield is sensitive to both -- and
// for a stacked batch)
#define VLCT_THREAD_IN_BOX(G, box, i, j, kl, k)                                \
  const unsigned nxb__ = (box).hi[0] - (box).lo[0];                            \
  const unsigned t__ = blockIdx.x * kBlock + threadIdx.x;                      \
  if (t__ >= nxb__ * (unsigned) ((box).hi[1] - (box).lo[1])) return;           \
  const int i = (box).lo[0] + (int) (t__ % nxb__);                             \
  const int j = (box).lo[1] + (int) (t__ / nxb__);                             \
  int kl, k;                                                                   \
  if constexpr (STACKED) unstack((G), (box), blockIdx.y, kl, k);               \
  else kl = k = (box).lo[2] + (int) blockIdx.y;

// ---------------------------------------------------------------------------
// specific passive scalars: EnzoPhysicsFluidProps::primitive_from_integration
// (fluid-props/EnzoPhysicsFluidProps.cpp:64-138). The pressure part of that
// routine is evaluated on the fly inside the flux kernels (vlct_flux.cu).
// ---------------------------------------------------------------------------
template <bool STACKED>
__global__ void __launch_bounds__(kBlock)
k_specific_scalars(const int nsc, const typename GeomFor<STACKED>::type G,
                   const __grid_constant__ State u,
                   const __grid_constant__ ScalarPtrs spec, const Box box)
{
  VLCT_THREAD_IN_BOX(G, box, i, j, kl, k);
  (void) kl;
  const size_t c = cidx(G, k, j, i);
  const double rho = __ldg(u.rho + c);
  // all quotients share one reciprocal chain of rho (vlct_fpops.cuh: the same
  // bits as the built-in division wherever its range guard passes; otherwise
  // the cell is redone with the built-in operator)
  FastOps op;
  const double r = op.prep(rho);
  for (int s = 0; s < nsc; s++)
    spec.p[s][c] = op.quotz(__ldg(u.sc[s] + c), rho, r);
  if (op.bad)
    for (int s = 0; s < nsc; s++)
      spec.p[s][c] = __ldg(u.sc[s] + c) / rho;
}

// ---------------------------------------------------------------------------
// constrained transport (toolkit/EnzoBfieldMethodCT.cpp)
//   identify_upwind :170-216, compute_center_efield :267-292,
//   compute_edge_ :384-457 (negate_Ej = true), update_bfield :617-687
// ---------------------------------------------------------------------------
#ifndef VLCT_EDGE_MINBLOCKS
#define VLCT_EDGE_MINBLOCKS 4
#endif
struct EdgeArgs {
  const double* v[3];
  const double* b[3];
  const double* frho[3];      // density flux of each sweep (upwind weights)
  const double* fb[3][3];     // fb[sweep][component]
  double* edge[3];
  Box box[3];
};

__device__ __forceinline__ double upwind_weight(double dflux)
{
  if (dflux > 0) return 1.0;
  if (dflux < 0) return 0.0;
  return 0.5;
}

template <int D, class GEOM>
__device__ __forceinline__ void edge_component(const GEOM& G, const EdgeArgs& A,
                                               int kl, int k, int j, int i)
{
  constexpr int JD = (D + 1) % 3, KD = (D + 2) % 3;
  const Box& bx = A.box[D];
  if (i < bx.lo[0] || i >= bx.hi[0] || j < bx.lo[1] || j >= bx.hi[1] ||
      kl < bx.lo[2] || kl >= bx.hi[2]) return;
  const ptrdiff_t st[3] = { 1, (ptrdiff_t) G.mx, (ptrdiff_t) G.mx * (ptrdiff_t) G.my };
  const ptrdiff_t sj = st[JD], sk = st[KD];
  const size_t c = cidx(G, k, j, i);

  // cell-centred E_d = -v_j B_k + v_k B_j
  auto ecen = [&](size_t n) {
    return (-__ldg(A.v[JD] + n) * __ldg(A.b[KD] + n) +
            __ldg(A.v[KD] + n) * __ldg(A.b[JD] + n));
  };
  const double Ec = ecen(c), Ec_jp1 = ecen(c + sj), Ec_kp1 = ecen(c + sk),
               Ec_jkp1 = ecen(c + sj + sk);
  // E_d on j-faces is -F_j(B_k) (negation applied below), on k-faces +F_k(B_j)
  const double* Fj = A.fb[JD][KD];
  const double* Fk = A.fb[KD][JD];
  const double Ej = __ldg(Fj + c), Ej_kp1 = __ldg(Fj + c + sk);
  const double Ek = __ldg(Fk + c), Ek_jp1 = __ldg(Fk + c + sj);
  const double Wj = upwind_weight(__ldg(A.frho[JD] + c));
  const double Wj_kp1 = upwind_weight(__ldg(A.frho[JD] + c + sk));
  const double Wk = upwind_weight(__ldg(A.frho[KD] + c));
  const double Wk_jp1 = upwind_weight(__ldg(A.frho[KD] + c + sj));

  const double dEdj_r = Wk_jp1 * (Ec_jp1 + Ej) + (1 - Wk_jp1) * (Ec_jkp1 + Ej_kp1);
  const double dEdj_l = Wk * (-Ej - Ec) + (1 - Wk) * (-Ej_kp1 - Ec_kp1);
  const double dEdk_r = Wj_kp1 * (Ec_kp1 - Ek) + (1 - Wj_kp1) * (Ec_jkp1 - Ek_jp1);
  const double dEdk_l = Wj * (Ek - Ec) + (1 - Wj) * (Ek_jp1 - Ec_jp1);

  double Ej_sum = Ej + Ej_kp1;
  Ej_sum *= -1;
  const double Ek_sum = Ek + Ek_jp1;
  A.edge[D][c] = 0.25 * (Ej_sum + Ek_sum + (dEdj_l - dEdj_r) + (dEdk_l - dEdk_r));
}

#ifndef VLCT_EDGE_V1
/// one edge value from its twelve inputs (compute_edge_, CT.cpp:384-457, with
/// negate_Ej = true): operand order as in the reference
__device__ __forceinline__ double
edge_value(double Ec, double Ec_jp1, double Ec_kp1, double Ec_jkp1, double Ej,
           double Ej_kp1, double Ek, double Ek_jp1, double Wj, double Wj_kp1,
           double Wk, double Wk_jp1)
{
  const double dEdj_r = Wk_jp1 * (Ec_jp1 + Ej) + (1 - Wk_jp1) * (Ec_jkp1 + Ej_kp1);
  const double dEdj_l = Wk * (-Ej - Ec) + (1 - Wk) * (-Ej_kp1 - Ec_kp1);
  const double dEdk_r = Wj_kp1 * (Ec_kp1 - Ek) + (1 - Wj_kp1) * (Ec_jkp1 - Ek_jp1);
  const double dEdk_l = Wj * (Ek - Ec) + (1 - Wj) * (Ek_jp1 - Ec_jp1);
  double Ej_sum = Ej + Ej_kp1;
  Ej_sum *= -1;
  const double Ek_sum = Ek + Ek_jp1;
  return 0.25 * (Ej_sum + Ek_sum + (dEdj_l - dEdj_r) + (dEdk_l - dEdk_r));
}

// All three components of the edge E of one cell in one pass. The components
// read v and B on seven cells of the cell's 2x2x2 cube (0, +x, +y, +z, +y+z,
// +z+x, +x+y) and the density flux of every sweep on three: each value is
// loaded once (57 loads where three separate passes issue 72), every array is
// addressed from one pointer at the cell (the +x neighbours are immediate
// offsets), and nothing branches: a thread on the lowest layer of a
// component's box evaluates that component too and only skips its store. The
// kernel is bound by instruction issue under the board's power cap, not by
// HBM (profiles/r2b_power_per_kernel_family.jsonl), so instructions are what
// counts.
template <bool STACKED>
__global__ void __launch_bounds__(kBlock, VLCT_EDGE_MINBLOCKS)
k_edge_efield(const typename GeomFor<STACKED>::type G, const EdgeArgs A, const Box box)
{
  VLCT_THREAD_IN_BOX(G, box, i, j, kl, k);
  const ptrdiff_t Y = (ptrdiff_t) G.mx, Z = (ptrdiff_t) G.mx * (ptrdiff_t) G.my;
  const size_t c = cidx(G, k, j, i);
  const double* const vx = A.v[0] + c; const double* const vy = A.v[1] + c;
  const double* const vz = A.v[2] + c; const double* const bx = A.b[0] + c;
  const double* const by = A.b[1] + c; const double* const bz = A.b[2] + c;
  // cell-centred E_d = -v_j B_k + v_k B_j  (compute_center_efield, CT.cpp:267-292)
#define VLCT_EX(o) (-__ldg(vy + (o)) * __ldg(bz + (o)) + __ldg(vz + (o)) * __ldg(by + (o)))
#define VLCT_EY(o) (-__ldg(vz + (o)) * __ldg(bx + (o)) + __ldg(vx + (o)) * __ldg(bz + (o)))
#define VLCT_EZ(o) (-__ldg(vx + (o)) * __ldg(by + (o)) + __ldg(vy + (o)) * __ldg(bx + (o)))
  // upwind weights from the density fluxes (identify_upwind, CT.cpp:170-216)
  const double* const rx = A.frho[0] + c;
  const double* const ry = A.frho[1] + c;
  const double* const rz = A.frho[2] + c;
  const double wx0 = upwind_weight(__ldg(rx)), wxY = upwind_weight(__ldg(rx + Y)),
               wxZ = upwind_weight(__ldg(rx + Z));
  const double wy0 = upwind_weight(__ldg(ry)), wyZ = upwind_weight(__ldg(ry + Z)),
               wyX = upwind_weight(__ldg(ry + 1));
  const double wz0 = upwind_weight(__ldg(rz)), wzX = upwind_weight(__ldg(rz + 1)),
               wzY = upwind_weight(__ldg(rz + Y));
  const bool in_x = (i < box.hi[0]), in_y = (j < box.hi[1]), in_z = (kl < box.hi[2]);
  (void) in_x; (void) in_y; (void) in_z;   // (the launch box is the union: always true)
  {
    // x component: (j, k) = (y, z); E_x on y-faces is -F_y(B_z), on z-faces +F_z(B_y)
    const double* const Fj = A.fb[1][2] + c;
    const double* const Fk = A.fb[2][1] + c;
    const double e = edge_value(VLCT_EX(0), VLCT_EX(Y), VLCT_EX(Z), VLCT_EX(Y + Z),
                                __ldg(Fj), __ldg(Fj + Z), __ldg(Fk), __ldg(Fk + Y),
                                wy0, wyZ, wz0, wzY);
    if (i >= A.box[0].lo[0]) A.edge[0][c] = e;
  }
  {
    // y component: (j, k) = (z, x)
    const double* const Fj = A.fb[2][0] + c;
    const double* const Fk = A.fb[0][2] + c;
    const double e = edge_value(VLCT_EY(0), VLCT_EY(Z), VLCT_EY(1), VLCT_EY(Z + 1),
                                __ldg(Fj), __ldg(Fj + 1), __ldg(Fk), __ldg(Fk + Z),
                                wz0, wzX, wx0, wxZ);
    if (j >= A.box[1].lo[1]) A.edge[1][c] = e;
  }
  {
    // z component: (j, k) = (x, y)
    const double* const Fj = A.fb[0][1] + c;
    const double* const Fk = A.fb[1][0] + c;
    const double e = edge_value(VLCT_EZ(0), VLCT_EZ(1), VLCT_EZ(Y), VLCT_EZ(Y + 1),
                                __ldg(Fj), __ldg(Fj + Y), __ldg(Fk), __ldg(Fk + 1),
                                wx0, wxY, wy0, wyX);
    if (kl >= A.box[2].lo[2]) A.edge[2][c] = e;
  }
#undef VLCT_EX
#undef VLCT_EY
#undef VLCT_EZ
}
#else
// (40 registers = 6 blocks of 256 threads per SM; measured faster than 5 blocks
// at 42-44 registers, 4.4 vs 5.1 ms at 512^3, and than 8 blocks at 32)
template <bool STACKED>
__global__ void __launch_bounds__(kBlock)
k_edge_efield(const typename GeomFor<STACKED>::type G, const EdgeArgs A, const Box box)
{
  VLCT_THREAD_IN_BOX(G, box, i, j, kl, k);
  edge_component<0>(G, A, kl, k, j, i);
  edge_component<1>(G, A, kl, k, j, i);
  edge_component<2>(G, A, kl, k, j, i);
}
#endif

struct FaceArgs {
  const double* edge[3];
  const double* bi0[3];
  double* bi_out[3];
  const double* sp;         // step parameters of the stage: dt/dx, dt/dy, dt/dz, dt
  Box box[3];
};

template <int D, class GEOM>
__device__ __forceinline__ void face_component(const GEOM& G, const FaceArgs& A,
                                               int kl, int k, int j, int i)
{
  constexpr int JD = (D + 1) % 3, KD = (D + 2) % 3;
  const Box& bx = A.box[D];
  if (i < bx.lo[0] || i >= bx.hi[0] || j < bx.lo[1] || j >= bx.hi[1] ||
      kl < bx.lo[2] || kl >= bx.hi[2]) return;
  const ptrdiff_t st[3] = { 1, (ptrdiff_t) G.mx, (ptrdiff_t) G.mx * (ptrdiff_t) G.my };
  // face f along D is edge index f-1 along D
  const size_t e = cidx(G, k - (D == 2), j - (D == 1), i - (D == 0));
  const double ek_Rj = __ldg(A.edge[KD] + e), ek_Lj = __ldg(A.edge[KD] + e - st[JD]);
  const double ej_Rk = __ldg(A.edge[JD] + e), ej_Lk = __ldg(A.edge[JD] + e - st[KD]);
  const double E_k_term = __ldg(A.sp + JD) * (ek_Rj - ek_Lj);
  const double E_j_term = __ldg(A.sp + KD) * (ej_Rk - ej_Lk);
  const size_t f = fidx(G, D, k, j, i);
  A.bi_out[D][f] = __ldg(A.bi0[D] + f) - E_k_term + E_j_term;
}

template <bool STACKED>
__global__ void __launch_bounds__(kBlock)
k_face_bfield(const typename GeomFor<STACKED>::type G, const FaceArgs A, const Box box)
{
  VLCT_THREAD_IN_BOX(G, box, i, j, kl, k);
  face_component<0>(G, A, kl, k, j, i);
  face_component<1>(G, A, kl, k, j, i);
  face_component<2>(G, A, kl, k, j, i);
}

// ---------------------------------------------------------------------------
// update: toolkit/EnzoBfieldMethodCT.cpp:702-728 (centred B),
// toolkit/EnzoIntegrationQuanUpdate.cpp:105-147,183-268,
// toolkit/EnzoSourceInternalEnergy.cpp:16-92, toolkit/EnzoSourceGravity.cpp:17-69,
// fluid-props/EnzoPhysicsFluidProps.cpp:162-290
// ---------------------------------------------------------------------------
struct UpdateArgs {
  State u0, out;
  FluxSet flux[3];
  const double* bi_out[3];
  const double* cur_rho;     // density / internal energy of the stage's input
  const double* cur_eint;    // state (dual-energy source term only)
  const double* accel[3];
  const double* sp;          // dt/dx, dt/dy, dt/dz, dt of the stage
  int gravity;
  Box inner;                 // [s+1, m-s-1)^3: conserved update + floors
  // passive scalars without flux arrays (Params::nsc_flux == 0): specific
  // scalars of the stage's input state, the stage's reconstruction
  const double* spec[kMaxPassive];
  int recon;
  int scalars_elsewhere;     // 1: k_scalar_update advances the scalars (single block)
  // CFL fold (last stage of vlct_compute_and_timestep*): the timestep() of the
  // NEXT cycle is evaluated on the freshly updated cells while they are still
  // in registers; the cells this kernel does not update go through
  // k_timestep_boxes
  double* pressure;          // the "pressure" field (written for every updated cell)
  unsigned long long* dt_bits;
  double dx, dy, dz;
};

template <bool DE, bool MHD>
__device__ __forceinline__ void
floor_energy_and_sync(const Params& P, double rho, double vx, double vy,
                      double vz, double bx, double by, double bz, double& etot,
                      double& eint)
{
  const double inv_gm1 = 1. / (P.gamma - 1.);
  const double inv_rho = 1. / rho;
  const double eint_floor = P.pressure_floor * inv_gm1 * inv_rho;
  const double v2 = (vx * vx + vy * vy + vz * vz);
  double non_thermal_e = 0.5 * v2;
  double b2 = 0;
  if (MHD) {
    b2 = (bx * bx + by * by + bz * bz);
    non_thermal_e += (0.5 * b2 * inv_rho);
  }
  if (DE) {
    const double eta = P.de_eta;
    const double half_factor = (eta != 0.) ? 0.5 : 0.;
    const double eint_1 = etot - non_thermal_e;
    double cur_eint = eint;
    const double cs2_1 = fmax(0., P.ggm1 * eint_1);
    if ((cs2_1 > fmax(eta * v2, eta * b2 * inv_rho)) &&
        (eint_1 > half_factor * cur_eint)) {
      cur_eint = eint_1;
    }
    cur_eint = apply_floor(cur_eint, eint_floor);
    eint = cur_eint;
    etot = cur_eint + non_thermal_e;
  } else {
    const double etot_floor = eint_floor + non_thermal_e;
    etot = apply_floor(etot, etot_floor);
  }
}

/// What EnzoMethodMHDVlct::timestep does with one cell
/// (EnzoMethodMHDVlct.cpp:551-588, EnzoMHDIntegratorStageCommands.cpp:299-366):
/// dual-energy sync (rewrites etot / eint), pressure, local CFL limit.
template <bool MHD, bool DE>
__device__ __forceinline__ double
timestep_of_cell(const Params& P, double rho, double vx, double vy, double vz,
                 double bx, double by, double bz, double& etot, double& eint,
                 double dx, double dy, double dz, double& p)
{
  if (DE) {
    floor_energy_and_sync<true, MHD>(P, rho, vx, vy, vz, bx, by, bz, etot, eint);
    p = (P.gamma - 1.0) * rho * eint;
  } else {
    const double ke = 0.5 * (vx * vx + vy * vy + vz * vz);
    double me_den = 0.;
    if (MHD) me_den = 0.5 * (bx * bx + by * by + bz * bz);
    p = (P.gamma - 1.0) * (rho * (etot - ke) - me_den);
  }
  double cs;
  ExactOps op;   // HBM-bound kernels: the built-in operators are fine here
  if (MHD) cs = eos_cfast_max(op, P.gamma, rho, p, bx, by, bz);
  else     cs = sqrt(eos_cs2(op, P.gamma, rho, p));
  return min3(dx / (fabs(vx) + cs), dy / (fabs(vy) + cs), dz / (fabs(vz) + cs));
}

/// minimum over the block (warp shuffles, then one atomicMin per block).
/// Non-negative doubles order like their bit patterns; NaNs compare above +inf
/// and so never win, which matches std::min(dtBaryons, local_dt) keeping the
/// old value. Every thread of the block must call this.
__device__ __forceinline__ void block_min_to(unsigned long long* dt_bits, double local_min)
{
  unsigned long long bits = (unsigned long long) __double_as_longlong(local_min);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const unsigned long long other = __shfl_down_sync(0xffffffffu, bits, off);
    bits = (other < bits) ? other : bits;
  }
  __shared__ unsigned long long warp_min[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) warp_min[warp] = bits;
  __syncthreads();
  if (warp == 0) {
    bits = (lane < (int) ((blockDim.x + 31) >> 5)) ? warp_min[lane]
                                                     : 0x7fefffffffffffffULL;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const unsigned long long other = __shfl_down_sync(0xffffffffu, bits, off);
      bits = (other < bits) ? other : bits;
    }
    if (lane == 0) atomicMin(dt_bits, bits);
  }
}

// SCAL: the kernel advances the passive scalars itself (stacked batches, or
// scalar flux arrays); otherwise k_scalar_update does, or there are none --
// and the scalar code with its registers is compiled out
#ifndef VLCT_UPDATE_MINBLOCKS
#define VLCT_UPDATE_MINBLOCKS 4
#endif
template <bool MHD, bool DE, bool STACKED, bool CFL, bool SCAL>
__global__ void __launch_bounds__(kBlock, SCAL ? 2 : VLCT_UPDATE_MINBLOCKS)
k_update(const Params P, const typename GeomFor<STACKED>::type G, const UpdateArgs A,
         const Box box)
{
  // (with the CFL fold every thread of the block reaches the block-wide
  // minimum at the end, so nothing returns early)
  const unsigned nxb = box.hi[0] - box.lo[0];
  const unsigned t = blockIdx.x * kBlock + threadIdx.x;
  bool active = t < nxb * (unsigned) (box.hi[1] - box.lo[1]);
  if (!CFL && !active) return;
  const int i = box.lo[0] + (int) (t % nxb);
  const int j = active ? box.lo[1] + (int) (t / nxb) : box.lo[1];
  int kl, k;
  if constexpr (STACKED) unstack(G, box, blockIdx.y, kl, k);
  else kl = k = box.lo[2] + (int) blockIdx.y;
  const size_t c = cidx(G, k, j, i);
  const ptrdiff_t st[3] = { 1, (ptrdiff_t) G.mx, (ptrdiff_t) G.mx * (ptrdiff_t) G.my };
  double local_dt = DBL_MAX;

  double bx = 0., by = 0., bz = 0.;
  if (MHD && active) {
    bx = 0.5 * (__ldg(A.bi_out[0] + fidx(G, 0, k, j, i)) +
                __ldg(A.bi_out[0] + fidx(G, 0, k, j, i + 1)));
    by = 0.5 * (__ldg(A.bi_out[1] + fidx(G, 1, k, j, i)) +
                __ldg(A.bi_out[1] + fidx(G, 1, k, j + 1, i)));
    bz = 0.5 * (__ldg(A.bi_out[2] + fidx(G, 2, k, j, i)) +
                __ldg(A.bi_out[2] + fidx(G, 2, k + 1, j, i)));
    A.out.bx[c] = bx;
    A.out.by[c] = by;
    A.out.bz[c] = bz;
  }

  const Box& in = A.inner;
  if (i < in.lo[0] || i >= in.hi[0] || j < in.lo[1] || j >= in.hi[1] ||
      kl < in.lo[2] || kl >= in.hi[2]) active = false;
  if (!CFL && !active) return;

  if (active) {
  const double dtd[3] = { __ldg(A.sp), __ldg(A.sp + 1), __ldg(A.sp + 2) };
  // accumulate dU = 0 - sum_d dt/dx_d (F_{c+1/2} - F_{c-1/2}) in x,y,z order
  double d_rho = 0., d_mx = 0., d_my = 0., d_mz = 0., d_e = 0., d_eint = 0.;
  // density fluxes: kept only where this kernel upwinds the scalars itself
  double frho_c[SCAL ? 3 : 1], frho_l[SCAL ? 3 : 1];
  double p_floored = 0.;
  if (DE) {
    // cell-centred primitive pressure of the current stage
    // (EnzoComputePressure.cpp: p = (gamma-1) rho eint with dual energy)
    const double p = (P.gamma - 1.0) * __ldg(A.cur_rho + c) * __ldg(A.cur_eint + c);
    p_floored = apply_floor(p, P.pressure_floor);
  }
#pragma unroll
  for (int d = 0; d < 3; d++) {
    const FluxSet& F = A.flux[d];
    const size_t l = c - st[d];
    const double dtdx = dtd[d];
    const double fr_c = __ldg(F.rho + c), fr_l = __ldg(F.rho + l);
    if (SCAL) { frho_c[SCAL ? d : 0] = fr_c; frho_l[SCAL ? d : 0] = fr_l; }
    d_rho -= dtdx * (fr_c - fr_l);
    d_mx -= dtdx * (__ldg(F.mx_ + c) - __ldg(F.mx_ + l));
    d_my -= dtdx * (__ldg(F.my_ + c) - __ldg(F.my_ + l));
    d_mz -= dtdx * (__ldg(F.mz_ + c) - __ldg(F.mz_ + l));
    d_e -= dtdx * (__ldg(F.e + c) - __ldg(F.e + l));
    if (DE) {
      d_eint -= dtdx * (__ldg(F.eint + c) - __ldg(F.eint + l));
      d_eint -= dtdx * p_floored * (__ldg(F.vbar + c) - __ldg(F.vbar + l));
    }
  }

  const double old_rho = __ldg(A.u0.rho + c);
  const double vx0 = __ldg(A.u0.vx + c), vy0 = __ldg(A.u0.vy + c),
               vz0 = __ldg(A.u0.vz + c);
  if (A.gravity) {
    const double ax = __ldg(A.accel[0] + c), ay = __ldg(A.accel[1] + c),
                 az = __ldg(A.accel[2] + c);
    const double dt = __ldg(A.sp + 3);
    d_mx += dt * old_rho * ax;
    d_my += dt * old_rho * ay;
    d_mz += dt * old_rho * az;
    d_e += dt * old_rho * ((vx0 * ax) + (vy0 * ay) + (vz0 * az));
  }

  // passive scalars (conserved form)
  if (!SCAL) {
    // none, or k_scalar_update (a z-marching kernel of its own) does them
  } else if (P.nsc_flux > 0) {
    for (int s = 0; s < P.nsc; s++) {
      double d_s = 0.;
#pragma unroll
      for (int d = 0; d < 3; d++) {
        const double* Fs = A.flux[d].sc[s];
        d_s -= dtd[d] * (__ldg(Fs + c) - __ldg(Fs + c - st[d]));
      }
      A.out.sc[s][c] = __ldg(A.u0.sc[s] + c) + d_s;
    }
  } else {
    // The scalar fluxes through the cell's two faces per direction, formed
    // here as the sweeps would (reconstruct the specific scalar on both sides
    // of a face, upwind by the sign of the density flux, times the density
    // flux: EnzoReconstructor*, riemann/EnzoRiemannUtils.hpp:224-249,267-314)
    // instead of being written by three sweeps and read back: 13 specific
    // values (mostly L1 / L2 hits) replace 3 stores and 6 loads per scalar,
    // and the sweeps carry no scalar work at all. Same expressions, same bits.
    const bool nn = (A.recon == VLCT_RECON_NN);
    const bool athena = (A.recon == VLCT_RECON_PLM_ATHENA);
    for (int s = 0; s < P.nsc; s++) {
      const double* const q = A.spec[s] + c;
      const double w0 = __ldg(q);
      double d_s = 0.;
#pragma unroll
      for (int d = 0; d < 3; d++) {
        const ptrdiff_t sd = st[d];
        const double wm1 = __ldg(q - sd), wp1 = __ldg(q + sd);
        double sl_c, sr_c, sl_l, sr_l;   // L / R states at the upper and the lower face
        if (nn) {
          sl_l = wm1; sr_l = w0; sl_c = w0; sr_c = wp1;
        } else {
          const double wm2 = __ldg(q - 2 * sd), wp2 = __ldg(q + 2 * sd);
          double dm, d0, dp;
          if (athena) {
            dm = limiter_athena(wm2, wm1, w0);
            d0 = limiter_athena(wm1, w0, wp1);
            dp = limiter_athena(w0, wp1, wp2);
          } else {
            dm = limiter_enzo(wm2, wm1, w0, P.theta);
            d0 = limiter_enzo(wm1, w0, wp1, P.theta);
            dp = limiter_enzo(w0, wp1, wp2, P.theta);
          }
          sl_l = wm1 + dm * 0.5;
          sr_l = w0 - d0 * 0.5;
          sl_c = w0 + d0 * 0.5;
          sr_c = wp1 - dp * 0.5;
        }
        d_s -= dtd[d] * (passive_flux(sl_c, sr_c, frho_c[SCAL ? d : 0]) -
                         passive_flux(sl_l, sr_l, frho_l[SCAL ? d : 0]));
      }
      A.out.sc[s][c] = __ldg(A.u0.sc[s] + c) + d_s;
    }
  }

  double new_rho = old_rho + d_rho;
  new_rho = apply_floor(new_rho, P.density_floor);
  const double inv_new_rho = 1. / new_rho;
  const double vx = (vx0 * old_rho + d_mx) * inv_new_rho;
  const double vy = (vy0 * old_rho + d_my) * inv_new_rho;
  const double vz = (vz0 * old_rho + d_mz) * inv_new_rho;
  double etot = (__ldg(A.u0.etot + c) * old_rho + d_e) * inv_new_rho;
  double eint = 0.;
  if (DE) eint = (__ldg(A.u0.eint + c) * old_rho + d_eint) * inv_new_rho;

  floor_energy_and_sync<DE, MHD>(P, new_rho, vx, vy, vz, bx, by, bz, etot, eint);

  if (CFL) {
    // timestep() of the next cycle on this cell: a second dual-energy sync
    // (what timestep() does to the field compute() left), "pressure", CFL
    double p;
    local_dt = timestep_of_cell<MHD, DE>(P, new_rho, vx, vy, vz, bx, by, bz, etot, eint,
                                         A.dx, A.dy, A.dz, p);
    A.pressure[c] = p;
  }

  A.out.rho[c] = new_rho;
  A.out.vx[c] = vx;
  A.out.vy[c] = vy;
  A.out.vz[c] = vz;
  A.out.etot[c] = etot;
  if (DE) A.out.eint[c] = eint;
  }
  if (CFL) block_min_to(A.dt_bits, local_dt);
}

// ---------------------------------------------------------------------------
// Pair variants of the cell kernels: a thread owns the two cells (i, i+1), i
// even, of a row. Rows of every cell-strided array start 16-byte aligned when
// mx is even (and the arrays themselves are: launch_* checks both), so each
// array is moved with one 128-bit load / store per pair (LDG.E.128 / STG.E.128)
// and the index arithmetic is paid once per two cells: the cell kernels run at
// the board's power cap with 60-75 % of their instructions being address
// arithmetic and loads, so instructions are what they cost. Only the x-face
// array (rows of mx + 1 entries) keeps 64-bit accesses. Values at odd x
// offsets (i-1, i+2) are single loads. Every cell evaluates exactly the
// expressions of the one-cell kernels above: bit-identical.
// A pair may stick out of the box by one cell on either side (odd lower bound,
// odd upper bound): that cell is computed from in-bounds memory and not stored.
// ---------------------------------------------------------------------------
constexpr int kPairBlock = 128;      // threads = 256 cells, like kBlock

__device__ __forceinline__ double2 ld2(const double* p)
{ return __ldg(reinterpret_cast<const double2*>(p)); }

/// store of a pair of which only the valid cells are written
__device__ __forceinline__ void st_pair(double* p, double a, double b, bool va, bool vb)
{
  if (va && vb) *reinterpret_cast<double2*>(p) = make_double2(a, b);
  else if (va) p[0] = a;
  else if (vb) p[1] = b;
}

inline dim3 pair_grid_for(const Geom& G, const Box& b)
{
  const unsigned npx = (unsigned) (b.hi[0] - (b.lo[0] & ~1) + 1) >> 1;
  const unsigned ny = b.hi[1] - b.lo[1];
  return dim3((npx * ny + kPairBlock - 1) / kPairBlock, stacked_nz(G, b), 1);
}

// i: the pair's first cell (even); va / vb: cell i / i+1 lies inside the box
#define VLCT_PAIR_IN_BOX(G, box, i, j, kl, k, va, vb)                          \
  const int i0__ = (box).lo[0] & ~1;                                           \
  const unsigned npx__ = (unsigned) ((box).hi[0] - i0__ + 1) >> 1;             \
  const unsigned t__ = blockIdx.x * kPairBlock + threadIdx.x;                  \
  if (t__ >= npx__ * (unsigned) ((box).hi[1] - (box).lo[1])) return;           \
  const unsigned jj__ = t__ / npx__;                                           \
  const int i = i0__ + 2 * (int) (t__ - jj__ * npx__);                         \
  const int j = (box).lo[1] + (int) jj__;                                      \
  const bool va = (i >= (box).lo[0]), vb = (i + 1 < (box).hi[0]);              \
  int kl, k;                                                                   \
  if constexpr (STACKED) unstack((G), (box), blockIdx.y, kl, k);               \
  else kl = k = (box).lo[2] + (int) blockIdx.y;

/// a row's values at x = i, i+1 (one 128-bit load) and i+2 (cell b's +x
/// neighbour: loaded only when cell b is computed for real)
struct Row3 { double x0, x1, x2; };
__device__ __forceinline__ Row3 row3(const double* p, bool vb)
{
  const double2 v = ld2(p);
  Row3 r;
  r.x0 = v.x; r.x1 = v.y; r.x2 = vb ? __ldg(p + 2) : 0.;
  return r;
}

#ifndef VLCT_EDGE2_MINBLOCKS
#define VLCT_EDGE2_MINBLOCKS 4
#endif
template <bool STACKED>
__global__ void __launch_bounds__(kPairBlock, VLCT_EDGE2_MINBLOCKS)
k_edge_efield2(const typename GeomFor<STACKED>::type G, const EdgeArgs A, const Box box)
{
  VLCT_PAIR_IN_BOX(G, box, i, j, kl, k, va, vb);
  const ptrdiff_t Y = (ptrdiff_t) G.mx, Z = (ptrdiff_t) G.mx * (ptrdiff_t) G.my;
  const size_t c = cidx(G, k, j, i);
  const double* const vx = A.v[0] + c; const double* const vy = A.v[1] + c;
  const double* const vz = A.v[2] + c; const double* const bx = A.b[0] + c;
  const double* const by = A.b[1] + c; const double* const bz = A.b[2] + c;
  // cell-centred E_d = -v_j B_k + v_k B_j  (compute_center_efield, CT.cpp:267-292)
#define VLCT_EC(vj, bk, vk, bj) (-(vj) * (bk) + (vk) * (bj))
  // the x component needs E_x on the rows 0, +y, +z, +y+z at x = i, i+1;
  // the y component E_y on the rows 0, +z at x = i .. i+2;
  // the z component E_z on the rows 0, +y at x = i .. i+2
  double ex[4][2], ey[2][3], ez[2][3];
  {
    const Row3 vy0 = row3(vy, vb), vz0 = row3(vz, vb), by0 = row3(by, vb), bz0 = row3(bz, vb);
    const Row3 vx0 = row3(vx, vb), bx0 = row3(bx, vb);
    ex[0][0] = VLCT_EC(vy0.x0, bz0.x0, vz0.x0, by0.x0);
    ex[0][1] = VLCT_EC(vy0.x1, bz0.x1, vz0.x1, by0.x1);
    ey[0][0] = VLCT_EC(vz0.x0, bx0.x0, vx0.x0, bz0.x0);
    ey[0][1] = VLCT_EC(vz0.x1, bx0.x1, vx0.x1, bz0.x1);
    ey[0][2] = VLCT_EC(vz0.x2, bx0.x2, vx0.x2, bz0.x2);
    ez[0][0] = VLCT_EC(vx0.x0, by0.x0, vy0.x0, bx0.x0);
    ez[0][1] = VLCT_EC(vx0.x1, by0.x1, vy0.x1, bx0.x1);
    ez[0][2] = VLCT_EC(vx0.x2, by0.x2, vy0.x2, bx0.x2);
  }
  {
    // row +y: E_x at x = i, i+1; E_z at x = i .. i+2
    const Row3 vxY = row3(vx + Y, vb), vyY = row3(vy + Y, vb);
    const Row3 bxY = row3(bx + Y, vb), byY = row3(by + Y, vb);
    const double2 vzY = ld2(vz + Y), bzY = ld2(bz + Y);
    ex[1][0] = VLCT_EC(vyY.x0, bzY.x, vzY.x, byY.x0);
    ex[1][1] = VLCT_EC(vyY.x1, bzY.y, vzY.y, byY.x1);
    ez[1][0] = VLCT_EC(vxY.x0, byY.x0, vyY.x0, bxY.x0);
    ez[1][1] = VLCT_EC(vxY.x1, byY.x1, vyY.x1, bxY.x1);
    ez[1][2] = VLCT_EC(vxY.x2, byY.x2, vyY.x2, bxY.x2);
  }
  {
    // row +z: E_x at x = i, i+1; E_y at x = i .. i+2
    const Row3 vxZ = row3(vx + Z, vb), vzZ = row3(vz + Z, vb);
    const Row3 bxZ = row3(bx + Z, vb), bzZ = row3(bz + Z, vb);
    const double2 vyZ = ld2(vy + Z), byZ = ld2(by + Z);
    ex[2][0] = VLCT_EC(vyZ.x, bzZ.x0, vzZ.x0, byZ.x);
    ex[2][1] = VLCT_EC(vyZ.y, bzZ.x1, vzZ.x1, byZ.y);
    ey[1][0] = VLCT_EC(vzZ.x0, bxZ.x0, vxZ.x0, bzZ.x0);
    ey[1][1] = VLCT_EC(vzZ.x1, bxZ.x1, vxZ.x1, bzZ.x1);
    ey[1][2] = VLCT_EC(vzZ.x2, bxZ.x2, vxZ.x2, bzZ.x2);
  }
  {
    // row +y+z: E_x only
    const double2 vyW = ld2(vy + Y + Z), vzW = ld2(vz + Y + Z);
    const double2 byW = ld2(by + Y + Z), bzW = ld2(bz + Y + Z);
    ex[3][0] = VLCT_EC(vyW.x, bzW.x, vzW.x, byW.x);
    ex[3][1] = VLCT_EC(vyW.y, bzW.y, vzW.y, byW.y);
  }
#undef VLCT_EC
  // upwind weights from the density fluxes (identify_upwind, CT.cpp:170-216)
  const double* const rx = A.frho[0] + c;
  const double* const ry = A.frho[1] + c;
  const double* const rz = A.frho[2] + c;
  const double2 rx0 = ld2(rx), rxY = ld2(rx + Y), rxZ = ld2(rx + Z);
  const Row3 ry0 = row3(ry, vb); const double2 ryZ = ld2(ry + Z);
  const Row3 rz0 = row3(rz, vb); const double2 rzY = ld2(rz + Y);
  const double wx0[2] = { upwind_weight(rx0.x), upwind_weight(rx0.y) };
  const double wxY[2] = { upwind_weight(rxY.x), upwind_weight(rxY.y) };
  const double wxZ[2] = { upwind_weight(rxZ.x), upwind_weight(rxZ.y) };
  const double wy0[3] = { upwind_weight(ry0.x0), upwind_weight(ry0.x1), upwind_weight(ry0.x2) };
  const double wyZ[2] = { upwind_weight(ryZ.x), upwind_weight(ryZ.y) };
  const double wz0[3] = { upwind_weight(rz0.x0), upwind_weight(rz0.x1), upwind_weight(rz0.x2) };
  const double wzY[2] = { upwind_weight(rzY.x), upwind_weight(rzY.y) };
  {
    // x component: (j, k) = (y, z); E_x on y-faces is -F_y(B_z), on z-faces +F_z(B_y)
    const double* const Fj = A.fb[1][2] + c;
    const double* const Fk = A.fb[2][1] + c;
    const double2 fj0 = ld2(Fj), fjZ = ld2(Fj + Z), fk0 = ld2(Fk), fkY = ld2(Fk + Y);
    const double ea = edge_value(ex[0][0], ex[1][0], ex[2][0], ex[3][0], fj0.x, fjZ.x, fk0.x,
                                 fkY.x, wy0[0], wyZ[0], wz0[0], wzY[0]);
    const double eb = edge_value(ex[0][1], ex[1][1], ex[2][1], ex[3][1], fj0.y, fjZ.y, fk0.y,
                                 fkY.y, wy0[1], wyZ[1], wz0[1], wzY[1]);
    const int lo = A.box[0].lo[0];
    st_pair(A.edge[0] + c, ea, eb, va && i >= lo, vb && i + 1 >= lo);
  }
  {
    // y component: (j, k) = (z, x)
    const double* const Fj = A.fb[2][0] + c;
    const double* const Fk = A.fb[0][2] + c;
    const Row3 fj = row3(Fj, vb);
    const double2 fk0 = ld2(Fk), fkZ = ld2(Fk + Z);
    const double ea = edge_value(ey[0][0], ey[1][0], ey[0][1], ey[1][1], fj.x0, fj.x1, fk0.x,
                                 fkZ.x, wz0[0], wz0[1], wx0[0], wxZ[0]);
    const double eb = edge_value(ey[0][1], ey[1][1], ey[0][2], ey[1][2], fj.x1, fj.x2, fk0.y,
                                 fkZ.y, wz0[1], wz0[2], wx0[1], wxZ[1]);
    const bool row = (j >= A.box[1].lo[1]);
    st_pair(A.edge[1] + c, ea, eb, va && row, vb && row);
  }
  {
    // z component: (j, k) = (x, y)
    const double* const Fj = A.fb[0][1] + c;
    const double* const Fk = A.fb[1][0] + c;
    const double2 fj0 = ld2(Fj), fjY = ld2(Fj + Y);
    const Row3 fk = row3(Fk, vb);
    const double ea = edge_value(ez[0][0], ez[0][1], ez[1][0], ez[1][1], fj0.x, fjY.x, fk.x0,
                                 fk.x1, wx0[0], wxY[0], wy0[0], wy0[1]);
    const double eb = edge_value(ez[0][1], ez[0][2], ez[1][1], ez[1][2], fj0.y, fjY.y, fk.x1,
                                 fk.x2, wx0[1], wxY[1], wy0[1], wy0[2]);
    const bool lev = (kl >= A.box[2].lo[2]);
    st_pair(A.edge[2] + c, ea, eb, va && lev, vb && lev);
  }
}

// ---------------------------------------------------------------------------
// Edge E with its inputs staged by the TMA unit (cp.async.bulk.tensor, SASS
// UTMALDG): a block owns a tile of 32 x 16 cells and marches along z. A
// producer warp asks the TMA unit for what the next level needs -- one box of
// 17 rows x 34 doubles from each of the 15 input arrays, completion counted on
// an mbarrier -- up to three levels ahead; the 16 consumer warps read
// everything from shared memory at immediate offsets and hand a level's stage
// back through a second mbarrier (no block-wide barrier). The loads in flight are no longer
// limited by registers x resident warps (what bounds k_edge_efield: 57 loads
// per thread, 4.8 TB/s), each level of v / B / fluxes crosses L2 once instead
// of twice, and a thread executes ~60 % of k_edge_efield's instructions.
// Same expressions in the same order: bit-identical. Single blocks, even row
// length, 16-byte aligned arrays (tensor-map strides are multiples of 16 bytes).
// ---------------------------------------------------------------------------
__device__ __forceinline__ double ecen_(double vj, double bk, double vk, double bj)
{ return (-vj * bk + vk * bj); }

#ifndef VLCT_TMA_TX
#define VLCT_TMA_TX 32
#endif
#ifndef VLCT_TMA_TY
#define VLCT_TMA_TY 16
#endif
#ifndef VLCT_TMA_STAGES
#define VLCT_TMA_STAGES 3
#endif
constexpr int kTmaTX = VLCT_TMA_TX, kTmaTY = VLCT_TMA_TY;
constexpr int kTmaRows = kTmaTY + 1;        // rows a level needs: +1 for the +y neighbours
constexpr int kTmaRowD = kTmaTX + 2;        // doubles per row: +1 for +x, +1 keeps 16 bytes
constexpr int kTmaArrays = 15;
// one staged array = one TMA box of (TY + 1) rows x (TX + 2) doubles, padded to a multiple
// of 128 bytes (destination alignment of cp.async.bulk.tensor)
constexpr int kTmaArrayD = ((kTmaRows * kTmaRowD * 8 + 127) / 128) * 128 / 8;
constexpr int kTmaLevelD = kTmaArrays * kTmaArrayD;            // doubles per level buffer
constexpr unsigned kTmaLevelTx = kTmaArrays * kTmaRows * kTmaRowD * 8;   // bytes that land
constexpr int kTmaStages = VLCT_TMA_STAGES;
constexpr size_t kTmaSmemBytes = (size_t) kTmaStages * kTmaLevelD * sizeof(double) + 128;
enum { TA_VX = 0, TA_VY, TA_VZ, TA_BX, TA_BY, TA_BZ, TA_RX, TA_RY, TA_RZ,
       TA_F12, TA_F21, TA_F20, TA_F02, TA_F01, TA_F10 };
struct EdgeTmaArgs {
  // (x, y, z) tensor maps of the 15 input arrays, box (TX + 2) x (TY + 1) x 1, zero fill
  // outside the array
  alignas(64) CUtensorMap map[kTmaArrays];
  double* edge[3];
  int s;              // stale depth: union box [s, m-s-1)^3, component d starts at s+1 along d
  int k0, kend;       // levels [k0, kend) of the (clipped) union box
};

__device__ __forceinline__ unsigned smem_u32(const void* p)
{ return (unsigned) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
               :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
/// wait for the phase with the given parity. try_wait suspends the thread in
/// hardware for a bounded time; a phase that never completes (a faulting copy)
/// traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
  const unsigned addr = smem_u32(bar);
  for (unsigned it = 0; it < (1u << 24); it++) {
    unsigned done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (done) return;
  }
  __trap();
}
/// one box of a 3-D tensor map, global -> shared; completes on the mbarrier
__device__ __forceinline__ void tma_box_g2s(void* dst, const CUtensorMap* map, int x, int y,
                                            int z, unsigned long long* bar)
{
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%2, %3, %4}], [%5];"
      :: "r"(smem_u32(dst)), "l"(reinterpret_cast<unsigned long long>(map)), "r"(x), "r"(y),
         "r"(z), "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_arrive(unsigned long long* bar)
{ asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory"); }

constexpr int kTmaConsumers = kTmaTX * kTmaTY;            // 16 warps, one cell per thread
constexpr int kTmaThreads = kTmaConsumers + 32;           // + the producer warp
constexpr int kTmaRowsPerLane = (kTmaArrays * kTmaRows + 31) / 32;

__global__ void __launch_bounds__(kTmaThreads, 1)
k_edge_efield_tma(const GeomLite G, const __grid_constant__ EdgeTmaArgs A, const int chunk)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* const buf = reinterpret_cast<double*>(smem_raw);
  // full[st]: the bulk copies of a level have landed; empty[st]: the 16
  // consumer warps are done with it
  unsigned long long* const full = reinterpret_cast<unsigned long long*>(
      smem_raw + (size_t) kTmaStages * kTmaLevelD * sizeof(double));
  unsigned long long* const empty = full + kTmaStages;
  const int tid = threadIdx.x;
  const int s = A.s;
  const int i0 = (s & ~1) + kTmaTX * (int) blockIdx.x;
  const int j0 = s + kTmaTY * (int) blockIdx.y;
  const int kc0 = A.k0 + (int) blockIdx.z * chunk;
  const int kc1 = min(kc0 + chunk, A.kend);
  const size_t Z = (size_t) G.mx * (size_t) G.my;
  if (tid == 0) {
    for (int b = 0; b < kTmaStages; b++) {
      mbar_init(full + b, 1);
      mbar_init(empty + b, kTmaConsumers / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (tid >= kTmaConsumers) {
    // ---- producer warp: levels kc0 .. kc1, one TMA box per array ----
    const int lane = tid - kTmaConsumers;
#pragma unroll 1
    for (int n = 0; n <= kc1 - kc0; n++) {
      const int st = n % kTmaStages, use = n / kTmaStages;
      if (use > 0) mbar_wait(empty + st, (unsigned) (use - 1) & 1u);
      if (lane == 0) mbar_expect_tx(full + st, kTmaLevelTx);
      __syncwarp();
      if (lane < kTmaArrays)
        tma_box_g2s(buf + (size_t) st * kTmaLevelD + lane * kTmaArrayD, &A.map[lane], i0, j0,
                    kc0 + n, full + st);
    }
    return;
  }

  // ---- consumers: warp w = row-pairs of the tile, no block-wide barrier ----
  const int lx = tid & (kTmaTX - 1), ly = tid / kTmaTX;
  const int i = i0 + lx, j = j0 + ly;
  const bool active = (i >= s && i < G.mx - s - 1 && j >= s && j < G.my - s - 1);
  const size_t cell0 = active ? cidx(G, kc0, j, i) : 0;
  const int o = ly * kTmaRowD + lx;           // the cell inside a staged array
  constexpr int AR = kTmaArrayD;              // doubles per staged array
  constexpr int Yo = kTmaRowD;                // +y inside a staged array
#pragma unroll 1
  for (int k = kc0; k < kc1; k++) {
    const int n = k - kc0;
    const int st0 = n % kTmaStages, st1 = (n + 1) % kTmaStages;
    mbar_wait(full + st0, (unsigned) (n / kTmaStages) & 1u);
    mbar_wait(full + st1, (unsigned) ((n + 1) / kTmaStages) & 1u);
    if (active) {
      const double* const c0 = buf + (size_t) st0 * kTmaLevelD + o;   // level k
      const double* const c1 = buf + (size_t) st1 * kTmaLevelD + o;   // level k+1
#define VLCT_EX(p, d) ecen_((p)[TA_VY * AR + (d)], (p)[TA_BZ * AR + (d)], (p)[TA_VZ * AR + (d)], (p)[TA_BY * AR + (d)])
#define VLCT_EY(p, d) ecen_((p)[TA_VZ * AR + (d)], (p)[TA_BX * AR + (d)], (p)[TA_VX * AR + (d)], (p)[TA_BZ * AR + (d)])
#define VLCT_EZ(p, d) ecen_((p)[TA_VX * AR + (d)], (p)[TA_BY * AR + (d)], (p)[TA_VY * AR + (d)], (p)[TA_BX * AR + (d)])
      const double wx0 = upwind_weight(c0[TA_RX * AR]), wxY = upwind_weight(c0[TA_RX * AR + Yo]),
                   wxZ = upwind_weight(c1[TA_RX * AR]);
      const double wy0 = upwind_weight(c0[TA_RY * AR]), wyZ = upwind_weight(c1[TA_RY * AR]),
                   wyX = upwind_weight(c0[TA_RY * AR + 1]);
      const double wz0 = upwind_weight(c0[TA_RZ * AR]), wzX = upwind_weight(c0[TA_RZ * AR + 1]),
                   wzY = upwind_weight(c0[TA_RZ * AR + Yo]);
      const size_t c = cell0 + (size_t) n * Z;
      {
        const double e = edge_value(VLCT_EX(c0, 0), VLCT_EX(c0, Yo), VLCT_EX(c1, 0),
                                    VLCT_EX(c1, Yo), c0[TA_F12 * AR], c1[TA_F12 * AR],
                                    c0[TA_F21 * AR], c0[TA_F21 * AR + Yo], wy0, wyZ, wz0, wzY);
        if (i >= s + 1) A.edge[0][c] = e;
      }
      {
        const double e = edge_value(VLCT_EY(c0, 0), VLCT_EY(c1, 0), VLCT_EY(c0, 1),
                                    VLCT_EY(c1, 1), c0[TA_F20 * AR], c0[TA_F20 * AR + 1],
                                    c0[TA_F02 * AR], c1[TA_F02 * AR], wz0, wzX, wx0, wxZ);
        if (j >= s + 1) A.edge[1][c] = e;
      }
      {
        const double e = edge_value(VLCT_EZ(c0, 0), VLCT_EZ(c0, 1), VLCT_EZ(c0, Yo),
                                    VLCT_EZ(c0, Yo + 1), c0[TA_F01 * AR], c0[TA_F01 * AR + Yo],
                                    c0[TA_F10 * AR], c0[TA_F10 * AR + 1], wx0, wxY, wy0, wyX);
        if (k >= s + 1) A.edge[2][c] = e;
      }
#undef VLCT_EX
#undef VLCT_EY
#undef VLCT_EZ
    }
    // this warp is done with the stage of level k
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(empty + st0);
  }
}

// ---------------------------------------------------------------------------
// Constrained transport in one TMA-staged kernel: k_edge_efield_tma whose
// consumers also update the faces, so that the edge E never reach HBM (-3
// stores and -6 loads of 18 + 9 doubles per cell and stage). A thread evaluates
// the three edge E of its cell at level k, publishes them in shared memory,
// and updates the x face on its +x side, the y face on its +y side (they need
// E_y / E_x of level k-1: kept in registers) and the z face above it; the
// edges of the -x / -y neighbours come from shared memory. A block is 32 x 16
// columns of which 30 x 15 update faces (column 0 and row 0 only provide the
// neighbours' edges; measured against 64 x 8: 8.95 vs 9.23 ms per step). Same expressions in the same order as k_edge_efield +
// k_face_bfield: bit-identical.
// ---------------------------------------------------------------------------
struct CtTmaArgs {
  alignas(64) CUtensorMap map[kTmaArrays];
  const double* bi0[3];
  double* bi_out[3];
  const double* sp;           // dt/dx, dt/dy, dt/dz of the stage
  int s;                      // stale depth: the edge and face boxes follow from it
  int zlo;                    // first face level (cells for x / y faces, face index for z)
  int zhi;                    // z faces: index < zhi
  int kend;                   // the march ends below this cell level
};
constexpr size_t kCtExchBytes = (size_t) 3 * kTmaTY * kTmaTX * sizeof(double);
constexpr size_t kCtSmemBytes = (size_t) kTmaStages * kTmaLevelD * sizeof(double) +
                                kCtExchBytes + 128;

__device__ __forceinline__ void consumer_barrier()
{ asm volatile("bar.sync 1, %0;" :: "n"(kTmaConsumers) : "memory"); }

__global__ void __launch_bounds__(kTmaThreads, 1)
k_ct_tma(const GeomLite G, const __grid_constant__ CtTmaArgs A, const int chunk)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* const buf = reinterpret_cast<double*>(smem_raw);
  double (*const X)[kTmaTY][kTmaTX] = reinterpret_cast<double (*)[kTmaTY][kTmaTX]>(
      smem_raw + (size_t) kTmaStages * kTmaLevelD * sizeof(double));
  unsigned long long* const full = reinterpret_cast<unsigned long long*>(
      smem_raw + (size_t) kTmaStages * kTmaLevelD * sizeof(double) + kCtExchBytes);
  unsigned long long* const empty = full + kTmaStages;
  const int tid = threadIdx.x;
  const int s = A.s;
  // column (0, 0) of the block is cell (xb, j0 - 1). The x coordinate of a TMA
  // box must be even (16-byte granularity along the contiguous axis), so a
  // block advances by TX - 2 columns: columns 1..TX-2 update faces, column 0
  // and row 0 only provide the neighbours' edges.
  const int xb = ((s - 1) & ~1) + (kTmaTX - 2) * (int) blockIdx.x;
  const int j0 = s + (kTmaTY - 1) * (int) blockIdx.y;
  // levels of this chunk; one warm-up level below it provides E_x, E_y of k-1
  const int K0 = A.zlo - 1;
  const int kc0 = K0 + (int) blockIdx.z * chunk;
  const int kc1 = min(kc0 + chunk, A.kend);
  const int kstart = (kc0 > K0) ? kc0 - 1 : kc0;
  if (tid == 0) {
    for (int b = 0; b < kTmaStages; b++) {
      mbar_init(full + b, 1);
      mbar_init(empty + b, kTmaConsumers / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (tid >= kTmaConsumers) {
    // ---- producer warp: levels kstart .. kc1, one TMA box per array ----
    const int lane = tid - kTmaConsumers;
#pragma unroll 1
    for (int n = 0; n <= kc1 - kstart; n++) {
      const int st = n % kTmaStages, use = n / kTmaStages;
      if (use > 0) mbar_wait(empty + st, (unsigned) (use - 1) & 1u);
      if (lane == 0) mbar_expect_tx(full + st, kTmaLevelTx);
      __syncwarp();
      if (lane < kTmaArrays)
        tma_box_g2s(buf + (size_t) st * kTmaLevelD + lane * kTmaArrayD, &A.map[lane], xb,
                    j0 - 1, kstart + n, full + st);
    }
    return;
  }

  // ---- consumers ----
  const int lx = tid & (kTmaTX - 1), ly = tid / kTmaTX;
  const int i = xb + lx, j = j0 - 1 + ly;
  // edges live inside the union of the three edge boxes (CT.cpp:548-556)
  const bool ein = (i >= s && i < G.mx - s - 1 && j >= s && j < G.my - s - 1);
  const bool own = ein && lx > 0 && lx < kTmaTX - 1 && ly > 0;
  // faces this column updates (CT.cpp:646-667)
  const bool fx_col = own && j >= s + 1;
  const bool fy_col = own && i >= s + 1;
  const bool fz_col = own && i >= s + 1 && j >= s + 1;
  const double sdx = __ldg(A.sp), sdy = __ldg(A.sp + 1), sdz = __ldg(A.sp + 2);
  const int o = ly * kTmaRowD + lx;
  constexpr int AR = kTmaArrayD;
  constexpr int Yo = kTmaRowD;
  double ex_prev = 0., ey_prev = 0.;
#define VLCT_EX(p, d) ecen_((p)[TA_VY * AR + (d)], (p)[TA_BZ * AR + (d)], (p)[TA_VZ * AR + (d)], (p)[TA_BY * AR + (d)])
#define VLCT_EY(p, d) ecen_((p)[TA_VZ * AR + (d)], (p)[TA_BX * AR + (d)], (p)[TA_VX * AR + (d)], (p)[TA_BZ * AR + (d)])
#define VLCT_EZ(p, d) ecen_((p)[TA_VX * AR + (d)], (p)[TA_BY * AR + (d)], (p)[TA_VY * AR + (d)], (p)[TA_BX * AR + (d)])
  // What level k shares with the "+z" rows of level k-1 is carried from one
  // iteration to the next (47 shared-memory reads per cell and level instead
  // of 69: the consumers are bound by shared-memory bandwidth)
  double EX0 = 0., EXY = 0., EY0 = 0., EY1 = 0., f12_0 = 0., f02_0 = 0., wx0 = 0., wy0 = 0.;
  mbar_wait(full + 0, 0u);
  if (ein) {
    const double* const c0 = buf + o;           // level kstart sits in stage 0
    EX0 = VLCT_EX(c0, 0);
    EXY = VLCT_EX(c0, Yo);
    EY0 = VLCT_EY(c0, 0);
    EY1 = VLCT_EY(c0, 1);
    f12_0 = c0[TA_F12 * AR];
    f02_0 = c0[TA_F02 * AR];
    wx0 = upwind_weight(c0[TA_RX * AR]);
    wy0 = upwind_weight(c0[TA_RY * AR]);
  }
#pragma unroll 1
  for (int k = kstart; k < kc1; k++) {
    const int n = k - kstart;
    const int st0 = n % kTmaStages, st1 = (n + 1) % kTmaStages;
    const bool faces = (k >= kc0);              // not the warm-up level (block-uniform)
    const bool xy_level = faces && (k > K0);    // k >= zlo; k < kend holds in the loop
    const bool z_level = faces && (k + 1 < A.zhi);
    // the faces' old values: issued before the edges are evaluated
    double b0x = 0., b0y = 0., b0z = 0.;
    size_t fxi = 0, fyi = 0, fzi = 0;
    if (fx_col && xy_level) { fxi = fidx(G, 0, k, j, i + 1); b0x = __ldg(A.bi0[0] + fxi); }
    if (fy_col && xy_level) { fyi = fidx(G, 1, k, j + 1, i); b0y = __ldg(A.bi0[1] + fyi); }
    if (fz_col && z_level)  { fzi = fidx(G, 2, k + 1, j, i); b0z = __ldg(A.bi0[2] + fzi); }
    mbar_wait(full + st0, (unsigned) (n / kTmaStages) & 1u);
    mbar_wait(full + st1, (unsigned) ((n + 1) / kTmaStages) & 1u);
    double ex = 0., ey = 0., ez = 0.;
    if (ein) {
      const double* const c0 = buf + (size_t) st0 * kTmaLevelD + o;   // level k
      const double* const c1 = buf + (size_t) st1 * kTmaLevelD + o;   // level k+1
      // level k+1: E_x on the rows 0, +y; E_y at x, x+1; two fluxes, two weights
      const double vzZ = c1[TA_VZ * AR], bzZ = c1[TA_BZ * AR];
      const double EXZ = ecen_(c1[TA_VY * AR], bzZ, vzZ, c1[TA_BY * AR]);
      const double EXW = VLCT_EX(c1, Yo);
      const double EYZ = ecen_(vzZ, c1[TA_BX * AR], c1[TA_VX * AR], bzZ);
      const double EYZ1 = VLCT_EY(c1, 1);
      const double f12_Z = c1[TA_F12 * AR], f02_Z = c1[TA_F02 * AR];
      const double wxZ = upwind_weight(c1[TA_RX * AR]);
      const double wyZ = upwind_weight(c1[TA_RY * AR]);
      // level k only: E_z at (x, x+1) x (y, y+1), the other fluxes and weights
      const double wxY = upwind_weight(c0[TA_RX * AR + Yo]);
      const double wyX = upwind_weight(c0[TA_RY * AR + 1]);
      const double wz0 = upwind_weight(c0[TA_RZ * AR]), wzX = upwind_weight(c0[TA_RZ * AR + 1]),
                   wzY = upwind_weight(c0[TA_RZ * AR + Yo]);
      ex = edge_value(EX0, EXY, EXZ, EXW, f12_0, f12_Z, c0[TA_F21 * AR], c0[TA_F21 * AR + Yo],
                      wy0, wyZ, wz0, wzY);
      ey = edge_value(EY0, EYZ, EY1, EYZ1, c0[TA_F20 * AR], c0[TA_F20 * AR + 1], f02_0, f02_Z,
                      wz0, wzX, wx0, wxZ);
      ez = edge_value(VLCT_EZ(c0, 0), VLCT_EZ(c0, 1), VLCT_EZ(c0, Yo), VLCT_EZ(c0, Yo + 1),
                      c0[TA_F01 * AR], c0[TA_F01 * AR + Yo], c0[TA_F10 * AR],
                      c0[TA_F10 * AR + 1], wx0, wxY, wy0, wyX);
      EX0 = EXZ; EXY = EXW; EY0 = EYZ; EY1 = EYZ1;
      f12_0 = f12_Z; f02_0 = f02_Z; wx0 = wxZ; wy0 = wyZ;
    }
    // this warp is done with the stage of level k
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(empty + st0);
    if (faces) {
      X[0][ly][lx] = ex;
      X[1][ly][lx] = ey;
      X[2][ly][lx] = ez;
      consumer_barrier();
      if (fx_col && xy_level) {
        // x face i+1: B -= dt/dy (E_z(j) - E_z(j-1)) - dt/dz (E_y(k) - E_y(k-1))
        const double ez_jm = X[2][ly - 1][lx];
        A.bi_out[0][fxi] = b0x - sdy * (ez - ez_jm) + sdz * (ey - ey_prev);
      }
      if (fy_col && xy_level) {
        // y face j+1: B -= dt/dz (E_x(k) - E_x(k-1)) - dt/dx (E_z(i) - E_z(i-1))
        const double ez_im = X[2][ly][lx - 1];
        A.bi_out[1][fyi] = b0y - sdz * (ex - ex_prev) + sdx * (ez - ez_im);
      }
      if (fz_col && z_level) {
        // z face k+1: B -= dt/dx (E_y(i) - E_y(i-1)) - dt/dy (E_x(j) - E_x(j-1))
        const double ey_im = X[1][ly][lx - 1];
        const double ex_jm = X[0][ly - 1][lx];
        A.bi_out[2][fzi] = b0z - sdx * (ey - ey_im) + sdy * (ex - ex_jm);
      }
      consumer_barrier();    // the exchange buffer may be rewritten
    }
    ex_prev = ex;
    ey_prev = ey;
  }
#undef VLCT_EX
#undef VLCT_EY
#undef VLCT_EZ
}

template <bool STACKED>
__global__ void __launch_bounds__(kPairBlock)
k_face_bfield2(const typename GeomFor<STACKED>::type G, const FaceArgs A, const Box box)
{
  VLCT_PAIR_IN_BOX(G, box, i, j, kl, k, va, vb);
  const ptrdiff_t Y = (ptrdiff_t) G.mx, Z = (ptrdiff_t) G.mx * (ptrdiff_t) G.my;
  // x faces: rows of mx + 1 entries, 64-bit accesses, one cell after the other
  if (va) face_component<0>(G, A, kl, k, j, i);
  if (vb) face_component<0>(G, A, kl, k, j, i + 1);
  const size_t c = cidx(G, k, j, i);
  {
    // y faces (D = 1: JD = z, KD = x); face j is edge index j-1 along y
    const Box& bx = A.box[1];
    if (j >= bx.lo[1] && j < bx.hi[1] && kl >= bx.lo[2] && kl < bx.hi[2]) {
      const size_t e = c - Y;
      const double2 ekR = ld2(A.edge[0] + e), ekL = ld2(A.edge[0] + e - Z);
      const double2 ejR = ld2(A.edge[2] + e);
      const double ejm = __ldg(A.edge[2] + e - 1);
      const double sj = __ldg(A.sp + 2), sk = __ldg(A.sp + 0);
      const size_t f = fidx(G, 1, k, j, i);
      const double2 b0 = ld2(A.bi0[1] + f);
      const double oa = b0.x - sj * (ekR.x - ekL.x) + sk * (ejR.x - ejm);
      const double ob = b0.y - sj * (ekR.y - ekL.y) + sk * (ejR.y - ejR.x);
      st_pair(A.bi_out[1] + f, oa, ob, va && i >= bx.lo[0] && i < bx.hi[0],
              vb && i + 1 >= bx.lo[0] && i + 1 < bx.hi[0]);
    }
  }
  {
    // z faces (D = 2: JD = x, KD = y); face k is edge index k-1 along z
    const Box& bx = A.box[2];
    if (j >= bx.lo[1] && j < bx.hi[1] && kl >= bx.lo[2] && kl < bx.hi[2]) {
      const size_t e = c - Z;
      const double2 ekR = ld2(A.edge[1] + e);
      const double ekm = __ldg(A.edge[1] + e - 1);
      const double2 ejR = ld2(A.edge[0] + e), ejL = ld2(A.edge[0] + e - Y);
      const double sj = __ldg(A.sp + 0), sk = __ldg(A.sp + 1);
      const size_t f = fidx(G, 2, k, j, i);
      const double2 b0 = ld2(A.bi0[2] + f);
      const double oa = b0.x - sj * (ekR.x - ekm) + sk * (ejR.x - ejL.x);
      const double ob = b0.y - sj * (ekR.y - ekR.x) + sk * (ejR.y - ejL.y);
      st_pair(A.bi_out[2] + f, oa, ob, va && i >= bx.lo[0] && i < bx.hi[0],
              vb && i + 1 >= bx.lo[0] && i + 1 < bx.hi[0]);
    }
  }
}

/// minimum over the block of two bit patterns per thread (see block_min_to)
__device__ __forceinline__ unsigned long long dt_bits_of(double v)
{ return (unsigned long long) __double_as_longlong(v); }

// (6 blocks of 128 threads: 80 registers, no spills; measured at 512^3 against
// the one-cell kernel's 11.58 ms per step: 4 blocks 11.98, 6 blocks 11.27,
// 8 blocks (64 registers, spills) 12.17 -- profiles/r2l_update_pair_occupancy.json)
#ifndef VLCT_UPDATE2_MINBLOCKS
#define VLCT_UPDATE2_MINBLOCKS 6
#endif
struct Pair { double a, b; };

template <bool MHD, bool DE, bool STACKED, bool CFL>
__global__ void __launch_bounds__(kPairBlock, VLCT_UPDATE2_MINBLOCKS)
k_update2(const Params P, const typename GeomFor<STACKED>::type G, const UpdateArgs A,
          const Box box)
{
  const int i0 = box.lo[0] & ~1;
  const unsigned npx = (unsigned) (box.hi[0] - i0 + 1) >> 1;
  const unsigned t = blockIdx.x * kPairBlock + threadIdx.x;
  const bool active = t < npx * (unsigned) (box.hi[1] - box.lo[1]);
  if (!CFL && !active) return;
  const unsigned jj = active ? t / npx : 0u;
  const int i = active ? i0 + 2 * (int) (t - jj * npx) : i0;
  const int j = box.lo[1] + (int) jj;
  const bool va = active && (i >= box.lo[0]), vb = active && (i + 1 < box.hi[0]);
  int kl, k;
  if constexpr (STACKED) unstack(G, box, blockIdx.y, kl, k);
  else kl = k = box.lo[2] + (int) blockIdx.y;
  const size_t c = cidx(G, k, j, i);
  const ptrdiff_t st[3] = { 1, (ptrdiff_t) G.mx, (ptrdiff_t) G.mx * (ptrdiff_t) G.my };
  double dt_a = DBL_MAX, dt_b = DBL_MAX;

  Pair bx = { 0., 0. }, by = { 0., 0. }, bz = { 0., 0. };
  if (MHD && active) {
    // x faces i, i+1, i+2 of the row (always inside the row of mx + 1 entries)
    const double* const fx = A.bi_out[0] + fidx(G, 0, k, j, i);
    const double f0 = __ldg(fx), f1 = __ldg(fx + 1), f2 = __ldg(fx + 2);
    bx.a = 0.5 * (f0 + f1);
    bx.b = 0.5 * (f1 + f2);
    const double2 y0 = ld2(A.bi_out[1] + fidx(G, 1, k, j, i));
    const double2 y1 = ld2(A.bi_out[1] + fidx(G, 1, k, j + 1, i));
    by.a = 0.5 * (y0.x + y1.x);
    by.b = 0.5 * (y0.y + y1.y);
    const double2 z0 = ld2(A.bi_out[2] + fidx(G, 2, k, j, i));
    const double2 z1 = ld2(A.bi_out[2] + fidx(G, 2, k + 1, j, i));
    bz.a = 0.5 * (z0.x + z1.x);
    bz.b = 0.5 * (z0.y + z1.y);
    st_pair(A.out.bx + c, bx.a, bx.b, va, vb);
    st_pair(A.out.by + c, by.a, by.b, va, vb);
    st_pair(A.out.bz + c, bz.a, bz.b, va, vb);
  }

  const Box& in = A.inner;
  const bool row_in = (j >= in.lo[1] && j < in.hi[1] && kl >= in.lo[2] && kl < in.hi[2]);
  const bool ia = va && row_in && i >= in.lo[0] && i < in.hi[0];
  const bool ib = vb && row_in && i + 1 >= in.lo[0] && i + 1 < in.hi[0];
  if (!CFL && !ia && !ib) return;

  if (ia || ib) {
  const double dtd[3] = { __ldg(A.sp), __ldg(A.sp + 1), __ldg(A.sp + 2) };
  // F_{c+1/2} - F_{c-1/2} of both cells along direction d
  auto diff = [&](const double* q, int d) {
    const double2 fc = ld2(q + c);
    Pair r;
    if (d == 0) {
      const double fl = ia ? __ldg(q + c - 1) : 0.;
      r.a = fc.x - fl; r.b = fc.y - fc.x;
    } else {
      const double2 fl = ld2(q + c - st[d]);
      r.a = fc.x - fl.x; r.b = fc.y - fl.y;
    }
    return r;
  };
  Pair d_rho = { 0., 0. }, d_mx = { 0., 0. }, d_my = { 0., 0. }, d_mz = { 0., 0. },
       d_e = { 0., 0. }, d_eint = { 0., 0. };
  Pair p_floored = { 0., 0. };
  if (DE) {
    const double2 r = ld2(A.cur_rho + c), e = ld2(A.cur_eint + c);
    p_floored.a = apply_floor((P.gamma - 1.0) * r.x * e.x, P.pressure_floor);
    p_floored.b = apply_floor((P.gamma - 1.0) * r.y * e.y, P.pressure_floor);
  }
#pragma unroll
  for (int d = 0; d < 3; d++) {
    const FluxSet& F = A.flux[d];
    const double dtdx = dtd[d];
    Pair f;
    f = diff(F.rho, d); d_rho.a -= dtdx * f.a; d_rho.b -= dtdx * f.b;
    f = diff(F.mx_, d); d_mx.a -= dtdx * f.a;  d_mx.b -= dtdx * f.b;
    f = diff(F.my_, d); d_my.a -= dtdx * f.a;  d_my.b -= dtdx * f.b;
    f = diff(F.mz_, d); d_mz.a -= dtdx * f.a;  d_mz.b -= dtdx * f.b;
    f = diff(F.e, d);   d_e.a -= dtdx * f.a;   d_e.b -= dtdx * f.b;
    if (DE) {
      f = diff(F.eint, d); d_eint.a -= dtdx * f.a; d_eint.b -= dtdx * f.b;
      f = diff(F.vbar, d);
      d_eint.a -= dtdx * p_floored.a * f.a;
      d_eint.b -= dtdx * p_floored.b * f.b;
    }
  }

  const double2 rho0 = ld2(A.u0.rho + c);
  const double2 vx0 = ld2(A.u0.vx + c), vy0 = ld2(A.u0.vy + c), vz0 = ld2(A.u0.vz + c);
  if (A.gravity) {
    const double2 ax = ld2(A.accel[0] + c), ay = ld2(A.accel[1] + c),
                  az = ld2(A.accel[2] + c);
    const double dt = __ldg(A.sp + 3);
    d_mx.a += dt * rho0.x * ax.x; d_mx.b += dt * rho0.y * ax.y;
    d_my.a += dt * rho0.x * ay.x; d_my.b += dt * rho0.y * ay.y;
    d_mz.a += dt * rho0.x * az.x; d_mz.b += dt * rho0.y * az.y;
    d_e.a += dt * rho0.x * ((vx0.x * ax.x) + (vy0.x * ay.x) + (vz0.x * az.x));
    d_e.b += dt * rho0.y * ((vx0.y * ax.y) + (vy0.y * ay.y) + (vz0.y * az.y));
  }
  const double2 et0 = ld2(A.u0.etot + c);
  double2 ei0 = make_double2(0., 0.);
  if (DE) ei0 = ld2(A.u0.eint + c);

  // conserved update, floors, dual-energy sync of one cell (as in k_update)
  auto finish = [&](double old_rho, double vx0_, double vy0_, double vz0_, double etot0,
                    double eint0, double drho, double dmx, double dmy, double dmz, double de,
                    double deint, double bx_, double by_, double bz_, double& new_rho,
                    double& vx, double& vy, double& vz, double& etot, double& eint,
                    double& p, double& local_dt) {
    new_rho = old_rho + drho;
    new_rho = apply_floor(new_rho, P.density_floor);
    const double inv_new_rho = 1. / new_rho;
    vx = (vx0_ * old_rho + dmx) * inv_new_rho;
    vy = (vy0_ * old_rho + dmy) * inv_new_rho;
    vz = (vz0_ * old_rho + dmz) * inv_new_rho;
    etot = (etot0 * old_rho + de) * inv_new_rho;
    eint = 0.;
    if (DE) eint = (eint0 * old_rho + deint) * inv_new_rho;
    floor_energy_and_sync<DE, MHD>(P, new_rho, vx, vy, vz, bx_, by_, bz_, etot, eint);
    if (CFL)
      local_dt = timestep_of_cell<MHD, DE>(P, new_rho, vx, vy, vz, bx_, by_, bz_, etot, eint,
                                           A.dx, A.dy, A.dz, p);
  };
  double ra, ua, va_, wa, ea, ia_, pa = 0., rb, ub, vb_, wb, eb, ib_, pb = 0.;
  double la = DBL_MAX, lb = DBL_MAX;
  finish(rho0.x, vx0.x, vy0.x, vz0.x, et0.x, ei0.x, d_rho.a, d_mx.a, d_my.a, d_mz.a, d_e.a,
         d_eint.a, bx.a, by.a, bz.a, ra, ua, va_, wa, ea, ia_, pa, la);
  finish(rho0.y, vx0.y, vy0.y, vz0.y, et0.y, ei0.y, d_rho.b, d_mx.b, d_my.b, d_mz.b, d_e.b,
         d_eint.b, bx.b, by.b, bz.b, rb, ub, vb_, wb, eb, ib_, pb, lb);
  if (CFL) {
    if (ia) dt_a = la;
    if (ib) dt_b = lb;
    st_pair(A.pressure + c, pa, pb, ia, ib);
  }
  st_pair(A.out.rho + c, ra, rb, ia, ib);
  st_pair(A.out.vx + c, ua, ub, ia, ib);
  st_pair(A.out.vy + c, va_, vb_, ia, ib);
  st_pair(A.out.vz + c, wa, wb, ia, ib);
  st_pair(A.out.etot + c, ea, eb, ia, ib);
  if (DE) st_pair(A.out.eint + c, ia_, ib_, ia, ib);
  }
  if (CFL) {
    // non-negative doubles order like their bit patterns, NaNs above +inf
    const unsigned long long ba = dt_bits_of(dt_a), bb = dt_bits_of(dt_b);
    block_min_to(A.dt_bits, __longlong_as_double((long long) (bb < ba ? bb : ba)));
  }
}

// ---------------------------------------------------------------------------
// passive scalars without flux arrays, single block: one thread = one scalar
// of one (x, y) column, marching along z
// ---------------------------------------------------------------------------
// Per face the sweeps of the reference reconstruct the specific scalar on both
// sides, upwind by the sign of the density flux and multiply by it
// (EnzoReconstructor*, riemann/EnzoRiemannUtils.hpp:224-249,267-314), and the
// update takes the divergence of those fluxes (EnzoIntegrationQuanUpdate.cpp).
// Here a thread does exactly that for the six faces of its cell, level after
// level: the five z neighbours and the z slopes roll through registers (each
// specific value crosses L2 once for the z direction, each z slope is evaluated
// once), the x neighbours are L1 hits, the y neighbours L1 / L2 hits, and the
// four blocks that work on the four scalars of a tile run side by side, so the
// density fluxes they share are L2 hits. Nothing but the scalars' own arrays
// and the density fluxes is touched: 3 + 3/nsc doubles per scalar and cell,
// where flux arrays written by the sweeps cost 14. Same expressions in the
// same order as the sweeps + update: bit-identical.
struct ScalarArgs {
  const double* spec[kMaxPassive];   // specific scalars of the stage's input
  const double* u0[kMaxPassive];     // conserved scalars at the start of the step
  double* out[kMaxPassive];
  const double* frho[3];             // density fluxes of the three sweeps
  const double* sp;                  // dt/dx, dt/dy, dt/dz of the stage
  double theta;
  int nsc;
};
constexpr int kScalarRows = 4;       // 32 x 4 columns per block

template <int RECON>
__device__ __forceinline__ double scalar_slope(double a, double b, double c, double theta)
{ return limited_slope<RECON>(a, b, c, theta); }

template <int RECON>
__global__ void __launch_bounds__(32 * kScalarRows)
k_scalar_update(const __grid_constant__ ScalarArgs A, const GeomLite G, const Box box,
                const int chunk)
{
  constexpr bool PLM = (RECON != RECON_NN);
  const int s = (int) (blockIdx.x % (unsigned) A.nsc);
  const int i = box.lo[0] + (int) (blockIdx.x / (unsigned) A.nsc) * 32 + (int) (threadIdx.x & 31);
  const int j = box.lo[1] + (int) blockIdx.y * kScalarRows + (int) (threadIdx.x >> 5);
  if (i >= box.hi[0] || j >= box.hi[1]) return;
  const int k0 = box.lo[2] + (int) blockIdx.z * chunk;
  const int k1 = min(k0 + chunk, box.hi[2]);
  const ptrdiff_t Y = (ptrdiff_t) G.mx, Z = (ptrdiff_t) G.mx * (ptrdiff_t) G.my;
  const double theta = A.theta;
  const double dtdx = __ldg(A.sp), dtdy = __ldg(A.sp + 1), dtdz = __ldg(A.sp + 2);
  size_t c = cidx(G, k0, j, i);
  const double* __restrict__ q = A.spec[s];
  const double* __restrict__ u0 = A.u0[s];
  double* __restrict__ out = A.out[s];
  const double* __restrict__ fx = A.frho[0];
  const double* __restrict__ fy = A.frho[1];
  const double* __restrict__ fz = A.frho[2];

  // the z window: values at k-2 .. k+1, slopes at k-1 and k, flux at the lower face
  double wm2 = 0., wm1 = __ldg(q + c - Z), w0 = __ldg(q + c), wp1 = __ldg(q + c + Z);
  double dm = 0., d0 = 0.;
  if (PLM) {
    wm2 = __ldg(q + c - 2 * Z);
    dm = scalar_slope<RECON>(wm2, wm1, w0, theta);
    d0 = scalar_slope<RECON>(wm1, w0, wp1, theta);
  }
  double fz_l = __ldg(fz + c - Z);
#pragma unroll 1
  for (int k = k0; k < k1; k++, c += Z) {
    double d_s = 0.;
    {   // x faces
      const double xm1 = __ldg(q + c - 1), xp1 = __ldg(q + c + 1);
      double sl_l = xm1, sr_l = w0, sl_c = w0, sr_c = xp1;
      if (PLM) {
        const double xm2 = __ldg(q + c - 2), xp2 = __ldg(q + c + 2);
        const double a = scalar_slope<RECON>(xm2, xm1, w0, theta);
        const double b = scalar_slope<RECON>(xm1, w0, xp1, theta);
        const double e = scalar_slope<RECON>(w0, xp1, xp2, theta);
        sl_l = xm1 + a * 0.5; sr_l = w0 - b * 0.5;
        sl_c = w0 + b * 0.5;  sr_c = xp1 - e * 0.5;
      }
      d_s -= dtdx * (passive_flux(sl_c, sr_c, __ldg(fx + c)) -
                     passive_flux(sl_l, sr_l, __ldg(fx + c - 1)));
    }
    {   // y faces
      const double ym1 = __ldg(q + c - Y), yp1 = __ldg(q + c + Y);
      double sl_l = ym1, sr_l = w0, sl_c = w0, sr_c = yp1;
      if (PLM) {
        const double ym2 = __ldg(q + c - 2 * Y), yp2 = __ldg(q + c + 2 * Y);
        const double a = scalar_slope<RECON>(ym2, ym1, w0, theta);
        const double b = scalar_slope<RECON>(ym1, w0, yp1, theta);
        const double e = scalar_slope<RECON>(w0, yp1, yp2, theta);
        sl_l = ym1 + a * 0.5; sr_l = w0 - b * 0.5;
        sl_c = w0 + b * 0.5;  sr_c = yp1 - e * 0.5;
      }
      d_s -= dtdy * (passive_flux(sl_c, sr_c, __ldg(fy + c)) -
                     passive_flux(sl_l, sr_l, __ldg(fy + c - Y)));
    }
    double wp2 = 0., dp = 0.;
    {   // z faces: the window
      double sl_l = wm1, sr_l = w0, sl_c = w0, sr_c = wp1;
      if (PLM) {
        wp2 = __ldg(q + c + 2 * Z);
        dp = scalar_slope<RECON>(w0, wp1, wp2, theta);
        sl_l = wm1 + dm * 0.5; sr_l = w0 - d0 * 0.5;
        sl_c = w0 + d0 * 0.5;  sr_c = wp1 - dp * 0.5;
      }
      const double fz_c = __ldg(fz + c);
      d_s -= dtdz * (passive_flux(sl_c, sr_c, fz_c) - passive_flux(sl_l, sr_l, fz_l));
      fz_l = fz_c;
    }
    out[c] = __ldg(u0 + c) + d_s;
    // roll
    wm2 = wm1; wm1 = w0; w0 = wp1;
    if (PLM) { wp1 = wp2; dm = d0; d0 = dp; }
    else if (k + 1 < k1) wp1 = __ldg(q + c + 2 * Z);
  }
}

// ---------------------------------------------------------------------------
// timestep: hydro-mhd/EnzoMethodMHDVlct.cpp:551-588,
//           hydro-mhd/EnzoMHDIntegratorStageCommands.cpp:299-366
// ---------------------------------------------------------------------------
template <bool MHD, bool DE>
__device__ __forceinline__ double timestep_cell_at(const Params& P, const State& u,
                                                   double* pressure, size_t c,
                                                   double dx, double dy, double dz)
{
  const double rho = u.rho[c];
  const double vx = u.vx[c], vy = u.vy[c], vz = u.vz[c];
  double bx = 0., by = 0., bz = 0.;
  if (MHD) { bx = u.bx[c]; by = u.by[c]; bz = u.bz[c]; }
  double etot = u.etot[c], eint = 0., p;
  if (DE) eint = u.eint[c];
  const double local_dt = timestep_of_cell<MHD, DE>(P, rho, vx, vy, vz, bx, by, bz, etot,
                                                    eint, dx, dy, dz, p);
  if (DE) { u.etot[c] = etot; u.eint[c] = eint; }
  pressure[c] = p;
  return local_dt;
}

template <bool MHD, bool DE>
__global__ void __launch_bounds__(256)
k_timestep(const Params P, const Geom G, const State u, double* pressure,
           double dx, double dy, double dz, unsigned long long* dt_bits,
           const size_t c_begin, const size_t c_end, const size_t rep_stride)
{
  // blockIdx.y = block of a stacked batch (cells [c_begin, c_end) of each)
  const size_t shift = (size_t) blockIdx.y * rep_stride;
  const size_t n = c_end + shift;
  double local_min = DBL_MAX;
  for (size_t c = c_begin + shift + (size_t) blockIdx.x * blockDim.x + threadIdx.x;
       c < n; c += (size_t) gridDim.x * blockDim.x)
    local_min = std_min(local_min,
                        timestep_cell_at<MHD, DE>(P, u, pressure, c, dx, dy, dz));
  block_min_to(dt_bits, local_min);
}

/// the same over a list of boxes (blockIdx.y = box): the cells a CFL-folding
/// update kernel did not touch -- the ghost shell around its inner box
struct BoxList { Box b[6]; int count; };

template <bool MHD, bool DE>
__global__ void __launch_bounds__(256)
k_timestep_boxes(const Params P, const GeomLite G, const State u, double* pressure,
                 double dx, double dy, double dz, unsigned long long* dt_bits,
                 const __grid_constant__ BoxList L)
{
  const Box& b = L.b[blockIdx.y];
  const size_t nx = (size_t) (b.hi[0] - b.lo[0]), ny = (size_t) (b.hi[1] - b.lo[1]);
  const size_t total = nx * ny * (size_t) (b.hi[2] - b.lo[2]);
  double local_min = DBL_MAX;
  for (size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (size_t) gridDim.x * blockDim.x) {
    const int i = b.lo[0] + (int) (t % nx);
    const int j = b.lo[1] + (int) ((t / nx) % ny);
    const int k = b.lo[2] + (int) (t / (nx * ny));
    local_min = std_min(local_min, timestep_cell_at<MHD, DE>(P, u, pressure,
                                                             cidx(G, k, j, i), dx, dy, dz));
  }
  block_min_to(dt_bits, local_min);
}

__global__ void k_set_u64(unsigned long long* p, unsigned long long v) { *p = v; }

/// dt_out = courant * (minimum found by k_timestep); cpp:585-587
__global__ void k_finish_dt(const unsigned long long* bits, double courant,
                            double* dt_out)
{ *dt_out = __longlong_as_double((long long) *bits) * courant; }

/// Per-stage constants of one step, from a dt that lives on the host or on the
/// device: out[4*stage + {0,1,2,3}] = dt_stage/dx, /dy, /dz, dt_stage, where
/// dt_stage = dt/2 for the predictor of the two-stage scheme
/// (EnzoMethodMHDVlct.cpp:462-464) and dt otherwise.
__global__ void k_step_params(const double* dt_dev, double dt_host, int nstages,
                              double wx, double wy, double wz, double* out)
{
  const double dt = dt_dev ? *dt_dev : dt_host;
  for (int stage = 0; stage < nstages; stage++) {
    const double cur = (stage + 1 < nstages) ? dt / 2. : dt;
    out[4 * stage + 0] = cur / wx;
    out[4 * stage + 1] = cur / wy;
    out[4 * stage + 2] = cur / wz;
    out[4 * stage + 3] = cur;
  }
}

// ---------------------------------------------------------------------------
// ghost-zone helpers (stand-ins for the refresh phase on a unigrid)
// ---------------------------------------------------------------------------
/// the periodic wrap of all fields of a block along one axis in one launch:
/// blockIdx.y = field (face: -1 cell-centred, else the axis it is face-centred on)
__global__ void __launch_bounds__(256)
k_wrap_axis_all(const __grid_constant__ WrapTable T, int mz, int my, int mx, int axis,
                int n, int g)
{
  const int face = T.face[blockIdx.y];
  double* const p = T.p[blockIdx.y];
  const int n0 = mz + (face == 2), n1 = my + (face == 1), n2 = mx + (face == 0);
  const int cen = (face == axis) ? 1 : 0;
  const int ext[3] = { n2, n1, n0 };
  int sh[3] = { ext[0], ext[1], ext[2] };
  sh[axis] = 2 * g;
  const size_t total = (size_t) sh[0] * sh[1] * sh[2];
  for (size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (size_t) gridDim.x * blockDim.x) {
    int idx[3];
    idx[0] = (int) (t % sh[0]);
    idx[1] = (int) ((t / sh[0]) % sh[1]);
    idx[2] = (int) (t / ((size_t) sh[0] * sh[1]));
    const int a = idx[axis];
    int dst, src;
    if (a < g) { dst = a; src = a + n; }
    else       { dst = g + n + cen + (a - g); src = dst - n; }
    idx[axis] = dst;
    const size_t d = ((size_t) idx[2] * n1 + idx[1]) * n2 + idx[0];
    idx[axis] = src;
    const size_t s = ((size_t) idx[2] * n1 + idx[1]) * n2 + idx[0];
    p[d] = p[s];
  }
}

/// outflow / reflecting domain boundary of one field along one axis
/// (enzo-core/EnzoBoundary.cpp:164-283 reflecting, :352-466 outflow): threads
/// enumerate the g ghost layers of one side; the other two axes run over their
/// full, ghost- and centering-including extent like the reference's loops.
/// (all fields of the boundary in one launch: blockIdx.y = field; T.sign holds
/// the field's sign under reflection resp. its inflow value)
__global__ void __launch_bounds__(256)
k_boundary_axis(const __grid_constant__ BoundaryTable T, int mz, int my, int mx,
                int axis, int n, int g, int side, int type)
{
  const int face = T.face[blockIdx.y];
  double* const p = T.p[blockIdx.y];
  const double sign = T.sign[blockIdx.y];
  const int n0 = mz + (face == 2), n1 = my + (face == 1), n2 = mx + (face == 0);
  const int cen = (face == axis) ? 1 : 0;
  const int ext[3] = { n2, n1, n0 };
  int sh[3] = { ext[0], ext[1], ext[2] };
  sh[axis] = g;
  const size_t total = (size_t) sh[0] * sh[1] * sh[2];
  for (size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (size_t) gridDim.x * blockDim.x) {
    int idx[3];
    idx[0] = (int) (t % sh[0]);
    idx[1] = (int) ((t / sh[0]) % sh[1]);
    idx[2] = (int) (t / ((size_t) sh[0] * sh[1]));
    const int ig = idx[axis];
    int src, dst;
    if (type == VLCT_BOUNDARY_INFLOW) {
      // BoundaryValue::enforce (Cello/problem_BoundaryValue.cpp:190-202): the g
      // outermost layers take the value (passed in `sign`)
      idx[axis] = (side == 0) ? ig : n + g + cen + ig;
      p[((size_t) idx[2] * n1 + idx[1]) * n2 + idx[0]] = sign;
      continue;
    }
    if (type == VLCT_BOUNDARY_OUTFLOW) {
      if (side == 0) { src = g;               dst = g - ig - 1; }
      else           { src = n + g - 1 + cen; dst = src + ig + 1; }
    } else {
      if (side == 0) { src = g + cen + ig;    dst = g - ig - 1; }
      else           { src = n + g - 1 - ig;  dst = n + g + ig + cen; }
    }
    idx[axis] = src;
    const double v = p[((size_t) idx[2] * n1 + idx[1]) * n2 + idx[0]];
    idx[axis] = dst;
    p[((size_t) idx[2] * n1 + idx[1]) * n2 + idx[0]] =
        (type == VLCT_BOUNDARY_OUTFLOW) ? v : sign * v;
  }
}

/// dt/dx * flux through one face of the block, over the active transverse
/// extent (EnzoMethodMHDVlct.cpp:250-330). flux: cell-strided array, entry at
/// index `at` along `dim`; out: packed (n1, n0), slower axis first.
__global__ void __launch_bounds__(256)
k_face_flux(const double* __restrict__ flux, const double* __restrict__ dtdx,
            double* __restrict__ out, int mx, int my, int dim, int at, int n0,
            int n1, int g0, int g1)
{
  const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (unsigned) n0 * (unsigned) n1) return;
  const int i0 = (int) (t % (unsigned) n0), i1 = (int) (t / (unsigned) n0);
  const int a0 = (dim == 0) ? 1 : 0, a1 = (dim == 2) ? 1 : 2;
  int idx[3];
  idx[dim] = at; idx[a0] = g0 + i0; idx[a1] = g1 + i1;
  const size_t c = ((size_t) idx[2] * my + idx[1]) * mx + idx[0];
  out[t] = __ldg(dtdx) * __ldg(flux + c);
}

/// gather the blocks of a batch into their stacked array, or scatter them back:
/// blockIdx.y = block; ptrs[block] = that block's own (device) array of `count`
/// elements, stacked at stride `stride`
__global__ void __launch_bounds__(256)
k_batch_copy(double* __restrict__ stacked, double* const* __restrict__ ptrs,
             size_t count, size_t stride, int to_stacked)
{
  double* p = ptrs[blockIdx.y];
  double* q = stacked + (size_t) blockIdx.y * stride;
  for (size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x; t < count;
       t += (size_t) gridDim.x * blockDim.x) {
    if (to_stacked) q[t] = p[t];
    else            p[t] = q[t];
  }
}

/// ghost-exchange slabs of all fields of a block along one axis, packed into /
/// unpacked from one contiguous buffer in one launch: blockIdx.y = field
__global__ void __launch_bounds__(256)
k_slab_copy_all(const __grid_constant__ SlabTable T, int mz, int my, int mx, int axis,
                int width, double* buffer, int pack)
{
  const int face = T.face[blockIdx.y];
  double* const field = T.p[blockIdx.y];
  double* const buf = buffer + T.off[blockIdx.y];
  const int lo = T.lo[blockIdx.y];
  const int n1 = my + (face == 1), n2 = mx + (face == 0);
  int sh[3] = { n2, n1, mz + (face == 2) };
  sh[axis] = width;
  const size_t total = (size_t) sh[0] * sh[1] * sh[2];
  for (size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (size_t) gridDim.x * blockDim.x) {
    int idx[3];
    idx[0] = (int) (t % sh[0]);
    idx[1] = (int) ((t / sh[0]) % sh[1]);
    idx[2] = (int) (t / ((size_t) sh[0] * sh[1]));
    idx[axis] += lo;
    const size_t f = ((size_t) idx[2] * n1 + idx[1]) * n2 + idx[0];
    if (pack) buf[t] = field[f];
    else      field[f] = buf[t];
  }
}

}  // namespace

// ---------------------------------------------------------------------------
// profiler
// ---------------------------------------------------------------------------
void Profiler::begin(cudaStream_t st, const char* name)
{
  Entry e;
  e.name = name;
  cudaEventCreate(&e.beg);
  cudaEventCreate(&e.end);
  cudaEventRecord(e.beg, st);
  pending.push_back(e);
}

void Profiler::end(cudaStream_t st)
{ cudaEventRecord(pending.back().end, st); }

void Profiler::collect()
{
  for (Entry& e : pending) {
    cudaEventSynchronize(e.end);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e.beg, e.end);
    size_t idx = 0;
    for (; idx < names.size(); idx++) if (names[idx] == e.name) break;
    if (idx == names.size()) {
      names.push_back(e.name); total_ms.push_back(0.); calls.push_back(0);
    }
    total_ms[idx] += (double) ms;
    calls[idx] += 1;
    cudaEventDestroy(e.beg);
    cudaEventDestroy(e.end);
  }
  pending.clear();
}

void Profiler::reset()
{
  collect();
  names.clear(); total_ms.clear(); calls.clear();
}

// ---------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------
namespace {

/// 3-D tensor map (x fastest) of a cell-strided fp64 array for boxes of
/// (TX + 2) x (TY + 1) x 1 elements; cached per (pointer, shape). False if the driver entry
/// point is missing or the encode fails.
bool tensor_map_for(const double* p, int mx, int my, int mz, CUtensorMap* out)
{
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                               const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static std::mutex mu;
  static EncodeFn encode = nullptr;
  static bool looked_up = false;
  static std::map<std::tuple<const void*, int, int, int>, CUtensorMap> cache;
  std::lock_guard<std::mutex> lock(mu);
  if (!looked_up) {
    looked_up = true;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) ==
            cudaSuccess && q == cudaDriverEntryPointSuccess)
      encode = (EncodeFn) fn;
    else
      cudaGetLastError();
  }
  if (encode == nullptr) return false;
  const auto key = std::make_tuple((const void*) p, mx, my, mz);
  auto it = cache.find(key);
  if (it == cache.end()) {
    CUtensorMap m;
    const cuuint64_t dims[3] = { (cuuint64_t) mx, (cuuint64_t) my, (cuuint64_t) mz };
    const cuuint64_t strides[2] = { (cuuint64_t) mx * 8u, (cuuint64_t) mx * (cuuint64_t) my * 8u };
    const cuuint32_t box[3] = { (cuuint32_t) kTmaRowD, (cuuint32_t) kTmaRows, 1u };
    const cuuint32_t estr[3] = { 1u, 1u, 1u };
    if (encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*) p, dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) !=
        CUDA_SUCCESS)
      return false;
    if (cache.size() > 4096) cache.clear();
    it = cache.emplace(key, m).first;
  }
  *out = it->second;
  return true;
}

/// The TMA-staged kernels keep one block per SM busy for a whole z chunk:
/// blocks of a launch = tiles x chunks. Small blocks of cells do not fill the
/// chip with them and keep the one-cell / pair kernels. The criterion uses the
/// block's full depth, never the clipped range of a launch: all z passes of a
/// step must take the same path (the fused kernel writes no edge arrays).
bool tma_tiles_fill_chip(long long tiles, int mz, bool force)
{
  if (force) return true;       // option "pair_kernels" bit 5 (tests)
  static const long long min_blocks = [] {
    const char* e = getenv("VLCT_TMA_MIN_BLOCKS");    // A/B runs
    return (e && atoll(e) >= 0) ? atoll(e) : 148LL;
  }();
  return tiles * ((mz + 15) / 16) >= min_blocks;
}

/// z chunk of a TMA-staged launch over nk levels: about two waves of blocks,
/// 16..64 levels per chunk (a chunk costs one or two extra levels)
int tma_chunk(long long tiles, int nk)
{
  const long long want = (296 + tiles - 1) / tiles;       // chunks for two waves
  long long chunk = nk / (want > 0 ? want : 1);
  if (chunk > 64) chunk = 64;
  if (chunk < 16) chunk = 16;
  if (nk < 2 * chunk) chunk = nk;
  return (int) chunk;
}

struct Align16 {
  bool ok = true;
  void operator()(const void* p) { if (((uintptr_t) p & 15u) != 0) ok = false; }
};

}  // namespace

int default_pair_kernels()
{
  const char* e = getenv("VLCT_PAIR_MASK");
  return e ? (atoi(e) & 63) : 30;
}

void launch_primitives(const LaunchCtx& ctx, const Params& P, const Geom& G,
                       const State& cur, const Scratch& S, int stage, int stale, ZClip zc)
{
  if (P.nsc == 0) return;   // pressure is computed inside the flux kernels
  Box box = full_box(G, stale);
  if (!clip_z(box, zc)) return;
  ScopedLaunch sl(ctx, "k_specific_scalars");
  if (G.nrep > 1)
    k_specific_scalars<true><<<grid_for(G, box), kBlock, 0, ctx.st>>>(
        P.nsc, G, cur, scalar_ptrs(S.prim_sc[stage], P.nsc), box);
  else
    k_specific_scalars<false><<<grid_for(G, box), kBlock, 0, ctx.st>>>(
        P.nsc, lite(G), cur, scalar_ptrs(S.prim_sc[stage], P.nsc), box);
}

void launch_ct(const LaunchCtx& ctx, const Params& P, const Geom& G,
               const State& cur, const Scratch& S, const FaceB& bi0,
               const FaceB& bi_out, const double* step_params, int s,
               ZClip z_edge, ZClip z_face)
{
  cudaStream_t st = ctx.st;
  const int m[3] = { G.mx, G.my, G.mz };
  const int block = kBlock;
  if ((ctx.pair_mask & 16) && G.nrep == 1 && G.mx % 2 == 0 && z_face.hi > z_face.lo &&
      z_edge.hi > z_edge.lo) {
    // edge E + face B in one TMA-staged kernel (k_ct_tma). The faces of the
    // face clip, the edges they need evaluated on the way: the pass tables of
    // vlct_api.cu keep the inputs of those edges intact.
    const double* in[kTmaArrays];
    in[TA_VX] = cur.vx; in[TA_VY] = cur.vy; in[TA_VZ] = cur.vz;
    in[TA_BX] = cur.bx; in[TA_BY] = cur.by; in[TA_BZ] = cur.bz;
    in[TA_RX] = S.flux[0].rho; in[TA_RY] = S.flux[1].rho; in[TA_RZ] = S.flux[2].rho;
    in[TA_F12] = S.flux[1].bz; in[TA_F21] = S.flux[2].by; in[TA_F20] = S.flux[2].bx;
    in[TA_F02] = S.flux[0].bz; in[TA_F01] = S.flux[0].by; in[TA_F10] = S.flux[1].bx;
    Align16 al;
    for (int a = 0; a < kTmaArrays; a++) al(in[a]);
    CtTmaArgs T;
    bool ok = al.ok &&
              tma_tiles_fill_chip((long long) ((G.mx - 2 * s + kTmaTX - 3) / (kTmaTX - 2)) *
                                  ((G.my - 2 * s - 1 + kTmaTY - 2) / (kTmaTY - 1)), G.mz,
                                  (ctx.pair_mask & 32) != 0);
    for (int a = 0; a < kTmaArrays && ok; a++)
      ok = tensor_map_for(in[a], G.mx, G.my, G.mz, &T.map[a]);
    if (ok) {
      const int zlo = (s + 1 > z_face.lo) ? s + 1 : z_face.lo;
      const int zhi = (G.mz - s < z_face.hi) ? G.mz - s : z_face.hi;
      const int kend = (G.mz - s - 1 < z_face.hi) ? G.mz - s - 1 : z_face.hi;
      const int K0 = zlo - 1;
      const int ncx = G.mx - 2 * s - 1, ncy = G.my - 2 * s - 1;
      if (kend <= K0 || ncx <= 0 || ncy <= 0) return;
      for (int d = 0; d < 3; d++) { T.bi0[d] = bi0.bi[d]; T.bi_out[d] = bi_out.bi[d]; }
      T.sp = step_params;
      T.s = s; T.zlo = zlo; T.zhi = zhi; T.kend = kend;
      const int nk = kend - K0;
      // (columns from ((s - 1) & ~1) + 1 on, TX - 2 per block: see the kernel)
      const int xfirst = ((s - 1) & ~1) + 1;
      const unsigned gx = (unsigned) ((G.mx - s - 1 - xfirst + kTmaTX - 3) / (kTmaTX - 2));
      const unsigned gy = (unsigned) ((ncy + kTmaTY - 2) / (kTmaTY - 1));
      const int chunk = tma_chunk((long long) gx * gy, nk);
      const dim3 grid(gx, gy, (unsigned) ((nk + chunk - 1) / chunk));
      static bool smem_set[64] = {};
      int device = 0;
      cudaGetDevice(&device);
      if (device < 0 || device >= 64 || !smem_set[device]) {
        cudaFuncSetAttribute(k_ct_tma, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int) kCtSmemBytes);
        if (device >= 0 && device < 64) smem_set[device] = true;
      }
      ScopedLaunch sl(ctx, "k_ct_tma");
      k_ct_tma<<<grid, kTmaThreads, kCtSmemBytes, st>>>(lite(G), T, chunk);
      return;
    }
  }
  {
    EdgeArgs A;
    A.v[0] = cur.vx; A.v[1] = cur.vy; A.v[2] = cur.vz;
    A.b[0] = cur.bx; A.b[1] = cur.by; A.b[2] = cur.bz;
    for (int d = 0; d < 3; d++) {
      A.frho[d] = S.flux[d].rho;
      A.fb[d][0] = S.flux[d].bx; A.fb[d][1] = S.flux[d].by; A.fb[d][2] = S.flux[d].bz;
      A.edge[d] = S.edge[d];
      // CT.cpp:548-556: start 1 along d, 0 along j,k; stop (extent-1)
      for (int a = 0; a < 3; a++) {
        A.box[d].lo[a] = s + ((a == d) ? 1 : 0);
        A.box[d].hi[a] = m[a] - s - 1;
      }
    }
    Box box;   // union of the three component boxes
    for (int a = 0; a < 3; a++) { box.lo[a] = s; box.hi[a] = m[a] - s - 1; }
    // (the per-component boxes are tested per thread: clipping the launch box
    // clips all three)
    if (clip_z(box, z_edge)) {
      ScopedLaunch sl(ctx, "k_edge_efield");
      Align16 al;
      for (int d = 0; d < 3; d++) {
        al(A.v[d]); al(A.b[d]); al(A.frho[d]); al(A.edge[d]);
        for (int q = 0; q < 3; q++) if (q != d) al(A.fb[d][q]);
      }
      bool use_tma = (ctx.pair_mask & 8) && G.nrep == 1 && G.mx % 2 == 0 && al.ok &&
                     tma_tiles_fill_chip((long long) ((G.mx - 2 * s + kTmaTX - 1) / kTmaTX) *
                                         ((G.my - 2 * s - 1 + kTmaTY - 1) / kTmaTY), G.mz,
                                         (ctx.pair_mask & 32) != 0);
      EdgeTmaArgs T;
      if (use_tma) {
        // inputs staged by TMA boxes (k_edge_efield_tma); without the driver's
        // encode entry point the other kernels run
        const double* in[kTmaArrays];
        in[TA_VX] = A.v[0]; in[TA_VY] = A.v[1]; in[TA_VZ] = A.v[2];
        in[TA_BX] = A.b[0]; in[TA_BY] = A.b[1]; in[TA_BZ] = A.b[2];
        in[TA_RX] = A.frho[0]; in[TA_RY] = A.frho[1]; in[TA_RZ] = A.frho[2];
        in[TA_F12] = A.fb[1][2]; in[TA_F21] = A.fb[2][1]; in[TA_F20] = A.fb[2][0];
        in[TA_F02] = A.fb[0][2]; in[TA_F01] = A.fb[0][1]; in[TA_F10] = A.fb[1][0];
        for (int a = 0; a < kTmaArrays && use_tma; a++)
          use_tma = tensor_map_for(in[a], G.mx, G.my, G.mz, &T.map[a]);
      }
      if (use_tma) {
        for (int d = 0; d < 3; d++) T.edge[d] = A.edge[d];
        T.s = s; T.k0 = box.lo[2]; T.kend = box.hi[2];
        const int nk = box.hi[2] - box.lo[2];
        const int x0 = s & ~1;
        const unsigned gx = (unsigned) ((G.mx - s - 1 - x0 + kTmaTX - 1) / kTmaTX);
        const unsigned gy = (unsigned) ((G.my - 2 * s - 1 + kTmaTY - 1) / kTmaTY);
        const int chunk = tma_chunk((long long) gx * gy, nk);
        const dim3 grid(gx, gy, (unsigned) ((nk + chunk - 1) / chunk));
        static bool smem_set[64] = {};
        int device = 0;
        cudaGetDevice(&device);
        if (device < 0 || device >= 64 || !smem_set[device]) {
          cudaFuncSetAttribute(k_edge_efield_tma, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int) kTmaSmemBytes);
          if (device >= 0 && device < 64) smem_set[device] = true;
        }
        k_edge_efield_tma<<<grid, kTmaThreads, kTmaSmemBytes, st>>>(lite(G), T, chunk);
      } else if ((ctx.pair_mask & 1) && G.mx % 2 == 0 && al.ok) {
        const dim3 grid = pair_grid_for(G, box);
        if (G.nrep > 1) k_edge_efield2<true><<<grid, kPairBlock, 0, st>>>(G, A, box);
        else            k_edge_efield2<false><<<grid, kPairBlock, 0, st>>>(lite(G), A, box);
      } else {
        if (G.nrep > 1) k_edge_efield<true><<<grid_for(G, box), block, 0, st>>>(G, A, box);
        else            k_edge_efield<false><<<grid_for(G, box), block, 0, st>>>(lite(G), A, box);
      }
    }
  }
  {
    FaceArgs A;
    A.sp = step_params;
    for (int d = 0; d < 3; d++) {
      A.edge[d] = S.edge[d];
      A.bi0[d] = bi0.bi[d];
      A.bi_out[d] = bi_out.bi[d];
      // CT.cpp:646-667: interior faces along d, inner cells along j,k
      for (int a = 0; a < 3; a++) {
        A.box[d].lo[a] = s + 1;
        A.box[d].hi[a] = (a == d) ? (m[a] - s) : (m[a] - s - 1);
      }
    }
    Box box;
    for (int a = 0; a < 3; a++) { box.lo[a] = s + 1; box.hi[a] = m[a] - s; }
    if (clip_z(box, z_face)) {
      ScopedLaunch sl(ctx, "k_face_bfield");
      Align16 al;
      for (int d = 0; d < 3; d++) al(A.edge[d]);
      for (int d = 1; d < 3; d++) { al(A.bi0[d]); al(A.bi_out[d]); }
      if ((ctx.pair_mask & 2) && G.mx % 2 == 0 && al.ok) {
        const dim3 grid = pair_grid_for(G, box);
        if (G.nrep > 1) k_face_bfield2<true><<<grid, kPairBlock, 0, st>>>(G, A, box);
        else            k_face_bfield2<false><<<grid, kPairBlock, 0, st>>>(lite(G), A, box);
      } else {
        if (G.nrep > 1) k_face_bfield<true><<<grid_for(G, box), block, 0, st>>>(G, A, box);
        else            k_face_bfield<false><<<grid_for(G, box), block, 0, st>>>(lite(G), A, box);
      }
    }
  }
}

void launch_update(const LaunchCtx& ctx, const Params& P, const Geom& G,
                   const State& u0, const State& cur, const State& out,
                   const Scratch& S, const FaceB& bi_out,
                   const double* accel[3], bool gravity,
                   const double* step_params, int stage, int recon, int s, ZClip zc,
                   const CflFold* cfl)
{
  cudaStream_t st = ctx.st;
  UpdateArgs A;
  A.u0 = u0; A.out = out;
  for (int d = 0; d < 3; d++) {
    A.flux[d] = S.flux[d];
    A.bi_out[d] = bi_out.bi[d];
    A.accel[d] = gravity ? accel[d] : nullptr;
  }
  A.cur_rho = cur.rho;
  A.cur_eint = cur.eint;
  A.sp = step_params;
  A.gravity = gravity ? 1 : 0;
  A.inner = full_box(G, s + 1);
  for (int n = 0; n < kMaxPassive; n++) A.spec[n] = (n < P.nsc) ? S.prim_sc[stage][n] : nullptr;
  A.recon = recon;
  // single blocks: the scalars get a z-marching kernel of their own
  const bool scalar_kernel = (P.nsc > 0 && P.nsc_flux == 0 && G.nrep == 1);
  A.scalars_elsewhere = scalar_kernel ? 1 : 0;
  if (scalar_kernel) {
    Box sbox = A.inner;
    if (clip_z(sbox, zc)) {
      ScalarArgs SA;
      for (int n = 0; n < kMaxPassive; n++) {
        SA.spec[n] = (n < P.nsc) ? S.prim_sc[stage][n] : nullptr;
        SA.u0[n] = (n < P.nsc) ? u0.sc[n] : nullptr;
        SA.out[n] = (n < P.nsc) ? out.sc[n] : nullptr;
      }
      for (int d = 0; d < 3; d++) SA.frho[d] = S.flux[d].rho;
      SA.sp = step_params;
      SA.theta = P.theta;
      SA.nsc = P.nsc;
      const int nz = sbox.hi[2] - sbox.lo[2];
      int chunk = 64;
      if (nz < 2 * chunk) chunk = nz;
      const dim3 sgrid((unsigned) ((sbox.hi[0] - sbox.lo[0] + 31) / 32) * (unsigned) P.nsc,
                       (unsigned) ((sbox.hi[1] - sbox.lo[1] + kScalarRows - 1) / kScalarRows),
                       (unsigned) ((nz + chunk - 1) / chunk));
      ScopedLaunch sl(ctx, recon == VLCT_RECON_NN ? "k_scalar_update_nn"
                                                  : "k_scalar_update_plm");
      const int threads = 32 * kScalarRows;
      if (recon == VLCT_RECON_NN)
        k_scalar_update<RECON_NN><<<sgrid, threads, 0, st>>>(SA, lite(G), sbox, chunk);
      else if (recon == VLCT_RECON_PLM_ATHENA)
        k_scalar_update<RECON_PLM_ATHENA><<<sgrid, threads, 0, st>>>(SA, lite(G), sbox, chunk);
      else
        k_scalar_update<RECON_PLM_ENZO><<<sgrid, threads, 0, st>>>(SA, lite(G), sbox, chunk);
    }
  }
  A.pressure = cfl ? cfl->pressure : nullptr;
  A.dt_bits = cfl ? cfl->dt_bits : nullptr;
  A.dx = cfl ? cfl->width[0] : 0.; A.dy = cfl ? cfl->width[1] : 0.;
  A.dz = cfl ? cfl->width[2] : 0.;
  // with CT the centred B is rewritten on the whole [s, m-s)^3 region
  Box box = P.mhd ? full_box(G, s) : A.inner;
  if (clip_z(box, zc)) {
    const int block = kBlock; const dim3 grid = grid_for(G, box);
    ScopedLaunch sl(ctx, cfl ? "k_update_cfl" : "k_update");
    const bool in_kernel_scalars = (P.nsc > 0 && !scalar_kernel);
    // pair kernel (128-bit accesses): even rows, 16-byte aligned arrays, no
    // scalar work inside the kernel
    Align16 al;
    {
      const State* sts[2] = { &A.u0, &A.out };
      for (const State* q : sts) {
        al(q->rho); al(q->vx); al(q->vy); al(q->vz); al(q->etot);
        if (P.de) al(q->eint);
      }
      if (P.mhd) { al(A.out.bx); al(A.out.by); al(A.out.bz); al(A.bi_out[1]); al(A.bi_out[2]); }
      for (int d = 0; d < 3; d++) {
        const FluxSet& F = A.flux[d];
        al(F.rho); al(F.mx_); al(F.my_); al(F.mz_); al(F.e);
        if (P.de) { al(F.eint); al(F.vbar); }
        if (gravity) al(A.accel[d]);
      }
      if (P.de) { al(A.cur_rho); al(A.cur_eint); }
      if (cfl) al(A.pressure);
    }
    const bool pair = (ctx.pair_mask & 4) && !in_kernel_scalars && G.mx % 2 == 0 && al.ok;
    const dim3 pgrid = pair_grid_for(G, box);
#define VLCT_UPDATE3(MHD_, DE_, CFL_, SCAL_)                                      \
  do {                                                                          \
    if (G.nrep > 1) k_update<MHD_, DE_, true, CFL_, SCAL_><<<grid, block, 0, st>>>(P, G, A, box);  \
    else            k_update<MHD_, DE_, false, CFL_, SCAL_><<<grid, block, 0, st>>>(P, lite(G), A, box); \
  } while (0)
#define VLCT_UPDATE2P(MHD_, DE_, CFL_)                                            \
  do {                                                                          \
    if (G.nrep > 1) k_update2<MHD_, DE_, true, CFL_><<<pgrid, kPairBlock, 0, st>>>(P, G, A, box);  \
    else            k_update2<MHD_, DE_, false, CFL_><<<pgrid, kPairBlock, 0, st>>>(P, lite(G), A, box); \
  } while (0)
#define VLCT_UPDATE2(MHD_, DE_, CFL_)                                             \
  do { if (pair) VLCT_UPDATE2P(MHD_, DE_, CFL_);                                 \
       else if (in_kernel_scalars) VLCT_UPDATE3(MHD_, DE_, CFL_, true);         \
       else VLCT_UPDATE3(MHD_, DE_, CFL_, false); } while (0)
#define VLCT_UPDATE(MHD_, DE_)                                                   \
  do { if (cfl) VLCT_UPDATE2(MHD_, DE_, true); else VLCT_UPDATE2(MHD_, DE_, false); } while (0)
    if (P.mhd) {
      if (P.de) VLCT_UPDATE(true, true);
      else      VLCT_UPDATE(true, false);
    } else {
      if (P.de) VLCT_UPDATE(false, true);
      else      VLCT_UPDATE(false, false);
    }
#undef VLCT_UPDATE
#undef VLCT_UPDATE2
#undef VLCT_UPDATE2P
#undef VLCT_UPDATE3
  }
  if (cfl == nullptr) return;
  // the cells of the levels [zc.lo, zc.hi) outside the inner box: up to six
  // slabs (whole levels below / above, then y slabs, then x slabs)
  const Box& in = A.inner;
  const int zlo = zc.lo < 0 ? 0 : zc.lo, zhi = zc.hi > G.mz ? G.mz : zc.hi;
  if (zhi <= zlo) return;
  BoxList L;
  L.count = 0;
  auto add = [&](int x0, int x1, int y0, int y1, int z0, int z1) {
    if (z0 < zlo) z0 = zlo;
    if (z1 > zhi) z1 = zhi;
    if (x1 <= x0 || y1 <= y0 || z1 <= z0) return;
    Box b;
    b.lo[0] = x0; b.hi[0] = x1; b.lo[1] = y0; b.hi[1] = y1; b.lo[2] = z0; b.hi[2] = z1;
    L.b[L.count++] = b;
  };
  add(0, G.mx, 0, G.my, 0, in.lo[2]);
  add(0, G.mx, 0, G.my, in.hi[2], G.mz);
  add(0, G.mx, 0, in.lo[1], in.lo[2], in.hi[2]);
  add(0, G.mx, in.hi[1], G.my, in.lo[2], in.hi[2]);
  add(0, in.lo[0], in.lo[1], in.hi[1], in.lo[2], in.hi[2]);
  add(in.hi[0], G.mx, in.lo[1], in.hi[1], in.lo[2], in.hi[2]);
  if (L.count == 0) return;
  size_t largest = 0;
  for (int n = 0; n < L.count; n++) {
    const Box& b = L.b[n];
    const size_t cnt = (size_t) (b.hi[0] - b.lo[0]) * (b.hi[1] - b.lo[1]) * (b.hi[2] - b.lo[2]);
    if (cnt > largest) largest = cnt;
  }
  int blocks_x = (int) ((largest + 255) / 256);
  if (blocks_x > 148 * 4) blocks_x = 148 * 4;
  const dim3 grid(blocks_x, L.count, 1);
  const State u = out;
  ScopedLaunch sl(ctx, "k_timestep_shell");
  if (P.mhd) {
    if (P.de) k_timestep_boxes<true, true><<<grid, 256, 0, st>>>(P, lite(G), u, cfl->pressure, cfl->width[0], cfl->width[1], cfl->width[2], cfl->dt_bits, L);
    else      k_timestep_boxes<true, false><<<grid, 256, 0, st>>>(P, lite(G), u, cfl->pressure, cfl->width[0], cfl->width[1], cfl->width[2], cfl->dt_bits, L);
  } else {
    if (P.de) k_timestep_boxes<false, true><<<grid, 256, 0, st>>>(P, lite(G), u, cfl->pressure, cfl->width[0], cfl->width[1], cfl->width[2], cfl->dt_bits, L);
    else      k_timestep_boxes<false, false><<<grid, 256, 0, st>>>(P, lite(G), u, cfl->pressure, cfl->width[0], cfl->width[1], cfl->width[2], cfl->dt_bits, L);
  }
}

void launch_timestep_reset(const LaunchCtx& ctx, unsigned long long* dt_bits)
{
  ScopedLaunch sl0(ctx, "k_set_u64");
  k_set_u64<<<1, 1, 0, ctx.st>>>(dt_bits, 0x7fefffffffffffffULL);   // DBL_MAX
}

void launch_timestep(const LaunchCtx& ctx, const Params& P, const Geom& G,
                     const State& u, double* pressure, const double* width,
                     unsigned long long* dt_bits, ZClip zc)
{
  cudaStream_t st = ctx.st;
  const int zlo = zc.lo < 0 ? 0 : zc.lo, zhi = zc.hi > G.mz ? G.mz : zc.hi;
  if (zhi <= zlo) return;
  const size_t plane = (size_t) G.mx * (size_t) G.my;
  const size_t c0 = plane * (size_t) zlo, c1 = plane * (size_t) zhi;
  const size_t n = c1 - c0;
  int blocks_x = (int) ((n + 255) / 256);
  const int max_blocks = (148 * 16 + G.nrep - 1) / G.nrep;
  if (blocks_x > max_blocks) blocks_x = max_blocks;
  const dim3 blocks(blocks_x, G.nrep, 1);
  const size_t rep_stride = plane * (size_t) G.zper;
  ScopedLaunch sl(ctx, "k_timestep");
  if (P.mhd) {
    if (P.de) k_timestep<true, true><<<blocks, 256, 0, st>>>(P, G, u, pressure, width[0], width[1], width[2], dt_bits, c0, c1, rep_stride);
    else      k_timestep<true, false><<<blocks, 256, 0, st>>>(P, G, u, pressure, width[0], width[1], width[2], dt_bits, c0, c1, rep_stride);
  } else {
    if (P.de) k_timestep<false, true><<<blocks, 256, 0, st>>>(P, G, u, pressure, width[0], width[1], width[2], dt_bits, c0, c1, rep_stride);
    else      k_timestep<false, false><<<blocks, 256, 0, st>>>(P, G, u, pressure, width[0], width[1], width[2], dt_bits, c0, c1, rep_stride);
  }
}

void launch_finish_dt(const LaunchCtx& ctx, const unsigned long long* bits,
                      double courant, double* dt_out)
{
  ScopedLaunch sl(ctx, "k_finish_dt");
  k_finish_dt<<<1, 1, 0, ctx.st>>>(bits, courant, dt_out);
}

void launch_step_params(const LaunchCtx& ctx, const double* dt_dev, double dt_host,
                        int nstages, const double* width, double* out)
{
  ScopedLaunch sl(ctx, "k_step_params");
  k_step_params<<<1, 1, 0, ctx.st>>>(dt_dev, dt_host, nstages, width[0], width[1],
                                     width[2], out);
}

void launch_wrap_axis_all(const LaunchCtx& ctx, const WrapTable& T, int mz, int my,
                          int mx, int axis, int n, int g)
{
  if (T.count == 0) return;
  // sized for the largest (face-centred) field; every field strides over its own
  const int ext[3] = { mx + 1, my + 1, mz + 1 };
  size_t total = (size_t) 2 * g;
  for (int a = 0; a < 3; a++) if (a != axis) total *= (size_t) ext[a];
  if (total == 0) return;
  int blocks = (int) ((total + 255) / 256);
  const int cap = (148 * 16 + T.count - 1) / T.count;
  if (blocks > cap) blocks = cap;
  ScopedLaunch sl(ctx, "k_wrap_axis");
  k_wrap_axis_all<<<dim3(blocks, T.count), 256, 0, ctx.st>>>(T, mz, my, mx, axis, n, g);
}

void launch_boundary_axis(const LaunchCtx& ctx, const BoundaryTable& T, int mz, int my,
                          int mx, int axis, int n, int g, int side, int type)
{
  if (T.count == 0) return;
  const int ext[3] = { mx + 1, my + 1, mz + 1 };
  size_t total = (size_t) g;
  for (int a = 0; a < 3; a++) if (a != axis) total *= (size_t) ext[a];
  if (total == 0) return;
  int blocks = (int) ((total + 255) / 256);
  const int cap = (148 * 16 + T.count - 1) / T.count;
  if (blocks > cap) blocks = cap;
  ScopedLaunch sl(ctx, "k_boundary_axis");
  k_boundary_axis<<<dim3(blocks, T.count), 256, 0, ctx.st>>>(T, mz, my, mx, axis, n, g,
                                                             side, type);
}

void launch_face_flux(const LaunchCtx& ctx, const Geom& G, const double* flux,
                      const double* dtdx, double* out, int dim, int at, int n0,
                      int n1, int g0, int g1)
{
  const unsigned total = (unsigned) n0 * (unsigned) n1;
  if (total == 0) return;
  ScopedLaunch sl(ctx, "k_face_flux");
  k_face_flux<<<(total + 255) / 256, 256, 0, ctx.st>>>(flux, dtdx, out, G.mx, G.my,
                                                       dim, at, n0, n1, g0, g1);
}

void launch_batch_copy(const LaunchCtx& ctx, double* stacked, double* const* ptrs,
                       int nblocks, size_t count, size_t stride, bool to_stacked,
                       bool over_pcie)
{
  if (count == 0 || nblocks == 0) return;
  int bx = (int) ((count + 255) / 256);
  // Device-to-device copies want the whole GPU. Copies that read or write
  // pinned host memory in place are bound by PCIe and must leave the SMs to
  // the kernels (and to the copy in the other direction) they overlap with:
  // ~2 thread blocks per SM keep far more bytes in flight than the link needs.
  const int total = over_pcie ? 148 * 2 : 148 * 16;
  int cap = (total + nblocks - 1) / nblocks;
  if (over_pcie && nblocks > total) cap = 1;
  if (bx > cap) bx = cap;
  ScopedLaunch sl(ctx, to_stacked ? "k_batch_gather" : "k_batch_scatter");
  k_batch_copy<<<dim3(bx, nblocks), 256, 0, ctx.st>>>(stacked, ptrs, count, stride,
                                                      to_stacked ? 1 : 0);
}

void launch_slab_copy_all(const LaunchCtx& ctx, const SlabTable& T, int mz, int my,
                          int mx, int axis, int width, double* buffer, bool pack)
{
  if (T.count == 0) return;
  const int ext[3] = { mx + 1, my + 1, mz + 1 };
  size_t total = (size_t) width;
  for (int a = 0; a < 3; a++) if (a != axis) total *= (size_t) ext[a];
  if (total == 0) return;
  int blocks = (int) ((total + 255) / 256);
  const int cap = (148 * 16 + T.count - 1) / T.count;
  if (blocks > cap) blocks = cap;
  ScopedLaunch sl(ctx, pack ? "k_slab_pack" : "k_slab_unpack");
  k_slab_copy_all<<<dim3(blocks, T.count), 256, 0, ctx.st>>>(T, mz, my, mx, axis, width,
                                                             buffer, pack ? 1 : 0);
}

}  // namespace vlct
