// vlct_api.cu -- the C ABI (include/vlct.h) over the CUDA kernels.
//
// Host-side mirror of EnzoMethodMHDVlct (hydro-mhd/EnzoMethodMHDVlct.cpp): the
// handle plays the role of the Method object (configuration + lazily
// allocated scratch that is reused for every block, cpp:236-246), vlct_compute
// is the two-stage loop of EnzoMethodMHDVlct::compute (cpp:459-496) and
// vlct_timestep is EnzoMethodMHDVlct::timestep (cpp:551-588).
//
// There is deliberately no CPU path in this file: without a CUDA device
// vlct_create fails with VLCT_ERR_NO_DEVICE.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include <cuda_runtime.h>

#include "vlct_kernels.cuh"

using namespace vlct;

struct vlct_handle {
  vlct_config cfg;
  Params P;
  int device = -1;
  cudaStream_t own_stream = nullptr;
  // HOST blocks: H2D / D2H streams of the staging pipeline, event pool
  cudaStream_t in_stream = nullptr, out_stream = nullptr;
  std::vector<cudaEvent_t> events;
  size_t events_used = 0;
  // options (vlct_set_option)
  long long host_pipeline_levels = -1;   // -1 auto, 0 off, n > 0: n z levels per pass
  long long device_pipeline_levels = 0;  // test hook: run DEVICE steps in passes too
  // measurement hook (scripts/gpu_power.py): bit k set = launch kernel family k
  // of a step (KernelId order). Anything but "all" leaves garbage in the fields.
  long long debug_kernel_mask = -1;
  int pair_kernels = default_pair_kernels();   // option "pair_kernels"
  Geom G{0, 0, 0};
  Scratch S;
  std::vector<void*> allocations;
  long long scratch_bytes = 0;
  unsigned long long* d_dt_bits = nullptr;
  double* d_step = nullptr;   // per-stage dt/dx, dt/dy, dt/dz, dt (k_step_params)
  unsigned long long* h_dt_bits = nullptr;   // pinned
  // device mirror of a HOST block (mem_space == VLCT_MEM_HOST)
  bool have_mirror = false;
  Geom mirror_G{0, 0, 0};     // shape the mirror was allocated for
  vlct_block mirror;
  // option "host_mirror_reuse": after vlct_compute of a HOST block the mirror
  // holds exactly what was copied back; a vlct_timestep of the same block that
  // follows immediately may use it instead of uploading the fields again
  long long host_mirror_reuse = 0;
  bool mirror_is_current = false;
  vlct_block mirror_of;          // the host block the mirror is a copy of
  std::vector<void*> mirror_allocs;
  // stacked device copy of a batch of blocks (vlct_compute_batch)
  // (two of them: HOST batches are double-buffered so that the copies of one
  // sub-batch overlap the kernels of another)
  vlct_block arena[2];
  int arena_capacity[2] = { 0, 0 };       // blocks each arena can hold
  Geom arena_G[2] = { Geom{0, 0, 0}, Geom{0, 0, 0} };   // block shape of each arena
  std::vector<void*> arena_allocs[2];
  long long host_batch_blocks = 0;        // option: HOST sub-batch size, 0 = auto
  // option: how HOST batches cross PCIe. 0 = one cudaMemcpyBatchAsync per
  // field set (copy engines), 1 = gather / scatter kernels on pinned memory
  // (zero-copy), 2 = one cudaMemcpyAsync per (block, field)
  long long host_batch_copy_mode = 0;
  bool batch_memcpy_works = true;         // cudaMemcpyBatchAsync (CUDA >= 12.8)
  double** d_ptr_table = nullptr;         // device: [field][block] user pointers
  double** h_ptr_table = nullptr;         // pinned staging of the same
  bool ptr_table_valid = false;           // device table == staging, for ptr_table_nb
  int ptr_table_nb = 0;
  bool ptr_table_usable = false;          // every pointer is device-accessible
  // host pointers known to be pinned / registered -> their device alias
  std::unordered_map<const void*, double*> mapped_host;
  std::unordered_map<void*, size_t> registered;   // vlct_host_register ranges
  size_t ptr_table_count = 0;
  size_t scratch_levels = 0;              // z levels the scratch arrays hold
  bool stepped = false;                   // a compute has filled the flux arrays
  double* d_face_stage = nullptr;         // device staging of vlct_save_face_fluxes
  size_t face_stage_count = 0;
  long long batch_max_blocks = 1024;
  long long launches = 0;
  long long copied_bytes[2] = { 0, 0 };   // H2D, D2H staged for HOST blocks
  Profiler prof;
  std::string last_error;
};

namespace {

int fail(vlct_handle* h, int code, const char* fmt, ...)
{
  char buf[512];
  va_list args;
  va_start(args, fmt);
  vsnprintf(buf, sizeof(buf), fmt, args);
  va_end(args);
  if (h) h->last_error = buf;
  return code;
}

#define CUDA_TRY(h, call)                                                    \
  do {                                                                       \
    cudaError_t err__ = (call);                                              \
    if (err__ != cudaSuccess)                                                \
      return fail((h), VLCT_ERR_CUDA, "%s failed: %s (%s:%d)", #call,        \
                  cudaGetErrorString(err__), __FILE__, __LINE__);            \
  } while (0)

/// Every entry point runs on the handle's device, whatever device the calling
/// thread has current (single-process multi-GPU hosts), and restores the
/// caller's device on return.
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(int device)
  {
    if (device < 0) return;
    if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
    if (prev != device) { cudaSetDevice(device); switched = true; }
  }
  ~DeviceGuard() { if (switched && prev >= 0) cudaSetDevice(prev); }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

enum Pool { POOL_SCRATCH = 1, POOL_MIRROR = 0, POOL_ARENA = 2 /* + arena index */ };

int dev_alloc(vlct_handle* h, double** out, size_t count, int pool = POOL_SCRATCH)
{
  void* p = nullptr;
  CUDA_TRY(h, cudaMalloc(&p, count * sizeof(double)));
  CUDA_TRY(h, cudaMemset(p, 0, count * sizeof(double)));
  (pool == POOL_SCRATCH ? h->allocations
   : pool == POOL_MIRROR ? h->mirror_allocs
   : h->arena_allocs[pool - POOL_ARENA]).push_back(p);
  if (pool == POOL_SCRATCH) h->scratch_bytes += (long long) (count * sizeof(double));
  *out = (double*) p;
  return VLCT_OK;
}

/// elements of a (possibly stacked, Geom::levels) cell- / face-centred array
size_t cell_count(const Geom& G)
{ return (size_t) G.mx * (size_t) G.my * G.levels(); }

size_t face_count(const Geom& G, int d)
{
  return (size_t) (G.mx + (d == 0)) * (size_t) (G.my + (d == 1)) *
         (G.levels() + (d == 2));
}

/// EnzoVlctScratchSpace (hydro-mhd/EnzoMethodMHDVlct.hpp:217-312) +
/// EnzoBfieldMethodCT scratch (toolkit/EnzoBfieldMethodCT.cpp:41-76), minus
/// everything the fused kernels never materialise.
int alloc_scratch(vlct_handle* h, const Geom& G)
{
  h->G = G;
  h->scratch_levels = G.levels();
  const size_t n = cell_count(G);
  const Params& P = h->P;
  Scratch& S = h->S;
  memset(&S, 0, sizeof(S));
  int rc;
#define ALLOC(ptr, count) if ((rc = dev_alloc(h, &(ptr), (count))) != VLCT_OK) return rc
  const bool two_stage = (h->cfg.time_scheme == VLCT_TIME_VL);
  if (two_stage) {
    ALLOC(S.temp.rho, n); ALLOC(S.temp.vx, n); ALLOC(S.temp.vy, n);
    ALLOC(S.temp.vz, n); ALLOC(S.temp.etot, n);
    if (P.de) ALLOC(S.temp.eint, n);
    if (P.mhd) {
      ALLOC(S.temp.bx, n); ALLOC(S.temp.by, n); ALLOC(S.temp.bz, n);
      for (int d = 0; d < 3; d++) ALLOC(S.tbi.bi[d], face_count(G, d));
    }
    for (int s = 0; s < P.nsc; s++) ALLOC(S.temp.sc[s], n);
  }
  for (int d = 0; d < 3; d++) {
    FluxSet& F = S.flux[d];
    ALLOC(F.rho, n); ALLOC(F.mx_, n); ALLOC(F.my_, n); ALLOC(F.mz_, n);
    ALLOC(F.e, n);
    if (P.mhd) {
      if (d != 0) ALLOC(F.bx, n);
      if (d != 1) ALLOC(F.by, n);
      if (d != 2) ALLOC(F.bz, n);
    }
    if (P.de) { ALLOC(F.eint, n); ALLOC(F.vbar, n); }
    for (int s = 0; s < P.nsc_flux; s++) ALLOC(F.sc[s], n);
  }
  for (int stage = 0; stage < (two_stage ? 2 : 1); stage++)
    for (int s = 0; s < P.nsc; s++) ALLOC(S.prim_sc[stage][s], n);
  if (P.mhd) for (int d = 0; d < 3; d++) ALLOC(S.edge[d], n);
#undef ALLOC
  // dev_alloc's memsets run on the legacy default stream; the kernels run on a
  // non-blocking stream that does not order against it
  CUDA_TRY(h, cudaStreamSynchronize(cudaStreamLegacy));
  return VLCT_OK;
}

/// The reference sizes its scratch from the first block and reuses it for
/// every later one (EnzoMethodMHDVlct.cpp:236-246): one shape per handle. A
/// batch of stacked blocks needs more z levels of the same shape; the scratch
/// then grows (it never shrinks).
int ensure_scratch(vlct_handle* h, const Geom& G)
{
  if (h->G.mx != 0 && (G.mx != h->G.mx || G.my != h->G.my || G.mz != h->G.mz))
    return fail(h, VLCT_ERR_INVALID_BLOCK,
                "all blocks handled by one handle must share one shape "
                "(first block was %dx%dx%d incl. ghosts, got %dx%dx%d)",
                h->G.mx, h->G.my, h->G.mz, G.mx, G.my, G.mz);
  if (h->G.mx != 0 && G.levels() <= h->scratch_levels) return VLCT_OK;
  if (h->G.mx != 0) {
    CUDA_TRY(h, cudaDeviceSynchronize());
    for (void* p : h->allocations) cudaFree(p);
    h->allocations.clear();
    h->scratch_bytes = 0;
  }
  return alloc_scratch(h, G);
}

int check_block(vlct_handle* h, const vlct_block* b, bool for_timestep)
{
  if (b == nullptr) return fail(h, VLCT_ERR_INVALID_BLOCK, "NULL block");
  if (b->nx <= 0 || b->ny <= 0 || b->nz <= 0)
    return fail(h, VLCT_ERR_INVALID_BLOCK,
                "block must be three-dimensional (nx,ny,nz > 0); this "
                "implementation supports rank 3 only");
  if (b->gx < 0 || b->gy < 0 || b->gz < 0)
    return fail(h, VLCT_ERR_INVALID_BLOCK, "negative ghost depth");
  {
    // the kernels count rows, columns and warps in 32 bits
    const double cells = (double) (b->nx + 2 * b->gx + 1) * (b->ny + 2 * b->gy + 1) *
                         (b->nz + 2 * b->gz + 1);
    if (cells >= 2147483648.0)
      return fail(h, VLCT_ERR_INVALID_BLOCK,
                  "block too large: %g cells incl. ghosts (limit 2^31)", cells);
    // one shape per handle (EnzoMethodMHDVlct.cpp:236-246), also for the entry
    // points that allocate no scratch themselves
    const int mx = b->nx + 2 * b->gx, my = b->ny + 2 * b->gy, mz = b->nz + 2 * b->gz;
    if (h->G.mx != 0 && (mx != h->G.mx || my != h->G.my || mz != h->G.mz))
      return fail(h, VLCT_ERR_INVALID_BLOCK,
                  "all blocks handled by one handle must share one shape "
                  "(first block was %dx%dx%d incl. ghosts, got %dx%dx%d)",
                  h->G.mx, h->G.my, h->G.mz, mx, my, mz);
  }
  if (!(b->dx > 0 && b->dy > 0 && b->dz > 0))
    return fail(h, VLCT_ERR_INVALID_BLOCK, "cell widths must be positive");
  const Params& P = h->P;
#define NEED(field)                                                          \
  if (b->field == nullptr)                                                   \
    return fail(h, VLCT_ERR_INVALID_BLOCK,                                   \
                "\"%s\" must be a permanent field", #field)
  NEED(density); NEED(velocity_x); NEED(velocity_y); NEED(velocity_z);
  NEED(total_energy);
  if (P.de) NEED(internal_energy);
  if (P.mhd) {
    NEED(bfield_x); NEED(bfield_y); NEED(bfield_z);
    if (!for_timestep) { NEED(bfieldi_x); NEED(bfieldi_y); NEED(bfieldi_z); }
  }
  if (for_timestep) NEED(pressure);
#undef NEED
  for (int s = 0; s < P.nsc; s++)
    if (b->passive[s] == nullptr)
      return fail(h, VLCT_ERR_INVALID_BLOCK, "passive scalar %d is NULL", s);
  if (!for_timestep) {
    // EnzoMethodMHDVlct.cpp:124-133: ghost depth >= sum of the stages' staling
    const int full = (h->cfg.reconstruct_method == VLCT_RECON_NN) ? 1 : 2;
    const int need = (h->cfg.time_scheme == VLCT_TIME_VL) ? 1 + full : full;
    const int gmin = b->gx < b->gy ? (b->gx < b->gz ? b->gx : b->gz)
                                   : (b->gy < b->gz ? b->gy : b->gz);
    if (gmin < need)
      return fail(h, VLCT_ERR_INVALID_BLOCK, "ghost depth must be at least %d.",
                  need);
    if (h->cfg.has_acceleration &&
        (b->acceleration_x == nullptr || b->acceleration_y == nullptr ||
         b->acceleration_z == nullptr))
      return fail(h, VLCT_ERR_INVALID_BLOCK,
                  "acceleration fields are configured but missing");
  }
  if (b->mem_space != VLCT_MEM_HOST && b->mem_space != VLCT_MEM_DEVICE)
    return fail(h, VLCT_ERR_INVALID_BLOCK, "unknown mem_space %d", b->mem_space);
  return VLCT_OK;
}

Geom geom_of(const vlct_block* b)
{
  Geom G;
  G.mx = b->nx + 2 * b->gx; G.my = b->ny + 2 * b->gy; G.mz = b->nz + 2 * b->gz;
  return G;
}

State state_of(const vlct_handle* h, const vlct_block* b)
{
  State u;
  memset(&u, 0, sizeof(u));
  u.rho = b->density; u.vx = b->velocity_x; u.vy = b->velocity_y;
  u.vz = b->velocity_z; u.etot = b->total_energy; u.eint = b->internal_energy;
  u.bx = b->bfield_x; u.by = b->bfield_y; u.bz = b->bfield_z;
  for (int s = 0; s < h->P.nsc; s++) u.sc[s] = b->passive[s];
  return u;
}

// ---- HOST mem_space: a device mirror of the block ---------------------------
struct FieldRef { double* vlct_block::*member; int face; };

const FieldRef kFields[] = {
  { &vlct_block::density, -1 }, { &vlct_block::velocity_x, -1 },
  { &vlct_block::velocity_y, -1 }, { &vlct_block::velocity_z, -1 },
  { &vlct_block::total_energy, -1 }, { &vlct_block::internal_energy, -1 },
  { &vlct_block::bfield_x, -1 }, { &vlct_block::bfield_y, -1 },
  { &vlct_block::bfield_z, -1 }, { &vlct_block::bfieldi_x, 0 },
  { &vlct_block::bfieldi_y, 1 }, { &vlct_block::bfieldi_z, 2 },
  { &vlct_block::pressure, -1 }, { &vlct_block::acceleration_x, -1 },
  { &vlct_block::acceleration_y, -1 }, { &vlct_block::acceleration_z, -1 },
};
constexpr int kNumFields = (int) (sizeof(kFields) / sizeof(kFields[0]));

size_t field_count(const Geom& G, int face)
{ return face < 0 ? cell_count(G) : face_count(G, face); }

/// The device mirror of HOST blocks: allocated from the first HOST block, one
/// shape per handle like the scratch. Entry points need different field sets
/// (vlct_timestep may omit the face fields, vlct_compute may omit pressure):
/// a field that a later block brings along and the mirror does not have yet is
/// added then, so no kernel ever sees a NULL mirror pointer for a field its
/// block supplies.
int ensure_mirror(vlct_handle* h, const vlct_block* b, const Geom& G)
{
  if (h->have_mirror &&
      (G.mx != h->mirror_G.mx || G.my != h->mirror_G.my || G.mz != h->mirror_G.mz))
    return fail(h, VLCT_ERR_INVALID_BLOCK,
                "all HOST blocks handled by one handle must share one shape "
                "(the device mirror holds %dx%dx%d incl. ghosts, got %dx%dx%d)",
                h->mirror_G.mx, h->mirror_G.my, h->mirror_G.mz, G.mx, G.my, G.mz);
  const bool fresh = !h->have_mirror;
  if (fresh) {
    h->mirror = *b;
    for (int f = 0; f < kNumFields; f++) h->mirror.*(kFields[f].member) = nullptr;
    for (int s = 0; s < VLCT_MAX_PASSIVE; s++) h->mirror.passive[s] = nullptr;
  }
  // geometry and cell widths follow the current block
  h->mirror.nx = b->nx; h->mirror.ny = b->ny; h->mirror.nz = b->nz;
  h->mirror.gx = b->gx; h->mirror.gy = b->gy; h->mirror.gz = b->gz;
  h->mirror.dx = b->dx; h->mirror.dy = b->dy; h->mirror.dz = b->dz;
  h->mirror.mem_space = VLCT_MEM_DEVICE;
  h->mirror.stream = nullptr;
  int rc;
  bool added = false;
  for (int f = 0; f < kNumFields; f++) {
    if (b->*(kFields[f].member) == nullptr) continue;
    if (h->mirror.*(kFields[f].member) != nullptr) continue;
    double* p;
    if ((rc = dev_alloc(h, &p, field_count(G, kFields[f].face), POOL_MIRROR)) != VLCT_OK)
      return rc;
    h->mirror.*(kFields[f].member) = p;
    added = true;
  }
  for (int s = 0; s < h->P.nsc; s++) {
    if (h->mirror.passive[s] != nullptr) continue;
    double* p;
    if ((rc = dev_alloc(h, &p, G.cells(), POOL_MIRROR)) != VLCT_OK) return rc;
    h->mirror.passive[s] = p;
    added = true;
  }
  // the zero-fill above must not overtake the H2D copies on the work stream
  if (added) CUDA_TRY(h, cudaStreamSynchronize(cudaStreamLegacy));
  if (added) h->mirror_is_current = false;
  h->mirror_G = G;
  h->have_mirror = true;
  return VLCT_OK;
}

/// which fields one entry point reads (H2D) and writes (D2H)
enum CopySet { COPY_COMPUTE_IN, COPY_COMPUTE_OUT, COPY_TIMESTEP_IN,
               COPY_FUSED_OUT /* compute's outputs + "pressure" */ };

bool in_copy_set(const vlct_handle* h, double* vlct_block::*m, CopySet set)
{
  const bool face = (m == &vlct_block::bfieldi_x || m == &vlct_block::bfieldi_y ||
                     m == &vlct_block::bfieldi_z);
  const bool accel = (m == &vlct_block::acceleration_x ||
                      m == &vlct_block::acceleration_y ||
                      m == &vlct_block::acceleration_z);
  if (m == &vlct_block::pressure) return set == COPY_FUSED_OUT;   // timestep's output
  switch (set) {
  case COPY_COMPUTE_IN:  return accel ? (h->cfg.has_acceleration != 0) : true;
  case COPY_FUSED_OUT:
  case COPY_COMPUTE_OUT: return !accel;            // compute never writes them
  case COPY_TIMESTEP_IN: return !face && !accel;   // cell-centred state only
  }
  return true;
}

/// copy the z levels [z0, z1) of a set of fields between the host block and
/// its device mirror (z is the slowest axis: a range of levels is contiguous).
/// z1 >= mz means "to the end", which for bfieldi_z includes its extra level.
int mirror_copy(vlct_handle* h, const vlct_block* host, const Geom& G,
                cudaStream_t st, bool to_device, CopySet set,
                int z0 = 0, int z1 = 1 << 30)
{
  if (z0 < 0) z0 = 0;
  auto copy = [&](double* hp, double* dp, int face) -> int {
    const size_t plane = (size_t) (G.mx + (face == 0)) * (size_t) (G.my + (face == 1));
    const int levels = G.mz + (face == 2);
    const int hi = (z1 >= G.mz) ? levels : z1;
    if (hi <= z0) return VLCT_OK;
    const size_t off = plane * (size_t) z0;
    const size_t bytes = plane * (size_t) (hi - z0) * sizeof(double);
    CUDA_TRY(h, cudaMemcpyAsync(to_device ? (void*) (dp + off) : (void*) (hp + off),
                                to_device ? (void*) (hp + off) : (void*) (dp + off),
                                bytes, to_device ? cudaMemcpyHostToDevice
                                                 : cudaMemcpyDeviceToHost, st));
    h->copied_bytes[to_device ? 0 : 1] += (long long) bytes;
    return VLCT_OK;
  };
  int rc;
  for (int f = 0; f < kNumFields; f++) {
    double* hp = host->*(kFields[f].member);
    double* dp = h->mirror.*(kFields[f].member);
    if (hp == nullptr || dp == nullptr) continue;
    if (!in_copy_set(h, kFields[f].member, set)) continue;
    if ((rc = copy(hp, dp, kFields[f].face)) != VLCT_OK) return rc;
  }
  for (int s = 0; s < h->P.nsc; s++)
    if ((rc = copy(host->passive[s], h->mirror.passive[s], -1)) != VLCT_OK) return rc;
  return VLCT_OK;
}

/// an event from the handle's pool (no timing), recorded on st
int record_event(vlct_handle* h, cudaStream_t st, cudaEvent_t* out)
{
  if (h->events_used == h->events.size()) {
    cudaEvent_t e;
    CUDA_TRY(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    h->events.push_back(e);
  }
  *out = h->events[h->events_used++];
  CUDA_TRY(h, cudaEventRecord(*out, st));
  return VLCT_OK;
}

int total_staling(int recon) { return recon == VLCT_RECON_NN ? 1 : 2; }
int immediate_staling(int recon) { return recon == VLCT_RECON_NN ? 0 : 1; }

// ---- z-skewed partial execution of a step -------------------------------------
// A step may be executed in several passes, each restricted to a range of z
// levels. Every kernel of the step is a pure function of its inputs at a given
// index, so the passes give bit-identical results as long as (i) each kernel's
// index box is tiled exactly once by the passes and (ii) a kernel only runs
// where everything it reads has already been produced. (ii) is what the lags
// below encode. With k the z index of a kernel (cells; faces for the z sweep
// and for the z component of the face-B update), the reads are
//   SCAL(k)    <- stage input k
//   FLUX_XY(k) <- stage input k, SCAL k, face B k
//   FLUX_Z(f)  <- stage input / SCAL f-1..f+2 (PLM; f..f+1 with NN), z-face f+1
//   EDGE(k)    <- FLUX_XY k..k+1, FLUX_Z k, stage input k..k+1
//   FACE(k)    <- EDGE k-1..k
//   UPDATE(k)  <- FLUX_XY k, FLUX_Z k-1..k, FACE k..k+1, SCAL k-2..k+2 (passive
//                 scalars without flux arrays: the update reconstructs them
//                 itself; each stage has its own SCAL arrays)
// and the second stage's input is the first stage's UPDATE / FACE output.
//   an END cut at z   : the pass covers indices below  z + kEndLag[stage][kernel]
//   a START cut at z  : the pass covers indices from   z - kStartLag[stage][kernel]
// Upstream kernels have the larger lag in both tables, so a pass never needs
// anything outside itself and the passes before it. The stage-1 rows carry
// one level of slack beyond the data dependence: flux / edge scratch arrays
// are shared by the two stages, and the slack keeps a later pass's stage-1
// reads clear of what the earlier pass's stage 2 has overwritten. The fields
// themselves are updated in place by the last stage only below / above
// everything the first stage of a later pass still reads.
enum KernelId { K_SCAL = 0, K_FLUX_XY, K_FLUX_Z, K_EDGE, K_FACE, K_UPDATE, K_COUNT };
enum { CUT_NONE = 0, CUT_END = 1, CUT_START = 2 };
struct ZCut { int kind; int z; };

//                                   SCAL XY  Z  EDGE FACE UPDATE
const int kEndLag[2][K_COUNT]   = { { 6,  6,  5,  5,  5,  4 },     // first of two stages
                                    { 3,  2,  1,  1,  1,  0 } };   // last stage
const int kStartLag[2][K_COUNT] = { { 5,  4,  4,  4,  3,  3 },
                                    { 2,  1,  1,  1,  0,  0 } };
/// input levels a pass needs beyond an END cut / before a START cut
constexpr int kEndReach = 6, kStartReach = 5;

int cut_index(const ZCut& c, int row, KernelId id)
{ return (c.kind == CUT_END) ? c.z + kEndLag[row][id] : c.z - kStartLag[row][id]; }

/// the clip of one kernel of one stage for the pass (lo, hi]
ZClip pass_clip(const ZCut& lo, const ZCut& hi, int row, KernelId id)
{
  ZClip z = kNoClip;
  if (lo.kind != CUT_NONE) z.lo = cut_index(lo, row, id);
  if (hi.kind != CUT_NONE) z.hi = cut_index(hi, row, id);
  return z;
}

/// the stage loop of EnzoMethodMHDVlct::compute on device pointers, restricted
/// to the pass between two cuts (CUT_NONE on both sides = the whole block)
/// fold_cfl: the last stage's update also evaluates the next cycle's
/// timestep() (launch_update / CflFold): "pressure" is written and the CFL
/// minimum of the levels this pass finishes is folded into h->d_dt_bits (reset
/// by the pass that sets the step parameters, i.e. the first one of a step).
int compute_on_device(vlct_handle* h, const vlct_block* b, const Geom& G,
                      double dt, const double* dt_dev, cudaStream_t st,
                      ZCut zlo = ZCut{ CUT_NONE, 0 }, ZCut zhi = ZCut{ CUT_NONE, 0 },
                      bool set_step_params = true, bool fold_cfl = false)
{
  const Params& P = h->P;
  const State ext = state_of(h, b);
  FaceB bi;
  bi.bi[0] = b->bfieldi_x; bi.bi[1] = b->bfieldi_y; bi.bi[2] = b->bfieldi_z;
  const double width[3] = { b->dx, b->dy, b->dz };
  const double* accel[3] = { b->acceleration_x, b->acceleration_y,
                             b->acceleration_z };
  const int nstages = (h->cfg.time_scheme == VLCT_TIME_EULER) ? 1 : 2;
  // dt may live on the device (vlct_compute_dev): the stage constants are
  // formed there, so a step never has to wait for the host
  if (set_step_params)
    launch_step_params(LaunchCtx{ st, &h->launches, &h->prof, h->pair_kernels }, dt_dev, dt, nstages,
                       width, h->d_step);
  if (fold_cfl && G.nrep != 1)
    return fail(h, VLCT_ERR_INTERNAL, "the CFL fold handles single blocks only");
  if (fold_cfl && set_step_params)
    launch_timestep_reset(LaunchCtx{ st, &h->launches, &h->prof, h->pair_kernels }, h->d_dt_bits);
  CflFold cfl;
  cfl.pressure = b->pressure;
  cfl.dt_bits = h->d_dt_bits;
  cfl.width[0] = b->dx; cfl.width[1] = b->dy; cfl.width[2] = b->dz;
  int stale = 0;
  for (int stage = 0; stage < nstages; stage++) {
    const bool final_stage = (stage + 1) == nstages;
    const double* step_params = h->d_step + 4 * stage;
    const int recon = (nstages == 2 && stage == 0) ? VLCT_RECON_NN
                                                   : h->cfg.reconstruct_method;
    const State& cur = (stage == 0) ? ext : h->S.temp;
    const State& out = final_stage ? ext : h->S.temp;
    const FaceB& bi_cur = (stage == 0) ? bi : h->S.tbi;
    const FaceB& bi_out = (stage == 1 || nstages == 1) ? bi : h->S.tbi;

    const LaunchCtx ctx{ st, &h->launches, &h->prof, h->pair_kernels };
    const int row = final_stage ? 1 : 0;
    const long long mask = h->debug_kernel_mask;
    const ZClip nothing{ 0, 0 };
    if (mask & (1 << K_SCAL))
      launch_primitives(ctx, P, G, cur, h->S, stage, stale,
                        pass_clip(zlo, zhi, row, K_SCAL));
    const int cs = stale + immediate_staling(recon);
    for (int dim = 0; dim < 3; dim++)
      if (mask & (1 << (dim == 2 ? K_FLUX_Z : K_FLUX_XY)))
        launch_flux(ctx, P, G, dim, recon, cur, h->S, stage, bi_cur, cs,
                    pass_clip(zlo, zhi, row, dim == 2 ? K_FLUX_Z : K_FLUX_XY));
    if (P.mhd)
      launch_ct(ctx, P, G, cur, h->S, bi, bi_out, step_params, cs,
                (mask & (1 << K_EDGE)) ? pass_clip(zlo, zhi, row, K_EDGE) : nothing,
                (mask & (1 << K_FACE)) ? pass_clip(zlo, zhi, row, K_FACE) : nothing);
    // gravity: full step only, i.e. stage index 1
    // (EnzoMHDIntegratorStageCommands.cpp:181,279)
    const bool gravity = (stage == 1) && h->cfg.has_acceleration &&
                         accel[0] != nullptr;
    if (mask & (1 << K_UPDATE))
      launch_update(ctx, P, G, ext, cur, out, h->S, bi_out, accel, gravity,
                    step_params, stage, recon, cs, pass_clip(zlo, zhi, row, K_UPDATE),
                    (fold_cfl && final_stage) ? &cfl : nullptr);
    stale += total_staling(recon);
  }
  CUDA_TRY(h, cudaGetLastError());
  h->stepped = (G.nrep == 1);   // face fluxes of a single block can be read back
  return VLCT_OK;
}

}  // namespace

extern "C" {

int vlct_create(const vlct_config* cfg, vlct_handle** out)
{
  if (out == nullptr) return VLCT_ERR_INVALID_CONFIG;
  *out = nullptr;
  vlct_handle* h = new vlct_handle;
  *out = h;   // returned even on failure so that vlct_last_error() works
  char err[512] = "";
  int rc = vlct_config_validate(cfg, err, (int) sizeof(err));
  if (rc != VLCT_OK) { h->last_error = err; return rc; }
  h->cfg = *cfg;
  if (h->cfg.courant < 0)   // EnzoMethodMHDVlct.cpp:100-101
    h->cfg.courant = (cfg->time_scheme == VLCT_TIME_VL) ? 0.3 : 1.0;

  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(h, VLCT_ERR_NO_DEVICE,
                "no CUDA device is visible; this library has no CPU fallback");
  }
  CUDA_TRY(h, cudaGetDevice(&h->device));
  cudaDeviceProp prop;
  CUDA_TRY(h, cudaGetDeviceProperties(&prop, h->device));
  if (prop.major < 10)
    return fail(h, VLCT_ERR_NO_DEVICE,
                "device %d (%s, sm_%d%d) is not a Blackwell-class GPU; the "
                "kernels are built for sm_100a only", h->device, prop.name,
                prop.major, prop.minor);

  Params& P = h->P;
  P.gamma = cfg->gamma;
  P.theta = cfg->theta_limiter;
  P.density_floor = cfg->density_floor;
  P.pressure_floor = cfg->pressure_floor;
  P.mhd = (cfg->mhd_choice == VLCT_MHD_CONSTRAINED_TRANSPORT);
  P.de = (cfg->dual_energy == VLCT_DE_MODERN);
  P.de_eta = P.de ? cfg->dual_energy_eta : 0.0;
  // `float ggm1 = gamma*(gamma-1.)` -- a float in the fp64 path
  // (fluid-props/EnzoPhysicsFluidProps.cpp:234)
  const float ggm1 = (float) (cfg->gamma * (cfg->gamma - 1.));
  P.ggm1 = (double) ggm1;
  {
    // through volatiles so the host compiler cannot fold or reassociate it
    volatile double gm1 = cfg->gamma - 1.0;
    volatile double one = 1.0;
    P.igm1 = one / gm1;
  }
  P.nsc = cfg->n_passive;
  P.nsc_flux = 0;          // option "scalar_flux_arrays"
  P.riemann = cfg->riemann_solver;
  P.recon = cfg->reconstruct_method;

  CUDA_TRY(h, cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
  CUDA_TRY(h, cudaStreamCreateWithFlags(&h->in_stream, cudaStreamNonBlocking));
  CUDA_TRY(h, cudaStreamCreateWithFlags(&h->out_stream, cudaStreamNonBlocking));
  CUDA_TRY(h, cudaMalloc((void**) &h->d_dt_bits, sizeof(unsigned long long)));
  CUDA_TRY(h, cudaMalloc((void**) &h->d_step, 8 * sizeof(double)));
  CUDA_TRY(h, cudaMallocHost((void**) &h->h_dt_bits, sizeof(unsigned long long)));
  return VLCT_OK;
}

void vlct_destroy(vlct_handle* h)
{
  if (h == nullptr) return;
  if (h->device >= 0) {
    DeviceGuard device_guard__(h->device);
    if (h->own_stream) cudaStreamSynchronize(h->own_stream);
    for (void* p : h->allocations) cudaFree(p);
    for (void* p : h->mirror_allocs) cudaFree(p);
    for (int a = 0; a < 2; a++)
      for (void* p : h->arena_allocs[a]) cudaFree(p);
    if (h->d_face_stage) cudaFree(h->d_face_stage);
    if (h->d_ptr_table) cudaFree(h->d_ptr_table);
    if (h->h_ptr_table) cudaFreeHost(h->h_ptr_table);
    if (h->d_dt_bits) cudaFree(h->d_dt_bits);
    if (h->d_step) cudaFree(h->d_step);
    if (h->h_dt_bits) cudaFreeHost(h->h_dt_bits);
    for (auto& r : h->registered) cudaHostUnregister(r.first);
    for (cudaEvent_t e : h->events) cudaEventDestroy(e);
    if (h->in_stream) cudaStreamDestroy(h->in_stream);
    if (h->out_stream) cudaStreamDestroy(h->out_stream);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
  }
  delete h;
}

}  // extern "C"

namespace {

/// levels per pass of the HOST staging pipeline (0 = one shot)
int host_levels(const vlct_handle* h, const Geom& G)
{
  if (h->cfg.time_scheme == VLCT_TIME_EULER) return 0;  // single-stage: one pass
  if (h->host_pipeline_levels == 0) return 0;
  if (h->host_pipeline_levels > 0) return (int) h->host_pipeline_levels;
  // auto: ~32 passes, and only when a field is large enough for the copies to
  // matter (>= 32 MB); small blocks go through in one shot
  if (G.cells() * sizeof(double) < ((size_t) 32 << 20)) return 0;
  const int lv = (G.mz + 31) / 32;
  return lv < 4 ? 4 : lv;
}

/// A whole step as consecutive z passes separated by END cuts, `levels` input
/// levels per pass. With a HOST block each pass's input levels are copied to
/// the device mirror on in_stream while the previous pass computes, and the
/// levels a pass has finished are copied back on out_stream while the next
/// one computes, so H2D, kernels and D2H overlap (PCIe is full duplex).
int timestep_launch(vlct_handle* h, const vlct_block* db, const Geom& G,
                    cudaStream_t st, ZClip zc = kNoClip, bool reset = true);

/// fused_timestep: the CFL kernel of the NEXT cycle follows the update on every
/// level a pass has finished (vlct_compute_and_timestep), before the level
/// goes back to the host.
int compute_in_passes(vlct_handle* h, const vlct_block* host, const vlct_block* dev,
                      const Geom& G, double dt, const double* dt_dev,
                      cudaStream_t st, int levels, bool fused_timestep = false)
{
  const bool staged = (host != nullptr);
  int rc;
  h->events_used = 0;
  ZCut prev{ CUT_NONE, 0 };
  int uploaded = 0, downloaded = 0;
  bool first = true;
  while (uploaded < G.mz) {
    const int up_to = (uploaded + levels < G.mz) ? uploaded + levels : G.mz;
    if (staged) {
      if ((rc = mirror_copy(h, host, G, h->in_stream, true, COPY_COMPUTE_IN,
                            uploaded, up_to)) != VLCT_OK) return rc;
      cudaEvent_t ev;
      if ((rc = record_event(h, h->in_stream, &ev)) != VLCT_OK) return rc;
      CUDA_TRY(h, cudaStreamWaitEvent(st, ev, 0));
    }
    uploaded = up_to;
    ZCut cut{ CUT_NONE, 0 };
    if (uploaded < G.mz) {
      cut = ZCut{ CUT_END, uploaded - kEndReach };
      // nothing new below the cut yet: keep uploading
      if (cut.z <= (prev.kind == CUT_END ? prev.z : 0)) continue;
    }
    if ((rc = compute_on_device(h, dev, G, dt, dt_dev, st, prev, cut, first,
                                fused_timestep)) != VLCT_OK)
      return rc;
    const int done = (cut.kind == CUT_END) ? cut.z : (1 << 30);
    first = false;
    if (staged) {
      cudaEvent_t ev;
      if ((rc = record_event(h, st, &ev)) != VLCT_OK) return rc;
      CUDA_TRY(h, cudaStreamWaitEvent(h->out_stream, ev, 0));
      if ((rc = mirror_copy(h, host, G, h->out_stream, false,
                            fused_timestep ? COPY_FUSED_OUT : COPY_COMPUTE_OUT,
                            downloaded, done)) != VLCT_OK) return rc;
    }
    downloaded = done;
    prev = cut;
  }
  if (staged) {
    CUDA_TRY(h, cudaStreamSynchronize(h->out_stream));
    CUDA_TRY(h, cudaStreamSynchronize(st));
  }
  return VLCT_OK;
}

/// dt_next != nullptr: vlct_compute_and_timestep
int compute_entry(vlct_handle* h, const vlct_block* b, double dt, const double* dt_dev,
                  double* dt_next = nullptr, double* dt_next_dev = nullptr)
{
  if (h == nullptr) return VLCT_ERR_INVALID_CONFIG;
  DeviceGuard device_guard__(h->device);
  if (h->device < 0) return fail(h, VLCT_ERR_NO_DEVICE, "handle has no device");
  int rc = check_block(h, b, false);
  if (rc != VLCT_OK) return rc;
  const bool fused = (dt_next != nullptr || dt_next_dev != nullptr);
  if (fused && b->pressure == nullptr)
    return fail(h, VLCT_ERR_INVALID_BLOCK, "\"pressure\" must be a permanent field");
  if (dt_next_dev != nullptr && b->mem_space != VLCT_MEM_DEVICE)
    return fail(h, VLCT_ERR_INVALID_BLOCK,
                "vlct_compute_and_timestep_dev needs a block in device memory");
  const Geom G = geom_of(b);
  if ((rc = ensure_scratch(h, G)) != VLCT_OK) return rc;
  // the fused call ends like vlct_timestep: the minimum comes back to the host
  auto finish_dt = [&](cudaStream_t st) -> int {
    CUDA_TRY(h, cudaMemcpyAsync(h->h_dt_bits, h->d_dt_bits, sizeof(unsigned long long),
                                cudaMemcpyDeviceToHost, st));
    CUDA_TRY(h, cudaStreamSynchronize(st));
    double dt_min;
    memcpy(&dt_min, h->h_dt_bits, sizeof(double));
    *dt_next = dt_min * h->cfg.courant;   // cpp:585-587
    return VLCT_OK;
  };
  if (b->mem_space == VLCT_MEM_DEVICE) {
    cudaStream_t st = b->stream ? (cudaStream_t) b->stream : h->own_stream;
    if (h->device_pipeline_levels > 0 && h->cfg.time_scheme != VLCT_TIME_EULER)
      rc = compute_in_passes(h, nullptr, b, G, dt, dt_dev, st,
                             (int) h->device_pipeline_levels, fused);
    else
      rc = compute_on_device(h, b, G, dt, dt_dev, st, ZCut{ CUT_NONE, 0 },
                             ZCut{ CUT_NONE, 0 }, true, fused);
    if (rc == VLCT_OK && dt_next_dev != nullptr) {
      // courant * minimum, left on the device: nothing waits for the host
      launch_finish_dt(LaunchCtx{ st, &h->launches, &h->prof, h->pair_kernels }, h->d_dt_bits,
                       h->cfg.courant, dt_next_dev);
      CUDA_TRY(h, cudaGetLastError());
    } else if (rc == VLCT_OK && fused) {
      rc = finish_dt(st);
    }
    return rc;
  }
  if (dt_dev != nullptr)
    return fail(h, VLCT_ERR_INVALID_BLOCK,
                "vlct_compute_dev needs a block in device memory");
  // HOST: stage through the device mirror; synchronous
  cudaStream_t st = h->own_stream;
  if ((rc = ensure_mirror(h, b, G)) != VLCT_OK) return rc;
  h->mirror_is_current = false;
  if (const int levels = host_levels(h, G)) {
    rc = compute_in_passes(h, b, &h->mirror, G, dt, nullptr, st, levels, fused);
  } else {
    if ((rc = mirror_copy(h, b, G, st, true, COPY_COMPUTE_IN)) != VLCT_OK) return rc;
    if ((rc = compute_on_device(h, &h->mirror, G, dt, nullptr, st, ZCut{ CUT_NONE, 0 },
                                ZCut{ CUT_NONE, 0 }, true, fused)) != VLCT_OK) return rc;
    if ((rc = mirror_copy(h, b, G, st, false,
                          fused ? COPY_FUSED_OUT : COPY_COMPUTE_OUT)) != VLCT_OK) return rc;
    CUDA_TRY(h, cudaStreamSynchronize(st));
  }
  if (rc == VLCT_OK && fused) rc = finish_dt(st);
  if (rc == VLCT_OK) {
    // every field timestep() reads was uploaded and/or written by this call
    h->mirror_is_current = true;
    h->mirror_of = *b;
  }
  return rc;
}

/// may vlct_timestep(b) read the device mirror instead of uploading b's fields?
bool mirror_serves(const vlct_handle* h, const vlct_block* b)
{
  if (!h->host_mirror_reuse || !h->mirror_is_current) return false;
  const vlct_block& m = h->mirror_of;
  if (b->nx != m.nx || b->ny != m.ny || b->nz != m.nz || b->gx != m.gx ||
      b->gy != m.gy || b->gz != m.gz) return false;
  for (int f = 0; f < kNumFields; f++)
    if (b->*(kFields[f].member) != m.*(kFields[f].member)) return false;
  for (int s = 0; s < h->P.nsc; s++)
    if (b->passive[s] != m.passive[s]) return false;
  return true;
}

/// launches DE sync + pressure + CFL minimum on the block's stream
int timestep_launch(vlct_handle* h, const vlct_block* db, const Geom& G,
                    cudaStream_t st, ZClip zc, bool reset)
{
  const double width[3] = { db->dx, db->dy, db->dz };
  const State u = state_of(h, db);
  const LaunchCtx ctx{ st, &h->launches, &h->prof, h->pair_kernels };
  if (reset) launch_timestep_reset(ctx, h->d_dt_bits);
  launch_timestep(ctx, h->P, G, u, db->pressure, width, h->d_dt_bits, zc);
  CUDA_TRY(h, cudaGetLastError());
  return VLCT_OK;
}

/// vlct_timestep of a HOST block as a pipeline over z: H2D of the next levels,
/// the CFL kernel on the current ones and D2H of what it wrote overlap
int timestep_host_pipelined(vlct_handle* h, const vlct_block* b, const Geom& G,
                            cudaStream_t st, int levels)
{
  int rc;
  h->events_used = 0;
  const vlct_block* db = &h->mirror;
  const size_t plane = (size_t) G.mx * (size_t) G.my;
  bool first = true;
  for (int z0 = 0; z0 < G.mz; z0 += levels) {
    const int z1 = (z0 + levels < G.mz) ? z0 + levels : G.mz;
    if ((rc = mirror_copy(h, b, G, h->in_stream, true, COPY_TIMESTEP_IN, z0, z1)) != VLCT_OK)
      return rc;
    cudaEvent_t ev;
    if ((rc = record_event(h, h->in_stream, &ev)) != VLCT_OK) return rc;
    CUDA_TRY(h, cudaStreamWaitEvent(st, ev, 0));
    if ((rc = timestep_launch(h, db, G, st, ZClip{ z0, z1 }, first)) != VLCT_OK) return rc;
    first = false;
    if ((rc = record_event(h, st, &ev)) != VLCT_OK) return rc;
    CUDA_TRY(h, cudaStreamWaitEvent(h->out_stream, ev, 0));
    const size_t off = plane * (size_t) z0;
    const size_t bytes = plane * (size_t) (z1 - z0) * sizeof(double);
    CUDA_TRY(h, cudaMemcpyAsync(b->pressure + off, db->pressure + off, bytes,
                                cudaMemcpyDeviceToHost, h->out_stream));
    h->copied_bytes[1] += (long long) bytes;
    if (h->P.de) {
      CUDA_TRY(h, cudaMemcpyAsync(b->total_energy + off, db->total_energy + off, bytes,
                                  cudaMemcpyDeviceToHost, h->out_stream));
      CUDA_TRY(h, cudaMemcpyAsync(b->internal_energy + off, db->internal_energy + off,
                                  bytes, cudaMemcpyDeviceToHost, h->out_stream));
      h->copied_bytes[1] += 2 * (long long) bytes;
    }
  }
  CUDA_TRY(h, cudaMemcpyAsync(h->h_dt_bits, h->d_dt_bits, sizeof(unsigned long long),
                              cudaMemcpyDeviceToHost, st));
  CUDA_TRY(h, cudaStreamSynchronize(st));
  CUDA_TRY(h, cudaStreamSynchronize(h->out_stream));
  return VLCT_OK;
}
}  // namespace

extern "C" {

int vlct_compute(vlct_handle* h, const vlct_block* b, double dt)
{ return compute_entry(h, b, dt, nullptr); }

int vlct_compute_and_timestep(vlct_handle* h, const vlct_block* b, double dt,
                              double* dt_next)
{
  if (h != nullptr && dt_next == nullptr)
    return fail(h, VLCT_ERR_INVALID_BLOCK, "dt_next is NULL");
  return compute_entry(h, b, dt, nullptr, dt_next);
}

int vlct_compute_and_timestep_dev(vlct_handle* h, const vlct_block* b,
                                  const double* dt_device, double* dt_next_device)
{
  if (h != nullptr && (dt_device == nullptr || dt_next_device == nullptr))
    return fail(h, VLCT_ERR_INVALID_BLOCK, "dt_device / dt_next_device is NULL");
  return compute_entry(h, b, 0.0, dt_device, nullptr, dt_next_device);
}

int vlct_compute_dev(vlct_handle* h, const vlct_block* b, const double* dt_device)
{
  if (h != nullptr && dt_device == nullptr)
    return fail(h, VLCT_ERR_INVALID_BLOCK, "dt_device is NULL");
  return compute_entry(h, b, 0.0, dt_device);
}

int vlct_timestep_dev(vlct_handle* h, const vlct_block* b, double* dt_device)
{
  if (h == nullptr) return VLCT_ERR_INVALID_CONFIG;
  DeviceGuard device_guard__(h->device);
  if (h->device < 0) return fail(h, VLCT_ERR_NO_DEVICE, "handle has no device");
  if (dt_device == nullptr) return fail(h, VLCT_ERR_INVALID_BLOCK, "dt_device is NULL");
  int rc = check_block(h, b, true);
  if (rc != VLCT_OK) return rc;
  if (b->mem_space != VLCT_MEM_DEVICE)
    return fail(h, VLCT_ERR_INVALID_BLOCK,
                "vlct_timestep_dev needs a block in device memory");
  const Geom G = geom_of(b);
  cudaStream_t st = b->stream ? (cudaStream_t) b->stream : h->own_stream;
  if ((rc = timestep_launch(h, b, G, st)) != VLCT_OK) return rc;
  // "Multiply resulting dt by CourantSafetyNumber" (cpp:585-587), on the device
  launch_finish_dt(LaunchCtx{ st, &h->launches, &h->prof, h->pair_kernels }, h->d_dt_bits,
                   h->cfg.courant, dt_device);
  CUDA_TRY(h, cudaGetLastError());
  return VLCT_OK;
}

int vlct_timestep(vlct_handle* h, const vlct_block* b, double* dt_out)
{
  if (h == nullptr) return VLCT_ERR_INVALID_CONFIG;
  DeviceGuard device_guard__(h->device);
  if (h->device < 0) return fail(h, VLCT_ERR_NO_DEVICE, "handle has no device");
  if (dt_out == nullptr) return fail(h, VLCT_ERR_INVALID_BLOCK, "dt_out is NULL");
  int rc = check_block(h, b, true);
  if (rc != VLCT_OK) return rc;
  const Geom G = geom_of(b);
  cudaStream_t st;
  const vlct_block* db = b;
  if (b->mem_space == VLCT_MEM_DEVICE) {
    st = b->stream ? (cudaStream_t) b->stream : h->own_stream;
  } else {
    st = h->own_stream;
    if ((rc = ensure_mirror(h, b, G)) != VLCT_OK) return rc;
    const bool reuse = mirror_serves(h, b);
    h->mirror_is_current = false;   // one timestep per compute; DE sync rewrites energies
    int levels = reuse ? 0 : host_levels(h, G);
    if (h->cfg.time_scheme == VLCT_TIME_EULER && h->host_pipeline_levels > 0)
      levels = (int) h->host_pipeline_levels;   // no stage coupling in timestep
    if (levels > 0) {
      if ((rc = timestep_host_pipelined(h, b, G, st, levels)) != VLCT_OK) return rc;
      double dt_min;
      memcpy(&dt_min, h->h_dt_bits, sizeof(double));
      *dt_out = dt_min * h->cfg.courant;
      return VLCT_OK;
    }
    if (!reuse &&
        (rc = mirror_copy(h, b, G, st, true, COPY_TIMESTEP_IN)) != VLCT_OK) return rc;
    db = &h->mirror;
  }
  if ((rc = timestep_launch(h, db, G, st)) != VLCT_OK) return rc;
  CUDA_TRY(h, cudaMemcpyAsync(h->h_dt_bits, h->d_dt_bits, sizeof(unsigned long long),
                              cudaMemcpyDeviceToHost, st));
  if (b->mem_space == VLCT_MEM_HOST) {
    // timestep writes "pressure" and (dual energy) total/internal energy
    const size_t bytes = G.cells() * sizeof(double);
    CUDA_TRY(h, cudaMemcpyAsync(b->pressure, db->pressure, bytes, cudaMemcpyDeviceToHost, st));
    h->copied_bytes[1] += (long long) bytes;
    if (h->P.de) {
      CUDA_TRY(h, cudaMemcpyAsync(b->total_energy, db->total_energy, bytes, cudaMemcpyDeviceToHost, st));
      CUDA_TRY(h, cudaMemcpyAsync(b->internal_energy, db->internal_energy, bytes, cudaMemcpyDeviceToHost, st));
      h->copied_bytes[1] += 2 * (long long) bytes;
    }
  }
  CUDA_TRY(h, cudaStreamSynchronize(st));
  double dt_baryons;
  memcpy(&dt_baryons, h->h_dt_bits, sizeof(double));
  // "Multiply resulting dt by CourantSafetyNumber" (cpp:585-587)
  *dt_out = dt_baryons * h->cfg.courant;
  return VLCT_OK;
}

// ---- batches of equally shaped blocks ---------------------------------------
}  // extern "C"

namespace {

/// every block of a batch must look like the first one
int check_batch(vlct_handle* h, const vlct_block* blocks, int nblocks,
                bool for_timestep)
{
  if (blocks == nullptr || nblocks <= 0)
    return fail(h, VLCT_ERR_INVALID_BLOCK, "empty batch");
  int rc;
  const vlct_block& b0 = blocks[0];
  if ((rc = check_block(h, &b0, for_timestep)) != VLCT_OK) return rc;
  {
    // a sub-batch is stacked along z: its cell count must fit 32 bits as well
    const Geom G0 = geom_of(&b0);
    long long cap = 65535 / (G0.mz + 1);
    if (cap > h->batch_max_blocks) cap = h->batch_max_blocks;
    if (cap > nblocks) cap = nblocks;
    if ((double) G0.mx * G0.my * (double) (G0.mz + 1) * (double) cap >= 2147483648.0)
      return fail(h, VLCT_ERR_INVALID_BLOCK,
                  "batch too large: lower the option batch_max_blocks");
  }
  for (int n = 0; n < nblocks; n++) {
    const vlct_block& b = blocks[n];
    if ((rc = check_block(h, &b, for_timestep)) != VLCT_OK) return rc;
    if (b.nx != b0.nx || b.ny != b0.ny || b.nz != b0.nz || b.gx != b0.gx ||
        b.gy != b0.gy || b.gz != b0.gz || b.dx != b0.dx || b.dy != b0.dy ||
        b.dz != b0.dz || b.mem_space != b0.mem_space || b.stream != b0.stream)
      return fail(h, VLCT_ERR_INVALID_BLOCK,
                  "block %d of the batch differs from block 0 in shape, cell "
                  "width, mem_space or stream", n);
    for (int f = 0; f < kNumFields; f++)
      if ((b.*(kFields[f].member) == nullptr) != (b0.*(kFields[f].member) == nullptr))
        return fail(h, VLCT_ERR_INVALID_BLOCK,
                    "block %d of the batch does not have the same fields as block 0", n);
  }
  return VLCT_OK;
}

/// stacked device arrays (arena `a`) for `nrep` blocks shaped like b0 (grow-only)
int ensure_arena(vlct_handle* h, int a, const vlct_block& b0, const Geom& G)
{
  vlct_block& arena = h->arena[a];
  bool enough = (h->arena_capacity[a] >= G.nrep) && h->arena_G[a].mx == G.mx &&
                h->arena_G[a].my == G.my && h->arena_G[a].mz == G.mz;
  for (int f = 0; enough && f < kNumFields; f++)
    if (b0.*(kFields[f].member) != nullptr && arena.*(kFields[f].member) == nullptr)
      enough = false;     // a field the arena was not built with
  if (enough) {
    // shape is fixed per handle; cell widths may change from batch to batch
    arena.dx = b0.dx; arena.dy = b0.dy; arena.dz = b0.dz;
    return VLCT_OK;
  }
  CUDA_TRY(h, cudaDeviceSynchronize());
  h->ptr_table_valid = false;
  for (void* p : h->arena_allocs[a]) cudaFree(p);
  h->arena_allocs[a].clear();
  h->arena_capacity[a] = 0;
  arena = b0;
  arena.mem_space = VLCT_MEM_DEVICE;
  int rc;
  for (int f = 0; f < kNumFields; f++) {
    arena.*(kFields[f].member) = nullptr;
    if (b0.*(kFields[f].member) == nullptr) continue;
    double* p;
    if ((rc = dev_alloc(h, &p, field_count(G, kFields[f].face), POOL_ARENA + a)) != VLCT_OK)
      return rc;
    arena.*(kFields[f].member) = p;
  }
  for (int s = 0; s < VLCT_MAX_PASSIVE; s++) {
    arena.passive[s] = nullptr;
    if (s < h->P.nsc) {
      double* p;
      if ((rc = dev_alloc(h, &p, cell_count(G), POOL_ARENA + a)) != VLCT_OK) return rc;
      arena.passive[s] = p;
    }
  }
  CUDA_TRY(h, cudaStreamSynchronize(cudaStreamLegacy));
  h->arena_capacity[a] = G.nrep;
  h->arena_G[a] = G;
  return VLCT_OK;
}

/// A pointer the device can dereference for p: p itself for device memory, the
/// device alias for pinned / registered host memory (zero-copy over PCIe),
/// nullptr for pageable host memory.
double* device_alias(vlct_handle* h, double* p, bool host)
{
  if (!host || p == nullptr) return p;
  auto it = h->mapped_host.find(p);
  if (it != h->mapped_host.end()) return it->second;
  cudaPointerAttributes attr;
  double* alias = nullptr;
  if (cudaPointerGetAttributes(&attr, p) == cudaSuccess) {
    if (attr.type == cudaMemoryTypeHost && attr.devicePointer != nullptr)
      alias = (double*) attr.devicePointer;
  } else {
    cudaGetLastError();   // pageable memory on older drivers: not an error for us
  }
  if (alias != nullptr) h->mapped_host[p] = alias;   // (pageable: looked up again next time)
  return alias;
}

/// Device table of every block's field pointers, row = field slot, column =
/// block (all nblocks of the batch): what the gather / scatter kernels index.
/// For HOST blocks it exists only if every array is pinned or registered
/// (ptr_table_usable); pageable arrays go through cudaMemcpyAsync instead.
int prepare_ptr_table(vlct_handle* h, const vlct_block* blocks, int nblocks)
{
  const bool host = (blocks[0].mem_space == VLCT_MEM_HOST);
  const int nslots = kNumFields + h->P.nsc;
  const size_t need = (size_t) (kNumFields + VLCT_MAX_PASSIVE) * (size_t) nblocks;
  if (need > h->ptr_table_count) {
    CUDA_TRY(h, cudaDeviceSynchronize());
    if (h->d_ptr_table) cudaFree(h->d_ptr_table);
    if (h->h_ptr_table) cudaFreeHost(h->h_ptr_table);
    CUDA_TRY(h, cudaMalloc((void**) &h->d_ptr_table, need * sizeof(double*)));
    CUDA_TRY(h, cudaMallocHost((void**) &h->h_ptr_table, need * sizeof(double*)));
    h->ptr_table_count = need;
    h->ptr_table_valid = false;
  }
  bool same = h->ptr_table_valid && h->ptr_table_nb == nblocks;
  bool usable = true;
  std::vector<double*> fresh;
  fresh.reserve((size_t) nslots * nblocks);
  for (int f = 0; f < nslots; f++)
    for (int n = 0; n < nblocks; n++) {
      double* raw = f < kNumFields ? blocks[n].*(kFields[f].member)
                                   : blocks[n].passive[f - kNumFields];
      double* p = device_alias(h, raw, host);
      if (raw != nullptr && p == nullptr) usable = false;
      fresh.push_back(p);
      if (same && h->h_ptr_table[(size_t) f * nblocks + n] != p) same = false;
    }
  h->ptr_table_usable = usable;
  if (!usable) { h->ptr_table_valid = false; return VLCT_OK; }
  if (!same) {
    // the staging buffer may still be the source of an upload in flight
    CUDA_TRY(h, cudaDeviceSynchronize());
    memcpy(h->h_ptr_table, fresh.data(), fresh.size() * sizeof(double*));
    CUDA_TRY(h, cudaMemcpy(h->d_ptr_table, h->h_ptr_table,
                           fresh.size() * sizeof(double*), cudaMemcpyHostToDevice));
    h->ptr_table_valid = true;
    h->ptr_table_nb = nblocks;
  }
  return VLCT_OK;
}

/// Move a set of fields between nb blocks of a batch (starting at block `first`
/// of the nblocks the pointer table was prepared for) and the stacked arena:
/// one gather / scatter kernel per field over the device table of the blocks'
/// pointers -- device memory, or pinned / registered host memory read and
/// written in place over PCIe --, or one cudaMemcpyAsync per (block, field) for
/// pageable host arrays.
int batch_copy(vlct_handle* h, int a, const vlct_block* blocks, int first, int nb,
               int nblocks, const Geom& G, cudaStream_t st, bool to_arena, CopySet set)
{
  const vlct_block& arena = h->arena[a];
  const bool host = (blocks[0].mem_space == VLCT_MEM_HOST);
  const Geom one{ G.mx, G.my, G.mz, 1, 0 };
  const LaunchCtx ctx{ st, &h->launches, &h->prof, h->pair_kernels };
  // DEVICE blocks: gather / scatter kernels. HOST blocks: see host_batch_copy_mode.
  const bool use_kernels = h->ptr_table_usable && (!host || h->host_batch_copy_mode == 1);
  const bool use_batch_memcpy = host && !use_kernels && h->host_batch_copy_mode != 2 &&
                                h->batch_memcpy_works;
  std::vector<void*> dsts, srcs;
  std::vector<size_t> sizes;
  auto move = [&](int slot, double* stacked, int face,
                  double* vlct_block::*member, int passive) -> int {
    const size_t count = field_count(one, face);
    const size_t plane = (size_t) (G.mx + (face == 0)) * (size_t) (G.my + (face == 1));
    const size_t stride = plane * (size_t) G.zper;
    if (use_kernels) {
      launch_batch_copy(ctx, stacked, h->d_ptr_table + (size_t) slot * nblocks + first,
                        nb, count, stride, to_arena, host);
      if (host) h->copied_bytes[to_arena ? 0 : 1] += (long long) (nb * count * sizeof(double));
      return VLCT_OK;
    }
    for (int n = 0; n < nb; n++) {
      const vlct_block& b = blocks[first + n];
      double* hp = passive >= 0 ? b.passive[passive] : b.*member;
      double* dp = stacked + (size_t) n * stride;
      if (use_batch_memcpy) {
        dsts.push_back(to_arena ? (void*) dp : (void*) hp);
        srcs.push_back(to_arena ? (void*) hp : (void*) dp);
        sizes.push_back(count * sizeof(double));
      } else {
        CUDA_TRY(h, cudaMemcpyAsync(to_arena ? (void*) dp : (void*) hp,
                                    to_arena ? (void*) hp : (void*) dp,
                                    count * sizeof(double),
                                    to_arena ? cudaMemcpyHostToDevice
                                             : cudaMemcpyDeviceToHost, st));
      }
      h->copied_bytes[to_arena ? 0 : 1] += (long long) (count * sizeof(double));
    }
    return VLCT_OK;
  };
  int rc;
  for (int f = 0; f < kNumFields; f++) {
    double* stacked = arena.*(kFields[f].member);
    if (stacked == nullptr || blocks[0].*(kFields[f].member) == nullptr) continue;
    if (!in_copy_set(h, kFields[f].member, set)) continue;
    if ((rc = move(f, stacked, kFields[f].face, kFields[f].member, -1)) != VLCT_OK)
      return rc;
  }
  for (int s = 0; s < h->P.nsc; s++)
    if ((rc = move(kNumFields + s, arena.passive[s], -1, nullptr, s)) != VLCT_OK)
      return rc;
  if (use_batch_memcpy && !dsts.empty()) {
    // all (block, field) copies of this set in ONE call, executed by the copy
    // engines: no per-copy launch cost, no SM time, H2D and D2H truly overlap
    cudaMemcpyAttributes attr;
    memset(&attr, 0, sizeof(attr));
    attr.srcAccessOrder = cudaMemcpySrcAccessOrderStream;
    size_t attr_idx = 0, fail_idx = 0;
    cudaError_t err = cudaMemcpyBatchAsync(dsts.data(), srcs.data(), sizes.data(),
                                           dsts.size(), &attr, &attr_idx, 1, &fail_idx, st);
    if (err != cudaSuccess) {
      // older driver: remember, and do these copies one by one
      cudaGetLastError();
      h->batch_memcpy_works = false;
      for (size_t i = 0; i < dsts.size(); i++)
        CUDA_TRY(h, cudaMemcpyAsync(dsts[i], srcs[i], sizes[i], cudaMemcpyDefault, st));
    }
  }
  CUDA_TRY(h, cudaGetLastError());
  return VLCT_OK;
}

/// sub-batch size: gridDim.y carries (z levels) x (blocks)
int batch_chunk(const vlct_handle* h, const Geom& G)
{
  long long cap = 65535 / (G.mz + 1);
  if (cap > h->batch_max_blocks) cap = h->batch_max_blocks;
  return cap < 1 ? 1 : (int) cap;
}

Geom stacked_geom(const vlct_block& b0, int nrep)
{
  Geom G = geom_of(&b0);
  G.nrep = nrep;
  G.zper = G.mz + 1;
  return G;
}

}  // namespace

extern "C" {

int vlct_save_face_fluxes(vlct_handle* h, const vlct_block* b,
                          const vlct_face_fluxes* out)
{
  if (h == nullptr) return VLCT_ERR_INVALID_CONFIG;
  DeviceGuard device_guard__(h->device);
  if (b == nullptr || out == nullptr)
    return fail(h, VLCT_ERR_INVALID_BLOCK, "NULL argument to vlct_save_face_fluxes");
  if (h->P.mhd)   // EnzoMethodMHDVlct.cpp:137-141
    return fail(h, VLCT_ERR_INVALID_CONFIG,
                "Flux corrections are currently only supported in hydro-mode");
  const Geom G = geom_of(b);
  if (h->G.mx == 0 || !h->stepped)
    return fail(h, VLCT_ERR_INVALID_BLOCK,
                "vlct_save_face_fluxes needs a preceding vlct_compute");
  if (G.mx != h->G.mx || G.my != h->G.my || G.mz != h->G.mz)
    return fail(h, VLCT_ERR_INVALID_BLOCK, "block shape differs from the last compute");
  const bool host = (out->mem_space == VLCT_MEM_HOST);
  cudaStream_t st = (b->mem_space == VLCT_MEM_DEVICE && b->stream)
                        ? (cudaStream_t) b->stream : h->own_stream;
  const LaunchCtx ctx{ st, &h->launches, &h->prof, h->pair_kernels };
  const int n[3] = { b->nx, b->ny, b->nz }, g[3] = { b->gx, b->gy, b->gz };
  const int m[3] = { G.mx, G.my, G.mz };
  const int nstages = (h->cfg.time_scheme == VLCT_TIME_EULER) ? 1 : 2;
  const double* sp = h->d_step + 4 * (nstages - 1);   // dt/dx, dt/dy, dt/dz of the final stage
  // device staging for HOST output: the largest face of one field at a time
  size_t max_face = 0;
  for (int d = 0; d < 3; d++) {
    const int a0 = (d == 0) ? 1 : 0, a1 = (d == 2) ? 1 : 2;
    const size_t cnt = (size_t) n[a0] * n[a1];
    if (cnt > max_face) max_face = cnt;
  }
  if (host && h->face_stage_count < 2 * max_face) {
    if (h->d_face_stage) cudaFree(h->d_face_stage);
    CUDA_TRY(h, cudaMalloc((void**) &h->d_face_stage, 2 * max_face * sizeof(double)));
    h->face_stage_count = 2 * max_face;
  }
  for (int d = 0; d < 3; d++) {
    const FluxSet& F = h->S.flux[d];
    const double* arrays[VLCT_FLUX_FIELDS] = { F.rho, F.mx_, F.my_, F.mz_, F.e,
                                               h->P.de ? F.eint : nullptr };
    if (h->P.nsc > 0 && h->P.nsc_flux == 0) {
      for (int s = 0; s < h->P.nsc; s++)
        for (int side = 0; side < 2; side++)
          if (out->face[d][side][6 + s] != nullptr)
            return fail(h, VLCT_ERR_INVALID_CONFIG,
                        "passive-scalar face fluxes need the option "
                        "\"scalar_flux_arrays\" = 1 (set before vlct_compute)");
    }
    for (int s = 0; s < h->P.nsc_flux; s++) arrays[6 + s] = F.sc[s];
    const int a0 = (d == 0) ? 1 : 0, a1 = (d == 2) ? 1 : 2;
    const size_t cnt = (size_t) n[a0] * n[a1];
    for (int f = 0; f < 6 + h->P.nsc; f++) {
      if (arrays[f] == nullptr) continue;
      for (int side = 0; side < 2; side++) {
        double* dst = out->face[d][side][f];
        if (dst == nullptr) continue;
        const int at = side ? m[d] - g[d] - 1 : g[d] - 1;
        double* ddst = host ? h->d_face_stage + (size_t) side * max_face : dst;
        launch_face_flux(ctx, G, arrays[f], sp + d, ddst, d, at, n[a0], n[a1],
                         g[a0], g[a1]);
        if (host) {
          CUDA_TRY(h, cudaMemcpyAsync(dst, ddst, cnt * sizeof(double),
                                      cudaMemcpyDeviceToHost, st));
          h->copied_bytes[1] += (long long) (cnt * sizeof(double));
        }
      }
      // the staging buffer is reused by the next field
      if (host) CUDA_TRY(h, cudaStreamSynchronize(st));
    }
  }
  CUDA_TRY(h, cudaGetLastError());
  return VLCT_OK;
}

}  // extern "C"

namespace {
/// vlct_compute_batch; with dt_next != nullptr vlct_compute_and_timestep_batch:
/// the CFL kernel follows the update of every sub-batch on its stacked arrays
/// and "pressure" (+ the dual-energy-synced energies) come back with the rest
int compute_batch_impl(vlct_handle* h, const vlct_block* blocks, int nblocks, double dt,
                       double* dt_next)
{
  if (h == nullptr) return VLCT_ERR_INVALID_CONFIG;
  DeviceGuard device_guard__(h->device);
  if (h->device < 0) return fail(h, VLCT_ERR_NO_DEVICE, "handle has no device");
  int rc = check_batch(h, blocks, nblocks, false);
  if (rc != VLCT_OK) return rc;
  const bool fused = (dt_next != nullptr);
  if (fused && blocks[0].pressure == nullptr)
    return fail(h, VLCT_ERR_INVALID_BLOCK, "\"pressure\" must be a permanent field");
  const CopySet out_set = fused ? COPY_FUSED_OUT : COPY_COMPUTE_OUT;
  auto finish_dt = [&](cudaStream_t st) -> int {
    CUDA_TRY(h, cudaMemcpyAsync(h->h_dt_bits, h->d_dt_bits, sizeof(unsigned long long),
                                cudaMemcpyDeviceToHost, st));
    CUDA_TRY(h, cudaStreamSynchronize(st));
    double dt_min;
    memcpy(&dt_min, h->h_dt_bits, sizeof(double));
    *dt_next = dt_min * h->cfg.courant;
    return VLCT_OK;
  };
  const bool host = (blocks[0].mem_space == VLCT_MEM_HOST);
  cudaStream_t st = (!host && blocks[0].stream) ? (cudaStream_t) blocks[0].stream
                                                 : h->own_stream;
  int chunk = batch_chunk(h, geom_of(&blocks[0]));
  if ((rc = prepare_ptr_table(h, blocks, nblocks)) != VLCT_OK) return rc;
  if (!host) {
    for (int first = 0; first < nblocks; first += chunk) {
      const int nb = (nblocks - first < chunk) ? nblocks - first : chunk;
      const Geom G = stacked_geom(blocks[0], nb);
      if ((rc = ensure_scratch(h, G)) != VLCT_OK) return rc;
      if ((rc = ensure_arena(h, 0, blocks[0], G)) != VLCT_OK) return rc;
      if ((rc = batch_copy(h, 0, blocks, first, nb, nblocks, G, st, true,
                           COPY_COMPUTE_IN)) != VLCT_OK) return rc;
      if ((rc = compute_on_device(h, &h->arena[0], G, dt, nullptr, st)) != VLCT_OK) return rc;
      if (fused && (rc = timestep_launch(h, &h->arena[0], G, st, kNoClip, first == 0)) != VLCT_OK)
        return rc;
      if ((rc = batch_copy(h, 0, blocks, first, nb, nblocks, G, st, false,
                           out_set)) != VLCT_OK) return rc;
      // (stream order protects the arena between consecutive sub-batches)
    }
    return fused ? finish_dt(st) : VLCT_OK;
  }
  // HOST blocks: a pipeline over sub-batches with two arenas -- the H2D copies
  // of sub-batch s+1 (in_stream), the kernels of sub-batch s (st) and the D2H
  // copies of sub-batch s-1 (out_stream) overlap; the kernels of consecutive
  // sub-batches share the scratch arrays and run one after the other on st.
  if (h->host_batch_blocks > 0) {
    if (h->host_batch_blocks < chunk) chunk = (int) h->host_batch_blocks;
  } else if (nblocks > 4) {
    // auto: at least ~4 sub-batches, each of at least ~16 MB per field
    const size_t per_block = geom_of(&blocks[0]).cells() * sizeof(double);
    int want = (nblocks + 3) / 4;
    const int min_blocks = (int) (((size_t) 16 << 20) / per_block) + 1;
    if (want < min_blocks) want = min_blocks;
    if (want < chunk) chunk = want;
  }
  const Geom Gmax = stacked_geom(blocks[0], nblocks < chunk ? nblocks : chunk);
  if ((rc = ensure_scratch(h, Gmax)) != VLCT_OK) return rc;
  const int narena = (nblocks > chunk) ? 2 : 1;
  for (int a = 0; a < narena; a++)
    if ((rc = ensure_arena(h, a, blocks[0], Gmax)) != VLCT_OK) return rc;
  h->events_used = 0;
  h->mirror_is_current = false;
  cudaEvent_t downloaded[2] = { nullptr, nullptr };   // arena free again
  int sub = 0;
  for (int first = 0; first < nblocks; first += chunk, sub++) {
    const int nb = (nblocks - first < chunk) ? nblocks - first : chunk;
    const int a = sub % narena;
    const Geom G = stacked_geom(blocks[0], nb);
    if (downloaded[a]) CUDA_TRY(h, cudaStreamWaitEvent(h->in_stream, downloaded[a], 0));
    if ((rc = batch_copy(h, a, blocks, first, nb, nblocks, G, h->in_stream, true,
                         COPY_COMPUTE_IN)) != VLCT_OK) return rc;
    cudaEvent_t ev;
    if ((rc = record_event(h, h->in_stream, &ev)) != VLCT_OK) return rc;
    CUDA_TRY(h, cudaStreamWaitEvent(st, ev, 0));
    if ((rc = compute_on_device(h, &h->arena[a], G, dt, nullptr, st)) != VLCT_OK) return rc;
    if (fused && (rc = timestep_launch(h, &h->arena[a], G, st, kNoClip, first == 0)) != VLCT_OK)
      return rc;
    if ((rc = record_event(h, st, &ev)) != VLCT_OK) return rc;
    CUDA_TRY(h, cudaStreamWaitEvent(h->out_stream, ev, 0));
    if ((rc = batch_copy(h, a, blocks, first, nb, nblocks, G, h->out_stream, false,
                         out_set)) != VLCT_OK) return rc;
    if ((rc = record_event(h, h->out_stream, &downloaded[a])) != VLCT_OK) return rc;
  }
  CUDA_TRY(h, cudaStreamSynchronize(h->out_stream));
  CUDA_TRY(h, cudaStreamSynchronize(st));
  return fused ? finish_dt(st) : VLCT_OK;
}
}  // namespace

extern "C" {

int vlct_compute_batch(vlct_handle* h, const vlct_block* blocks, int nblocks, double dt)
{ return compute_batch_impl(h, blocks, nblocks, dt, nullptr); }

int vlct_compute_and_timestep_batch(vlct_handle* h, const vlct_block* blocks, int nblocks,
                                    double dt, double* dt_next)
{
  if (h != nullptr && dt_next == nullptr)
    return fail(h, VLCT_ERR_INVALID_BLOCK, "dt_next is NULL");
  return compute_batch_impl(h, blocks, nblocks, dt, dt_next);
}


int vlct_timestep_batch(vlct_handle* h, const vlct_block* blocks, int nblocks,
                        double* dt_out)
{
  if (h == nullptr) return VLCT_ERR_INVALID_CONFIG;
  DeviceGuard device_guard__(h->device);
  if (h->device < 0) return fail(h, VLCT_ERR_NO_DEVICE, "handle has no device");
  if (dt_out == nullptr) return fail(h, VLCT_ERR_INVALID_BLOCK, "dt_out is NULL");
  int rc = check_batch(h, blocks, nblocks, true);
  if (rc != VLCT_OK) return rc;
  const bool host = (blocks[0].mem_space == VLCT_MEM_HOST);
  cudaStream_t st = (!host && blocks[0].stream) ? (cudaStream_t) blocks[0].stream
                                                 : h->own_stream;
  const int chunk = batch_chunk(h, geom_of(&blocks[0]));
  const LaunchCtx ctx{ st, &h->launches, &h->prof, h->pair_kernels };
  if ((rc = prepare_ptr_table(h, blocks, nblocks)) != VLCT_OK) return rc;
  for (int first = 0; first < nblocks; first += chunk) {
    const int nb = (nblocks - first < chunk) ? nblocks - first : chunk;
    const Geom G = stacked_geom(blocks[0], nb);
    if ((rc = ensure_arena(h, 0, blocks[0], G)) != VLCT_OK) return rc;
    if ((rc = batch_copy(h, 0, blocks, first, nb, nblocks, G, st, true,
                         COPY_TIMESTEP_IN)) != VLCT_OK) return rc;
    // the minimum accumulates over the sub-batches
    if ((rc = timestep_launch(h, &h->arena[0], G, st, kNoClip, first == 0)) != VLCT_OK)
      return rc;
    // timestep writes "pressure" and (dual energy) total / internal energy
    {
      const vlct_block* bb = blocks + first;
      const size_t stride = (size_t) G.mx * G.my * (size_t) G.zper;
      const size_t count = (size_t) G.mx * G.my * (size_t) G.mz;
      double* vlct_block::* const outs[3] = { &vlct_block::pressure,
                                              &vlct_block::total_energy,
                                              &vlct_block::internal_energy };
      const int nout = h->P.de ? 3 : 1;
      for (int o = 0; o < nout; o++) {
        if (h->ptr_table_usable && (!host || h->host_batch_copy_mode == 1)) {
          int slot = 0;
          for (int f = 0; f < kNumFields; f++) if (kFields[f].member == outs[o]) slot = f;
          launch_batch_copy(ctx, h->arena[0].*(outs[o]),
                            h->d_ptr_table + (size_t) slot * nblocks + first, nb, count,
                            stride, false, host);
          if (host) h->copied_bytes[1] += (long long) (nb * count * sizeof(double));
        } else {
          for (int n = 0; n < nb; n++) {
            CUDA_TRY(h, cudaMemcpyAsync(bb[n].*(outs[o]),
                                        h->arena[0].*(outs[o]) + (size_t) n * stride,
                                        count * sizeof(double), cudaMemcpyDeviceToHost, st));
            h->copied_bytes[1] += (long long) (count * sizeof(double));
          }
        }
      }
    }
    if (first + chunk < nblocks) CUDA_TRY(h, cudaStreamSynchronize(st));
  }
  CUDA_TRY(h, cudaMemcpyAsync(h->h_dt_bits, h->d_dt_bits, sizeof(unsigned long long),
                              cudaMemcpyDeviceToHost, st));
  CUDA_TRY(h, cudaStreamSynchronize(st));
  double dt_min;
  memcpy(&dt_min, h->h_dt_bits, sizeof(double));
  *dt_out = dt_min * h->cfg.courant;
  return VLCT_OK;
}

int vlct_host_register(vlct_handle* h, void* ptr, unsigned long long bytes)
{
  if (h == nullptr) return VLCT_ERR_INVALID_CONFIG;
  DeviceGuard device_guard__(h->device);
  if (ptr == nullptr || bytes == 0)
    return fail(h, VLCT_ERR_INVALID_BLOCK, "vlct_host_register: empty range");
  CUDA_TRY(h, cudaHostRegister(ptr, (size_t) bytes, cudaHostRegisterPortable |
                                                     cudaHostRegisterMapped));
  h->registered[ptr] = (size_t) bytes;
  return VLCT_OK;
}

int vlct_host_unregister(vlct_handle* h, void* ptr)
{
  if (h == nullptr) return VLCT_ERR_INVALID_CONFIG;
  DeviceGuard device_guard__(h->device);
  auto it = h->registered.find(ptr);
  if (it == h->registered.end())
    return fail(h, VLCT_ERR_INVALID_BLOCK,
                "vlct_host_unregister: range was not registered through this handle");
  // nothing of this handle may still be copying from / to the range
  CUDA_TRY(h, cudaDeviceSynchronize());
  const char* lo = (const char*) ptr;
  const char* hi = lo + it->second;
  for (auto m = h->mapped_host.begin(); m != h->mapped_host.end();) {
    const char* p = (const char*) m->first;
    if (p >= lo && p < hi) m = h->mapped_host.erase(m);
    else ++m;
  }
  h->ptr_table_valid = false;
  h->mirror_is_current = false;
  h->registered.erase(it);
  CUDA_TRY(h, cudaHostUnregister(ptr));
  return VLCT_OK;
}

int vlct_set_option(vlct_handle* h, const char* key, long long value)
{
  if (h == nullptr || key == nullptr) return VLCT_ERR_INVALID_CONFIG;
  DeviceGuard device_guard__(h->device);
  if (strcmp(key, "host_pipeline_levels") == 0) {
    if (value < -1) return fail(h, VLCT_ERR_INVALID_CONFIG, "host_pipeline_levels >= -1");
    h->host_pipeline_levels = value;
  } else if (strcmp(key, "host_mirror_reuse") == 0) {
    h->host_mirror_reuse = (value != 0);
    h->mirror_is_current = false;
  } else if (strcmp(key, "host_batch_copy_mode") == 0) {
    if (value < 0 || value > 2)
      return fail(h, VLCT_ERR_INVALID_CONFIG, "host_batch_copy_mode in {0, 1, 2}");
    h->host_batch_copy_mode = value;
  } else if (strcmp(key, "host_batch_blocks") == 0) {
    if (value < 0) return fail(h, VLCT_ERR_INVALID_CONFIG, "host_batch_blocks >= 0");
    h->host_batch_blocks = value;
  } else if (strcmp(key, "batch_max_blocks") == 0) {
    if (value < 1) return fail(h, VLCT_ERR_INVALID_CONFIG, "batch_max_blocks >= 1");
    h->batch_max_blocks = value;
  } else if (strcmp(key, "scalar_flux_arrays") == 0) {
    const int want = (value != 0) ? h->P.nsc : 0;
    if (want != h->P.nsc_flux) {
      h->P.nsc_flux = want;
      if (h->G.mx != 0) {      // the scratch layout changes: rebuild it lazily
        CUDA_TRY(h, cudaDeviceSynchronize());
        for (void* p : h->allocations) cudaFree(p);
        h->allocations.clear();
        h->scratch_bytes = 0;
        h->G = Geom{0, 0, 0};
        h->scratch_levels = 0;
        h->stepped = false;
      }
    }
  } else if (strcmp(key, "pair_kernels") == 0) {
    if (value < 0 || value > 63) return fail(h, VLCT_ERR_INVALID_CONFIG, "pair_kernels in 0..63");
    h->pair_kernels = (int) value;
  } else if (strcmp(key, "debug_kernel_mask") == 0) {
    h->debug_kernel_mask = value;
  } else if (strcmp(key, "device_pipeline_levels") == 0) {
    if (value < 0) return fail(h, VLCT_ERR_INVALID_CONFIG, "device_pipeline_levels >= 0");
    h->device_pipeline_levels = value;
  } else {
    return fail(h, VLCT_ERR_UNKNOWN_KEY, "unknown option \"%s\"", key);
  }
  return VLCT_OK;
}

}  // extern "C"

namespace {
int compute_dev_part(vlct_handle* h, const vlct_block* b, const double* dt_device,
                     int part, int z_lo, int z_hi, double* dt_next_device);
}

extern "C" {

int vlct_compute_dev_part(vlct_handle* h, const vlct_block* b,
                          const double* dt_device, int part, int z_lo, int z_hi)
{ return compute_dev_part(h, b, dt_device, part, z_lo, z_hi, nullptr); }

int vlct_compute_and_timestep_dev_part(vlct_handle* h, const vlct_block* b,
                                       const double* dt_device, int part, int z_lo,
                                       int z_hi, double* dt_next_device)
{
  if (h != nullptr && dt_next_device == nullptr)
    return fail(h, VLCT_ERR_INVALID_BLOCK, "dt_next_device is NULL");
  return compute_dev_part(h, b, dt_device, part, z_lo, z_hi, dt_next_device);
}

}  // extern "C"

namespace {
int compute_dev_part(vlct_handle* h, const vlct_block* b, const double* dt_device,
                     int part, int z_lo, int z_hi, double* dt_next_device)
{
  if (h == nullptr) return VLCT_ERR_INVALID_CONFIG;
  DeviceGuard device_guard__(h->device);
  if (h->device < 0) return fail(h, VLCT_ERR_NO_DEVICE, "handle has no device");
  if (dt_device == nullptr) return fail(h, VLCT_ERR_INVALID_BLOCK, "dt_device is NULL");
  int rc = check_block(h, b, false);
  if (rc != VLCT_OK) return rc;
  if (b->mem_space != VLCT_MEM_DEVICE)
    return fail(h, VLCT_ERR_INVALID_BLOCK,
                "vlct_compute_dev_part needs a block in device memory");
  if (h->cfg.time_scheme == VLCT_TIME_EULER)
    return fail(h, VLCT_ERR_INVALID_CONFIG,
                "vlct_compute_dev_part supports the two-stage \"vl\" scheme only");
  const Geom G = geom_of(b);
  // the interior pass must not read a z ghost level, the other two must not
  // overlap each other
  if (z_lo < b->gz + kStartReach || z_hi > G.mz - b->gz - kEndReach || z_hi <= z_lo)
    return fail(h, VLCT_ERR_INVALID_BLOCK,
                "interior range [%d,%d) must satisfy gz+%d <= z_lo < z_hi <= mz-gz-%d",
                z_lo, z_hi, kStartReach, kEndReach);
  if ((rc = ensure_scratch(h, G)) != VLCT_OK) return rc;
  cudaStream_t st = b->stream ? (cudaStream_t) b->stream : h->own_stream;
  const bool fold = (dt_next_device != nullptr);
  if (fold && b->pressure == nullptr)
    return fail(h, VLCT_ERR_INVALID_BLOCK, "\"pressure\" must be a permanent field");
  const ZCut none{ CUT_NONE, 0 }, start{ CUT_START, z_lo }, end{ CUT_END, z_hi };
  switch (part) {
  case VLCT_PART_INTERIOR:
    return compute_on_device(h, b, G, 0.0, dt_device, st, start, end, true, fold);
  case VLCT_PART_LOWER:
    return compute_on_device(h, b, G, 0.0, dt_device, st, none, start, false, fold);
  case VLCT_PART_UPPER:
    rc = compute_on_device(h, b, G, 0.0, dt_device, st, end, none, false, fold);
    if (rc == VLCT_OK && fold) {
      // the three parts are done (INTERIOR, LOWER, UPPER, in this order): the
      // minimum is complete
      launch_finish_dt(LaunchCtx{ st, &h->launches, &h->prof, h->pair_kernels }, h->d_dt_bits,
                       h->cfg.courant, dt_next_device);
      CUDA_TRY(h, cudaGetLastError());
    }
    return rc;
  default: return fail(h, VLCT_ERR_INVALID_BLOCK, "unknown part %d", part);
  }
}
}  // namespace

extern "C" {


const char* vlct_last_error(const vlct_handle* h)
{ return h ? h->last_error.c_str() : "NULL handle"; }

long long vlct_kernel_launches(const vlct_handle* h)
{ return h ? h->launches : 0; }

long long vlct_scratch_bytes(const vlct_handle* h)
{ return h ? h->scratch_bytes : 0; }

long long vlct_staged_bytes(const vlct_handle* h, int direction)
{ return (h && (direction == 0 || direction == 1)) ? h->copied_bytes[direction] : 0; }

int vlct_synchronize(vlct_handle* h)
{
  if (h == nullptr) return VLCT_ERR_INVALID_CONFIG;
  DeviceGuard device_guard__(h->device);
  CUDA_TRY(h, cudaDeviceSynchronize());
  return VLCT_OK;
}

// ---- per-kernel timing ---------------------------------------------------------
int vlct_profile_enable(vlct_handle* h, int on)
{
  if (h == nullptr) return VLCT_ERR_INVALID_CONFIG;
  DeviceGuard device_guard__(h->device);
  h->prof.collect();
  h->prof.enabled = (on != 0);
  return VLCT_OK;
}

int vlct_profile_reset(vlct_handle* h)
{
  if (h == nullptr) return VLCT_ERR_INVALID_CONFIG;
  DeviceGuard device_guard__(h->device);
  h->prof.reset();
  return VLCT_OK;
}

int vlct_profile_count(vlct_handle* h)
{
  if (h == nullptr) return 0;
  h->prof.collect();
  return (int) h->prof.names.size();
}

int vlct_profile_get(vlct_handle* h, int index, char* name, int name_len,
                     double* total_ms, long long* calls)
{
  if (h == nullptr) return VLCT_ERR_INVALID_CONFIG;
  DeviceGuard device_guard__(h->device);
  h->prof.collect();
  if (index < 0 || index >= (int) h->prof.names.size())
    return fail(h, VLCT_ERR_INTERNAL, "profile index out of range");
  if (name && name_len > 0) snprintf(name, (size_t) name_len, "%s", h->prof.names[index].c_str());
  if (total_ms) *total_ms = h->prof.total_ms[index];
  if (calls) *calls = h->prof.calls[index];
  return VLCT_OK;
}

// ---- ghost-zone refresh ------------------------------------------------------

namespace {
struct RefreshField { double* p; int n0, n1, n2; int face; };

int collect_fields(vlct_handle* h, const vlct_block* b, const Geom& G,
                   std::vector<RefreshField>& out)
{
  for (int f = 0; f < kNumFields; f++) {
    double* p = b->*(kFields[f].member);
    if (p == nullptr) continue;
    const int face = kFields[f].face;
    out.push_back({ p, G.mz + (face == 2), G.my + (face == 1),
                    G.mx + (face == 0), face });
  }
  for (int s = 0; s < h->P.nsc; s++)
    if (b->passive[s]) out.push_back({ b->passive[s], G.mz, G.my, G.mx, -1 });
  return VLCT_OK;
}
}  // namespace

int vlct_refresh_periodic(vlct_handle* h, const vlct_block* b, int axes)
{
  if (h == nullptr) return VLCT_ERR_INVALID_CONFIG;
  DeviceGuard device_guard__(h->device);
  if (b == nullptr || b->mem_space != VLCT_MEM_DEVICE)
    return fail(h, VLCT_ERR_INVALID_BLOCK,
                "vlct_refresh_periodic needs a DEVICE block");
  const Geom G = geom_of(b);
  cudaStream_t st = b->stream ? (cudaStream_t) b->stream : h->own_stream;
  std::vector<RefreshField> fields;
  collect_fields(h, b, G, fields);
  const int n[3] = { b->nx, b->ny, b->nz }, g[3] = { b->gx, b->gy, b->gz };
  WrapTable table;
  table.count = 0;
  for (const RefreshField& f : fields) {
    if (table.count == kMaxWrapFields)
      return fail(h, VLCT_ERR_INTERNAL, "too many fields for the wrap table");
    table.p[table.count] = f.p;
    table.face[table.count++] = f.face;
  }
  for (int axis = 0; axis < 3; axis++) {
    if (!(axes & (1 << axis))) continue;
    if (n[axis] < g[axis])
      return fail(h, VLCT_ERR_INVALID_BLOCK,
                  "periodic refresh needs n >= ghost depth along every axis");
    launch_wrap_axis_all(LaunchCtx{ st, &h->launches, &h->prof, h->pair_kernels }, table, G.mz, G.my,
                         G.mx, axis, n[axis], g[axis]);
  }
  CUDA_TRY(h, cudaGetLastError());
  return VLCT_OK;
}

int vlct_boundary(vlct_handle* h, const vlct_block* b, int axis, int side, int type)
{
  if (h == nullptr) return VLCT_ERR_INVALID_CONFIG;
  DeviceGuard device_guard__(h->device);
  if (b == nullptr || b->mem_space != VLCT_MEM_DEVICE)
    return fail(h, VLCT_ERR_INVALID_BLOCK, "vlct_boundary needs a DEVICE block");
  if (axis < 0 || axis > 2 || (side != 0 && side != 1) ||
      (type != VLCT_BOUNDARY_OUTFLOW && type != VLCT_BOUNDARY_REFLECTING))
    return fail(h, VLCT_ERR_INVALID_BLOCK, "bad arguments to vlct_boundary");
  const Geom G = geom_of(b);
  cudaStream_t st = b->stream ? (cudaStream_t) b->stream : h->own_stream;
  const int n[3] = { b->nx, b->ny, b->nz }, g[3] = { b->gx, b->gy, b->gz };
  if (type == VLCT_BOUNDARY_REFLECTING && n[axis] < g[axis])
    return fail(h, VLCT_ERR_INVALID_BLOCK,
                "reflecting boundary needs n >= ghost depth along the axis");
  // the vector component along `axis` flips under reflection
  // (has_vector_name_, EnzoBoundary.cpp:80-86)
  double* const vec[3][3] = { { b->velocity_x, b->bfield_x, b->bfieldi_x },
                              { b->velocity_y, b->bfield_y, b->bfieldi_y },
                              { b->velocity_z, b->bfield_z, b->bfieldi_z } };
  std::vector<RefreshField> fields;
  collect_fields(h, b, G, fields);
  BoundaryTable table;
  table.count = 0;
  for (const RefreshField& f : fields) {
    if (table.count == kMaxWrapFields)
      return fail(h, VLCT_ERR_INTERNAL, "too many fields for the boundary table");
    double sign = 1.0;
    for (int c = 0; c < 3; c++) if (f.p == vec[axis][c]) sign = -1.0;
    table.p[table.count] = f.p;
    table.face[table.count] = f.face;
    table.sign[table.count++] = sign;
  }
  launch_boundary_axis(LaunchCtx{ st, &h->launches, &h->prof, h->pair_kernels }, table, G.mz, G.my, G.mx,
                       axis, n[axis], g[axis], side, type);
  CUDA_TRY(h, cudaGetLastError());
  return VLCT_OK;
}

int vlct_boundary_inflow(vlct_handle* h, const vlct_block* b, int axis, int side,
                         const vlct_inflow_values* v)
{
  if (h == nullptr) return VLCT_ERR_INVALID_CONFIG;
  DeviceGuard device_guard__(h->device);
  if (b == nullptr || b->mem_space != VLCT_MEM_DEVICE)
    return fail(h, VLCT_ERR_INVALID_BLOCK, "vlct_boundary_inflow needs a DEVICE block");
  if (axis < 0 || axis > 2 || (side != 0 && side != 1) || v == nullptr)
    return fail(h, VLCT_ERR_INVALID_BLOCK, "bad arguments to vlct_boundary_inflow");
  const Geom G = geom_of(b);
  cudaStream_t st = b->stream ? (cudaStream_t) b->stream : h->own_stream;
  const int n[3] = { b->nx, b->ny, b->nz }, g[3] = { b->gx, b->gy, b->gz };
  struct Item { double* p; double value; int face; };
  const Item items[] = {
    { b->density, v->density, -1 },
    { b->velocity_x, v->velocity_x, -1 }, { b->velocity_y, v->velocity_y, -1 },
    { b->velocity_z, v->velocity_z, -1 },
    { b->total_energy, v->total_energy, -1 },
    { b->internal_energy, v->internal_energy, -1 },
    { b->bfield_x, v->bfield_x, -1 }, { b->bfield_y, v->bfield_y, -1 },
    { b->bfield_z, v->bfield_z, -1 },
    { b->bfieldi_x, v->bfieldi_x, 0 }, { b->bfieldi_y, v->bfieldi_y, 1 },
    { b->bfieldi_z, v->bfieldi_z, 2 },
    { b->pressure, v->pressure, -1 } };
  BoundaryTable table;
  table.count = 0;
  for (const Item& it : items) {
    if (it.p == nullptr || it.value != it.value) continue;
    table.p[table.count] = it.p;
    table.face[table.count] = it.face;
    table.sign[table.count++] = it.value;
  }
  for (int s = 0; s < h->P.nsc; s++) {
    const double value = v->passive[s];
    if (b->passive[s] == nullptr || value != value) continue;
    table.p[table.count] = b->passive[s];
    table.face[table.count] = -1;
    table.sign[table.count++] = value;
  }
  launch_boundary_axis(LaunchCtx{ st, &h->launches, &h->prof, h->pair_kernels }, table, G.mz, G.my, G.mx,
                       axis, n[axis], g[axis], side, VLCT_BOUNDARY_INFLOW);
  CUDA_TRY(h, cudaGetLastError());
  return VLCT_OK;
}

long long vlct_halo_bytes(const vlct_handle* h, const vlct_block* b, int axis)
{
  if (h == nullptr || b == nullptr || axis < 0 || axis > 2) return -1;
  const Geom G = geom_of(b);
  std::vector<RefreshField> fields;
  collect_fields(const_cast<vlct_handle*>(h), b, G, fields);
  const int g[3] = { b->gx, b->gy, b->gz };
  long long total = 0;
  for (const RefreshField& f : fields) {
    const int ext[3] = { f.n2, f.n1, f.n0 };
    long long cnt = g[axis];
    for (int a = 0; a < 3; a++) if (a != axis) cnt *= ext[a];
    total += cnt;
  }
  return total * (long long) sizeof(double);
}

namespace {
int halo_copy(vlct_handle* h, const vlct_block* b, int axis, int side,
              double* buffer, bool pack)
{
  if (h == nullptr) return VLCT_ERR_INVALID_CONFIG;
  DeviceGuard device_guard__(h->device);
  if (b == nullptr || b->mem_space != VLCT_MEM_DEVICE || axis < 0 || axis > 2 ||
      (side != 0 && side != 1) || buffer == nullptr)
    return fail(h, VLCT_ERR_INVALID_BLOCK, "bad arguments to halo pack/unpack");
  const Geom G = geom_of(b);
  cudaStream_t st = b->stream ? (cudaStream_t) b->stream : h->own_stream;
  std::vector<RefreshField> fields;
  collect_fields(h, b, G, fields);
  const int n[3] = { b->nx, b->ny, b->nz }, g[3] = { b->gx, b->gy, b->gz };
  size_t off = 0;
  SlabTable table;
  table.count = 0;
  for (const RefreshField& f : fields) {
    if (table.count == kMaxWrapFields)
      return fail(h, VLCT_ERR_INTERNAL, "too many fields for the slab table");
    // Along `axis` a cell-centred field has ghosts [0,g) and [g+n, 2g+n); a
    // field that is face-centred along `axis` (cen = 1) has n+1 active faces
    // [g, g+n] and ghosts [0,g), [g+n+1, 2g+n+1). The shared boundary face is
    // computed identically on both sides and is not exchanged.
    //   pack   side 0 (goes to the lower neighbour's upper ghosts): [g+cen, 2g+cen)
    //   pack   side 1 (goes to the upper neighbour's lower ghosts): [n, n+g)
    //   unpack side 0 (my lower ghosts): [0, g)
    //   unpack side 1 (my upper ghosts): [g+n+cen, 2g+n+cen)
    const int cen = (f.face == axis) ? 1 : 0;
    const int width = g[axis];
    int lo;
    if (pack) lo = (side == 0) ? g[axis] + cen : n[axis];
    else      lo = (side == 0) ? 0 : g[axis] + n[axis] + cen;
    table.p[table.count] = f.p;
    table.face[table.count] = f.face;
    table.lo[table.count] = lo;
    table.off[table.count++] = (long long) off;
    const int ext[3] = { f.n2, f.n1, f.n0 };
    size_t cnt = (size_t) width;
    for (int a = 0; a < 3; a++) if (a != axis) cnt *= (size_t) ext[a];
    off += cnt;
  }
  launch_slab_copy_all(LaunchCtx{ st, &h->launches, &h->prof, h->pair_kernels }, table, G.mz, G.my, G.mx,
                       axis, g[axis], buffer, pack);
  CUDA_TRY(h, cudaGetLastError());
  return VLCT_OK;
}
}  // namespace

int vlct_halo_pack(vlct_handle* h, const vlct_block* b, int axis, int side,
                   double* buffer)
{ return halo_copy(h, b, axis, side, buffer, true); }

int vlct_halo_unpack(vlct_handle* h, const vlct_block* b, int axis, int side,
                     const double* buffer)
{ return halo_copy(h, b, axis, side, const_cast<double*>(buffer), false); }

}  // extern "C"
