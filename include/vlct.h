/* vlct.h -- C ABI of the B200-native VL+CT hydro/MHD block update.
 *
 * This is the drop-in boundary for ONE hot path of Enzo-E: the Method plugin
 * EnzoMethodMHDVlct (van Leer predictor-corrector + constrained transport).
 * Every entry point below replaces one piece of the reference's plugin
 * surface; the reference file:line it stands in for is cited next to it.
 * (All citations are relative to the reference checkout's root.)
 *
 * Conventions
 *  - plain C types only: no C++/torch/CUDA types in any signature;
 *  - every function returns an int status (VLCT_OK == 0) -- the reference
 *    reports failures through ASSERT/ERROR macros that abort
 *    (src/Cello/error_Error.hpp:52-59,264-273); a host adapter turns a
 *    non-zero status + vlct_last_error() into the same ERROR(...) call;
 *  - all field data are fp64 (enzo_float == double under
 *    CONFIG_PRECISION_DOUBLE, src/Enzo/enzo_typedefs.hpp:14-18), stored as
 *    C-order (z,y,x) arrays with x contiguous and ghost zones included,
 *    exactly what Field::view<enzo_float>(name) wraps
 *    (src/Cello/data_FieldData.cpp:1318-1383);
 *  - a handle is NOT thread-safe: one handle per host thread / GPU, like one
 *    Method instance per Charm++ PE.
 *  - there is no CPU fallback: if no CUDA device is usable, vlct_create fails.
 */
#ifndef VLCT_H
#define VLCT_H

#ifdef __cplusplus
extern "C" {
#endif

#define VLCT_ABI_VERSION 1
#define VLCT_MAX_PASSIVE 16

/* ---- status codes ---------------------------------------------------- */
enum {
  VLCT_OK = 0,
  VLCT_ERR_INVALID_CONFIG = 1,  /* a reference ASSERT/ERROR on parameters   */
  VLCT_ERR_INVALID_BLOCK  = 2,  /* missing field pointer / bad shape        */
  VLCT_ERR_CUDA           = 3,  /* CUDA runtime failure (sticky)            */
  VLCT_ERR_NO_DEVICE      = 4,  /* no usable sm_100 device: no CPU fallback */
  VLCT_ERR_UNKNOWN_KEY    = 5,  /* vlct_config_set: unknown parameter key   */
  VLCT_ERR_INTERNAL       = 6
};

/* ---- parameter values ------------------------------------------------ */
/* Method:mhd_vlct:riemann_solver  (src/Enzo/hydro-mhd/riemann/EnzoRiemann.cpp:26-68) */
enum { VLCT_RIEMANN_HLL = 0, VLCT_RIEMANN_HLLE = 1,
       VLCT_RIEMANN_HLLC = 2, VLCT_RIEMANN_HLLD = 3 };
/* Method:mhd_vlct:reconstruct_method
 * (src/Enzo/hydro-mhd/toolkit/EnzoReconstructor.cpp:14-44) */
enum { VLCT_RECON_NN = 0, VLCT_RECON_PLM_ENZO = 1, VLCT_RECON_PLM_ATHENA = 2 };
/* Method:mhd_vlct:mhd_choice
 * (src/Enzo/hydro-mhd/EnzoMHDIntegratorStageCommands.cpp:76-98) */
enum { VLCT_MHD_UNSET = -1, VLCT_MHD_NO_BFIELD = 0,
       VLCT_MHD_CONSTRAINED_TRANSPORT = 1 };
/* Method:mhd_vlct:time_scheme (src/Enzo/hydro-mhd/EnzoMethodMHDVlct.cpp:46-60) */
enum { VLCT_TIME_VL = 0, VLCT_TIME_EULER = 1 };
/* Physics:fluid_props:dual_energy:type
 * (src/Enzo/fluid-props/EnzoDualEnergyConfig.hpp; only "disabled" and
 *  "modern" are accepted by VL+CT, EnzoMHDIntegratorStageCommands.cpp:31-34) */
enum { VLCT_DE_DISABLED = 0, VLCT_DE_MODERN = 1, VLCT_DE_BRYAN95 = 2 };
enum { VLCT_MEM_HOST = 0, VLCT_MEM_DEVICE = 1 };

/* The configuration the reference reads in the EnzoMethodMHDVlct constructor
 * (src/Enzo/hydro-mhd/EnzoMethodMHDVlct.cpp:38-152) and from
 * Physics:fluid_props (src/Enzo/enzo-core/EnzoConfig.cpp:952-1262).
 * Trivially copyable: this is also what a pup() routine serialises
 * (EnzoMethodMHDVlct.cpp:170-197). */
typedef struct vlct_config {
  int    riemann_solver;      /* default VLCT_RIEMANN_HLLD                      */
  int    reconstruct_method;  /* full-step reconstructor; default PLM_ENZO.
                                 The half step ALWAYS uses NN (cpp:54-55)       */
  double theta_limiter;       /* default 1.5, must lie in [1,2]                 */
  int    mhd_choice;          /* REQUIRED (cpp:74-77); default VLCT_MHD_UNSET   */
  int    time_scheme;         /* default VLCT_TIME_VL                           */
  double courant;             /* < 0 => default (0.3 for vl, 1.0 for euler)     */
  double gamma;               /* Physics:fluid_props:eos:gamma, default 5/3     */
  int    dual_energy;         /* default VLCT_DE_DISABLED                       */
  double dual_energy_eta;     /* default 0.001                                  */
  double density_floor;       /* must be > 0 (StageCommands.cpp:37-40)          */
  double pressure_floor;      /* must be > 0                                    */
  int    n_passive;           /* number of fields in group "color"              */
  int    has_acceleration;    /* 1 if acceleration_{x,y,z} fields exist
                                 (EnzoMethodMHDVlct.cpp:219-232)                */
} vlct_config;

/* One Cello Block as seen by Method::compute / Method::timestep:
 * the field pointers Field::view would return, the active size, the ghost
 * depth, EnzoBlock::CellWidth (src/Enzo/enzo-core/EnzoBlock.cpp:236-245).
 *
 * Array shapes, with m? = n? + 2 g?:
 *   cell-centred fields              (mz,   my,   mx)
 *   bfieldi_x                        (mz,   my,   mx+1)
 *   bfieldi_y                        (mz,   my+1, mx)
 *   bfieldi_z                        (mz+1, my,   mx)
 * Unused pointers (e.g. bfield_* in hydro mode) may be NULL. */
typedef struct vlct_block {
  int nx, ny, nz;
  int gx, gy, gz;
  double dx, dy, dz;
  double *density;
  double *velocity_x, *velocity_y, *velocity_z;
  double *total_energy;          /* specific total energy                   */
  double *internal_energy;       /* specific; only with dual energy         */
  double *bfield_x, *bfield_y, *bfield_z;     /* cell-centred, MHD only     */
  double *bfieldi_x, *bfieldi_y, *bfieldi_z;  /* face-centred, MHD only     */
  double *pressure;              /* permanent scratch field, written by
                                    timestep (EnzoMethodMHDVlct.cpp:571-573)*/
  double *acceleration_x, *acceleration_y, *acceleration_z; /* optional      */
  double *passive[VLCT_MAX_PASSIVE]; /* "color" group scalars, as densities  */
  int   mem_space;               /* VLCT_MEM_HOST: pointers are host memory,
                                    staged H2D/D2H inside the call;
                                    VLCT_MEM_DEVICE: device pointers         */
  void *stream;                  /* cudaStream_t (DEVICE only); NULL = the
                                    library's own non-blocking stream, which
                                    does NOT wait for work the caller queued
                                    on other streams: pass the stream that
                                    produced the data (cudaStreamLegacy for
                                    the default stream) or synchronise first */
} vlct_block;

typedef struct vlct_handle vlct_handle;

/* ---- configuration ---------------------------------------------------- */

/* Fill *cfg with the reference's defaults (see field comments above). */
int vlct_config_init(vlct_config *cfg);

/* Set one parameter from its parameter-file key and textual value, e.g.
 *   vlct_config_set(&cfg, "Method:mhd_vlct:riemann_solver", "hlld");
 *   vlct_config_set(&cfg, "Physics:fluid_props:eos:gamma", "1.4");
 * Keys: the Method:mhd_vlct:* keys parsed at
 * src/Enzo/hydro-mhd/EnzoMethodMHDVlct.cpp:38-101 and the
 * Physics:fluid_props:{eos:gamma, dual_energy:{type,eta}, floors:{density,
 * pressure}} keys of src/Enzo/enzo-core/EnzoConfig.cpp:952-1262. The removed
 * keys half_dt_reconstruct_method / full_dt_reconstruct_method are rejected
 * like the reference does (EnzoMethodMHDVlct.cpp:62-70). errbuf (may be
 * NULL) receives a message on failure. */
int vlct_config_set(vlct_config *cfg, const char *key, const char *value,
                    char *errbuf, int errbuf_len);

/* Validate a configuration the way the reference's constructors do, without
 * touching a GPU (EnzoMethodMHDVlct.cpp:90-152,
 * EnzoMHDIntegratorStageCommands.cpp:18-64, EnzoRiemann.cpp:26-68,
 * EnzoReconstructor.cpp:14-44, EnzoBfieldMethod.cpp:14-27). */
int vlct_config_validate(const vlct_config *cfg, char *errbuf, int errbuf_len);

/* ---- the Method plugin surface ----------------------------------------- */

/* EnzoMethodMHDVlct::EnzoMethodMHDVlct(ParameterGroup, bool)
 * (src/Enzo/hydro-mhd/EnzoMethodMHDVlct.cpp:90-152). Binds to the current
 * CUDA device. Scratch space is allocated lazily from the first block and
 * reused for every later block (cpp:236-246), so all blocks given to one
 * handle must share one shape. */
int vlct_create(const vlct_config *cfg, vlct_handle **out);

/* ~EnzoMethodMHDVlct (cpp:156-166) */
void vlct_destroy(vlct_handle *h);

/* Method::name() (src/Enzo/hydro-mhd/EnzoMethodMHDVlct.hpp:123-124): "mhd_vlct" */
const char *vlct_name(void);

/* EnzoMethodMHDVlct::compute(Block*) (cpp:356-500) with dt = block->dt().
 * Advances the block's fields in place by one full VL+CT step. On return the
 * arrays hold exactly what the reference leaves behind, ghost zones included:
 * hydro fields change only in the active zone, centred B on [2,m-2)^3, face
 * B on the active-zone faces. The caller still calls block->compute_done(). */
int vlct_compute(vlct_handle *h, const vlct_block *block, double dt);

/* EnzoMethodMHDVlct::timestep(Block*) (cpp:551-588): dual-energy sync, writes
 * the "pressure" field, returns courant * min over ALL cells (ghosts included,
 * EnzoMHDIntegratorStageCommands.cpp:299-366) of dx_i/(|v_i| + c_signal). */
int vlct_timestep(vlct_handle *h, const vlct_block *block, double *dt_out);

/* The same two entry points with the timestep kept in DEVICE memory, for
 * drivers whose fields are device-resident (VLCT_MEM_DEVICE blocks only):
 * vlct_timestep_dev writes courant * min(...) to *dt_device without waiting for
 * the host, vlct_compute_dev reads the step's dt from *dt_device (after, e.g.,
 * an NCCL min-all-reduce over blocks -- the stand-in for the reduction of
 * src/Cello/control_stopping.cpp:96-142). Both are asynchronous on
 * block->stream, so consecutive cycles queue back to back. Results are
 * bit-identical to vlct_timestep / vlct_compute. */
int vlct_timestep_dev(vlct_handle *h, const vlct_block *block, double *dt_device);
int vlct_compute_dev(vlct_handle *h, const vlct_block *block,
                     const double *dt_device);

/* One step in three parts, for drivers that overlap the ghost-zone exchange
 * with the update (the reference overlaps the refresh messages of one block
 * with other blocks' compute through Charm++'s message-driven scheduling,
 * src/Cello/control_refresh.cpp:243-359; with one large block per GPU the
 * overlap has to happen inside the block):
 *   VLCT_PART_INTERIOR  everything that can be computed from the cell levels
 *                       z_lo-5 .. z_hi+5 alone (z in ghost-including indices),
 *                       i.e. without the z ghost levels when
 *                       gz+5 <= z_lo < z_hi <= mz-gz-6;
 *   VLCT_PART_LOWER / VLCT_PART_UPPER  the rest, below / above it; these read
 *                       the z ghost levels and may be issued in either order
 *                       once the exchange has delivered them.
 * INTERIOR must be issued first (it also latches dt), all three with the same
 * (z_lo, z_hi). Every kernel of the step is a pure function of its inputs at a
 * given index and the parts tile each kernel's index box exactly once, so the
 * three calls together leave bit-for-bit what vlct_compute_dev leaves
 * (tests/test_gpu_parts.py). Two-stage "vl" scheme, DEVICE blocks only. */
enum { VLCT_PART_INTERIOR = 0, VLCT_PART_LOWER = 1, VLCT_PART_UPPER = 2 };
int vlct_compute_dev_part(vlct_handle *h, const vlct_block *block,
                          const double *dt_device, int part, int z_lo, int z_hi);
/* The three parts with the next cycle's timestep folded in (see
 * vlct_compute_and_timestep_dev): every part adds the CFL minimum of the levels
 * it finishes; issue INTERIOR first and UPPER last -- UPPER completes the
 * minimum and writes courant * min to *dt_next_device. */
int vlct_compute_and_timestep_dev_part(vlct_handle *h, const vlct_block *block,
                                       const double *dt_device, int part, int z_lo,
                                       int z_hi, double *dt_next_device);

/* compute(block) and the timestep(block) of the cycle that follows, in ONE
 * call. In Enzo-E's cycle the stopping phase calls Method::timestep on every
 * block right after the compute phase (src/Cello/control_compute.cpp:42-157 ->
 * control_stopping.cpp:44-142), on exactly the fields compute left behind. For
 * a VLCT_MEM_HOST block two separate calls move the cell-centred state across
 * PCIe twice (compute downloads it, timestep uploads it again). Here the CFL
 * kernel (EnzoMethodMHDVlct.cpp:551-588: dual-energy sync, "pressure" field,
 * minimum over all cells incl. ghost zones, times courant) runs on the device
 * copy right behind the update, z pass by z pass inside the staging pipeline,
 * and "pressure" plus the (dual-energy-synced) energies come back with the
 * other fields: one upload and one download per cycle. The result -- all
 * fields, "pressure", *dt_next -- is bit for bit what vlct_compute followed by
 * vlct_timestep on the same block leaves (tests/test_gpu_fused_timestep.py).
 * block->pressure must be non-NULL. The reference-side adapter
 * (integration/EnzoMethodMHDVlctGpu.cpp) calls this from compute() and hands
 * the cached value to the timestep() that follows when it is the last Method
 * touching the fields. Two-stage "vl" and single-stage "euler" schemes, HOST
 * and DEVICE blocks. */
int vlct_compute_and_timestep(vlct_handle *h, const vlct_block *block, double dt,
                              double *dt_next);
/* The same with both timesteps in device memory (DEVICE blocks): nothing waits
 * for the host, cycles queue back to back. For a single block the CFL work is
 * folded into the last stage's update kernel -- the freshly updated cells are
 * still in registers, so the separate pass over eight fields that
 * vlct_timestep_dev makes disappears -- plus one small launch for the ghost
 * shell that compute never updates. dt_next_device may alias dt_device (dt is
 * latched by the first kernel of the step). */
int vlct_compute_and_timestep_dev(vlct_handle *h, const vlct_block *block,
                                  const double *dt_device, double *dt_next_device);

/* Flux-correction output:
 * EnzoMethodMHDVlct::save_fluxes_for_corrections_
 * (src/Enzo/hydro-mhd/EnzoMethodMHDVlct.cpp:250-330, called for the final
 * stage at :480-490). After vlct_compute / vlct_compute_dev of a block,
 * vlct_save_face_fluxes writes dt/dx * (final-stage flux) through the block's
 * two faces along every dimension -- what the reference deposits in the
 * block's FluxData for Method "flux_correct" -- over the active transverse
 * extent:
 *   face[dim][side][field]  packed 2-D array, slower axis first:
 *        dim 0: (nz, ny)   dim 1: (nz, nx)   dim 2: (ny, nx);
 *        side 0 = lower face (flux index g-1), 1 = upper face (index m-g-1);
 *        field 0 density, 1..3 velocity_x..z (momentum fluxes), 4 total_energy,
 *        5 internal_energy (dual energy only), 6+s passive scalar s.
 * NULL entries are skipped. Like the reference (cpp:137-141) this is
 * supported in pure-hydro mode only (mhd_choice = "no_bfield"). The fluxes are
 * those of the LAST compute call on this handle; mem_space says where the
 * output arrays live. */
#define VLCT_FLUX_FIELDS (6 + VLCT_MAX_PASSIVE)
typedef struct vlct_face_fluxes {
  double *face[3][2][VLCT_FLUX_FIELDS];
  int mem_space;
} vlct_face_fluxes;
int vlct_save_face_fluxes(vlct_handle *h, const vlct_block *block,
                          const vlct_face_fluxes *out);

/* Many blocks of one shape in ONE set of kernel launches.
 * The reference calls compute(Block*) once per block
 * (src/Cello/control_compute.cpp:42-124), typically for 16^3..32^3-cell blocks;
 * one launch sequence per such block would leave a B200 idle. Here the blocks
 * (all with the same active size, ghost depth, cell widths, fields, mem_space
 * and stream -- e.g. all leaf blocks of one refinement level on this PE) are
 * stacked along z in device memory and every kernel of the step covers all of
 * them at once; stencils never leave a block's own levels, so each block gets
 * bit for bit what vlct_compute / vlct_timestep give it alone
 * (tests/test_gpu_batch.py). vlct_timestep_batch returns the minimum over the
 * blocks (the reference min-reduces the per-block values anyway,
 * src/Cello/control_stopping.cpp:96-142) and fills every block's "pressure".
 * `blocks` is an array of nblocks structs. HOST blocks are staged by one copy
 * per (block, field); DEVICE blocks are gathered / scattered by one kernel per
 * field. Batches larger than "batch_max_blocks" run in sub-batches. */
int vlct_compute_batch(vlct_handle *h, const vlct_block *blocks, int nblocks,
                       double dt);
int vlct_timestep_batch(vlct_handle *h, const vlct_block *blocks, int nblocks,
                        double *dt_out);
/* vlct_compute_batch followed by vlct_timestep_batch on the same blocks in one
 * call (see vlct_compute_and_timestep): the CFL kernel runs on each sub-batch's
 * stacked device arrays right behind the update; *dt_next is the minimum over
 * the batch; every block's "pressure" is filled. One upload and one download
 * per cycle for HOST blocks. */
int vlct_compute_and_timestep_batch(vlct_handle *h, const vlct_block *blocks,
                                    int nblocks, double dt, double *dt_next);

/* Pinning of host field memory. Staging of VLCT_MEM_HOST blocks is fastest
 * from page-locked memory: the copies of the single-block pipeline and of the
 * double-buffered batch pipeline become truly asynchronous and run at full
 * PCIe speed in both directions at once. Cello allocates a
 * block's fields once, as one pageable array
 * (FieldData::array_permanent_, src/Cello/data_FieldData.hpp:386-398): an
 * adapter registers that range when it first sees the block and unregisters
 * it before the block is destroyed. Thin wrappers of cudaHostRegister /
 * cudaHostUnregister, so that the host code need not link the CUDA runtime;
 * pageable memory keeps working without them. */
int vlct_host_register(vlct_handle *h, void *ptr, unsigned long long bytes);
int vlct_host_unregister(vlct_handle *h, void *ptr);

/* Tuning knobs of a handle (none changes any result bit):
 *   "scalar_flux_arrays" (0 = off, default) Passive-scalar fluxes are normally
 *        never stored: the update kernel forms the fluxes through a cell's
 *        faces itself from the specific scalars and the density fluxes (the
 *        sweeps then carry no scalar work and three arrays per scalar are
 *        neither written nor read). 1 brings the flux arrays back;
 *        vlct_save_face_fluxes needs that for its passive-scalar slots. Set it
 *        before the first vlct_compute of the handle (setting it later
 *        rebuilds the scratch).
 *   "host_mirror_reuse"  (0 = off, default) A PROMISE BY THE CALLER: between
 *        vlct_compute of a VLCT_MEM_HOST block and the vlct_timestep of the
 *        same block that follows it, nobody writes the block's fields. That is
 *        the order of Enzo-E's cycle on a unigrid when "mhd_vlct" is the last
 *        Method that touches them: compute, then the stopping phase's
 *        timestep, and only then the next refresh
 *        (src/Cello/control_charm.cpp:111-150, control_stopping.cpp:44-142).
 *        With the option on, that timestep reads the device copy the compute
 *        call left behind instead of uploading eight fields again. Any other
 *        call order falls back to uploading.
 *   "host_batch_blocks"  VLCT_MEM_HOST batches run as a pipeline over
 *        sub-batches of this many blocks, double-buffered on the device: the
 *        H2D copies of one sub-batch, the kernels of the previous one and the
 *        D2H copies of the one before overlap. 0 (default): about a quarter of
 *        the batch, at least 16 MB per field.
 *   "host_batch_copy_mode"  how VLCT_MEM_HOST batches cross PCIe: 0 (default)
 *        all (block, field) copies of a set in one cudaMemcpyBatchAsync, run by
 *        the copy engines; 1 gather / scatter kernels that read / write pinned
 *        host arrays in place; 2 one cudaMemcpyAsync per (block, field).
 *        Measured for 512 pinned blocks of 32^3: 113 / 142 / 192 ms per cycle.
 *   "batch_max_blocks"  blocks stacked per launch set by the *_batch entry
 *        points (default 1024; also limited by 65535 / (mz + 1)).
 *   "host_pipeline_levels"   VLCT_MEM_HOST blocks are staged through the GPU
 *        as a pipeline over z: the H2D copy of the next levels, the kernels on
 *        the current ones and the D2H copy of the finished ones overlap.
 *        -1 (default): ~mz/32 levels per pass for blocks of >= 32 MB per
 *        field, one shot below that; 0: always one shot; n > 0: n levels.
 *   "pair_kernels"  variants of the cell kernels, a bit mask. Bits 0-2: edge E,
 *        face B, update as pair kernels (a thread owns two x-neighbours and
 *        moves them with 128-bit loads / stores). Bit 3: edge E of a single
 *        block with its inputs staged by the TMA unit (cp.async.bulk.tensor
 *        boxes, a producer warp and 16 consumer warps over mbarriers). Bit 4:
 *        edge E and face B of a single block in one such kernel (the edge E
 *        stay on chip); it replaces both. All need an even row length mx and
 *        16-byte aligned arrays, else the one-cell kernels run; the TMA-staged
 *        kernels also need a block whose 30 x 15 tiles fill the chip (from about
 *        96^3 cells; bit 5 lifts that, for tests). Default 30.
 *        Measured at 512^3 per stage: edge E 4.45 -> 3.46 ms TMA-staged (86 %
 *        of the HBM peak on algorithmic bytes), edge E + face B 6.0 -> 4.8 ms
 *        fused, face-B pair kernel -9 %, update -3 %; the edge-E pair kernel
 *        (bit 0) is 3 % slower than the one-cell kernel.
 *   "device_pipeline_levels" run VLCT_MEM_DEVICE steps in passes of n levels
 *        too (0 = off, default; a test hook for the pass machinery). */
int vlct_set_option(vlct_handle *h, const char *key, long long value);

/* Message of the last failure on this handle (never NULL). */
const char *vlct_last_error(const vlct_handle *h);
const char *vlct_status_string(int status);

/* ---- instrumentation ---------------------------------------------------- */

/* Number of CUDA kernels this handle has launched since creation / last reset
 * (bench.py reports the difference over the timed region as gpu_launches). */
long long vlct_kernel_launches(const vlct_handle *h);

/* Bytes of device scratch currently owned by the handle. */
long long vlct_scratch_bytes(const vlct_handle *h);

/* Bytes staged so far for VLCT_MEM_HOST blocks: direction 0 = host->device,
 * 1 = device->host (bench.py reports them per step as e2e.h2d/d2h bytes). */
long long vlct_staged_bytes(const vlct_handle *h, int direction);

/* Block until all work submitted through this handle has finished. */
int vlct_synchronize(vlct_handle *h);

/* Per-kernel timing with CUDA events on the launching stream -- the device-side
 * analogue of the reference's Performance regions around Method::compute
 * (src/Cello/performance_Performance.hpp:36-65, control_compute.cpp:74,104).
 * Off by default (two event records per launch when on). */
int vlct_profile_enable(vlct_handle *h, int on);
int vlct_profile_reset(vlct_handle *h);
int vlct_profile_count(vlct_handle *h);
int vlct_profile_get(vlct_handle *h, int index, char *name, int name_len,
                     double *total_ms, long long *calls);

/* Device self-test of the straight-line fp64 division / reciprocal / square
 * root used by the flux kernels (csrc/vlct_fpops.cuh): evaluates n operand
 * tuples generated from `seed` (mode 0: solver-like magnitudes, 1: arbitrary
 * bit patterns, 2: specials) and compares with the built-in IEEE operators.
 * counters_out[10] = for op in (div, rcp, sqrt, shared-reciprocal pair,
 * divz = division with zero-numerator select): { results that passed
 * the range guard but differ from the built-in (must be 0), results flagged
 * for re-evaluation with the built-in }. No reference counterpart: the
 * reference's arithmetic is the compiler's (value-safe, OPTIMIZE_FP=OFF,
 * CMakeLists.txt:275-283), which is what this test pins the kernels to. */
int vlct_selftest_fpops(long long n, unsigned long long seed, int mode,
                        long long *counters_out);

/* ---- ghost-zone refresh on the device --------------
 * Stand-ins for the refresh phase that precedes compute() on a unigrid
 * (src/Cello/control_refresh.cpp:243-359, src/Cello/data_FieldFace.cpp).
 * All pointers are DEVICE pointers. */

/* Fill every ghost zone of every field of the block from the block's own
 * active zone (a single periodic block: root_blocks = [1,1,1] with
 * Boundary:type = "periodic"). Face-centred fields follow FieldFace's rule
 * that the shared face belongs to both sides. axes: bit 0/1/2 = x/y/z. */
int vlct_refresh_periodic(vlct_handle *h, const vlct_block *block, int axes);

/* Non-periodic domain boundaries: EnzoBoundary::enforce for one face of the
 * domain (src/Enzo/enzo-core/EnzoBoundary.cpp:33-76), applied to every field
 * of the block like Block::update_boundary_ does at the end of a refresh
 * (src/Cello/mesh_Block.cpp:1057-1077, control_refresh.cpp:229-232 -- i.e.
 * AFTER the periodic / neighbour ghost copies of the other axes).
 *   outflow     ghost layers copy the outermost active cell (for a field that
 *               is face-centred along `axis`: the boundary face)
 *   reflecting  ghost layers mirror the active zone; the vector component
 *               along `axis` (velocity_, bfield_, bfieldi_) changes sign
 * The layers span the full ghost-including extent of the other two axes.
 * Masks are not covered. */
enum { VLCT_BOUNDARY_OUTFLOW = 0, VLCT_BOUNDARY_REFLECTING = 1,
       VLCT_BOUNDARY_INFLOW = 2 /* vlct_boundary_inflow only */ };
int vlct_boundary(vlct_handle *h, const vlct_block *block, int axis, int side,
                  int type);

/* "inflow" boundary, BoundaryValue::enforce
 * (src/Cello/problem_BoundaryValue.cpp:131-273), for value-expressions that
 * are constants (as in input/vlct/dual_energy_cloud/initial_cloud_HD.in:77-95):
 * the g ghost layers of one face of the domain are set to a value, for the
 * fields in the boundary's field list only. For a field that is face-centred
 * along `axis` these are its g outermost layers; the boundary face itself is
 * not touched (ix0 = 0 resp. ndx - gx with ndx = nx + 2 gx + 1, cpp:190-202).
 * A field is in the list when its pointer in `block` is non-NULL and its entry
 * in `values` is not a NaN. */
typedef struct vlct_inflow_values {
  double density;
  double velocity_x, velocity_y, velocity_z;
  double total_energy, internal_energy;
  double bfield_x, bfield_y, bfield_z;
  double bfieldi_x, bfieldi_y, bfieldi_z;
  double pressure;
  double passive[VLCT_MAX_PASSIVE];
} vlct_inflow_values;
int vlct_boundary_inflow(vlct_handle *h, const vlct_block *block, int axis,
                         int side, const vlct_inflow_values *values);

/* Pack / unpack the ghost-exchange slab of all fields along one axis
 * (axis 0,1,2 = x,y,z; side 0 = lower, 1 = upper) into / from a contiguous
 * device buffer: the payload of one MsgRefresh FieldFace
 * (src/Cello/control_refresh.cpp:331-359). vlct_halo_bytes gives its size.
 * "send" slabs are the g outermost ACTIVE layers; "recv" slabs are the ghost
 * layers. Slabs span the full (ghost-including) extent of the other axes so
 * that exchanging x, then y, then z also fills edges and corners. */
long long vlct_halo_bytes(const vlct_handle *h, const vlct_block *block, int axis);
int vlct_halo_pack(vlct_handle *h, const vlct_block *block, int axis, int side,
                   double *buffer);
int vlct_halo_unpack(vlct_handle *h, const vlct_block *block, int axis,
                     int side, const double *buffer);

#ifdef __cplusplus
}
#endif
#endif /* VLCT_H */
