// TEST INFRASTRUCTURE ONLY -- not part of the product.
//
// Runtime half of the shim (written for this repository): definitions of the
// cello:: / enzo:: accessors declared in the shim headers, and a small C ABI
// (vlct_ref_*) that drives the reference's *own* EnzoMethodMHDVlct object --
// compiled unmodified from /root/reference/src -- on caller-provided arrays.
//
// The result, oracle/_ref/libvlct_ref.so, is the ground truth the C
// restatement in oracle/vlct_oracle.c is validated against, and the
// "reference" CPU baseline that bench.py times. The product never links it.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "Cello/cello.hpp"
#include "Enzo/enzo.hpp"
#include "Enzo/hydro-mhd/hydro-mhd.hpp"
#ifndef VLCT_SHIM_GPU_ADAPTER
#include "Enzo/initial/initial.hpp"     // the reference's EnzoInitialCloud
#endif

#include "../../include/vlct.h"

// Two flavours of this file are built:
//   default                  drives the reference's own EnzoMethodMHDVlct
//                            (libvlct_ref.so, entry points vlct_ref_*)
//   -DVLCT_SHIM_GPU_ADAPTER  drives integration/EnzoMethodMHDVlctGpu -- the
//                            reference-side binding of the CUDA library --
//                            through the very same Block/Field stand-ins
//                            (libvlct_adapter.so, entry points vlct_adapter_*)
#ifdef VLCT_SHIM_GPU_ADAPTER
#include "../../integration/EnzoMethodMHDVlctGpu.hpp"
typedef EnzoMethodMHDVlctGpu ShimMethod;
#define SHIM_FN(name) vlct_adapter_##name
#else
typedef EnzoMethodMHDVlct ShimMethod;
#define SHIM_FN(name) vlct_ref_##name
#endif

//----------------------------------------------------------------------
// global state the reference reaches through cello:: / enzo:: accessors
//----------------------------------------------------------------------

namespace {
  FieldDescr* g_field_descr = nullptr;
  EnzoPhysicsFluidProps* g_fluid_props = nullptr;
  Simulation g_simulation;
  Refresh g_refresh;
  Monitor g_monitor;
  Problem g_problem;
}

namespace cello {
  void message(FILE* fp, const char* type, const char* file, int line,
               const char* function, const char* message, ...)
  {
    va_list args;
    va_start(args, message);
    fprintf(fp, "[vlct_ref] %s %s:%d %s: ", type, file, line, function);
    vfprintf(fp, message, args);
    fprintf(fp, "\n");
    va_end(args);
    fflush(fp);
  }
  [[noreturn]] void error() { fflush(stdout); fflush(stderr); abort(); }

  int rank() { return 3; }
  FieldDescr* field_descr() { return g_field_descr; }
  Simulation* simulation() { return &g_simulation; }
  Refresh* refresh(int) { return &g_refresh; }
  Monitor* monitor() { return &g_monitor; }
  // returning false skips EnzoMethodMHDVlct::post_init_checks_, which only
  // inspects other Methods of a full simulation (gravity, flux_correct, ...)
  bool is_initial_cycle(InitCycleKind) noexcept { return false; }
}

namespace enzo {
  EnzoBlock* block(Block* block) { return static_cast<EnzoBlock*>(block); }
  Problem* problem() { return &g_problem; }
  EnzoPhysicsCosmology* cosmology() { return nullptr; }
  EnzoPhysicsFluidProps* fluid_props() { return g_fluid_props; }
  const EnzoMethodGrackle* grackle_method() { return nullptr; }
  const GrackleChemistryData* grackle_chemistry() { return nullptr; }
  double grav_constant_codeU() noexcept { return 1.0; }
}

//----------------------------------------------------------------------

namespace {

// The reference reaches its field descriptor and fluid properties through
// process-wide accessors. Handles created with the same configuration share one
// context, so several host threads (one handle each) can run concurrently
// without ever rewriting those globals (bench.py's CPU baseline does that).
struct RefContext {
  vlct_config cfg;
  int g[3];
  FieldDescr descr;
  EnzoPhysicsFluidProps* fluid_props;
  std::vector<std::string> passive_names;
  int refcount;
};

struct RefHandle {
  RefContext* ctx;
  ShimMethod* method;
  // blocks of the *_many entry points: they live as long as the handle, like
  // Cello's blocks live across cycles (the adapter keys its cache on Block*)
  std::vector<std::unique_ptr<EnzoBlock>> blocks;
  // flux-correction output of the last compute: [field slot][axis][face]
  bool store_fluxes = false;
  std::vector<double> saved[6 + VLCT_MAX_PASSIVE][3][2];
};

std::vector<RefContext*> g_contexts;
std::mutex g_mutex;

void activate(RefHandle* h) {
  if (g_field_descr != &h->ctx->descr) g_field_descr = &h->ctx->descr;
  if (g_fluid_props != h->ctx->fluid_props) g_fluid_props = h->ctx->fluid_props;
}

std::string fmt_double(double v) {
  char buf[64];
  snprintf(buf, sizeof(buf), "%.17g", v);
  return buf;
}

const char* riemann_name(int v) {
  switch (v) {
  case VLCT_RIEMANN_HLL:  return "hll";
  case VLCT_RIEMANN_HLLE: return "hlle";
  case VLCT_RIEMANN_HLLC: return "hllc";
  default:                return "hlld";
  }
}
const char* recon_name(int v) {
  switch (v) {
  case VLCT_RECON_NN:         return "nn";
  case VLCT_RECON_PLM_ATHENA: return "plm_athena";
  default:                    return "plm";
  }
}

// attach the caller's arrays to a shim block
void bind_block(RefHandle* h, EnzoBlock& blk, const vlct_block* b) {
  FieldData& fd = blk.data()->field_data;
  fd.nx = b->nx; fd.ny = b->ny; fd.nz = b->nz;
  fd.ptrs.assign(h->ctx->descr.field_count(), nullptr);
  auto set = [&](const char* name, double* p) {
    int id = h->ctx->descr.field_id(name);
    if (id >= 0) {
      if (p == nullptr) {
        fprintf(stderr, "[vlct_ref] missing pointer for field %s\n", name);
        abort();
      }
      fd.ptrs[id] = p;
    }
  };
  set("density", b->density);
  set("velocity_x", b->velocity_x);
  set("velocity_y", b->velocity_y);
  set("velocity_z", b->velocity_z);
  set("total_energy", b->total_energy);
  set("internal_energy", b->internal_energy);
  set("bfield_x", b->bfield_x);
  set("bfield_y", b->bfield_y);
  set("bfield_z", b->bfield_z);
  set("bfieldi_x", b->bfieldi_x);
  set("bfieldi_y", b->bfieldi_y);
  set("bfieldi_z", b->bfieldi_z);
  set("pressure", b->pressure);
  set("acceleration_x", b->acceleration_x);
  set("acceleration_y", b->acceleration_y);
  set("acceleration_z", b->acceleration_z);
  for (std::size_t i = 0; i < h->ctx->passive_names.size(); i++)
    set(h->ctx->passive_names[i].c_str(), b->passive[i]);
  blk.CellWidth[0] = b->dx; blk.CellWidth[1] = b->dy; blk.CellWidth[2] = b->dz;
  blk.data()->h[0] = b->dx; blk.data()->h[1] = b->dy; blk.data()->h[2] = b->dz;
}

} // namespace

//----------------------------------------------------------------------

extern "C" {

static void* create_(const vlct_config* cfg, int gx, int gy, int gz,
                     bool store_fluxes, int gpu_batch = -1, int gpu_fused = -1)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  RefContext* c = nullptr;
  for (RefContext* cand : g_contexts) {
    if (memcmp(&cand->cfg, cfg, sizeof(vlct_config)) == 0 &&
        cand->g[0] == gx && cand->g[1] == gy && cand->g[2] == gz) { c = cand; break; }
  }
  if (c == nullptr) {
    c = new RefContext;
    memcpy(&c->cfg, cfg, sizeof(vlct_config));
    c->g[0] = gx; c->g[1] = gy; c->g[2] = gz;
    c->refcount = 0;
    const bool mhd = (cfg->mhd_choice == VLCT_MHD_CONSTRAINED_TRANSPORT);
    const bool de = (cfg->dual_energy != VLCT_DE_DISABLED);
    c->descr.set_ghost_depth(gx, gy, gz);
    c->descr.insert("density", 0,0,0);
    c->descr.insert("velocity_x", 0,0,0);
    c->descr.insert("velocity_y", 0,0,0);
    c->descr.insert("velocity_z", 0,0,0);
    c->descr.insert("total_energy", 0,0,0);
    if (de) c->descr.insert("internal_energy", 0,0,0);
    if (mhd) {
      c->descr.insert("bfield_x", 0,0,0);
      c->descr.insert("bfield_y", 0,0,0);
      c->descr.insert("bfield_z", 0,0,0);
      c->descr.insert("bfieldi_x", 1,0,0);
      c->descr.insert("bfieldi_y", 0,1,0);
      c->descr.insert("bfieldi_z", 0,0,1);
    }
    c->descr.insert("pressure", 0,0,0);
    if (cfg->has_acceleration) {
      c->descr.insert("acceleration_x", 0,0,0);
      c->descr.insert("acceleration_y", 0,0,0);
      c->descr.insert("acceleration_z", 0,0,0);
    }
    for (int i = 0; i < cfg->n_passive; i++) {
      char name[32];
      snprintf(name, sizeof(name), "passive_%d", i);
      c->passive_names.push_back(name);
      c->descr.insert(name, 0,0,0);
      c->descr.groups()->add(name, "color");
    }
    // what an input file lists in Group "conserved" for a VL+CT run with
    // Method "flux_correct" (only looked at when fluxes are stored)
    c->descr.groups()->add("density", "conserved");
    c->descr.groups()->add("velocity_x", "conserved");
    c->descr.groups()->add("velocity_y", "conserved");
    c->descr.groups()->add("velocity_z", "conserved");
    c->descr.groups()->add("total_energy", "conserved");
    if (de) c->descr.groups()->add("internal_energy", "conserved");
    for (const std::string& name : c->passive_names)
      c->descr.groups()->add(name, "conserved");
    EnzoDualEnergyConfig de_config = EnzoDualEnergyConfig::build_disabled();
    if (cfg->dual_energy == VLCT_DE_MODERN) {
      de_config = EnzoDualEnergyConfig::build_modern_formulation
        (cfg->dual_energy_eta);
    } else if (cfg->dual_energy == VLCT_DE_BRYAN95) {
      de_config = EnzoDualEnergyConfig::build_bryan95_formulation
        (cfg->dual_energy_eta, cfg->dual_energy_eta);
    }
    EnzoFluidFloorConfig floors(cfg->density_floor, cfg->pressure_floor, 0., 0.);
    EnzoEOSVariant eos(EnzoEOSIdeal::construct(cfg->gamma));
    c->fluid_props = new EnzoPhysicsFluidProps(de_config, floors, eos, 0.6);
    g_contexts.push_back(c);
  }
  c->refcount++;

  RefHandle* h = new RefHandle;
  h->ctx = c;
  activate(h);

  const bool mhd = (cfg->mhd_choice == VLCT_MHD_CONSTRAINED_TRANSPORT);
  ParameterGroup p;
  p.set("riemann_solver", riemann_name(cfg->riemann_solver));
  p.set("reconstruct_method", recon_name(cfg->reconstruct_method));
  p.set("theta_limiter", fmt_double(cfg->theta_limiter));
  p.set("time_scheme", cfg->time_scheme == VLCT_TIME_EULER ? "euler" : "vl");
  if (cfg->mhd_choice != VLCT_MHD_UNSET)
    p.set("mhd_choice", mhd ? "constrained_transport" : "no_bfield");
  if (cfg->courant >= 0) p.set("courant", fmt_double(cfg->courant));
  // keys of the GPU binding only (integration/EnzoMethodMHDVlctGpu.hpp)
  if (gpu_batch >= 0) p.set("gpu_batch_blocks", gpu_batch ? "true" : "false");
  if (gpu_fused >= 0) p.set("gpu_fused_timestep", gpu_fused ? "true" : "false");

  h->store_fluxes = store_fluxes;
  h->method = new ShimMethod(p, store_fluxes);
  return h;
}

void* SHIM_FN(create)(const vlct_config* cfg, int gx, int gy, int gz)
{ return create_(cfg, gx, gy, gz, false); }

/// the same Method constructed with store_fluxes_for_corrections = true
void* SHIM_FN(create_fc)(const vlct_config* cfg, int gx, int gy, int gz)
{ return create_(cfg, gx, gy, gz, true); }

void SHIM_FN(destroy)(void* handle)
{
  RefHandle* h = static_cast<RefHandle*>(handle);
  if (h == nullptr) return;
  std::lock_guard<std::mutex> lock(g_mutex);
  activate(h);
  delete h->method;
  RefContext* c = h->ctx;
  if (--c->refcount == 0) {
    if (g_field_descr == &c->descr) g_field_descr = nullptr;
    if (g_fluid_props == c->fluid_props) g_fluid_props = nullptr;
    delete c->fluid_props;
    for (std::size_t i = 0; i < g_contexts.size(); i++)
      if (g_contexts[i] == c) { g_contexts.erase(g_contexts.begin() + i); break; }
    delete c;
  }
  delete h;
}

int SHIM_FN(compute)(void* handle, const vlct_block* b, double dt)
{
  RefHandle* h = static_cast<RefHandle*>(handle);
  activate(h);
  EnzoBlock blk(&h->ctx->descr);
  bind_block(h, blk, b);
  blk.set_dt(dt);
  h->method->compute(&blk);
  if (h->store_fluxes) {
    // keep what save_fluxes_for_corrections_ deposited in the block's FluxData
    FluxData* fd = blk.data()->flux_data();
    Field field = blk.data()->field();
    for (int i_f = 0; i_f < fd->num_fields(); i_f++) {
      const std::string name = field.field_name(fd->index_field(i_f));
      int slot = -1;
      const char* fixed[6] = { "density", "velocity_x", "velocity_y", "velocity_z",
                               "total_energy", "internal_energy" };
      for (int k = 0; k < 6; k++) if (name == fixed[k]) slot = k;
      for (std::size_t k = 0; k < h->ctx->passive_names.size(); k++)
        if (name == h->ctx->passive_names[k]) slot = 6 + (int) k;
      if (slot < 0) continue;
      for (int axis = 0; axis < 3; axis++)
        for (int face = 0; face < 2; face++) {
          FaceFluxes* ff = fd->block_fluxes(axis, face, i_f);
          const int m = ff->get_size();
          double* p = ff->flux_array();
          h->saved[slot][axis][face].assign(p, p + m);
        }
    }
  }
  return (blk.compute_done_count == 1) ? 0 : 1;
}

/// out[axis][face][slot] (NULL = skip) <- the face fluxes of the last compute
int SHIM_FN(face_fluxes)(void* handle, double* const out[3][2][6 + VLCT_MAX_PASSIVE])
{
  RefHandle* h = static_cast<RefHandle*>(handle);
  if (!h->store_fluxes) return 1;
  for (int axis = 0; axis < 3; axis++)
    for (int face = 0; face < 2; face++)
      for (int slot = 0; slot < 6 + VLCT_MAX_PASSIVE; slot++) {
        double* dst = out[axis][face][slot];
        if (dst == nullptr) continue;
        const std::vector<double>& src = h->saved[slot][axis][face];
        if (src.empty()) return 2;
        memcpy(dst, src.data(), src.size() * sizeof(double));
      }
  return 0;
}

int SHIM_FN(timestep)(void* handle, const vlct_block* b, double* dt_out)
{
  RefHandle* h = static_cast<RefHandle*>(handle);
  activate(h);
  EnzoBlock blk(&h->ctx->descr);
  bind_block(h, blk, b);
  *dt_out = h->method->timestep(&blk);
  return 0;
}

#ifndef VLCT_SHIM_GPU_ADAPTER
/// The reference's own EnzoInitialCloud::enforce_block
/// (src/Enzo/initial/EnzoInitialCloud.cpp:606-748, compiled unmodified) on the
/// caller's arrays. lower: coordinates of the block's first active cell;
/// p[] = { cloud_radius, center_x, center_y, center_z, cloud_density,
///         wind_density, wind_velocity, wind_total_energy, wind_internal_energy }
/// (the Initial:cloud parameters of input/vlct/dual_energy_cloud).
int vlct_ref_ic_cloud_perturbed(void* handle, const vlct_block* b, const double* lower,
                                int subsample_n, const double* p, int nwaves,
                                unsigned int seed, double amplitude, double min_lambda,
                                double max_lambda);

int vlct_ref_ic_cloud(void* handle, const vlct_block* b, const double* lower,
                      int subsample_n, const double* p)
{ return vlct_ref_ic_cloud_perturbed(handle, b, lower, subsample_n, p, 0, 0u, 0., 0., 0.); }

/// ... with Initial:cloud:perturb_Nwaves / perturb_seed / perturb_amplitude /
/// perturb_min_lambda / perturb_max_lambda (EnzoInitialCloud.hpp:39-58)
int vlct_ref_ic_cloud_perturbed(void* handle, const vlct_block* b, const double* lower,
                                int subsample_n, const double* p, int nwaves,
                                unsigned int seed, double amplitude, double min_lambda,
                                double max_lambda)
{
  RefHandle* h = static_cast<RefHandle*>(handle);
  activate(h);
  EnzoBlock blk(&h->ctx->descr);
  bind_block(h, blk, b);
  for (int a = 0; a < 3; a++) blk.data()->xm[a] = lower[a];
  ParameterGroup pg;
  pg.set("subsample_n", std::to_string(subsample_n));
  const char* keys[9] = { "cloud_radius", "cloud_center_x", "cloud_center_y",
                          "cloud_center_z", "cloud_density", "wind_density",
                          "wind_velocity", "wind_total_energy",
                          "wind_internal_energy" };
  for (int k = 0; k < 9; k++) pg.set(keys[k], fmt_double(p[k]));
  if (nwaves > 0) {
    pg.set("perturb_Nwaves", std::to_string(nwaves));
    pg.set("perturb_seed", std::to_string(seed));
    pg.set("perturb_amplitude", fmt_double(amplitude));
    pg.set("perturb_min_lambda", fmt_double(min_lambda));
    pg.set("perturb_max_lambda", fmt_double(max_lambda));
  }
  EnzoInitialCloud initial(0, 0.0, pg);
  initial.enforce_block(&blk, nullptr);
  return 0;
}

/// The reference's own EnzoInitialShockTube::enforce_block
/// (src/Enzo/initial/EnzoInitialShockTube.cpp, compiled unmodified); gamma comes
/// from the handle's fluid properties like in the reference.
int vlct_ref_ic_shock_tube(void* handle, const vlct_block* b, const double* lower,
                           const char* setup, int aligned_ax, double axis_velocity,
                           double trans_velocity, int flipped)
{
  RefHandle* h = static_cast<RefHandle*>(handle);
  activate(h);
  EnzoBlock blk(&h->ctx->descr);
  bind_block(h, blk, b);
  for (int a = 0; a < 3; a++) blk.data()->xm[a] = lower[a];
  ParameterGroup pg;
  pg.set("setup_name", setup);
  pg.set("aligned_ax", aligned_ax == 0 ? "x" : (aligned_ax == 1 ? "y" : "z"));
  pg.set("axis_velocity", fmt_double(axis_velocity));
  pg.set("transverse_velocity", fmt_double(trans_velocity));
  pg.set("flip_initialize", flipped ? "true" : "false");
  EnzoInitialShockTube initial(0, 0.0, pg);
  initial.enforce_block(&blk, nullptr);
  return 0;
}

/// The reference's own EnzoInitialInclinedWave::enforce_block
/// (src/Enzo/initial/EnzoInitialInclinedWave.cpp, compiled unmodified).
int vlct_ref_ic_inclined_wave(void* handle, const vlct_block* b, const double* lower,
                              const char* wave_type, double alpha, double beta,
                              double amplitude, double lambda, int positive_vel)
{
  RefHandle* h = static_cast<RefHandle*>(handle);
  activate(h);
  EnzoBlock blk(&h->ctx->descr);
  bind_block(h, blk, b);
  for (int a = 0; a < 3; a++) blk.data()->xm[a] = lower[a];
  ParameterGroup pg;
  pg.set("wave_type", wave_type);
  pg.set("alpha", fmt_double(alpha));
  pg.set("beta", fmt_double(beta));
  pg.set("amplitude", fmt_double(amplitude));
  pg.set("lambda", fmt_double(lambda));
  pg.set("positive_vel", positive_vel ? "true" : "false");
  EnzoInitialInclinedWave initial(0, 0.0, pg);
  initial.enforce_block(&blk, nullptr);
  return 0;
}

/// The reference's own EnzoBoundary::enforce (src/Enzo/enzo-core/EnzoBoundary.cpp,
/// compiled unmodified) for one face of the domain, on every field of the
/// block: type 0 = "outflow", 1 = "reflecting" (the numbering of vlct.h).
int vlct_ref_boundary(void* handle, const vlct_block* b, int axis, int side, int type)
{
  RefHandle* h = static_cast<RefHandle*>(handle);
  activate(h);
  EnzoBlock blk(&h->ctx->descr);
  bind_block(h, blk, b);
  EnzoBoundary boundary(axis_all, face_all, nullptr,
                        type == VLCT_BOUNDARY_OUTFLOW ? boundary_type_outflow
                                                      : boundary_type_reflecting);
  boundary.enforce(&blk, side == 0 ? face_lower : face_upper, (axis_enum) axis);
  return 0;
}
#endif

#ifdef VLCT_SHIM_GPU_ADAPTER
/// the adapter constructed with its own two keys set
void* vlct_adapter_create_opts(const vlct_config* cfg, int gx, int gy, int gz,
                               int gpu_batch_blocks, int gpu_fused_timestep)
{ return create_(cfg, gx, gy, gz, false, gpu_batch_blocks, gpu_fused_timestep); }

static void bind_many(RefHandle* h, const vlct_block* blocks, int n, int cycle)
{
  while ((int) h->blocks.size() < n)
    h->blocks.emplace_back(new EnzoBlock(&h->ctx->descr));
  for (int i = 0; i < n; i++) {
    bind_block(h, *h->blocks[i], &blocks[i]);
    h->blocks[i]->set_cycle(cycle);
  }
  g_simulation.hierarchy()->set_num_blocks((size_t) n);
}

/// The compute phase of one cycle the way Cello runs it on a process with n
/// blocks: Method::compute(block) for one block after the other
/// (src/Cello/control_compute.cpp:72-112). Returns 0 if every block reported
/// compute_done() exactly once by the end; *deferred = how many blocks had NOT
/// reported it when the last call began (n - 1 when the adapter batches).
int vlct_adapter_compute_many(void* handle, const vlct_block* blocks, int n, double dt,
                              int cycle, int* deferred)
{
  RefHandle* h = static_cast<RefHandle*>(handle);
  activate(h);
  bind_many(h, blocks, n, cycle);
  for (int i = 0; i < n; i++) {
    h->blocks[i]->set_dt(dt);
    h->blocks[i]->compute_done_count = 0;
  }
  int not_done = 0;
  for (int i = 0; i < n; i++) {
    if (i == n - 1)
      for (int j = 0; j < n - 1; j++) not_done += (h->blocks[j]->compute_done_count == 0);
    h->method->compute(h->blocks[i].get());
  }
  if (deferred) *deferred = not_done;
  g_simulation.hierarchy()->set_num_blocks(1);
  for (int i = 0; i < n; i++)
    if (h->blocks[i]->compute_done_count != 1) return 1;
  return h->method->queued_blocks() == 0 ? 0 : 2;
}

/// Method::timestep on the same n blocks (stopping phase of cycle `cycle`)
int vlct_adapter_timestep_many(void* handle, const vlct_block* blocks, int n, int cycle,
                               double* dts)
{
  RefHandle* h = static_cast<RefHandle*>(handle);
  activate(h);
  bind_many(h, blocks, n, cycle);
  g_simulation.hierarchy()->set_num_blocks(1);
  for (int i = 0; i < n; i++) dts[i] = h->method->timestep(h->blocks[i].get());
  return 0;
}

/// Charm++ migration / checkpoint of the Method: size, pack, construct a new
/// object with the migration constructor, unpack into it, delete the old one
/// (EnzoMethodMHDVlct.cpp:170-197, hpp:102-112). Returns the packed size.
long long vlct_adapter_pup_roundtrip(void* handle)
{
  RefHandle* h = static_cast<RefHandle*>(handle);
  activate(h);
  std::vector<char> buffer;
  PUP::er sizer(PUP::er::SIZING, &buffer);
  h->method->pup(sizer);
  PUP::er packer(PUP::er::PACKING, &buffer);
  h->method->pup(packer);
  if (buffer.size() != sizer.size()) return -1;
  CkMigrateMessage msg;
  ShimMethod* fresh = new ShimMethod(&msg);
  PUP::er unpacker(PUP::er::UNPACKING, &buffer);
  fresh->pup(unpacker);
  if (unpacker.size() != buffer.size()) { delete fresh; return -2; }
  delete h->method;
  h->method = fresh;
  return (long long) buffer.size();
}
#endif

const char* SHIM_FN(name)(void* handle)
{
  static std::string name;
  name = static_cast<RefHandle*>(handle)->method->name();
  return name.c_str();
}

} // extern "C"
