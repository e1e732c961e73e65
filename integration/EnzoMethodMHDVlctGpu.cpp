// See EnzoMethodMHDVlctGpu.hpp. Mirrors EnzoMethodMHDVlct.cpp of the reference
// step by step; citations are to src/Enzo/hydro-mhd/EnzoMethodMHDVlct.cpp.
#include "Cello/cello.hpp"
#include "Enzo/enzo.hpp"
#include "EnzoMethodMHDVlctGpu.hpp"

#include <cstdio>
#include <cstring>

//----------------------------------------------------------------------

static void check_status_(int rc, const vlct_handle* h, const char* where)
{
  // the reference aborts through ERROR (src/Cello/error_Error.hpp:52-59)
  if (rc != VLCT_OK) {
    ERROR2(where, "libvlct_b200: %s (%s)", vlct_last_error(h),
           vlct_status_string(rc));
  }
}

static void set_key_(vlct_config* cfg, const char* key, const std::string& val)
{
  char err[512];
  if (vlct_config_set(cfg, key, val.c_str(), err, (int) sizeof(err)) != VLCT_OK) {
    ERROR1("EnzoMethodMHDVlctGpu", "%s", err);
  }
}

//----------------------------------------------------------------------

EnzoMethodMHDVlctGpu::EnzoMethodMHDVlctGpu(ParameterGroup p,
                                           bool store_fluxes_for_corrections)
  : Method(), handle_(nullptr), passive_names_(),
    store_fluxes_for_corrections_(store_fluxes_for_corrections),
    batch_blocks_(true), fused_timestep_(false)
{
  vlct_config_init(&config_);

  // Method:mhd_vlct:* (cpp:38-101): forward whatever the user wrote; the
  // library applies the reference's defaults and error messages. Like the
  // reference (cpp:42-43) a parameter counts as defined when
  // ParameterGroup::param() finds it; values are read with value_string /
  // value_float.
  auto is_defined = [&](const char* key) -> bool { return p.param(key) != nullptr; };
  static const char* const string_keys[] = {
    "mhd_choice", "riemann_solver", "reconstruct_method", "time_scheme",
    "half_dt_reconstruct_method", "full_dt_reconstruct_method" };
  for (const char* key : string_keys) {
    if (is_defined(key)) {
      set_key_(&config_, (std::string("Method:mhd_vlct:") + key).c_str(),
               p.value_string(key, ""));
    }
  }
  static const char* const float_keys[] = { "theta_limiter", "courant" };
  for (const char* key : float_keys) {
    if (is_defined(key)) {
      char buf[64];
      snprintf(buf, sizeof(buf), "%.17g", p.value_float(key, 0.));
      set_key_(&config_, (std::string("Method:mhd_vlct:") + key).c_str(), buf);
    }
  }
  // keys of this binding only (no counterpart in the reference)
  batch_blocks_ = p.value_logical("gpu_batch_blocks", true);
  fused_timestep_ = p.value_logical("gpu_fused_timestep", false);

  // Physics:fluid_props (what EnzoMHDIntegratorStageCommands reads through
  // enzo::fluid_props(), EnzoMHDIntegratorStageCommands.cpp:18-64)
  const EnzoPhysicsFluidProps* fluid_props = enzo::fluid_props();
  ASSERT("EnzoMethodMHDVlctGpu", "can't currently handle the case with a "
         "non-ideal EOS",
         fluid_props->eos_variant().holds_alternative<EnzoEOSIdeal>());
  config_.gamma = fluid_props->eos_variant().get<EnzoEOSIdeal>().get_gamma();
  const EnzoDualEnergyConfig& de = fluid_props->dual_energy_config();
  enzo_float eta = 0.;
  if (de.is_disabled()) {
    config_.dual_energy = VLCT_DE_DISABLED;
  } else if (de.modern_formulation(&eta)) {
    config_.dual_energy = VLCT_DE_MODERN;
    config_.dual_energy_eta = eta;
  } else {
    config_.dual_energy = VLCT_DE_BRYAN95;   // rejected by validation
  }
  const EnzoFluidFloorConfig& floors = fluid_props->fluid_floor_config();
  config_.density_floor = floors.has_density_floor() ? floors.density() : 0.;
  config_.pressure_floor = floors.has_pressure_floor() ? floors.pressure() : 0.;

  // passive scalars = all fields in group "color"
  // (toolkit/EnzoLazyPassiveScalarFieldList.cpp:15-31)
  FieldDescr* field_descr = cello::field_descr();
  const Grouping* groups = field_descr->groups();
  const int n_color = groups->size("color");
  ASSERT1("EnzoMethodMHDVlctGpu", "at most %d passive scalars are supported",
          VLCT_MAX_PASSIVE, n_color <= VLCT_MAX_PASSIVE);
  for (int i = 0; i < n_color; i++) passive_names_.push_back(groups->item("color", i));
  config_.n_passive = n_color;

  // gravity source terms only when the acceleration fields exist (cpp:219-232)
  config_.has_acceleration = field_descr->is_field("acceleration_x") ? 1 : 0;

  if (store_fluxes_for_corrections) {       // cpp:137-141
    ASSERT("EnzoMethodMHDVlctGpu",
           "Flux corrections are currently only supported in hydro-mode",
           config_.mhd_choice == VLCT_MHD_NO_BFIELD);
  }
  ASSERT("EnzoMethodMHDVlctGpu", "\"pressure\" must be a permanent field",
         field_descr->is_field("pressure"));

  // every field handed to the library is fp64 with one common ghost depth
  // (bind_block_ passes a single (gx, gy, gz) and raw double pointers)
  {
    static const char* const names[] = {
      "density", "velocity_x", "velocity_y", "velocity_z", "total_energy",
      "internal_energy", "bfield_x", "bfield_y", "bfield_z", "bfieldi_x",
      "bfieldi_y", "bfieldi_z", "pressure", "acceleration_x", "acceleration_y",
      "acceleration_z" };
    std::vector<std::string> all(names, names + sizeof(names) / sizeof(names[0]));
    all.insert(all.end(), passive_names_.begin(), passive_names_.end());
    int g0[3] = { -1, -1, -1 };
    for (const std::string& field_name : all) {
      if (!field_descr->is_field(field_name)) continue;
      const int id = field_descr->field_id(field_name);
      const int prec = field_descr->precision(id);
      ASSERT1("EnzoMethodMHDVlctGpu", "field \"%s\" must be double precision",
              field_name.c_str(),
              prec == precision_double ||
              (prec == precision_default && sizeof(enzo_float) == sizeof(double)));
      int g[3];
      field_descr->ghost_depth(id, &g[0], &g[1], &g[2]);
      if (g0[0] < 0) { g0[0] = g[0]; g0[1] = g[1]; g0[2] = g[2]; }
      ASSERT1("EnzoMethodMHDVlctGpu",
              "field \"%s\" must have the ghost depth of \"density\"",
              field_name.c_str(), g[0] == g0[0] && g[1] == g0[1] && g[2] == g0[2]);
    }
  }

  create_handle_();
  this->set_courant(config_.courant < 0
                    ? (config_.time_scheme == VLCT_TIME_VL ? 0.3 : 1.0)
                    : config_.courant);

  // default Refresh: all fields, like the reference (cpp:141-151)
  cello::simulation()->refresh_set_name(ir_post_, name());
  Refresh* refresh = cello::refresh(ir_post_);
  refresh->add_all_fields();
}

//----------------------------------------------------------------------

void EnzoMethodMHDVlctGpu::create_handle_()
{
  const int rc = vlct_create(&config_, &handle_);
  check_status_(rc, handle_, "EnzoMethodMHDVlctGpu");
  if (store_fluxes_for_corrections_ && config_.n_passive > 0) {
    // FluxData wants the passive scalars' face fluxes, too
    check_status_(vlct_set_option(handle_, "scalar_flux_arrays", 1), handle_,
                  "EnzoMethodMHDVlctGpu");
  }
}

//----------------------------------------------------------------------

EnzoMethodMHDVlctGpu::~EnzoMethodMHDVlctGpu()
{
  if (!queue_.empty()) {
    WARNING1("EnzoMethodMHDVlctGpu::~EnzoMethodMHDVlctGpu",
             "%d queued blocks were never advanced", (int) queue_.size());
  }
  vlct_destroy(handle_);
}

//----------------------------------------------------------------------

void EnzoMethodMHDVlctGpu::pup(PUP::er& p)
{
  Method::pup(p);
  // vlct_config is plain data; scratch space is never serialised (cpp:170-197)
  PUParray(p, reinterpret_cast<char*>(&config_), sizeof(vlct_config));
  p | passive_names_;
  p | store_fluxes_for_corrections_;
  p | batch_blocks_;
  p | fused_timestep_;
  // the queue and the cached timesteps live within one compute phase resp.
  // one cycle; Charm++ migrates / checkpoints Methods between cycles
  ASSERT("EnzoMethodMHDVlctGpu::pup", "pup() with blocks still queued",
         queue_.empty());
  if (p.isUnpacking()) {
    cached_dt_.clear();
    create_handle_();          // the library handle (device scratch) is rebuilt
  }
}

//----------------------------------------------------------------------

void EnzoMethodMHDVlctGpu::bind_block_(Block* block, vlct_block* out) noexcept
{
  memset(out, 0, sizeof(*out));
  Field field = block->data()->field();
  field.size(&out->nx, &out->ny, &out->nz);
  field.ghost_depth(field.field_id("density"), &out->gx, &out->gy, &out->gz);
  EnzoBlock* enzo_block = enzo::block(block);
  out->dx = enzo_block->CellWidth[0];
  out->dy = enzo_block->CellWidth[1];
  out->dz = enzo_block->CellWidth[2];
  auto ptr = [&](const char* name) -> double* {
    return field.is_field(name) ? (double*) field.values(name) : nullptr;
  };
  out->density = ptr("density");
  out->velocity_x = ptr("velocity_x");
  out->velocity_y = ptr("velocity_y");
  out->velocity_z = ptr("velocity_z");
  out->total_energy = ptr("total_energy");
  out->internal_energy = ptr("internal_energy");
  out->bfield_x = ptr("bfield_x");
  out->bfield_y = ptr("bfield_y");
  out->bfield_z = ptr("bfield_z");
  out->bfieldi_x = ptr("bfieldi_x");
  out->bfieldi_y = ptr("bfieldi_y");
  out->bfieldi_z = ptr("bfieldi_z");
  out->pressure = ptr("pressure");
  out->acceleration_x = ptr("acceleration_x");
  out->acceleration_y = ptr("acceleration_y");
  out->acceleration_z = ptr("acceleration_z");
  for (std::size_t i = 0; i < passive_names_.size(); i++)
    out->passive[i] = ptr(passive_names_[i].c_str());
  // Cello's permanent fields live in host memory (data_FieldData.hpp:386-398)
  out->mem_space = VLCT_MEM_HOST;
  out->stream = nullptr;
}

//----------------------------------------------------------------------

/// allocate_FC_flux_buffer_ (cpp:332-352): one single-flux-array FluxData
/// entry per field of the "conserved" group, every cycle
static void allocate_FC_flux_buffer_(Block* block) throw()
{
  Field field = block->data()->field();
  auto field_names = field.groups()->group_list("conserved");
  const int nf = (int) field_names.size();
  std::vector<int> field_list(nf);
  for (int i = 0; i < nf; i++) field_list[i] = field.field_id(field_names[i]);
  int nx, ny, nz;
  field.size(&nx, &ny, &nz);
  block->data()->flux_data()->allocate(nx, ny, nz, field_list, true);
}

//----------------------------------------------------------------------

/// save_fluxes_for_corrections_ (cpp:250-330): the library returns
/// dt/dx * (final-stage flux) on the block's six faces as packed arrays whose
/// layout is FaceFluxes' own (ix + mx*(iy + my*iz) with the normal extent 1)
void EnzoMethodMHDVlctGpu::save_fluxes_for_corrections_(Block* block,
                                                        const vlct_block& b) noexcept
{
  Field field = block->data()->field();
  FluxData* flux_data = block->data()->flux_data();
  const int nf = flux_data->num_fields();

  vlct_face_fluxes ff;
  memset(&ff, 0, sizeof(ff));
  ff.mem_space = VLCT_MEM_HOST;
  struct Target { FaceFluxes* face; std::vector<double> staging; };
  std::vector<Target> targets;
  targets.reserve((std::size_t) nf * 6);       // pointers into it stay valid

  for (int i_f = 0; i_f < nf; i_f++) {
    const std::string field_name = field.field_name(flux_data->index_field(i_f));
    int slot = -1;
    if (field_name == "density") slot = 0;
    else if (field_name == "velocity_x") slot = 1;
    else if (field_name == "velocity_y") slot = 2;
    else if (field_name == "velocity_z") slot = 3;
    else if (field_name == "total_energy") slot = 4;
    else if (field_name == "internal_energy") slot = 5;
    for (std::size_t k = 0; k < passive_names_.size(); k++)
      if (field_name == passive_names_[k]) slot = 6 + (int) k;
    if (slot < 0) {
      ERROR1("EnzoMethodMHDVlctGpu::save_fluxes_for_corrections_",
             "no flux is computed for the conserved field \"%s\"",
             field_name.c_str());
    }
    for (int dim = 0; dim < 3; dim++) {
      for (int side = 0; side < 2; side++) {
        FaceFluxes* face = flux_data->block_fluxes(dim, side, i_f);
        int mx, my, mz;
        face->get_size(&mx, &my, &mz);
        targets.push_back(Target{ face, std::vector<double>((std::size_t) mx * my * mz) });
        ff.face[dim][side][slot] = targets.back().staging.data();
      }
    }
  }
  const int rc = vlct_save_face_fluxes(handle_, &b, &ff);
  check_status_(rc, handle_, "EnzoMethodMHDVlctGpu::save_fluxes_for_corrections_");
  for (Target& t : targets) {
    enzo_float* dest = t.face->flux_array();
    for (std::size_t i = 0; i < t.staging.size(); i++) dest[i] = (enzo_float) t.staging[i];
  }
}

//----------------------------------------------------------------------

void EnzoMethodMHDVlctGpu::compute_one_(Block* block) throw()
{
  vlct_block b;
  bind_block_(block, &b);
  int rc;
  if (fused_timestep_ && !store_fluxes_for_corrections_) {
    double dt_next = 0.;
    rc = vlct_compute_and_timestep(handle_, &b, block->dt(), &dt_next);
    for (std::size_t i = 0; i < cached_dt_.size();) {       // one entry per block
      if (cached_dt_[i].block == block) cached_dt_.erase(cached_dt_.begin() + i);
      else i++;
    }
    cached_dt_.push_back(CachedDt{ block, block->cycle() + 1, dt_next });
  } else {
    rc = vlct_compute(handle_, &b, block->dt());
  }
  check_status_(rc, handle_, "EnzoMethodMHDVlctGpu::compute");
  if (store_fluxes_for_corrections_) save_fluxes_for_corrections_(block, b);  // cpp:480-490
}

//----------------------------------------------------------------------

void EnzoMethodMHDVlctGpu::compute(Block* block) throw()
{
  if (store_fluxes_for_corrections_) allocate_FC_flux_buffer_(block);   // cpp:362

  // Cello calls compute() once per block of this process, block after block
  // (src/Cello/control_compute.cpp:72-112), and a block only moves on when it
  // reports compute_done(). Entry methods run to completion, so with several
  // blocks per process the calls can be gathered: the blocks are queued, the
  // call for the process's last block advances them all in ONE set of kernel
  // launches (vlct_compute_batch) and then every block reports compute_done().
  // No block needs another block's compute_done() to reach its own compute()
  // (the refresh that precedes compute() only needs the neighbours' data of the
  // previous phase), so nothing can deadlock on the delay.
  const std::size_t expected = cello::simulation()->hierarchy()->num_blocks();
  const bool batching = batch_blocks_ && expected > 1 && !store_fluxes_for_corrections_;
  if (!batching) {
    if (block->is_leaf()) compute_one_(block);       // cpp:364
    block->compute_done();                           // cpp:499
    return;
  }
  if (queue_.empty()) cached_dt_.clear();            // a new compute phase
  queue_.push_back(Queued{ block, block->is_leaf() });
  if (queue_.size() >= expected) flush_queue_();
}

//----------------------------------------------------------------------

void EnzoMethodMHDVlctGpu::flush_queue_() throw()
{
  // group the leaf blocks by (dt, cell widths): blocks of one refinement level
  std::vector<char> done(queue_.size(), 0);
  for (std::size_t first = 0; first < queue_.size(); first++) {
    if (done[first] || !queue_[first].leaf) continue;
    std::vector<vlct_block> batch;
    std::vector<Block*> members;
    vlct_block b0;
    bind_block_(queue_[first].block, &b0);
    const double dt0 = queue_[first].block->dt();
    for (std::size_t i = first; i < queue_.size(); i++) {
      if (done[i] || !queue_[i].leaf) continue;
      vlct_block b;
      bind_block_(queue_[i].block, &b);
      if (queue_[i].block->dt() != dt0 || b.dx != b0.dx || b.dy != b0.dy || b.dz != b0.dz)
        continue;
      batch.push_back(b);
      members.push_back(queue_[i].block);
      done[i] = 1;
    }
    int rc;
    if (fused_timestep_) {
      // the batch minimum serves every member: Cello min-reduces the blocks'
      // timesteps anyway (src/Cello/control_stopping.cpp:96-142)
      double dt_next = 0.;
      rc = vlct_compute_and_timestep_batch(handle_, batch.data(), (int) batch.size(),
                                           dt0, &dt_next);
      for (Block* member : members)
        cached_dt_.push_back(CachedDt{ member, member->cycle() + 1, dt_next });
    } else {
      rc = vlct_compute_batch(handle_, batch.data(), (int) batch.size(), dt0);
    }
    check_status_(rc, handle_, "EnzoMethodMHDVlctGpu::compute (batch)");
  }
  // every block of the process moves on (cpp:499), in the order Cello called
  std::vector<Queued> queued;
  queued.swap(queue_);          // compute_done() may re-enter compute()
  for (const Queued& q : queued) q.block->compute_done();
}

//----------------------------------------------------------------------

double EnzoMethodMHDVlctGpu::timestep(Block* block) throw()
{
  for (const CachedDt& c : cached_dt_)
    if (c.block == block && c.cycle == block->cycle()) return c.dt;
  vlct_block b;
  bind_block_(block, &b);
  double dt = 0.;
  const int rc = vlct_timestep(handle_, &b, &dt);
  check_status_(rc, handle_, "EnzoMethodMHDVlctGpu::timestep");
  return dt;   // already multiplied by courant (cpp:585-587)
}
