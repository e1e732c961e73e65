// See EnzoMethodMHDVlctGpu.hpp. Mirrors EnzoMethodMHDVlct.cpp of the reference
// step by step; citations are to src/Enzo/hydro-mhd/EnzoMethodMHDVlct.cpp.
#include "Cello/cello.hpp"
#include "Enzo/enzo.hpp"
#include "EnzoMethodMHDVlctGpu.hpp"

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
    store_fluxes_for_corrections_(store_fluxes_for_corrections)
{
  vlct_config_init(&config_);

  // Method:mhd_vlct:* (cpp:38-101): forward whatever the user wrote; the
  // library applies the reference's defaults and error messages
  static const char* const keys[] = {
    "mhd_choice", "riemann_solver", "reconstruct_method", "theta_limiter",
    "time_scheme", "courant",
    "half_dt_reconstruct_method", "full_dt_reconstruct_method" };
  for (const char* key : keys) {
    const std::string* val = p.param(key);
    if (val != nullptr) {
      set_key_(&config_, (std::string("Method:mhd_vlct:") + key).c_str(), *val);
    }
  }

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
}

//----------------------------------------------------------------------

EnzoMethodMHDVlctGpu::~EnzoMethodMHDVlctGpu()
{
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
  if (p.isUnpacking()) create_handle_();
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

void EnzoMethodMHDVlctGpu::compute(Block* block) throw()
{
  if (store_fluxes_for_corrections_) allocate_FC_flux_buffer_(block);   // cpp:362
  if (block->is_leaf()) {           // cpp:364
    vlct_block b;
    bind_block_(block, &b);
    const int rc = vlct_compute(handle_, &b, block->dt());
    check_status_(rc, handle_, "EnzoMethodMHDVlctGpu::compute");
    if (store_fluxes_for_corrections_) save_fluxes_for_corrections_(block, b);  // cpp:480-490
  }
  block->compute_done();            // cpp:499
}

//----------------------------------------------------------------------

double EnzoMethodMHDVlctGpu::timestep(Block* block) throw()
{
  vlct_block b;
  bind_block_(block, &b);
  double dt = 0.;
  const int rc = vlct_timestep(handle_, &b, &dt);
  check_status_(rc, handle_, "EnzoMethodMHDVlctGpu::timestep");
  return dt;   // already multiplied by courant (cpp:585-587)
}
