// vlct_config.cpp -- host-only part of the C ABI: parameter parsing and the
// validation the reference performs in its constructors. No CUDA here, so the
// "does the configuration make sense" logic can be unit-tested without a GPU.
//
// Mirrors (reference, src/Enzo/):
//   hydro-mhd/EnzoMethodMHDVlct.cpp:38-152          parameter keys / defaults
//   hydro-mhd/EnzoMHDIntegratorStageCommands.cpp:18-98   EOS/DE/floor checks
//   hydro-mhd/riemann/EnzoRiemann.cpp:26-68         valid solver/physics combos
//   hydro-mhd/toolkit/EnzoReconstructor.cpp:14-44   reconstructor names, theta
//   hydro-mhd/toolkit/EnzoBfieldMethod.cpp:14-27    CT needs two stages
//   enzo-core/EnzoConfig.cpp:952-1262               Physics:fluid_props keys
#include <cctype>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/vlct.h"

namespace {

void set_err(char* buf, int len, const char* fmt, ...)
{
  if (buf == nullptr || len <= 0) return;
  va_list args;
  va_start(args, fmt);
  vsnprintf(buf, (size_t) len, fmt, args);
  va_end(args);
}

std::string lower(const char* s)
{
  std::string out(s ? s : "");
  for (char& c : out) c = (char) tolower((unsigned char) c);
  // strip surrounding quotes / blanks, as a parameter-file value may carry them
  while (!out.empty() && (out.front() == '"' || isspace((unsigned char) out.front()))) out.erase(out.begin());
  while (!out.empty() && (out.back() == '"' || out.back() == ';' || isspace((unsigned char) out.back()))) out.pop_back();
  return out;
}

bool parse_double(const std::string& v, double* out)
{
  char* end = nullptr;
  double d = strtod(v.c_str(), &end);
  if (end == v.c_str()) return false;
  while (*end && isspace((unsigned char) *end)) end++;
  if (*end != '\0') return false;
  *out = d;
  return true;
}

bool parse_bool(const std::string& v, int* out)
{
  if (v == "true" || v == "1" || v == "yes") { *out = 1; return true; }
  if (v == "false" || v == "0" || v == "no") { *out = 0; return true; }
  return false;
}

bool ends_with(const std::string& s, const char* suffix)
{
  const size_t n = strlen(suffix);
  return s.size() >= n && s.compare(s.size() - n, n, suffix) == 0;
}

}  // namespace

extern "C" {

int vlct_config_init(vlct_config* cfg)
{
  if (cfg == nullptr) return VLCT_ERR_INVALID_CONFIG;
  cfg->riemann_solver = VLCT_RIEMANN_HLLD;          // cpp:80 "hlld"
  cfg->reconstruct_method = VLCT_RECON_PLM_ENZO;    // cpp:49,55 "plm"
  cfg->theta_limiter = 1.5;                         // cpp:82
  cfg->mhd_choice = VLCT_MHD_UNSET;                 // cpp:74-77 required
  cfg->time_scheme = VLCT_TIME_VL;                  // cpp:46
  cfg->courant = -1.0;                              // cpp:100-101 (0.3 / 1.0)
  cfg->gamma = 5.0 / 3.0;                           // EnzoConfig.cpp eos default
  cfg->dual_energy = VLCT_DE_DISABLED;
  cfg->dual_energy_eta = 0.001;
  cfg->density_floor = 0.0;
  cfg->pressure_floor = 0.0;
  cfg->n_passive = 0;
  cfg->has_acceleration = 0;
  return VLCT_OK;
}

int vlct_config_set(vlct_config* cfg, const char* key_, const char* value_,
                    char* errbuf, int errbuf_len)
{
  if (cfg == nullptr || key_ == nullptr || value_ == nullptr) {
    set_err(errbuf, errbuf_len, "vlct_config_set: NULL argument");
    return VLCT_ERR_INVALID_CONFIG;
  }
  const std::string key(key_);
  const std::string val = lower(value_);
  double d;

  // removed parameters: EnzoMethodMHDVlct.cpp:62-70
  if (ends_with(key, "half_dt_reconstruct_method") ||
      ends_with(key, "full_dt_reconstruct_method")) {
    set_err(errbuf, errbuf_len,
            "In the \"Method:mhd_vlct\" parameter-group, "
            "\"half_dt_reconstruct_method\" & \"full_dt_reconstruct_method\" "
            "have been removed. The former was only allowed to have a value of "
            "\"nn\" and the latter was replaced with \"reconstruct_method\".");
    return VLCT_ERR_INVALID_CONFIG;
  }

  if (key == "Method:mhd_vlct:riemann_solver") {
    if (val == "hll") cfg->riemann_solver = VLCT_RIEMANN_HLL;
    else if (val == "hlle") cfg->riemann_solver = VLCT_RIEMANN_HLLE;
    else if (val == "hllc") cfg->riemann_solver = VLCT_RIEMANN_HLLC;
    else if (val == "hlld") cfg->riemann_solver = VLCT_RIEMANN_HLLD;
    else {
      set_err(errbuf, errbuf_len,
              "The only known solvers are HLL, HLLE, HLLC, & HLLD");
      return VLCT_ERR_INVALID_CONFIG;
    }
  } else if (key == "Method:mhd_vlct:reconstruct_method") {
    if (val == "nn") cfg->reconstruct_method = VLCT_RECON_NN;
    else if (val == "plm" || val == "plm_enzo") cfg->reconstruct_method = VLCT_RECON_PLM_ENZO;
    else if (val == "plm_athena") cfg->reconstruct_method = VLCT_RECON_PLM_ATHENA;
    else {
      set_err(errbuf, errbuf_len,
              "The only allowed solvers are NN, PLM, PLM_ENZO, & PLM_ATHENA");
      return VLCT_ERR_INVALID_CONFIG;
    }
  } else if (key == "Method:mhd_vlct:theta_limiter") {
    if (!parse_double(val, &d)) goto bad_value;
    cfg->theta_limiter = d;
  } else if (key == "Method:mhd_vlct:mhd_choice") {
    if (val == "no_bfield") cfg->mhd_choice = VLCT_MHD_NO_BFIELD;
    else if (val == "constrained_transport") cfg->mhd_choice = VLCT_MHD_CONSTRAINED_TRANSPORT;
    else if (val == "unsafe_constant_uniform") {
      set_err(errbuf, errbuf_len,
              "constant_uniform is primarilly for debugging purposes. DON'T "
              "use for science runs (things can break).");
      return VLCT_ERR_INVALID_CONFIG;
    } else {
      set_err(errbuf, errbuf_len,
              "Unrecognized choice. Known options include \"no_bfield\" and "
              "\"constrained_transport\"");
      return VLCT_ERR_INVALID_CONFIG;
    }
  } else if (key == "Method:mhd_vlct:time_scheme") {
    if (val == "vl") cfg->time_scheme = VLCT_TIME_VL;
    else if (val == "euler") cfg->time_scheme = VLCT_TIME_EULER;
    else {
      set_err(errbuf, errbuf_len,
              "\"Method:mhd_vlct:time_scheme\" must be \"vl\" or \"euler\"");
      return VLCT_ERR_INVALID_CONFIG;
    }
  } else if (key == "Method:mhd_vlct:courant") {
    if (!parse_double(val, &d)) goto bad_value;
    cfg->courant = d;
  } else if (key == "Physics:fluid_props:eos:gamma" || key == "Field:gamma") {
    if (!parse_double(val, &d)) goto bad_value;
    cfg->gamma = d;
  } else if (key == "Physics:fluid_props:eos:type") {
    if (val != "ideal") {
      set_err(errbuf, errbuf_len,
              "can't currently handle the case with a non-ideal EOS");
      return VLCT_ERR_INVALID_CONFIG;
    }
  } else if (key == "Physics:fluid_props:dual_energy:type") {
    if (val == "disabled") cfg->dual_energy = VLCT_DE_DISABLED;
    else if (val == "modern") cfg->dual_energy = VLCT_DE_MODERN;
    else if (val == "bryan95") cfg->dual_energy = VLCT_DE_BRYAN95;
    else goto bad_value;
  } else if (key == "Method:mhd_vlct:dual_energy") {       // legacy alias
    int b;
    if (!parse_bool(val, &b)) goto bad_value;
    cfg->dual_energy = b ? VLCT_DE_MODERN : VLCT_DE_DISABLED;
  } else if (key == "Physics:fluid_props:dual_energy:eta" ||
             key == "Method:mhd_vlct:dual_energy_eta") {
    if (!parse_double(val, &d)) goto bad_value;
    cfg->dual_energy_eta = d;
  } else if (key == "Physics:fluid_props:floors:density" ||
             key == "Method:mhd_vlct:density_floor") {
    if (!parse_double(val, &d)) goto bad_value;
    cfg->density_floor = d;
  } else if (key == "Physics:fluid_props:floors:pressure" ||
             key == "Method:mhd_vlct:pressure_floor") {
    if (!parse_double(val, &d)) goto bad_value;
    cfg->pressure_floor = d;
  } else {
    set_err(errbuf, errbuf_len, "unknown parameter \"%s\"", key_);
    return VLCT_ERR_UNKNOWN_KEY;
  }
  return VLCT_OK;

bad_value:
  set_err(errbuf, errbuf_len, "invalid value \"%s\" for parameter \"%s\"",
          value_, key_);
  return VLCT_ERR_INVALID_CONFIG;
}

int vlct_config_validate(const vlct_config* cfg, char* errbuf, int errbuf_len)
{
  if (cfg == nullptr) {
    set_err(errbuf, errbuf_len, "NULL configuration");
    return VLCT_ERR_INVALID_CONFIG;
  }
#define FAIL(...) do { set_err(errbuf, errbuf_len, __VA_ARGS__); \
                       return VLCT_ERR_INVALID_CONFIG; } while (0)

  // EnzoMethodMHDVlct.cpp:46-60
  if (cfg->time_scheme != VLCT_TIME_VL && cfg->time_scheme != VLCT_TIME_EULER)
    FAIL("\"Method:mhd_vlct:time_scheme\" must be \"vl\" or \"euler\"");
  // EnzoMethodMHDVlct.cpp:74-77
  if (cfg->mhd_choice == VLCT_MHD_UNSET)
    FAIL("Method:mhd_vlct:mhd_choice wasn't specified");
  if (cfg->mhd_choice != VLCT_MHD_NO_BFIELD &&
      cfg->mhd_choice != VLCT_MHD_CONSTRAINED_TRANSPORT)
    FAIL("Unrecognized choice. Known options include \"no_bfield\" and "
         "\"constrained_transport\"");
  const bool mhd = cfg->mhd_choice == VLCT_MHD_CONSTRAINED_TRANSPORT;

  // EnzoEOSIdeal::construct (fluid-props/EnzoEOSIdeal.hpp:69-73)
  if (!(cfg->gamma > 1.0)) FAIL("gamma should exceed 1.0");

  // EnzoMHDIntegratorStageCommands.cpp:31-40
  if (cfg->dual_energy != VLCT_DE_DISABLED && cfg->dual_energy != VLCT_DE_MODERN)
    FAIL("selected formulation of dual energy formalism is incompatible");
  if (cfg->dual_energy == VLCT_DE_MODERN && !(cfg->dual_energy_eta >= 0))
    FAIL("eta must be non-negative");
  if (!(cfg->density_floor > 0) || !(cfg->pressure_floor > 0))
    FAIL("density and pressure floors must be defined");

  // EnzoRiemann.cpp:26-68
  switch (cfg->riemann_solver) {
  case VLCT_RIEMANN_HLL:
    if (!mhd) FAIL("An \"HLL\" Riemann solver without magnetic fields isn't "
                   "currently supported.");
    // riemann/EnzoRiemannHLL.hpp:201 (DavisWavespeed)
    FAIL("EnzoHLLEWavespeed: This hasn't been tested yet");
  case VLCT_RIEMANN_HLLE:
    if (!mhd) FAIL("The \"HLLE\" Riemann solver without magnetic fields is "
                   "untested");
    break;
  case VLCT_RIEMANN_HLLC:
    if (mhd) FAIL("The \"HLLC\" Riemann Solver can't support mhd");
    break;
  case VLCT_RIEMANN_HLLD:
    if (!mhd) FAIL("The \"HLLD\" Riemann Solver requires magnetic fields");
    break;
  default:
    FAIL("The only known solvers are HLL, HLLE, HLLC, & HLLD");
  }

  // EnzoReconstructor.cpp:19-41
  if (!((1. <= cfg->theta_limiter) && (cfg->theta_limiter <= 2.)))
    FAIL("theta_limiter must satisfy 1<=theta_limiter<=2");
  if (cfg->reconstruct_method != VLCT_RECON_NN &&
      cfg->reconstruct_method != VLCT_RECON_PLM_ENZO &&
      cfg->reconstruct_method != VLCT_RECON_PLM_ATHENA)
    FAIL("The only allowed solvers are NN, PLM, PLM_ENZO, & PLM_ATHENA");

  // EnzoBfieldMethod.cpp:14-27: CT is only tested with two partial timesteps
  if (mhd && cfg->time_scheme == VLCT_TIME_EULER)
    FAIL("This machinery hasn't been tested for cases when "
         "num_partial_timesteps!=2.");

  if (cfg->n_passive < 0 || cfg->n_passive > VLCT_MAX_PASSIVE)
    FAIL("n_passive must lie in [0, %d]", VLCT_MAX_PASSIVE);
#undef FAIL
  return VLCT_OK;
}

const char* vlct_name(void) { return "mhd_vlct"; }

const char* vlct_status_string(int status)
{
  switch (status) {
  case VLCT_OK: return "ok";
  case VLCT_ERR_INVALID_CONFIG: return "invalid configuration";
  case VLCT_ERR_INVALID_BLOCK: return "invalid block";
  case VLCT_ERR_CUDA: return "CUDA error";
  case VLCT_ERR_NO_DEVICE: return "no usable CUDA device (there is no CPU fallback)";
  case VLCT_ERR_UNKNOWN_KEY: return "unknown parameter key";
  default: return "internal error";
  }
}

}  // extern "C"
