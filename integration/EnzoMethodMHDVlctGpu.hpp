// EnzoMethodMHDVlctGpu -- the reference-side binding of the B200 VL+CT library.
//
// This is the file a maintainer of Enzo-E adds next to
// src/Enzo/hydro-mhd/EnzoMethodMHDVlct.hpp: a `Method` plugin with the same
// name ("mhd_vlct"), constructor signature and parameter keys as
// EnzoMethodMHDVlct, whose compute()/timestep() forward to the C ABI of
// include/vlct.h (libvlct_b200.so). Nothing else of Enzo-E changes:
// EnzoProblem::create_method_ (src/Enzo/enzo-core/EnzoProblem.cpp:653-656)
// constructs this class instead of EnzoMethodMHDVlct when the library is
// available.
//
// It is compiled in this repository against the same Cello stand-in headers
// that build the CPU reference (oracle/ref_shim) and exercised on the GPU by
// tests/test_gpu_adapter.py; against a real Enzo-E tree it needs only the
// usual Cello/Enzo umbrella headers.
#ifndef ENZO_ENZO_METHOD_MHD_VLCT_GPU_HPP
#define ENZO_ENZO_METHOD_MHD_VLCT_GPU_HPP

#include <string>
#include <vector>

#include "vlct.h"

class EnzoMethodMHDVlctGpu : public Method {
  /// @class    EnzoMethodMHDVlctGpu
  /// @ingroup  Enzo
  /// @brief    [\ref Enzo] VL+CT MHD, computed by libvlct_b200 on a B200.

public:
  /// same signature as EnzoMethodMHDVlct (EnzoMethodMHDVlct.cpp:90)
  EnzoMethodMHDVlctGpu(ParameterGroup p, bool store_fluxes_for_corrections);

  /// Charm++ PUP::able declarations (only the configuration is serialised,
  /// like EnzoMethodMHDVlct::pup, EnzoMethodMHDVlct.cpp:170-197)
  PUPable_decl(EnzoMethodMHDVlctGpu);
  EnzoMethodMHDVlctGpu(CkMigrateMessage* m)
    : Method(m), handle_(nullptr), passive_names_(),
      store_fluxes_for_corrections_(false) {}
  void pup(PUP::er& p);

  virtual ~EnzoMethodMHDVlctGpu();

  virtual void compute(Block* block) throw();
  virtual std::string name() throw() { return vlct_name(); }
  virtual double timestep(Block* block) throw();

protected:
  /// (re)creates the library handle from config_
  void create_handle_();
  /// fills a vlct_block with the pointers Field::values() returns
  void bind_block_(Block* block, vlct_block* out) noexcept;
  /// deposits dt/dx * face fluxes in the block's FluxData (for "flux_correct")
  void save_fluxes_for_corrections_(Block* block, const vlct_block& b) noexcept;

  vlct_config config_;
  vlct_handle* handle_;
  std::vector<std::string> passive_names_;
  bool store_fluxes_for_corrections_;
};

#endif /* ENZO_ENZO_METHOD_MHD_VLCT_GPU_HPP */
