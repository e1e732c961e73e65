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
#include <type_traits>
#include <vector>

#include "vlct.h"

// The library computes in fp64 and reads Cello's field arrays in place: an
// Enzo-E build with CONFIG_PRECISION_SINGLE must not bind them.
static_assert(std::is_same<enzo_float, double>::value,
              "EnzoMethodMHDVlctGpu needs a double-precision Enzo-E build "
              "(CONFIG_PRECISION_DOUBLE): libvlct_b200 reads the fields as fp64");

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
      store_fluxes_for_corrections_(false), batch_blocks_(true),
      fused_timestep_(false) {}
  void pup(PUP::er& p);

  virtual ~EnzoMethodMHDVlctGpu();

  virtual void compute(Block* block) throw();
  virtual std::string name() throw() { return vlct_name(); }
  virtual double timestep(Block* block) throw();

  /// blocks queued by compute() and not yet advanced (0 between cycles)
  std::size_t queued_blocks() const { return queue_.size(); }

protected:
  /// (re)creates the library handle from config_
  void create_handle_();
  /// advances every queued block: one vlct_compute[_and_timestep]_batch per
  /// group of blocks with equal dt and cell widths, then compute_done() on all
  void flush_queue_() throw();
  /// one block, right away
  void compute_one_(Block* block) throw();
  /// fills a vlct_block with the pointers Field::values() returns
  void bind_block_(Block* block, vlct_block* out) noexcept;
  /// deposits dt/dx * face fluxes in the block's FluxData (for "flux_correct")
  void save_fluxes_for_corrections_(Block* block, const vlct_block& b) noexcept;

  vlct_config config_;
  vlct_handle* handle_;
  std::vector<std::string> passive_names_;
  bool store_fluxes_for_corrections_;

  /// "Method:mhd_vlct:gpu_batch_blocks" (default true): with more than one
  /// block on this process, compute() only queues a block; the call for the
  /// process's last block advances all of them in one set of kernel launches
  /// (vlct_compute_batch) and then reports compute_done() for each.
  bool batch_blocks_;
  /// "Method:mhd_vlct:gpu_fused_timestep" (default false): compute() also
  /// evaluates the timestep() of the next cycle on the device
  /// (vlct_compute_and_timestep[_batch]) and timestep() returns that value.
  /// Only valid when no other Method / Boundary changes the hydro / MHD fields
  /// between this Method's compute and the stopping phase's timestep.
  bool fused_timestep_;

  // -- transient state, never serialised
  struct Queued { Block* block; bool leaf; };
  std::vector<Queued> queue_;
  struct CachedDt { Block* block; int cycle; double dt; };
  std::vector<CachedDt> cached_dt_;     // filled by compute(), read by timestep()
};

#endif /* ENZO_ENZO_METHOD_MHD_VLCT_GPU_HPP */
