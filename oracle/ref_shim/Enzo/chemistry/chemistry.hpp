// TEST INFRASTRUCTURE ONLY. Grackle is not part of the VL+CT hot path; the
// oracle always runs with enzo::grackle_method() == nullptr, so this inert
// declaration only has to satisfy the compiler.
#ifndef VLCT_SHIM_CHEMISTRY_HPP
#define VLCT_SHIM_CHEMISTRY_HPP
#include "Cello/cello.hpp"
#include "Enzo/enzo_typedefs.hpp"
class EnzoFieldAdaptor;
class EnzoMethodGrackle {
public:
  void calculate_pressure(const EnzoFieldAdaptor&, enzo_float*, int) const {}
};
#endif
