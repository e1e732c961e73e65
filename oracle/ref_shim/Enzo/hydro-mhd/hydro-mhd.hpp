// TEST INFRASTRUCTURE ONLY -- not part of the product.
// Replaces the reference's hydro-mhd umbrella header so that only the VL+CT
// pieces (toolkit, riemann, EnzoMethodMHDVlct) are pulled in -- the PPM/PPML
// Fortran-backed methods are not built.
#ifndef VLCT_SHIM_HYDROMHD_HPP
#define VLCT_SHIM_HYDROMHD_HPP
#include <array>
#include <string>
#include <vector>
#include "Cello/cello.hpp"
#include "Enzo/enzo.hpp"
#include "Enzo/hydro-mhd/toolkit/toolkit.hpp"          // real
#include "Enzo/hydro-mhd/riemann/EnzoRiemann.hpp"      // real
#include "Enzo/hydro-mhd/EnzoMethodMHDVlct.hpp"        // real
#endif
