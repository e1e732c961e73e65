// TEST INFRASTRUCTURE ONLY -- not part of the product.
//
// Stand-in for the reference's Enzo/initial/initial.hpp umbrella header (written
// for this repo): pulls in only the *real* header of the one initialiser that
// oracle/_ref compiles, EnzoInitialCloud (from /root/reference/src, via the
// include path).
#ifndef VLCT_SHIM_ENZO_INITIAL_HPP
#define VLCT_SHIM_ENZO_INITIAL_HPP

#include <limits>

#include "Cello/cello.hpp"
#include "Enzo/enzo.hpp"

#include "Enzo/initial/EnzoInitialCloud.hpp"

#endif
