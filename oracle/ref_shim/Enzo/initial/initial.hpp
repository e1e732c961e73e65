// TEST INFRASTRUCTURE ONLY -- not part of the product.
//
// Stand-in for the reference's Enzo/initial/initial.hpp umbrella header (written
// for this repo): pulls in only the *real* headers of the initialisers that
// oracle/_ref compiles -- the ones the vlct answer tests use -- from
// /root/reference/src, via the include path.
#ifndef VLCT_SHIM_ENZO_INITIAL_HPP
#define VLCT_SHIM_ENZO_INITIAL_HPP

#include <limits>

#include "Cello/cello.hpp"
#include "Enzo/enzo.hpp"

#include "Enzo/initial/EnzoInitialCloud.hpp"
#include "Enzo/initial/EnzoInitialBCenter.hpp"
#include "Enzo/initial/EnzoInitialShockTube.hpp"
#include "Enzo/initial/EnzoInitialInclinedWave.hpp"

#endif
