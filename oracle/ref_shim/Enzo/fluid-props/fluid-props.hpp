// TEST INFRASTRUCTURE ONLY. Replaces the reference's fluid-props umbrella
// header so that EnzoComputeTemperature (which needs Grackle types and is not
// on the VL+CT path) is left out. Everything included below is the reference's
// own, unmodified header.
#ifndef VLCT_SHIM_FLUIDPROPS_HPP
#define VLCT_SHIM_FLUIDPROPS_HPP
#include <string>
#include "Cello/cello.hpp"
#include "Enzo/enzo.hpp"
#include "fluid-props/EnzoEOSIdeal.hpp"
#include "fluid-props/EnzoEOSIsothermal.hpp"
#include "fluid-props/EnzoEOSVariant.hpp"
#include "fluid-props/EnzoDualEnergyConfig.hpp"
#include "fluid-props/EnzoFluidFloorConfig.hpp"
#include "fluid-props/EnzoPhysicsFluidProps.hpp"
#include "fluid-props/EnzoComputePressure.hpp"
#endif
