// TEST INFRASTRUCTURE ONLY -- not part of the product.
//
// Stand-in for Enzo-E's umbrella header (written for this repo). It declares
// the few enzo:: accessors the VL+CT sources call, a minimal EnzoBlock (just
// CellWidth), and then pulls in the *real* fluid-props / utils headers of the
// reference (from /root/reference/src, via the include path).
#ifndef VLCT_SHIM_ENZO_HPP
#define VLCT_SHIM_ENZO_HPP

#include "Cello/cello.hpp"

#include "Enzo/enzo_typedefs.hpp"   // real: enzo_float, EFlt3DArray, ...

#define MAX_DIMENSION 3

class EnzoBlock : public Block {
public:
  EnzoBlock(FieldDescr* d) : Block(d) { CellWidth[0]=CellWidth[1]=CellWidth[2]=0; }
  enzo_float CellWidth[MAX_DIMENSION];
};

class EnzoPhysicsCosmology;
class EnzoPhysicsFluidProps;
class EnzoMethodGrackle;
class GrackleChemistryData;

namespace enzo {
  EnzoBlock*                  block(Block* block);
  Problem*                    problem();
  EnzoPhysicsCosmology*       cosmology();
  EnzoPhysicsFluidProps*      fluid_props();
  const EnzoMethodGrackle*    grackle_method();
  const GrackleChemistryData* grackle_chemistry();
  double grav_constant_codeU() noexcept;
}

#include "Enzo/chemistry/chemistry.hpp"   // shim: inert EnzoMethodGrackle

// real reference headers
#include "Enzo/utils/utils.hpp"
#include "Enzo/fluid-props/fluid-props.hpp"
#include "Enzo/enzo-core/EnzoBoundary.hpp"   // real: outflow / reflecting

#endif /* VLCT_SHIM_ENZO_HPP */
