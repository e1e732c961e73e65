// Test infrastructure only: forwards bare "enzo.hpp" includes to the shim.
#include "Enzo/enzo.hpp"
