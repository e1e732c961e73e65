// Test infrastructure only: forwards bare "cello.hpp" includes to the shim.
#include "Cello/cello.hpp"
