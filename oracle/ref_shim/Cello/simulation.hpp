// Test infrastructure only: forwards to the shim's single Cello header.
#ifndef VLCT_SHIM_CELLO_simulation_HPP
#define VLCT_SHIM_CELLO_simulation_HPP
#include "Cello/cello.hpp"
#endif
