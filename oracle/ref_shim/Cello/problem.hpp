// Test infrastructure only: forwards to the shim's single Cello header.
#ifndef VLCT_SHIM_CELLO_problem_HPP
#define VLCT_SHIM_CELLO_problem_HPP
#include "Cello/cello.hpp"
#endif
