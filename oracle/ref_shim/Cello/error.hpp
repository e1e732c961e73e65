// Test infrastructure only: forwards to the shim's single Cello header.
#ifndef VLCT_SHIM_CELLO_error_HPP
#define VLCT_SHIM_CELLO_error_HPP
#include "Cello/cello.hpp"
#endif
