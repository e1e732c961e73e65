// Test infrastructure only: forwards to the shim's single Cello header.
#ifndef VLCT_SHIM_CELLO_view_HPP
#define VLCT_SHIM_CELLO_view_HPP
#include "Cello/cello.hpp"
#endif
