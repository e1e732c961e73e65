// Test infrastructure only: empty stand-in for Charm++'s pup_stl.h
#ifndef VLCT_SHIM_PUP_STL_H
#define VLCT_SHIM_PUP_STL_H
#include "charm++.h"
#endif
