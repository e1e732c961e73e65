// Test infrastructure only (see oracle/README.md): a stand-in for the Charm++
// headers so that the reference's VL+CT sources compile without Charm++.
// Written for this repo; contains no reference code.
#ifndef VLCT_SHIM_CHARMPP_H
#define VLCT_SHIM_CHARMPP_H
#include <cstdio>
#include <cstdarg>
#include <string>
#include <vector>
#include <array>
#include <map>
#include <memory>
#include <type_traits>

class CkMigrateMessage {};

#define CkPrintf printf
inline int CkMyPe() { return 0; }
inline int CkNumPes() { return 1; }

namespace PUP {
  class er {
  public:
    bool isPacking() const { return false; }
    bool isUnpacking() const { return false; }
    bool isSizing() const { return true; }
    bool isDeleting() const { return false; }
  };
  class able {
  public:
    able() {}
    able(CkMigrateMessage*) {}
    virtual ~able() {}
    virtual void pup(PUP::er&) {}
  };
}

// every "p | x" is a no-op: nothing is ever serialised by the oracle
template <class T> inline PUP::er& operator|(PUP::er& p, T&) { return p; }
template <class T> inline void PUParray(PUP::er&, T*, std::size_t) {}

#define PUPable_decl(className) /* no registration needed */
#define PUPable_def(className)  /* no registration needed */

#endif
