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
#include <cstring>

class CkMigrateMessage {};

#define CkPrintf printf
inline int CkMyPe() { return 0; }
inline int CkNumPes() { return 1; }

namespace PUP {
  /// A byte-buffer PUP::er. Default-constructed it is the old "sizing" no-op
  /// (nothing the reference's own sources pup is ever serialised here). In
  /// the three real modes it sizes / packs / unpacks what the GPU adapter's
  /// pup() passes through it: trivially copyable values, raw arrays,
  /// std::string and std::vector of those (tests/test_gpu_adapter.py drives a
  /// pack -> unpack round trip of EnzoMethodMHDVlctGpu).
  class er {
  public:
    enum Mode { NOOP, SIZING, PACKING, UNPACKING };
    er() : mode_(NOOP), pos_(0), buf_(nullptr) {}
    er(Mode mode, std::vector<char>* buf) : mode_(mode), pos_(0), buf_(buf) {}
    bool isPacking() const { return mode_ == PACKING; }
    bool isUnpacking() const { return mode_ == UNPACKING; }
    bool isSizing() const { return mode_ == SIZING || mode_ == NOOP; }
    bool isDeleting() const { return false; }
    std::size_t size() const { return pos_; }
    void bytes(void* p, std::size_t n) {
      if (mode_ == PACKING) buf_->insert(buf_->end(), (char*) p, (char*) p + n);
      else if (mode_ == UNPACKING) memcpy(p, buf_->data() + pos_, n);
      if (mode_ != NOOP) pos_ += n;
    }
  private:
    Mode mode_;
    std::size_t pos_;
    std::vector<char>* buf_;
  };
  class able {
  public:
    able() {}
    able(CkMigrateMessage*) {}
    virtual ~able() {}
    virtual void pup(PUP::er&) {}
  };
}

// "p | x": trivially copyable values, strings and vectors of them travel
// through the buffer; anything else (the reference's own aggregate members,
// which this shim never serialises) is a no-op
inline PUP::er& operator|(PUP::er& p, std::string& s) {
  std::size_t n = s.size();
  p.bytes(&n, sizeof(n));
  if (p.isUnpacking()) s.resize(n);
  if (n) p.bytes(&s[0], n);
  return p;
}
template <class T> inline PUP::er& operator|(PUP::er& p, T& v) {
  if constexpr (std::is_trivially_copyable<T>::value) p.bytes(&v, sizeof(T));
  return p;
}
template <class T> inline PUP::er& operator|(PUP::er& p, std::vector<T>& v) {
  if constexpr (std::is_trivially_copyable<T>::value || std::is_same<T, std::string>::value) {
    std::size_t n = v.size();
    p.bytes(&n, sizeof(n));
    if (p.isUnpacking()) v.resize(n);
    for (std::size_t i = 0; i < n; i++) p | v[i];
  }
  return p;
}
template <class T> inline void PUParray(PUP::er& p, T* a, std::size_t n) {
  if constexpr (std::is_trivially_copyable<T>::value) p.bytes(a, n * sizeof(T));
}

#define PUPable_decl(className) /* no registration needed */
#define PUPable_def(className)  /* no registration needed */

#endif
