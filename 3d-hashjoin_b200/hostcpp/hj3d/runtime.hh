// hj3d/runtime.hh -- C++ side of the C ABI (include/hj3d.h): error handling, the process-wide device
// context, device buffers, and the derivation of a device-describable key specification from a
// reference-style hash functor (concepts.hh:22-28).
#pragma once

#include <cstdint>
#include <cstring>
#include <random>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "../../../include/hj3d.h"

namespace hj3d {

struct Error : std::runtime_error { using std::runtime_error::runtime_error; };

inline int check(int rc) {
  if (rc < 0) throw Error(std::string("hj3d: ") + hj3d_last_error());
  return rc;
}

// One device context per process (the reference is single-threaded, algebra.hh has no thread safety).
class Runtime {
  public:
    static Runtime& instance() { static Runtime r; return r; }
    hj3d_ctx* ctx() {
      if (!_ctx) check(hj3d_ctx_create(_device, &_ctx));   // throws without a GPU: there is no CPU fallback
      return _ctx;
    }
    void device(int d) { _device = d; }
    void cache_uploads(bool on) { _cache = on; }
    bool cache_uploads() const { return _cache; }
    ~Runtime() { if (_ctx) hj3d_ctx_destroy(_ctx); }
  private:
    Runtime() = default;
    hj3d_ctx* _ctx = nullptr;
    int       _device = 0;
    bool      _cache = false;
};

// grow-only device buffer
class DevBuf {
  public:
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void* ensure(uint64_t bytes) {
      if (bytes > _cap) {
        release();
        check(hj3d_mem_alloc(Runtime::instance().ctx(), bytes, &_p));
        _cap = bytes;
      }
      return _p;
    }
    void release() { if (_p) { hj3d_mem_free(Runtime::instance().ctx(), _p); _p = nullptr; _cap = 0; } tag_ptr = nullptr; tag_n = 0; }
    void* get() const { return _p; }
    template <class T> T* as() const { return static_cast<T*>(_p); }
    const void* tag_ptr = nullptr; uint64_t tag_n = 0;          // what the buffer holds a copy of (upload cache)
  private:
    void*    _p = nullptr;
    uint64_t _cap = 0;
};

// ---------------------------------------------------------------- key specification of a hash functor
inline uint32_t ref_murmur32(uint32_t x) { x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16; return x; }
inline uint64_t ref_murmur64(uint64_t x) { x ^= x >> 33; x *= 0xFF51AFD7ED558CCDull; x ^= x >> 33; x *= 0xC4CEB9FE1A95EC63ull; x ^= x >> 33; return x; }

template <class T> concept has_explicit_keyspec = requires { { T::hj3d_keyspec() } -> std::same_as<hj3d_keyspec>; };

// A functor over a flat, trivially copyable row-store tuple is probed on the host: which naturally aligned
// 4/8-byte field, hashed with which murmur finaliser, reproduces Thashfun::eval on random tuples?  This
// makes the reference's functor structs usable unmodified; functors that reach the key through a pointer
// (main_experiment4.cc:413-419) must provide `static hj3d_keyspec hj3d_keyspec()` plus `hj3d_base`.
template <class Thashfun>
hj3d_keyspec keyspec_of() {
  if constexpr (has_explicit_keyspec<Thashfun>) {
    return Thashfun::hj3d_keyspec();
  } else {
    using tuple_t = std::remove_const_t<typename Thashfun::input_t>;
    using out_t = typename Thashfun::output_t;
    static_assert(std::is_trivially_copyable_v<tuple_t>, "hj3d: hash functor input must be a trivially copyable tuple");
    static const hj3d_keyspec cached = [] {
      constexpr size_t B = sizeof(tuple_t);
      std::mt19937_64 rng(42);
      std::vector<hj3d_keyspec> cand;
      auto consider = [&](uint32_t off, uint32_t kb, uint32_t hid) {
        for (int trial = 0; trial < 16; ++trial) {
          alignas(tuple_t) unsigned char raw[B];
          for (size_t i = 0; i < B; ++i) raw[i] = (unsigned char)rng();
          tuple_t t; std::memcpy(&t, raw, B);
          uint64_t want = (uint64_t)Thashfun::eval(&t), got;
          if (hid == HJ3D_HASH_MURMUR32) { uint32_t k; std::memcpy(&k, raw + off, 4); got = ref_murmur32(k); }
          else if (hid == HJ3D_HASH_MURMUR64) { uint64_t k; std::memcpy(&k, raw + off, 8); got = ref_murmur64(k); }
          else { int32_t k; std::memcpy(&k, raw + off, 4); got = ref_murmur64((uint64_t)(int64_t)k); }
          if (got != want) return;
        }
        cand.push_back(hj3d_keyspec{(uint32_t)B, off, kb, hid, HJ3D_NO_ROWID});
      };
      if constexpr (sizeof(out_t) == 4) {
        for (uint32_t off = 0; off + 4 <= B; off += 4) consider(off, 4, HJ3D_HASH_MURMUR32);
      } else {
        for (uint32_t off = 0; off + 8 <= B; off += 8) consider(off, 8, HJ3D_HASH_MURMUR64);
        for (uint32_t off = 0; off + 4 <= B; off += 4) consider(off, 4, HJ3D_HASH_MURMUR64_SEXT32);
      }
      if (cand.size() != 1)
        throw Error("hj3d: cannot derive a device key description for this hash functor (not murmur_hash of one "
                    "aligned 4/8-byte attribute); give it `static hj3d_keyspec hj3d_keyspec()`");
      return cand[0];
    }();
    return cached;
  }
}

// The device compares the raw bits of the key attributes that the two hash functors name (hj3d_keyspec); the
// reference additionally evaluates a predicate functor per visited node (joinpred_t::eval, algebra.hh:447,647-648;
// eqfun_t::eval through isMainNodeMatch, ht_nested.hh:241-243).  A predicate that is not equality of those two
// attributes would silently give different results, so it is checked on random tuple pairs the same way the hash
// functor is, and rejected otherwise.  Functors over pointer-reaching intermediates (explicit keyspec) cannot be
// synthesised here; they are checked through their base functors (`hj3d_base`) when they name one.
template <class Tpred, class HashL, class HashR>
void check_key_equality_predicate(const char* what) {
  using left_t = std::remove_const_t<typename Tpred::left_t>;
  using right_t = std::remove_const_t<typename Tpred::right_t>;
  if constexpr (has_explicit_keyspec<HashL> || has_explicit_keyspec<HashR> || requires { typename HashL::hj3d_base; } ||
                requires { typename HashR::hj3d_base; }) {
    return;   // the functor reaches its key through a pointer: random tuples cannot be synthesised (its base functor is checked)
  } else if constexpr (!std::is_same_v<left_t, std::remove_const_t<typename HashL::input_t>> ||
                       !std::is_same_v<right_t, std::remove_const_t<typename HashR::input_t>> ||
                       !std::is_trivially_copyable_v<left_t> || !std::is_trivially_copyable_v<right_t>) {
    return;   // the predicate sees other tuple types than the hash functors (nested intermediates)
  } else {
    static const bool ok = [what] {
      const hj3d_keyspec kl = keyspec_of<HashL>(), kr = keyspec_of<HashR>();
      if (kl.key_bytes != kr.key_bytes || kl.hash_id != kr.hash_id)
        throw Error(std::string("hj3d: ") + what + ": probe and build hash functors hash different key types");
      std::mt19937_64 rng(4242);
      for (int trial = 0; trial < 64; ++trial) {
        alignas(left_t) unsigned char lraw[sizeof(left_t)];
        alignas(right_t) unsigned char rraw[sizeof(right_t)];
        for (auto& b : lraw) b = (unsigned char)rng();
        for (auto& b : rraw) b = (unsigned char)rng();
        const bool equal_keys = (trial & 1) != 0;
        if (equal_keys) std::memcpy(rraw + kr.key_offset, lraw + kl.key_offset, kl.key_bytes);
        else if (std::memcmp(rraw + kr.key_offset, lraw + kl.key_offset, kl.key_bytes) == 0) rraw[kr.key_offset] ^= 1;
        left_t l; right_t r;
        std::memcpy(&l, lraw, sizeof l); std::memcpy(&r, rraw, sizeof r);
        if (Tpred::eval(&l, &r) != equal_keys)
          throw Error(std::string("hj3d: ") + what + " is not equality of the hashed key attributes; the device engine "
                      "compares exactly the attribute the hash functors name (see INTEGRATION.md)");
      }
      return true;
    }();
    (void)ok;
  }
}

}  // namespace hj3d
