// hot_set.cuh -- the hot-key set of a skewed probe side (see hot.cuh): types and the lookup the partition kernel inlines.
#pragma once

#include "common.cuh"

namespace hj3d {

constexpr int kHotMax    = 128;     // hot keys at most
constexpr int kHotSlots  = 256;    // open-addressing table of the hot keys (shared memory of the partition kernel)
constexpr int kHotAns    = 8;       // build rows per hot key at most
constexpr int kHotSample = 16384;   // sample size over all ranks (a power of two: bitonic sort in shared memory)
constexpr uint32_t kHotMinCount = 4;   // a key is a candidate if it shows up this often in the sample (the most frequent kHotMax of them are taken)

template <class KeyT> struct HotEntry;
template <> struct __align__(8)  HotEntry<uint32_t> { uint32_t key; uint32_t idx; };                 // idx: 0 = empty, else hot index + 1
template <> struct __align__(16) HotEntry<uint64_t> { uint64_t key; uint32_t idx; uint32_t pad; };

template <class KeyT> struct HotTable {
  uint32_t n; uint32_t pad_[3];
  HotEntry<KeyT> slot[kHotSlots];
  KeyT key[kHotMax];
};

struct HotAns { uint32_t nm; uint32_t cmps; uint32_t row[kHotAns]; };   // per hot key; summed over the ranks (non-owners hold zeros)
struct HotAnswers { HotAns a[kHotMax]; uint32_t too_many; uint32_t pad_[3]; };
static_assert(sizeof(HotAnswers) % 4 == 0, "all-reduced as uint32");

__device__ __forceinline__ uint32_t hot_slot_of(uint32_t h) { return (h * 0x9E3779B1u) >> 24; }   // 8 bits
static_assert(kHotSlots == 256, "hot_slot_of yields 8 bits");

// index of `key` in the hot set or -1; sl = the table's slots (shared or global memory)
template <class KeyT>
__device__ __forceinline__ int hot_find(const HotEntry<KeyT>* sl, KeyT key, uint32_t h) {
  uint32_t s = hot_slot_of(h);
  while (true) {
    const HotEntry<KeyT> e = sl[s];
    if (e.idx == 0) return -1;
    if (e.key == key) return (int)e.idx - 1;
    s = (s + 1) & (kHotSlots - 1);
  }
}

}  // namespace hj3d
