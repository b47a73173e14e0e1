// probe_packed.cuh -- shared-memory probe over a COMPRESSED table slice (at-most-one-result chaining probes with
// 32-bit murmur keys: plan Csr, algebra.hh:625-659 with IsBuildKeyUnique).
//
// Why: the probe side of a large join is partitioned until a partition's slice of the table fits in shared memory; the
// fewer partitions that takes, the longer the runs the partition passes write and the fewer fixed costs per record.  The
// plain slice (probe_fine.cuh) costs 4 B of directory + 8 B of slot per bucket.  murmur32 is a bijection, so inside a
// slice a key is identified by (bucket, quotient) with quotient = hash / numBuckets -- log2(2^32 / D) bits -- and the slot
// shrinks to ONE 32-bit word  (quotient << rowid_bits) | row id  whenever quotient bits + row id bits <= 32 (always so
// for a key/foreign-key build side with D = |R| buckets); run starts shrink to 16-bit offsets relative to the slice.
// 6 B per bucket instead of 12: a 96 KB slice holds 2^14 buckets, two 512-thread blocks per SM.
// Measured (tools/ubench.cu, B200): 3.97 ms per 2^30 probes at 8192 partitions against 5.39 ms for the plain slice
// at 65536 partitions.
//
// The global table keeps its layout (off[], Slot[]); a block compresses its slice while staging it.  Buckets of
// <= kOrderedMax entries are stored in the reference's chain order, so the probe is the walk of algebra.hh:644-657
// (++cmps per node, stop at the first match); longer buckets take the row-id rules of probe.cuh.
// Probe records are read with 128-bit loads (two (key, id) records per load) one tile ahead of their use.
#pragma once

#include "common.cuh"
#include "probe.cuh"

namespace hj3d {

constexpr int kPkThreads = 512;
#ifndef HJ3D_PK_ITEMS
#define HJ3D_PK_ITEMS 8
#endif
constexpr int kPkItems   = HJ3D_PK_ITEMS;                 // records per thread and tile (even: 128-bit loads take two)
constexpr int kPkTile    = kPkThreads * kPkItems;

struct PackCfg {
  uint32_t width;        // buckets per fine partition
  uint32_t n_local;      // buckets of the (shard) directory
  uint32_t smem_bytes;   // dynamic shared memory of a block
  uint32_t rowid_bits;   // row ids of the table are < 2^rowid_bits
  uint32_t qshift;       // D a power of two: quotient = hash >> qshift
  uint32_t pow2;
  uint64_t qmagic;       // else quotient = umul64hi(qmagic, hash), qmagic = ceil(2^64 / D)
};

// probe_packed.cu
cudaError_t launch_probe_packed(cudaStream_t st, bool checksum, bool write, uint32_t n_work, size_t smem, const Slot<uint32_t>* recs, Dir d,
                                PackCfg pc, const uint2* work, const uint32_t* work_part, const uint32_t* off, const Slot<uint32_t>* slots,
                                uint2* out, unsigned long long out_cap, DevCounters* ctr);

}  // namespace hj3d
