/*
 * oracle_join.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C) of the reference's chaining / nested ("3D") hash
 * join hot path.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product
 * (libhj3d.so) never links or calls it.
 *
 * Parity pinning: this restatement is checked (tests/test_oracle_*.py) against
 *   - the golden vectors captured from the unmodified reference binaries
 *     (fixtures under tests/golden/), and
 *   - oracle/_ref/libhj3d_ref.so, which instantiates the reference's own
 *     operator templates from /root/reference (built by oracle/Makefile).
 *
 * Every function cites the reference file:line it follows.
 */
#ifndef HJ3D_ORACLE_JOIN_H
#define HJ3D_ORACLE_JOIN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same plain-data descriptors as include/hj3d.h (kept binary compatible). */
typedef struct {
  uint32_t tuple_bytes;  /* stride of one row-store tuple                              */
  uint32_t key_offset;   /* byte offset of the join attribute inside the tuple         */
  uint32_t key_bytes;    /* 4 or 8                                                     */
  uint32_t hash_id;      /* 0 murmur32(u32), 1 murmur64(u64), 2 murmur64((u64)(int32)) */
  uint32_t rowid_offset; /* 0xFFFFFFFF: row id = position; else u32 row id at offset   */
} orc_keyspec;

typedef struct {
  uint64_t matches;      /* probe operator AlgBase::_count  (algebra.hh:456,651)       */
  uint64_t num_cmps;     /* _numCmps                        (algebra.hh:449,658)       */
  uint64_t out_tuples;   /* flat result tuples (== matches for chaining / unnest)      */
  uint64_t checksum_sum; /* order-independent checksum over (left id, right id)        */
  uint64_t checksum_xor;
  uint64_t out_written;
  uint64_t overflow;
} orc_counters;

typedef struct {
  uint64_t num_buckets, num_empty, num_entries, num_distinct_keys;
  uint64_t cc_min, cc_max, cc_sum, cc_sumsq, cc_count;           /* _collisionChainLen         */
  uint64_t ccne_min, ccne_max, ccne_sum, ccne_sumsq, ccne_count; /* _collisionChainLenNonempty */
  uint64_t rsv_main, rsv_sub;        /* getRsvSize / getRsvMainSize, getRsvSubSize */
  uint64_t mem_dir, mem_main, mem_sub;
} orc_stats;

typedef struct orc_table orc_table;

/* util/hasht.hh:52-72 */
uint32_t orc_murmur32(uint32_t x);
uint64_t orc_murmur64(uint64_t x);
/* checksum contribution of one (left, right) result pair -- shared definition with the GPU engine */
uint64_t orc_pair_mix(uint32_t left, uint32_t right);

/* kind 0 = HtChaining1 (ht_chaining.hh), 1 = HtNested1 (ht_nested.hh) */
orc_table* orc_build(int kind, const void* tuples, uint64_t n, orc_keyspec ks, uint64_t num_buckets);
void       orc_table_free(orc_table*);
void       orc_table_stats(const orc_table*, orc_stats* out);

/* AlgHashJoinProbe::step (algebra.hh:625-659).  gather: optional indirection, probe tuple i =
 * tuples[gather[i]] and the emitted left id is i.  out_pairs (nullable) receives (left,right)
 * pairs in the reference's emission order. */
void orc_probe_chaining(const orc_table*, const void* tuples, uint64_t n, orc_keyspec ks,
                        const uint32_t* gather, int build_key_unique,
                        uint32_t* out_pairs, uint64_t out_cap, orc_counters* c);

/* AlgNestJoinProbe::step (algebra.hh:435-459) + HtNested1::findMainNodeByOther (ht_nested.hh:354-382).
 * out_pairs receives (left, group_ref); group_ref = index of the MainNode in creation order. */
void orc_probe_nested(const orc_table*, const void* tuples, uint64_t n, orc_keyspec ks,
                      const uint32_t* gather,
                      uint32_t* out_pairs, uint64_t out_cap, orc_counters* c);

/* AlgUnnestHt::step (algebra.hh:510-541): (left, group_ref) -> (left, build row id) per group member,
 * MainNode first, then the sub chain head to tail. */
void orc_unnest(const orc_table*, const uint32_t* left, const uint32_t* gref, uint64_t n,
                uint32_t* out_pairs, uint64_t out_cap, orc_counters* c);

/* number of MainNodes (distinct keys) and the size of one group */
uint64_t orc_num_groups(const orc_table*);
uint64_t orc_group_len(const orc_table*, uint32_t gref);

#ifdef __cplusplus
}
#endif
#endif
