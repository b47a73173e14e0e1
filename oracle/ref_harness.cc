/*
 * ref_harness.cc -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Thin C-ABI wrapper that instantiates the UNMODIFIED reference operator
 * templates (algebra.hh, ht_chaining.hh, ht_nested.hh, util/*) found under
 * $(REF) (= /root/reference, passed with -I by oracle/Makefile).  No reference
 * source is copied into this repository; this file only *uses* the headers
 * where they lie.  The build product (oracle/_ref/libhj3d_ref.so) is git-ignored
 * and is used
 *   (1) to validate oracle/oracle_join.c on random inputs,
 *   (2) to generate the golden fixtures in tests/golden/ (oracle/gen_golden.py),
 *   (3) as the "reference" CPU baseline of bench.py.
 *
 * The functor structs below have the same shape as the drivers' own
 * (main_experiment1.cc:287-410, main_experiment4.cc:346-491,
 * main_algebra_example.cc:31-145), generalised over the tuple layout.
 */
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <sstream>
#include <functional>
#include <iostream>
#include <memory>
#include <numeric>
#include <random>
#include <unordered_set>
#include <vector>

#include "algebra.hh"
#include "ht_chaining.hh"
#include "ht_nested.hh"
#include "util/hasht.hh"
#include "util/GenRandIntVec.hh"

#include "oracle_join.h"

namespace {

// ---------------------------------------------------------------- tuple layouts
template <size_t Bytes> struct Tup { unsigned char raw[Bytes]; };
template <size_t Bytes>
std::ostream& operator<<(std::ostream& os, const Tup<Bytes>&) { return os << "[tuple]"; }

struct GlobStat {};

// the drivers route every hash through a std::function (main_experiment1.cc:231, main_experiment4.cc:263)
static std::function<uint32_t(uint32_t)> g_hash32 = ht::murmur_hash<uint32_t>;
static std::function<uint64_t(uint64_t)> g_hash64 = ht::murmur_hash<uint64_t>;

template <int HashId> struct HashTraits;
template <> struct HashTraits<0> { using key_t = uint32_t; using out_t = uint32_t;
  static out_t h(key_t k) { return g_hash32(k); } };
template <> struct HashTraits<1> { using key_t = uint64_t; using out_t = uint64_t;
  static out_t h(key_t k) { return g_hash64(k); } };
template <> struct HashTraits<2> { using key_t = int32_t;  using out_t = uint64_t;   // main_algebra_example.cc:53-65
  static out_t h(key_t k) { return ht::murmur_hash<uint64_t>(k); } };

template <class T, size_t Off, int HashId>
inline typename HashTraits<HashId>::key_t key_of(const T* t) {
  typename HashTraits<HashId>::key_t k;
  std::memcpy(&k, t->raw + Off, sizeof(k));
  return k;
}

template <class T, size_t Off, int HashId>
struct Hashfun {                                   // concepts.hh:22-28
  using input_t = T;
  using output_t = typename HashTraits<HashId>::out_t;
  inline static output_t eval(const input_t* t) { return HashTraits<HashId>::h(key_of<T, Off, HashId>(t)); }
};

template <class L, size_t OffL, class R, size_t OffR, int HashId>
struct Eqfun {                                     // concepts.hh:49-56
  using left_t = L;
  using right_t = R;
  inline static bool eval(const left_t* l, const right_t* r) {
    return key_of<L, OffL, HashId>(l) == key_of<R, OffR, HashId>(r);
  }
};

// ---------------------------------------------------------------- result sinks
// consumer that folds every result tuple into orc_counters (alg_consumer_c, algebra.hh:62-73)
template <class Tin, class Extract>
class Sink : public AlgBase {
  public:
    using globstat_t = GlobStat;
    using input_t    = Tin;
    using output_t   = void;
    Sink() : AlgBase("Sink") {}
    void init(globstat_t*) { reset(); }
    void step(input_t* t, globstat_t*) {
      inc();
      auto [l, r] = Extract::ids(t, _lbase, _rbase);
      uint64_t m = orc_pair_mix(l, r);
      _c.checksum_sum += m; _c.checksum_xor ^= m;
      if (_out) { if (_c.out_tuples < _cap) { _out[2 * _c.out_tuples] = l; _out[2 * _c.out_tuples + 1] = r; ++_c.out_written; } else _c.overflow = 1; }
      ++_c.out_tuples;
    }
    void fin(globstat_t*) { stopTimer(); }
    orc_counters _c{};
    const void*  _lbase = nullptr;
    const void*  _rbase = nullptr;
    uint32_t*    _out = nullptr;
    uint64_t     _cap = 0;
};

template <class L, class R> struct FlatTuple { const L* _left; const R* _right; };
template <class L, class R>
std::ostream& operator<<(std::ostream& os, const FlatTuple<L, R>&) { return os << "[flat]"; }

template <class L, class R> struct FlatIds {
  static std::pair<uint32_t, uint32_t> ids(const FlatTuple<L, R>* t, const void* lb, const void* rb) {
    return { (uint32_t)(t->_left - (const L*)lb), (uint32_t)(t->_right - (const R*)rb) };
  }
};

// ---------------------------------------------------------------- type-erased handle
struct RefTable {
  virtual ~RefTable() {}
  virtual void stats(orc_stats*) const = 0;
  virtual std::string stats_text() const = 0;
  virtual int  kind() const = 0;
  virtual uint32_t tuple_bytes() const = 0;
  virtual uint32_t key_offset() const = 0;
  virtual int  hash_id() const = 0;
};

static std::string stats_text_of(const HtStatistics& hs) {
  std::ostringstream os;
  hs.print(os);
  os << '\x1e' << hs.toCsvString() << '\x1e' << HtStatistics::toCsvStringHeader();
  return os.str();
}

static void fill_stats(const HtStatistics& hs, orc_stats* s) {
  s->num_buckets = hs._numBuckets; s->num_empty = hs._numEmptyBuckets;
  s->num_entries = hs._numEntries; s->num_distinct_keys = hs._numDistinctKeys;
  s->cc_min = hs._collisionChainLen.min(); s->cc_max = hs._collisionChainLen.max();
  s->cc_sum = hs._collisionChainLen.sum(); s->cc_sumsq = hs._collisionChainLen.sumsq();
  s->cc_count = hs._collisionChainLen.count();
  s->ccne_min = hs._collisionChainLenNonempty.min(); s->ccne_max = hs._collisionChainLenNonempty.max();
  s->ccne_sum = hs._collisionChainLenNonempty.sum(); s->ccne_sumsq = hs._collisionChainLenNonempty.sumsq();
  s->ccne_count = hs._collisionChainLenNonempty.count();
}

template <size_t BB, size_t BOff, int HashId>
struct ChainingTable : RefTable {
  using tup_t   = Tup<BB>;
  using hf_t    = Hashfun<tup_t, BOff, HashId>;
  using eq_t    = Eqfun<tup_t, BOff, tup_t, BOff, HashId>;
  using build_t = AlgHashJoinBuild<hf_t, eq_t, GlobStat>;
  RelationRS<tup_t> rel;
  build_t           op;
  ChainingTable(const void* tuples, uint64_t n, uint64_t D) : rel(), op(D, 10) {   // log2 chunk size 10: main_experiment1.cc:218
    rel._tuples.resize(n);
    if (n) std::memcpy((void*)rel._tuples.data(), tuples, n * BB);
  }
  void run_build() { GlobStat gs; AlgScan<build_t> scan(&op, &rel); scan.run(&gs); }
  std::string stats_text() const override { return stats_text_of(op.hashtable().makeStatistics()); }
  void stats(orc_stats* s) const override {
    std::memset(s, 0, sizeof(*s));
    fill_stats(op.hashtable().makeStatistics(), s);
    s->rsv_main = op.hashtable().getRsvSize();
    s->mem_dir = op.hashtable().memoryConsupmtionDir();
    s->mem_main = op.hashtable().memoryConsupmtionChains();
  }
  int kind() const override { return 0; }
  uint32_t tuple_bytes() const override { return BB; }
  uint32_t key_offset() const override { return BOff; }
  int hash_id() const override { return HashId; }
};

template <size_t BB, size_t BOff, int HashId>
struct NestedTable : RefTable {
  using tup_t   = Tup<BB>;
  using hf_t    = Hashfun<tup_t, BOff, HashId>;
  using eq_t    = Eqfun<tup_t, BOff, tup_t, BOff, HashId>;
  using build_t = AlgNestJoinBuild<hf_t, eq_t, GlobStat>;
  using main_node_t = typename build_t::hashtable_t::MainNode;
  RelationRS<tup_t> rel;
  build_t           op;
  NestedTable(const void* tuples, uint64_t n, uint64_t D) : rel(), op(D, 10, 10) {
    rel._tuples.resize(n);
    if (n) std::memcpy((void*)rel._tuples.data(), tuples, n * BB);
  }
  void run_build() { GlobStat gs; AlgScan<build_t> scan(&op, &rel); scan.run(&gs); }
  std::string stats_text() const override { return stats_text_of(op.hashtable().makeStatistics()); }
  void stats(orc_stats* s) const override {
    std::memset(s, 0, sizeof(*s));
    fill_stats(op.hashtable().makeStatistics(), s);
    s->rsv_main = op.hashtable().getRsvMainSize();
    s->rsv_sub  = op.hashtable().getRsvSubSize();
    s->mem_dir  = op.hashtable().memoryConsupmtionDir();
    s->mem_main = op.hashtable().memoryConsupmtionMainChains();
    s->mem_sub  = op.hashtable().memoryConsupmtionSubChains();
  }
  int kind() const override { return 1; }
  uint32_t tuple_bytes() const override { return BB; }
  uint32_t key_offset() const override { return BOff; }
  int hash_id() const override { return HashId; }
};

// nested intermediate (probe tuple, MainNode*)  -- cf. nested_tuple_RS_t main_experiment1.cc:331-334
template <class L, class MN> struct NestedTuple { L* _left; const MN* _right; };
template <class L, class MN>
std::ostream& operator<<(std::ostream& os, const NestedTuple<L, MN>&) { return os << "[nested]"; }

template <class L, class R, class MN> struct NestedIds {
  static std::pair<uint32_t, uint32_t> ids(const NestedTuple<L, MN>* t, const void* lb, const void* rb) {
    return { (uint32_t)(t->_left - (const L*)lb), (uint32_t)(t->_right->data() - (const R*)rb) };
  }
};

struct ProbeTimes { int64_t probe_ns; };

// mode 0 chaining non-unique, 1 chaining unique   (AlgHashJoinProbe, algebra.hh:600-672)
template <size_t BB, size_t BOff, size_t PB, size_t POff, int HashId>
static void probe_chaining(ChainingTable<BB, BOff, HashId>* tab, const void* tuples, uint64_t n,
                           bool unique, uint32_t* out, uint64_t cap, orc_counters* c, bool timing_top, int64_t* ns) {
  using btup_t = Tup<BB>; using ptup_t = Tup<PB>;
  using build_t = typename ChainingTable<BB, BOff, HashId>::build_t;
  using phf_t = Hashfun<ptup_t, POff, HashId>;
  using jp_t  = Eqfun<ptup_t, POff, btup_t, BOff, HashId>;
  using res_t = FlatTuple<ptup_t, btup_t>;
  struct Concat { using left_t = ptup_t; using right_t = btup_t; using output_t = res_t;
    inline static output_t eval(left_t* l, const right_t* r) { return {l, r}; } };
  RelationRS<ptup_t> prel;
  prel._tuples.resize(n);
  if (n) std::memcpy((void*)prel._tuples.data(), tuples, n * PB);
  GlobStat gs;
  auto run = [&](auto* probe_tag, auto& sink) {
    using probe_t = std::remove_pointer_t<decltype(probe_tag)>;
    probe_t probe(&sink, &tab->op);
    AlgScan<probe_t> scan(&probe, &prel);
    auto t0 = std::chrono::steady_clock::now();
    scan.run(&gs);
    auto t1 = std::chrono::steady_clock::now();
    if (ns) *ns = std::chrono::duration_cast<std::chrono::nanoseconds>(t1 - t0).count();
    c->matches = probe.count(); c->num_cmps = probe.numCmps();
  };
  if (timing_top) {   // the drivers' own sink: AlgTop with printing off (main_experiment1.cc:655)
    using top_t = AlgTop<res_t, GlobStat>;
    top_t top(std::cout, false, [](const res_t*, std::ostream&) {});
    std::memset(c, 0, sizeof(*c));
    if (unique) run((AlgHashJoinProbe<top_t, build_t, phf_t, jp_t, Concat, true>*)nullptr, top);
    else        run((AlgHashJoinProbe<top_t, build_t, phf_t, jp_t, Concat, false>*)nullptr, top);
    c->out_tuples = top.count();
  } else {
    using sink_t = Sink<res_t, FlatIds<ptup_t, btup_t>>;
    sink_t sink; sink._lbase = prel._tuples.data(); sink._rbase = tab->rel._tuples.data(); sink._out = out; sink._cap = cap;
    orc_counters tmp{};
    orc_counters* keep = c; c = &tmp;
    if (unique) run((AlgHashJoinProbe<sink_t, build_t, phf_t, jp_t, Concat, true>*)nullptr, sink);
    else        run((AlgHashJoinProbe<sink_t, build_t, phf_t, jp_t, Concat, false>*)nullptr, sink);
    *keep = sink._c; keep->matches = tmp.matches; keep->num_cmps = tmp.num_cmps;
  }
}

// mode 2 nested probe only, 3 nested probe + unnest  (AlgNestJoinProbe :411-473, AlgUnnestHt :489-552)
template <size_t BB, size_t BOff, size_t PB, size_t POff, int HashId>
static void probe_nested(NestedTable<BB, BOff, HashId>* tab, const void* tuples, uint64_t n, bool unnest,
                         uint32_t* out, uint64_t cap, orc_counters* c, orc_counters* cu, bool timing_top, int64_t* ns) {
  using btup_t = Tup<BB>; using ptup_t = Tup<PB>;
  using build_t = typename NestedTable<BB, BOff, HashId>::build_t;
  using ht_t   = typename build_t::hashtable_t;
  using mn_t   = typename ht_t::MainNode;
  using phf_t = Hashfun<ptup_t, POff, HashId>;
  using jp_t  = Eqfun<ptup_t, POff, btup_t, BOff, HashId>;
  using nest_t = NestedTuple<ptup_t, mn_t>;
  using res_t  = FlatTuple<ptup_t, btup_t>;
  struct Concat { using left_t = ptup_t; using right_t = mn_t; using output_t = nest_t;
    inline static output_t eval(left_t* l, const right_t* r) { return {l, r}; } };
  struct Unnest {                                   // cf. UnnestFunRS main_experiment1.cc:377-393
    using input_t = nest_t; using output_t = res_t; using MainNode = mn_t; using data_t = typename ht_t::data_t;
    inline static const MainNode* getMainNode(input_t* t) { return t->_right; }
    inline static void eval_left(output_t* o, input_t* i) { o->_left = i->_left; }
    inline static void eval_right(output_t* o, input_t*, const data_t* d) { o->_right = d; }
  };
  RelationRS<ptup_t> prel;
  prel._tuples.resize(n);
  if (n) std::memcpy((void*)prel._tuples.data(), tuples, n * PB);
  GlobStat gs;
  std::memset(c, 0, sizeof(*c));
  if (cu) std::memset(cu, 0, sizeof(*cu));
  auto timed = [&](auto& scan) {
    auto t0 = std::chrono::steady_clock::now();
    scan.run(&gs);
    auto t1 = std::chrono::steady_clock::now();
    if (ns) *ns = std::chrono::duration_cast<std::chrono::nanoseconds>(t1 - t0).count();
  };
  if (!unnest) {
    if (timing_top) {
      using top_t = AlgTop<nest_t, GlobStat>;
      top_t top(std::cout, false, [](const nest_t*, std::ostream&) {});
      using probe_t = AlgNestJoinProbe<top_t, build_t, phf_t, jp_t, Concat>;
      probe_t probe(&top, &tab->op); AlgScan<probe_t> scan(&probe, &prel); timed(scan);
      c->matches = probe.count(); c->num_cmps = probe.numCmps(); c->out_tuples = top.count();
    } else {
      using sink_t = Sink<nest_t, NestedIds<ptup_t, btup_t, mn_t>>;
      sink_t sink; sink._lbase = prel._tuples.data(); sink._rbase = tab->rel._tuples.data(); sink._out = out; sink._cap = cap;
      using probe_t = AlgNestJoinProbe<sink_t, build_t, phf_t, jp_t, Concat>;
      probe_t probe(&sink, &tab->op); AlgScan<probe_t> scan(&probe, &prel); timed(scan);
      *c = sink._c; c->matches = probe.count(); c->num_cmps = probe.numCmps();
    }
  } else {
    if (timing_top) {
      using top_t = AlgTop<res_t, GlobStat>;
      top_t top(std::cout, false, [](const res_t*, std::ostream&) {});
      using unnest_t = AlgUnnestHt<top_t, Unnest, ht_t>;
      using probe_t = AlgNestJoinProbe<unnest_t, build_t, phf_t, jp_t, Concat>;
      unnest_t un(&top); probe_t probe(&un, &tab->op); AlgScan<probe_t> scan(&probe, &prel); timed(scan);
      c->matches = probe.count(); c->num_cmps = probe.numCmps(); c->out_tuples = probe.count();
      if (cu) { cu->matches = un.count(); cu->out_tuples = top.count(); }
    } else {
      using sink_t = Sink<res_t, FlatIds<ptup_t, btup_t>>;
      sink_t sink; sink._lbase = prel._tuples.data(); sink._rbase = tab->rel._tuples.data(); sink._out = out; sink._cap = cap;
      using unnest_t = AlgUnnestHt<sink_t, Unnest, ht_t>;
      using probe_t = AlgNestJoinProbe<unnest_t, build_t, phf_t, jp_t, Concat>;
      unnest_t un(&sink); probe_t probe(&un, &tab->op); AlgScan<probe_t> scan(&probe, &prel); timed(scan);
      c->matches = probe.count(); c->num_cmps = probe.numCmps(); c->out_tuples = probe.count();
      if (cu) { *cu = sink._c; cu->matches = un.count(); }
    }
  }
}

// ---------------------------------------------------------------- runtime -> template dispatch
// supported layouts (tuple_bytes, key_offset) per hash id
//   hash 0 (u32, murmur32):          (12,0) (12,4) (8,0) (8,4)
//   hash 2 (int32 -> murmur64):      (8,0) (8,4)
//   hash 1 (u64, murmur64):          (16,0) (16,8) (24,0) (24,8)
template <int HashId, class F> static bool with_layout(uint32_t tb, uint32_t ko, F&& f) {
  if constexpr (HashId == 0) {
    if (tb == 12 && ko == 0) { f(std::integral_constant<size_t, 12>{}, std::integral_constant<size_t, 0>{}); return true; }
    if (tb == 12 && ko == 4) { f(std::integral_constant<size_t, 12>{}, std::integral_constant<size_t, 4>{}); return true; }
    if (tb == 8 && ko == 0)  { f(std::integral_constant<size_t, 8>{},  std::integral_constant<size_t, 0>{}); return true; }
    if (tb == 8 && ko == 4)  { f(std::integral_constant<size_t, 8>{},  std::integral_constant<size_t, 4>{}); return true; }
  } else if constexpr (HashId == 2) {
    if (tb == 8 && ko == 0)  { f(std::integral_constant<size_t, 8>{},  std::integral_constant<size_t, 0>{}); return true; }
    if (tb == 8 && ko == 4)  { f(std::integral_constant<size_t, 8>{},  std::integral_constant<size_t, 4>{}); return true; }
  } else {
    if (tb == 16 && ko == 0) { f(std::integral_constant<size_t, 16>{}, std::integral_constant<size_t, 0>{}); return true; }
    if (tb == 16 && ko == 8) { f(std::integral_constant<size_t, 16>{}, std::integral_constant<size_t, 8>{}); return true; }
    if (tb == 24 && ko == 0) { f(std::integral_constant<size_t, 24>{}, std::integral_constant<size_t, 0>{}); return true; }
    if (tb == 24 && ko == 8) { f(std::integral_constant<size_t, 24>{}, std::integral_constant<size_t, 8>{}); return true; }
  }
  return false;
}

template <class F> static bool with_hash(uint32_t hash_id, F&& f) {
  switch (hash_id) {
    case 0: f(std::integral_constant<int, 0>{}); return true;
    case 1: f(std::integral_constant<int, 1>{}); return true;
    case 2: f(std::integral_constant<int, 2>{}); return true;
  }
  return false;
}

static std::vector<unsigned char> gathered(const void* tuples, uint64_t n, uint32_t tb, const uint32_t* gather) {
  std::vector<unsigned char> v((size_t)n * tb);
  for (uint64_t i = 0; i < n; ++i) std::memcpy(v.data() + i * tb, (const unsigned char*)tuples + (uint64_t)gather[i] * tb, tb);
  return v;
}

}  // namespace

extern "C" {

void* ref_build(int kind, const void* tuples, uint64_t n, orc_keyspec ks, uint64_t D) {
  RefTable* res = nullptr;
  with_hash(ks.hash_id, [&](auto H) {
    with_layout<H.value>(ks.tuple_bytes, ks.key_offset, [&](auto BB, auto BO) {
      if (kind == 0) { auto* t = new ChainingTable<BB.value, BO.value, H.value>(tuples, n, D); t->run_build(); res = t; }
      else           { auto* t = new NestedTable<BB.value, BO.value, H.value>(tuples, n, D);   t->run_build(); res = t; }
    });
  });
  return res;
}

void ref_free(void* h) { delete (RefTable*)h; }

void ref_stats(void* h, orc_stats* s) { ((RefTable*)h)->stats(s); }

/* the reference's own HtStatistics::print / toCsvString / toCsvStringHeader bytes (ht_statistics.cc:16-79), '\x1e' separated */
uint64_t ref_stats_text(void* h, char* buf, uint64_t cap) {
  const std::string s = ((RefTable*)h)->stats_text();
  if (buf && cap) { const uint64_t n = s.size() < cap - 1 ? s.size() : cap - 1; std::memcpy(buf, s.data(), n); buf[n] = 0; }
  return s.size();
}

/* mode: 0 chaining, 1 chaining IsBuildKeyUnique, 2 nested (no unnest), 3 nested + unnest.
 * timing_top != 0: use the drivers' AlgTop sink (no checksum / materialisation) and report the
 * probe strand's wall time in *ns.  Returns 0 on success, -1 for an unsupported layout. */
int ref_probe(void* h, const void* tuples, uint64_t n, orc_keyspec ks, const uint32_t* gather, int mode,
              uint32_t* out, uint64_t cap, orc_counters* c, orc_counters* c_unnest, int timing_top, int64_t* ns) {
  RefTable* t = (RefTable*)h;
  std::vector<unsigned char> tmp;
  if (gather) { tmp = gathered(tuples, n, ks.tuple_bytes, gather); tuples = tmp.data(); }
  bool ok = false;
  with_hash((uint32_t)t->hash_id(), [&](auto H) {
    with_layout<H.value>(t->tuple_bytes(), t->key_offset(), [&](auto BB, auto BO) {
      with_layout<H.value>(ks.tuple_bytes, ks.key_offset, [&](auto PB, auto PO) {
        if (t->kind() == 0 && mode <= 1) {
          probe_chaining<BB.value, BO.value, PB.value, PO.value, H.value>(
              (ChainingTable<BB.value, BO.value, H.value>*)t, tuples, n, mode == 1, out, cap, c, timing_top != 0, ns);
          ok = true;
        } else if (t->kind() == 1 && mode >= 2) {
          probe_nested<BB.value, BO.value, PB.value, PO.value, H.value>(
              (NestedTable<BB.value, BO.value, H.value>*)t, tuples, n, mode == 3, out, cap, c, c_unnest, timing_top != 0, ns);
          ok = true;
        }
      });
    });
  });
  return ok ? 0 : -1;
}

/* build strand timed like the drivers (steady_clock around AlgScan::run, main_experiment1.cc:665-667) */
void* ref_build_timed(int kind, const void* tuples, uint64_t n, orc_keyspec ks, uint64_t D, int64_t* ns) {
  RefTable* res = nullptr;
  with_hash(ks.hash_id, [&](auto H) {
    with_layout<H.value>(ks.tuple_bytes, ks.key_offset, [&](auto BB, auto BO) {
      auto go = [&](auto* t) {
        auto t0 = std::chrono::steady_clock::now();
        t->run_build();
        auto t1 = std::chrono::steady_clock::now();
        if (ns) *ns = std::chrono::duration_cast<std::chrono::nanoseconds>(t1 - t0).count();
        res = t;
      };
      if (kind == 0) go(new ChainingTable<BB.value, BO.value, H.value>(tuples, n, D));
      else           go(new NestedTable<BB.value, BO.value, H.value>(tuples, n, D));
    });
  });
  return res;
}

/* ---- input generators: Experiment1::init (main_experiment1.cc:415-457) ---- */
uint64_t ref_gen_exp1(uint32_t log2R, uint32_t log2S, int skew, uint32_t t, uint32_t* Rk, uint32_t* Sk, uint32_t* Sa) {
  const size_t cardR = 1U << log2R, cardS = 1U << log2S;
  std::mt19937 rng;
  std::vector<uint32_t> keysR(cardR);
  for (uint32_t i = 0; i < cardR; ++i) keysR[i] = i;
  std::shuffle(keysR.begin(), keysR.end(), rng);
  std::vector<uint32_t> fk;
  const uint32_t fkMax = (1 << (log2R - t));
  GenRandIntVec griv;
  GenRandIntVec::param_t p = skew
    ? GenRandIntVec::param_t(GenRandIntVec::dist_t::kZipf, fkMax, 0, 1.0, 0, -1)
    : GenRandIntVec::param_t(GenRandIntVec::dist_t::kUni,  fkMax, 0, 0.0, 0, -1);
  griv.generate(fk, cardS, p, rng);
  std::memcpy(Rk, keysR.data(), cardR * 4);
  for (uint32_t i = 0; i < cardS; ++i) { Sk[i] = i; Sa[i] = fk[i]; }
  return std::unordered_set<uint32_t>(fk.cbegin(), fk.cend()).size();   // _numDvSa
}

/* ---- Experiment4::init (main_experiment4.cc:517-575), aShuffle = true ---- */
void ref_gen_exp4(uint32_t log2R, uint32_t alpha, uint32_t mA, uint32_t beta, uint32_t mB,
                  uint32_t* Rk, uint32_t* Sk, uint32_t* Sa, uint32_t* Tk, uint32_t* Ta) {
  const size_t cardR = 1U << log2R;
  const size_t nC = cardR / (1U << alpha), nE = cardR / (1U << beta);
  const size_t cC = nC * mA, cE = nE * mB, cF = cC + cE;
  std::mt19937 rng;
  std::vector<uint32_t> fkC(cC), fkES(cE), fkET(cE);
  uint32_t val = 0; size_t idx = 0;
  for (; val < nC; ++val) for (size_t i = 0; i < mA; ++i) fkC[idx++] = val;
  idx = 0;
  for (; val < nC + nE; ++val) for (size_t i = 0; i < mB; ++i) fkES[idx++] = val;
  idx = 0;
  for (; val < nC + 2 * nE; ++val) for (size_t i = 0; i < mB; ++i) fkET[idx++] = val;
  for (size_t i = 0; i < cardR; ++i) Rk[i] = (uint32_t)i;
  std::shuffle(fkES.begin(), fkES.end(), rng);
  std::shuffle(fkET.begin(), fkET.end(), rng);
  std::shuffle(fkC.begin(), fkC.end(), rng);
  for (size_t i = 0; i < cF; ++i) { Sk[i] = (uint32_t)i; Sa[i] = i < cC ? fkC[i] : fkES[i - cC]; }
  std::shuffle(fkC.begin(), fkC.end(), rng);
  for (size_t i = 0; i < cF; ++i) { Tk[i] = (uint32_t)i; Ta[i] = i < cC ? fkC[i] : fkET[i - cC]; }
}

}  // extern "C"

/* ---- Experiment4 plans on 8-byte {k,a} tuples (main_experiment4.cc:831-1043) ---- */
namespace exp4 {
using base_t = Tup<8>;
using HashfunR     = Hashfun<base_t, 0, 0>;
using HashfunFkRel = Hashfun<base_t, 4, 0>;
using EqfunBuild   = Eqfun<base_t, 4, base_t, 4, 0>;
using JoinpredRS   = Eqfun<base_t, 0, base_t, 4, 0>;
using nbuild_t = AlgNestJoinBuild<HashfunFkRel, EqfunBuild, GlobStat>;
using cbuild_t = AlgHashJoinBuild<HashfunFkRel, EqfunBuild, GlobStat>;
using mn_t = nbuild_t::hashtable_t::MainNode;

struct nested_RS  { base_t* _r; const mn_t* _s; };
struct nested_RST { base_t* _r; const mn_t* _s; const mn_t* _t; };
struct R_nS_xT    { base_t* _r; const mn_t* _s; const base_t* _t; };
struct result_t   { const base_t* _r; const base_t* _s; const base_t* _t; };
struct result_RS  { const base_t* _r; const base_t* _s; };
std::ostream& operator<<(std::ostream& os, const nested_RS&)  { return os; }
std::ostream& operator<<(std::ostream& os, const nested_RST&) { return os; }
std::ostream& operator<<(std::ostream& os, const R_nS_xT&)    { return os; }
std::ostream& operator<<(std::ostream& os, const result_t&)   { return os; }
std::ostream& operator<<(std::ostream& os, const result_RS&)  { return os; }

struct ConcatNested_RS { using left_t = base_t; using right_t = const mn_t; using output_t = nested_RS;
  static output_t eval(left_t* l, const right_t* r) { return {l, r}; } };
struct HashfunNestedRS { using input_t = nested_RS; using output_t = uint32_t;
  static output_t eval(const input_t* t) { return HashfunR::eval(t->_r); } };
struct JoinpredRTnested { using left_t = nested_RS; using right_t = base_t;
  static bool eval(const left_t* l, const right_t* r) { return JoinpredRS::eval(l->_r, r); } };
struct ConcatNested_RST { using left_t = nested_RS; using right_t = const mn_t; using output_t = nested_RST;
  static output_t eval(left_t* l, const right_t* r) { return {l->_r, l->_s, r}; } };
struct Unnest_R_nS_xT { using input_t = nested_RST; using output_t = R_nS_xT; using MainNode = mn_t; using data_t = base_t;
  static const MainNode* getMainNode(input_t* t) { return t->_t; }
  static void eval_left(output_t* o, input_t* i) { o->_r = i->_r; o->_s = i->_s; }
  static void eval_right(output_t* o, input_t*, const data_t* d) { o->_t = d; } };
struct Unnest_R_xS_xT { using input_t = R_nS_xT; using output_t = result_t; using MainNode = mn_t; using data_t = base_t;
  static const MainNode* getMainNode(input_t* t) { return t->_s; }
  static void eval_left(output_t* o, input_t* i) { o->_r = i->_r; o->_t = i->_t; }
  static void eval_right(output_t* o, input_t*, const data_t* d) { o->_s = d; } };
struct HashfunRS { using input_t = result_RS; using output_t = uint32_t;
  static output_t eval(const input_t* t) { return HashfunR::eval(t->_r); } };
struct Joinpred_RS_T { using left_t = result_RS; using right_t = base_t;
  static bool eval(const left_t* l, const right_t* r) { return JoinpredRS::eval(l->_r, r); } };
struct ConcatCh_RS { using left_t = base_t; using right_t = base_t; using output_t = result_RS;
  static output_t eval(left_t* l, const right_t* r) { return {l, r}; } };
struct ConcatCh_RS_T { using left_t = result_RS; using right_t = base_t; using output_t = result_t;
  static output_t eval(left_t* l, const right_t* r) { return {l->_r, l->_s, r}; } };

// sink folding (r,s,t) row ids into a checksum: mix(mix32(r,s), t)
class Sink3 : public AlgBase {
  public:
    using globstat_t = GlobStat; using input_t = result_t; using output_t = void;
    Sink3() : AlgBase("Sink3") {}
    void init(globstat_t*) { reset(); }
    void step(input_t* t, globstat_t*) {
      inc();
      uint32_t r = (uint32_t)(t->_r - R), s = (uint32_t)(t->_s - S), tt = (uint32_t)(t->_t - T);
      uint64_t m = orc_pair_mix((uint32_t)orc_pair_mix(r, s), tt);
      sum += m; x ^= m;
    }
    void fin(globstat_t*) { stopTimer(); }
    const base_t *R = nullptr, *S = nullptr, *T = nullptr;
    uint64_t sum = 0, x = 0;
};
}  // namespace exp4

extern "C" {
/* plan 0 = Ndu (main_experiment4.cc:831-941), 1 = Chj (:943-1043).
 * out[0..11] = c_probe_RS, c_probe_RS_cmp, c_probe_RT, c_probe_RT_cmp, c_unnest1, c_unnest2, c_top,
 *              checksum_sum, checksum_xor, t_build_S ns, t_build_T ns, t_probe ns */
void ref_exp4_run(int plan, const void* R, uint64_t nR, const void* S, uint64_t nS, const void* T, uint64_t nT,
                  uint64_t D, uint64_t* out) {
  using namespace exp4;
  using clk = std::chrono::steady_clock;
  RelationRS<base_t> rR, rS, rT;
  rR._tuples.resize(nR); rS._tuples.resize(nS); rT._tuples.resize(nT);
  std::memcpy((void*)rR._tuples.data(), R, nR * 8); std::memcpy((void*)rS._tuples.data(), S, nS * 8); std::memcpy((void*)rT._tuples.data(), T, nT * 8);
  GlobStat gs;
  Sink3 top; top.R = rR._tuples.data(); top.S = rS._tuples.data(); top.T = rT._tuples.data();
  auto ns = [](clk::time_point a, clk::time_point b) { return (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(b - a).count(); };
  if (plan == 0) {
    using unnest_2_t = AlgUnnestHt<Sink3, Unnest_R_xS_xT, nbuild_t::hashtable_t>;
    using unnest_1_t = AlgUnnestHt<unnest_2_t, Unnest_R_nS_xT, nbuild_t::hashtable_t>;
    using probe_RT_t = AlgNestJoinProbe<unnest_1_t, nbuild_t, HashfunNestedRS, JoinpredRTnested, ConcatNested_RST>;
    using probe_RS_t = AlgNestJoinProbe<probe_RT_t, nbuild_t, HashfunR, JoinpredRS, ConcatNested_RS>;
    nbuild_t bS(D, 10, 10), bT(D, 10, 10);
    AlgScan<nbuild_t> scS(&bS, &rS), scT(&bT, &rT);
    unnest_2_t u2(&top); unnest_1_t u1(&u2); probe_RT_t pRT(&u1, &bT); probe_RS_t pRS(&pRT, &bS);
    AlgScan<probe_RS_t> scR(&pRS, &rR);
    auto t0 = clk::now(); scS.run(&gs); auto t1 = clk::now(); scT.run(&gs); auto t2 = clk::now(); scR.run(&gs); auto t3 = clk::now();
    out[0] = pRS.count(); out[1] = pRS.numCmps(); out[2] = pRT.count(); out[3] = pRT.numCmps();
    out[4] = u1.count(); out[5] = u2.count(); out[6] = top.count();
    out[9] = ns(t0, t1); out[10] = ns(t1, t2); out[11] = ns(t2, t3);
  } else {
    using probe_RT_t = AlgHashJoinProbe<Sink3, cbuild_t, HashfunRS, Joinpred_RS_T, ConcatCh_RS_T>;
    using probe_RS_t = AlgHashJoinProbe<probe_RT_t, cbuild_t, HashfunR, JoinpredRS, ConcatCh_RS>;
    cbuild_t bS(D, 10), bT(D, 10);
    AlgScan<cbuild_t> scS(&bS, &rS), scT(&bT, &rT);
    probe_RT_t pRT(&top, &bT); probe_RS_t pRS(&pRT, &bS);
    AlgScan<probe_RS_t> scR(&pRS, &rR);
    auto t0 = clk::now(); scS.run(&gs); auto t1 = clk::now(); scT.run(&gs); auto t2 = clk::now(); scR.run(&gs); auto t3 = clk::now();
    out[0] = pRS.count(); out[1] = pRS.numCmps(); out[2] = pRT.count(); out[3] = pRT.numCmps();
    out[4] = 0; out[5] = 0; out[6] = top.count();
    out[9] = ns(t0, t1); out[10] = ns(t1, t2); out[11] = ns(t2, t3);
  }
  out[7] = top.sum; out[8] = top.x;
}
}  // extern "C"
