// hj3d/algebra.hh -- the physical algebra of the reference (algebra.hh:14-672: templated, push based, same
// operator names, template parameters, constructors and accessors) as thin shims over the device engine.
//
//   AlgScan -> Alg{Hash,Nest}JoinBuild                       build strand: the table is built on the device in fin()
//   AlgScan -> [AlgSelection] -> Alg{Hash,Nest}JoinProbe -> [AlgUnnestHt] -> AlgTop      probe strand
//
// The tuple-at-a-time protocol (init / step / fin) is kept; join operators collect the tuples pushed into
// them and run the device kernels for the whole batch in fin().  Between two device operators batches stay
// on the device (probe -> unnest hands over (left, group_ref) columns); a consumer that is not a device
// operator is fed one tuple at a time, exactly like in the reference.  AlgTop without printing only counts,
// so no result tuple is materialised on the host for it.
#pragma once

#include <algorithm>
#include <chrono>
#include <deque>
#include <functional>
#include <iomanip>
#include <iostream>
#include <string>
#include <vector>

#include "concepts.hh"
#include "ht.hh"
#include "runtime.hh"

class AlgBase;

template <typename T>
concept alg_operator_c = std::derived_from<T, AlgBase> && requires {
  typename T::globstat_t; typename T::input_t; typename T::output_t;
};
template <typename T>
concept alg_consumer_c = alg_operator_c<T> && requires(T c) {
  { c.init(static_cast<typename T::globstat_t*>(nullptr)) } -> std::same_as<void>;
  { c.step(static_cast<typename T::input_t*>(nullptr), static_cast<typename T::globstat_t*>(nullptr)) } -> std::same_as<void>;
  { c.fin(static_cast<typename T::globstat_t*>(nullptr)) } -> std::same_as<void>;
};
template <typename T>
concept alg_producer_c = alg_operator_c<T> && requires(T c) {
  { c.run(static_cast<typename T::globstat_t*>(nullptr)) } -> std::same_as<void>;
};
template <typename T>
concept alg_buildop_c = alg_consumer_c<T> && requires(T t) {
  typename T::hashtable_t;
  { t.hashtable() } -> std::same_as<const typename T::hashtable_t&>;
};

// a row store relation (algebra.hh:98-106)
template <typename Ttuple>
struct RelationRS {
  using tuple_t = Ttuple;
  using tuple_vt = std::vector<tuple_t>;
  tuple_vt _tuples;
  inline size_t card() const { return _tuples.size(); }
};
template <typename Ttuple>
std::ostream& operator<<(std::ostream& os, const RelationRS<Ttuple>& aRel) {
  for (const auto& t : aRel._tuples) os << t << "\n";
  return os;
}

struct GlobStat0 {   // algebra.hh:118-123
  size_t _ht_num_buckets, _ht_rsv_log2_chunksize_main, _ht_rsv_log2_chunksize_sub, _ht_rsv_log2_chunksize;
};

// common base class of all operators (algebra.hh:166-200)
class AlgBase {
  public:
    using clock_t = std::chrono::steady_clock;
    using time_point_t = std::chrono::time_point<clock_t>;
    using duration_t = std::chrono::nanoseconds;
    explicit AlgBase(const std::string& aName) : _count(0), _ok(true), _startTime(), _stopTime(), _name(aName), _runs(0) {}
    AlgBase() : AlgBase("") {}
    inline void     reset() { _count = 0; _ok = true; startTimer(); ++_runs; }
    inline void     inc() { ++_count; }
    inline void     inc(uint64_t n) { _count += n; }
    inline uint64_t count() const { return _count; }
    inline bool     ok() const { return _ok; }
    inline bool     ok(const bool b) { return (_ok = b); }
    inline void     startTimer() { _startTime = clock_t::now(); }
    inline void     stopTimer()  { _stopTime = clock_t::now(); }
    inline const std::string& name() const { return _name; }
    inline uint64_t runs() const { return _runs; }
    inline duration_t getRuntime() const { return std::chrono::duration_cast<duration_t>(_stopTime - _startTime); }
    inline std::string getRuntimeStr() const { return std::to_string(getRuntime().count()) + " ns"; }
    // Device operators do their work in one batch when their strand ends, so the reference's inclusive wall-clock
    // timers (init .. fin) and their differences (get_runtime_excl, algebra.hh:129-138) say nothing about them.  They
    // record the CUDA-event time of their own kernels instead (hj3d_ctx_timings); that IS an exclusive runtime.
    inline void       setDeviceTime(double ms) { _deviceNs = (int64_t)(ms * 1e6); }
    inline bool       hasDeviceTime() const { return _deviceNs >= 0; }
    inline duration_t deviceTime() const { return duration_t(_deviceNs < 0 ? 0 : _deviceNs); }
  protected:
    int64_t      _deviceNs = -1;
    uint64_t     _count;
    bool         _ok;
    time_point_t _startTime, _stopTime;
    std::string  _name;
    uint64_t     _runs;
};

template <alg_operator_c Toperator>
auto get_runtime_excl(const Toperator* aOp) -> typename Toperator::duration_t {   // algebra.hh:129-138
  if (aOp->hasDeviceTime()) return aOp->deviceTime();                    // device operator: its kernels' CUDA-event time
  if constexpr (requires(Toperator t) { t.consumer(); }) {
    if (aOp->consumer()->hasDeviceTime()) return typename Toperator::duration_t(0);   // host hand-over stage in front of a device operator
    return (aOp->getRuntime() - aOp->consumer()->getRuntime());
  } else return aOp->getRuntime();
}
template <alg_operator_c Toperator>
void print_strand(const Toperator* aOp, const size_t aIndentLvl = 0, std::ostream& os = std::cout) {   // algebra.hh:148-162
  if constexpr (requires(Toperator t) { t.consumer(); }) print_strand(aOp->consumer(), aIndentLvl, os);
  os << std::setw((int)(aIndentLvl * 2)) << "" << aOp->name() << "|" << aOp->count() << "|"
     << std::to_string(get_runtime_excl(aOp).count()) << " ns|" << aOp->runs() << "\n";
}

namespace hj3d::detail {
struct DevClock { double ms = 0; inline void tick(); };
inline double last_call_device_ms() {
  hj3d_timings t{};
  check(hj3d_ctx_timings(Runtime::instance().ctx(), &t));
  return t.total_ms;
}
inline void DevClock::tick() { ms += last_call_device_ms(); }
// bulk / device hand-over hooks an operator may offer to its producer
template <class C, class In> concept takes_bulk = requires(C c, In* p, size_t n, typename C::globstat_t* g) { c.step_bulk(p, n, g); };
template <class C> concept counts_only = requires(C c, uint64_t n) { { c.wants_tuples() } -> std::same_as<bool>; c.add_count(n); };
}  // namespace hj3d::detail

// Top operator (algebra.hh:204-243)
template <typename Tinput, typename Tglobstat>
class AlgTop : public AlgBase {
  public:
    using globstat_t = Tglobstat; using input_t = Tinput; using output_t = void;
    using print_fun_t = std::function<void(const input_t*, std::ostream& os)>;
    inline AlgTop() : AlgTop(std::cout, true) {}
    inline AlgTop(std::ostream& aOs, const bool aPrintResult) : AlgBase("AlgTop"), _os(aOs), _print_result(aPrintResult) {}
    inline AlgTop(std::ostream& aOs, const bool aPrintResult, print_fun_t aPrintFunction)
      : AlgBase("AlgTop"), _os(aOs), _print_result(aPrintResult), _print_fun(aPrintFunction) {}
    inline void init([[maybe_unused]] globstat_t* aGlobstat) { reset(); }
    inline void step(input_t* aInput, [[maybe_unused]] globstat_t* aGlobstat) {
      inc();
      if (_print_result && runs() == 1) { _print_fun(aInput, _os); _os << "\n"; }
    }
    inline void fin([[maybe_unused]] globstat_t* aGlobstat) { stopTimer(); }
    bool printResult() const { return _print_result; }
    void printResult(const bool aPrint) { _print_result = aPrint; }
    // device producers: does this run need the tuples themselves, or only their number?
    bool wants_tuples() const { return _print_result && runs() == 1; }
    void add_count(uint64_t n) { inc(n); }
  private:
    std::ostream& _os;
    bool          _print_result;
    print_fun_t   _print_fun = [](const input_t* aInput, std::ostream& aOs) { aOs << aInput; };
};

// Table scan (algebra.hh:247-275)
template <alg_consumer_c Tconsumer>
class AlgScan : public AlgBase {
  public:
    using consumer_t = Tconsumer; using globstat_t = typename consumer_t::globstat_t;
    using input_t = typename consumer_t::input_t; using output_t = typename consumer_t::input_t;
    using input_rel_t = RelationRS<input_t>;
    inline AlgScan(consumer_t* aConsumer, input_rel_t* aRelation) : AlgBase("AlgScan"), _consumer(aConsumer), _relation(aRelation) {}
    inline void run(globstat_t* aGlobstat) {
      reset();
      _consumer->init(aGlobstat);
      if constexpr (hj3d::detail::takes_bulk<consumer_t, input_t>) {     // device operator: hand over the whole slab
        inc(_relation->_tuples.size());
        _consumer->step_bulk(_relation->_tuples.data(), _relation->_tuples.size(), aGlobstat);
      } else {
        for (auto& t : _relation->_tuples) { inc(); _consumer->step(&t, aGlobstat); }
      }
      _consumer->fin(aGlobstat);
      stopTimer();
    }
    inline const consumer_t* consumer() const { return _consumer; }
  private:
    consumer_t*  _consumer;
    input_rel_t* _relation;
};

// Selection (algebra.hh:279-315)
template <alg_consumer_c Tconsumer, alg_predicate_c Tpredicate>
class AlgSelection : public AlgBase {
  public:
    using consumer_t = Tconsumer; using globstat_t = typename consumer_t::globstat_t;
    using input_t = typename consumer_t::input_t; using output_t = typename consumer_t::input_t; using predicate_t = Tpredicate;
    inline AlgSelection(Tconsumer* aConsumer) : AlgBase("AlgSelection"), _consumer(aConsumer) {}
    inline void init(globstat_t* aGlobstat) { reset(); _consumer->init(aGlobstat); }
    inline void step(input_t* aInput, globstat_t* aGlobstat) {
      if (predicate_t::eval(aInput)) { inc(); _consumer->step(aInput, aGlobstat); }
    }
    inline void fin(globstat_t* aGlobstat) { _consumer->fin(aGlobstat); stopTimer(); }
    inline const consumer_t* consumer() const { return _consumer; }
  private:
    consumer_t* _consumer;
};

// Dynamic selection (algebra.hh:319-358)
template <alg_consumer_c Tconsumer, alg_dyn_predicate_c Tpredicate>
class AlgDynSelection : public AlgBase {
  public:
    using consumer_t = Tconsumer; using globstat_t = typename consumer_t::globstat_t;
    using input_t = typename consumer_t::input_t; using output_t = typename consumer_t::input_t; using predicate_t = Tpredicate;
    inline AlgDynSelection(consumer_t* aConsumer, predicate_t aPredicate) : AlgBase("AlgDynSelection"), _consumer(aConsumer), _pred(aPredicate) {}
    inline AlgDynSelection(consumer_t* aConsumer) : AlgDynSelection(aConsumer, predicate_t()) {}
    inline void init(globstat_t* aGlobstat) { reset(); _consumer->init(aGlobstat); }
    inline void step(input_t* aInput, globstat_t* aGlobstat) {
      if (_pred(aInput)) { inc(); _consumer->step(aInput, aGlobstat); }
    }
    inline void fin(globstat_t* aGlobstat) { _consumer->fin(aGlobstat); stopTimer(); }
    inline const consumer_t* consumer() const { return _consumer; }
  private:
    consumer_t* _consumer;
    predicate_t _pred;
};

// ---- build operators (algebra.hh:362-401, 556-586) ---------------------------------------------------
namespace hj3d::detail {
template <class Ttable, class Tglobstat>
class BuildOp : public AlgBase {
  public:
    using globstat_t = Tglobstat; using hashtable_t = Ttable; using input_t = typename Ttable::data_t; using output_t = void;
    template <class... A> BuildOp(const char* name, A... a) : AlgBase(name), _hashtable(a...) {}
    inline void init([[maybe_unused]] globstat_t* g) { reset(); }
    inline void step(input_t* aInput, [[maybe_unused]] globstat_t* g) { inc(); _hashtable.insert(aInput); }
    inline void step_bulk(input_t* first, size_t n, [[maybe_unused]] globstat_t* g) { inc(n); _hashtable.insert_bulk(first, n); }
    inline void fin([[maybe_unused]] globstat_t* g) {                                     // the device build happens here
      if (_hashtable.seal()) setDeviceTime(hj3d::detail::last_call_device_ms());
      stopTimer();
    }
    inline const hashtable_t& hashtable() const { return _hashtable; }
    inline void clear_ht() { _hashtable.clear(); }
  protected:
    hashtable_t _hashtable;
};
}  // namespace hj3d::detail

template <alg_hashfun_c Thashfun, alg_binary_predicate_c Tequalfun, typename Tglobstat>
class AlgNestJoinBuild : public hj3d::detail::BuildOp<HtNested1<typename Thashfun::input_t, Thashfun, Tequalfun>, Tglobstat> {
  using base_t = hj3d::detail::BuildOp<HtNested1<typename Thashfun::input_t, Thashfun, Tequalfun>, Tglobstat>;
  public:
    using hashfun_t = Thashfun; using eqfun_t = Tequalfun;
    AlgNestJoinBuild(const size_t aHashDirSize, const uint32_t aHtLog2ChunkSizeMain, const uint32_t aHtLog2ChunkSizeSub)
      : base_t("AlgNestJoinBuild", aHashDirSize, aHtLog2ChunkSizeMain, aHtLog2ChunkSizeSub) {}
    AlgNestJoinBuild(const Tglobstat* g)
      : AlgNestJoinBuild(g->_ht_num_buckets, g->_ht_rsv_log2_chunksize_main, g->_ht_rsv_log2_chunksize_sub) {}
};

template <alg_hashfun_c Thashfun, alg_binary_predicate_c Tequalfun, typename Tglobstat>
class AlgHashJoinBuild : public hj3d::detail::BuildOp<HtChaining1<typename Thashfun::input_t, Thashfun, Tequalfun>, Tglobstat> {
  using base_t = hj3d::detail::BuildOp<HtChaining1<typename Thashfun::input_t, Thashfun, Tequalfun>, Tglobstat>;
  public:
    using hashfun_t = Thashfun; using eqfun_t = Tequalfun;
    AlgHashJoinBuild(const size_t aHashDirSize, const uint32_t aHtLog2ChunkSize) : base_t("AlgHashJoinBuild", aHashDirSize, aHtLog2ChunkSize) {}
    AlgHashJoinBuild(const Tglobstat* g) : AlgHashJoinBuild(g->_ht_num_buckets, g->_ht_rsv_log2_chunksize) {}
};

// ---- probe side ----------------------------------------------------------------------------------------
namespace hj3d::detail {

// A probe hash functor may reach its key through a pointer of an intermediate tuple
// (HashfunNestedRS / HashfunRS, main_experiment4.cc:355-361,413-419).  Such functors name the functor of the
// base tuple (`hj3d_base`) and how to get there (`hj3d_deref`); the base tuples are staged and probed.
template <class H> concept derefs_to_base = requires(const typename H::input_t* t) {
  typename H::hj3d_base;
  { H::hj3d_deref(t) } -> std::convertible_to<const typename H::hj3d_base::input_t*>;
};

// device columns a nested probe hands to a device unnest
struct NestedBatch { const uint32_t* d_left; const uint32_t* d_gref; uint64_t n; };

template <class Hashfun>
struct ProbeInput {
  using input_t = typename Hashfun::input_t;
  TupleSeq<input_t> seq;
  std::deque<std::remove_const_t<input_t>> owned;   // tuples pushed one at a time may be the producer's scratch tuple
  DevBuf dbuf;
  std::vector<std::remove_const_t<input_t>> staging;
  hj3d_keyspec ks{};
  void clear() { seq.clear(); owned.clear(); }
  // step(): the pointee is only valid during the call (algebra.hh:455-457,650-652 reuse one _outputTuple) -> keep a copy
  void push_copy(input_t* t) { owned.push_back(*t); seq.push(&owned.back()); }
  // returns the device pointer of the tuples whose key the functor hashes
  const void* upload() {
    if constexpr (derefs_to_base<Hashfun>) {
      using base_in = typename Hashfun::hj3d_base::input_t;
      ks = keyspec_of<typename Hashfun::hj3d_base>();
      std::vector<std::remove_const_t<base_in>> tmp(seq.size());
      for (size_t i = 0; i < seq.size(); ++i) tmp[i] = *Hashfun::hj3d_deref(seq.at(i));
      void* d = dbuf.ensure(tmp.size() * sizeof(base_in));
      check(hj3d_memcpy_h2d(Runtime::instance().ctx(), d, tmp.data(), tmp.size() * sizeof(base_in)));
      check(hj3d_ctx_sync(Runtime::instance().ctx()));          // tmp dies at scope exit
      return d;
    } else {
      ks = keyspec_of<Hashfun>();
      const void* d = seq.upload(dbuf, staging);
      return d;
    }
  }
};

}  // namespace hj3d::detail

// 3D hash join unnest (algebra.hh:489-552)
template <alg_consumer_c Tconsumer, alg_unnestfun_c Tunnestfun, typename Thtnested>
class AlgUnnestHt : public AlgBase {
  public:
    using consumer_t = Tconsumer; using globstat_t = typename consumer_t::globstat_t; using unnestfun_t = Tunnestfun;
    using output_t = typename consumer_t::input_t; using input_t = typename unnestfun_t::input_t; using ht_nested_t = Thtnested;
    static_assert(std::is_same_v<output_t, typename unnestfun_t::output_t>,
                  "AlgUnnestHt::output_t (aka consumer_t::input_t) does not match unnestfun_t::output_t");
    inline AlgUnnestHt(consumer_t* aConsumer) : AlgBase("AlgUnnest"), _consumer(aConsumer), _outputTuple() {}
    inline void init([[maybe_unused]] globstat_t* g) { reset(); _clk.ms = 0; _consumer->init(g); _in.clear(); _dev = {nullptr, nullptr, 0}; _table = nullptr; }
    inline void step(input_t* aNestedTuple, [[maybe_unused]] globstat_t* g) { _in.push_back(*aNestedTuple); }
    // device hand-over from AlgNestJoinProbe: (left id, group_ref) columns + how to rebuild a nested tuple
    using make_nested_t = std::function<std::remove_const_t<input_t>(uint32_t)>;
    inline void step_device(const hj3d::detail::NestedBatch& b, const ht_nested_t* table, make_nested_t make, globstat_t* g) {
      run_device(table, b.d_left, b.d_gref, b.n, make, g);
    }
    // AlgNestJoinProbe ran probe + unnest as one device call (hj3d_probe_nested_unnest): only the count arrives here
    inline void take_fused_count(uint64_t m) {
      inc(m);
      if constexpr (hj3d::detail::counts_only<consumer_t>) _consumer->add_count(m);
    }
    // false: only the number of unnested tuples is needed downstream, nothing has to come back to the host
    inline bool needs_host_tuples() const {
      if constexpr (hj3d::detail::counts_only<consumer_t>) return _consumer->wants_tuples();
      else return true;
    }
    inline void fin(globstat_t* g) {
      if (!_in.empty()) {                                        // nested tuples came one at a time from a host operator
        using namespace hj3d;
        const size_t n = _in.size();
        std::vector<uint32_t> left(n), gref(n);
        const ht_nested_t* table = nullptr;
        for (size_t i = 0; i < n; ++i) {
          const auto* m = unnestfun_t::getMainNode(&_in[i]);
          left[i] = (uint32_t)i; gref[i] = m->group_ref(); table = m->owner();
        }
        hj3d_ctx* c = Runtime::instance().ctx();
        auto* dl = (uint32_t*)_dl.ensure(n * 4); auto* dg = (uint32_t*)_dg.ensure(n * 4);
        check(hj3d_memcpy_h2d(c, dl, left.data(), n * 4)); check(hj3d_memcpy_h2d(c, dg, gref.data(), n * 4));
        check(hj3d_ctx_sync(c));
        run_device(table, dl, dg, n, make_nested_t([this](uint32_t i) { return _in[i]; }), g);
      }
      _consumer->fin(g);
      setDeviceTime(_clk.ms); stopTimer();
    }
    inline const consumer_t* consumer() const { return _consumer; }
  private:
    hj3d::detail::DevClock _clk;
    void run_device(const ht_nested_t* table, const uint32_t* d_left, const uint32_t* d_gref, uint64_t n, const make_nested_t& make, globstat_t* g) {
      using namespace hj3d;
      hj3d_ctx* c = Runtime::instance().ctx();
      hj3d_counters cnt{};
      bool need_tuples = true;
      if constexpr (detail::counts_only<consumer_t>) need_tuples = _consumer->wants_tuples();
      if (!need_tuples) {
        check(hj3d_unnest(c, table->handle(), d_left, d_gref, n, 0, nullptr, 0, &cnt)); _clk.tick();
        inc(cnt.out_tuples);
        if constexpr (detail::counts_only<consumer_t>) _consumer->add_count(cnt.out_tuples);
        return;
      }
      check(hj3d_unnest(c, table->handle(), d_left, d_gref, n, 0, nullptr, 0, &cnt)); _clk.tick();           // size of the result
      const uint64_t m = cnt.out_tuples;
      auto* dout = (uint32_t*)_dout.ensure(m * 8 + 8);
      check(hj3d_unnest(c, table->handle(), d_left, d_gref, n, 0, dout, m, &cnt)); _clk.tick();
      std::vector<uint32_t> pairs(2 * m);
      check(hj3d_memcpy_d2h(c, pairs.data(), dout, m * 8));
      // emission order of the reference (algebra.hh:526-539): nested tuples in input order; inside a group the
      // MainNode's own tuple (oldest) first, then the sub chain, which is LIFO = newest first
      std::vector<uint64_t> order(m);
      for (uint64_t i = 0; i < m; ++i) order[i] = ((uint64_t)pairs[2 * i] << 32) | pairs[2 * i + 1];
      std::sort(order.begin(), order.end());
      for (uint64_t lo = 0; lo < m;) {
        uint64_t hi = lo + 1;
        while (hi < m && (order[hi] >> 32) == (order[lo] >> 32)) ++hi;
        std::reverse(order.begin() + lo + 1, order.begin() + hi);
        lo = hi;
      }
      for (uint64_t i = 0; i < m; ++i) {
        auto nested = make((uint32_t)(order[i] >> 32));
        unnestfun_t::eval_left(&_outputTuple, &nested);
        unnestfun_t::eval_right(&_outputTuple, &nested, table->row((uint32_t)order[i]));
        _consumer->step(&_outputTuple, g);
        inc();
      }
    }
    consumer_t* _consumer;
    output_t    _outputTuple;
    std::vector<std::remove_const_t<input_t>> _in;
    hj3d::detail::NestedBatch _dev{nullptr, nullptr, 0};
    const ht_nested_t* _table = nullptr;
    hj3d::DevBuf _dl, _dg, _dout;
};

namespace hj3d::detail {
template <class C> struct is_device_unnest : std::false_type {};
template <class A, class B, class T> struct is_device_unnest<AlgUnnestHt<A, B, T>> : std::true_type {};
}  // namespace hj3d::detail

// 3D hash join probe (algebra.hh:411-473)
template <alg_consumer_c Tconsumer, alg_buildop_c Tbuild, alg_hashfun_c Thashfun, alg_binary_predicate_c Tjoinpred,
          alg_concatfun_c Tconcatfun>
class AlgNestJoinProbe : public AlgBase {
  public:
    using consumer_t = Tconsumer; using build_t = Tbuild; using hashfun_t = Thashfun;
    using globstat_t = typename consumer_t::globstat_t; using input_t = typename hashfun_t::input_t;
    using output_t = typename consumer_t::input_t; using joinpred_t = Tjoinpred; using concatfun_t = Tconcatfun;
    AlgNestJoinProbe(consumer_t* aConsumer, build_t* aBuildOperator)
      : AlgBase("AlgNestJoinProbe"), _consumer(aConsumer), _buildOperator(aBuildOperator), _outputTuple(), _numCmps() {
      hj3d::check_key_equality_predicate<joinpred_t, hashfun_t, typename build_t::hashfun_t>("AlgNestJoinProbe: Tjoinpred");
    }
    inline void init(globstat_t* g) { reset(); _clk.ms = 0; _numCmps = 0; _in.clear(); _consumer->init(g); }
    inline void step(input_t* aProbeTuple, [[maybe_unused]] globstat_t* g) { _in.push_copy(aProbeTuple); }
    inline void step_bulk(input_t* first, size_t n, [[maybe_unused]] globstat_t* g) { _in.seq.push_bulk(first, n); }
    inline void fin(globstat_t* g) {
      using namespace hj3d;
      hj3d_ctx* c = Runtime::instance().ctx();
      const auto& table = _buildOperator->hashtable();
      const uint64_t n = _in.seq.size();
      const void* dprobe = _in.upload();
      hj3d_counters cnt{};
      bool need_tuples = true;
      if constexpr (detail::counts_only<consumer_t>) need_tuples = _consumer->wants_tuples();
      if (!need_tuples) {
        check(hj3d_probe_nested(c, table.handle(), dprobe, n, _in.ks, nullptr, 0, nullptr, 0, &cnt)); _clk.tick();
        if constexpr (detail::counts_only<consumer_t>) _consumer->add_count(cnt.matches);
      } else {
        if constexpr (detail::is_device_unnest<consumer_t>::value) {
          if (!_consumer->needs_host_tuples()) {                     // probe -> unnest -> count: one fused device call
            hj3d_counters ucnt{};
            check(hj3d_probe_nested_unnest(c, table.handle(), dprobe, n, _in.ks, 0, nullptr, 0, &cnt, &ucnt)); _clk.tick();
            _consumer->take_fused_count(ucnt.out_tuples);
            inc(cnt.matches); _numCmps += cnt.num_cmps;
            _consumer->fin(g); setDeviceTime(_clk.ms); stopTimer();
            return;
          }
        }
        auto* dout = (uint32_t*)_dout.ensure(n * 8 + 8);
        check(hj3d_probe_nested(c, table.handle(), dprobe, n, _in.ks, nullptr, 0, dout, n, &cnt)); _clk.tick();
        const uint64_t m = cnt.out_written;
        auto* dl = (uint32_t*)_dl.ensure(m * 4 + 4); auto* dg = (uint32_t*)_dg.ensure(m * 4 + 4);
        check(hj3d_split_pairs(c, dout, m, dl, dg));
        bool host_tuples = true;
        if constexpr (detail::is_device_unnest<consumer_t>::value) host_tuples = _consumer->needs_host_tuples();
        if (!host_tuples) {                                          // probe -> unnest -> count: everything stays on the device
          if constexpr (detail::is_device_unnest<consumer_t>::value)
            _consumer->step_device(detail::NestedBatch{dl, dg, m}, &table, typename consumer_t::make_nested_t(), g);
          inc(cnt.matches); _numCmps += cnt.num_cmps;
          _consumer->fin(g); setDeviceTime(_clk.ms); stopTimer();
          return;
        }
        auto* dfirst = (uint32_t*)_dfirst.ensure(m * 4 + 4);
        check(hj3d_group_first_row(c, table.handle(), dg, m, dfirst));
        std::vector<uint32_t> left(m), gref(m), first(m);
        check(hj3d_memcpy_d2h(c, left.data(), dl, m * 4)); check(hj3d_memcpy_d2h(c, gref.data(), dg, m * 4));
        check(hj3d_memcpy_d2h(c, first.data(), dfirst, m * 4));
        // group handles live as long as this operator's batch
        _nodes.resize(m);
        std::vector<uint32_t> at(n, 0xFFFFFFFFu);                  // probe position -> index of its nested tuple
        for (uint64_t i = 0; i < m; ++i) { _nodes[i] = table.make_node(gref[i], first[i]); at[left[i]] = (uint32_t)i; }
        auto make = [this, &at](uint32_t probe_pos) { return concatfun_t::eval(_in.seq.at(probe_pos), &_nodes[at[probe_pos]]); };
        if constexpr (detail::is_device_unnest<consumer_t>::value) {
          _consumer->step_device(detail::NestedBatch{dl, dg, m}, &table, typename consumer_t::make_nested_t(make), g);
        } else {
          // the reference emits in probe order (AlgScan pushes tuple by tuple)
          std::vector<uint32_t> order(m);
          for (uint32_t i = 0; i < m; ++i) order[i] = i;
          std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return left[a] < left[b]; });
          for (uint32_t i : order) { _outputTuple = make(left[i]); _consumer->step(&_outputTuple, g); }
        }
      }
      inc(cnt.matches);
      _numCmps += cnt.num_cmps;
      _consumer->fin(g);
      setDeviceTime(_clk.ms); stopTimer();
    }
    inline const consumer_t* consumer() const { return _consumer; }
    inline uint64_t numCmps() const { return _numCmps; }
  private:
    hj3d::detail::DevClock _clk;
    consumer_t* _consumer;
    build_t*    _buildOperator;
    output_t    _outputTuple;
    uint64_t    _numCmps;
    hj3d::detail::ProbeInput<hashfun_t> _in;
    hj3d::DevBuf _dout, _dl, _dg, _dfirst;
    std::vector<typename build_t::hashtable_t::MainNode> _nodes;
};

// Regular hash join probe (algebra.hh:600-672)
template <alg_consumer_c Tconsumer, alg_buildop_c Tbuild, alg_hashfun_c Thashfun, alg_binary_predicate_c Tjoinpred,
          alg_concatfun_c Tconcatfun, bool IsBuildKeyUnique = false>
class AlgHashJoinProbe : public AlgBase {
  public:
    using consumer_t = Tconsumer; using build_t = Tbuild; using hashfun_t = Thashfun;
    using globstat_t = typename consumer_t::globstat_t; using input_t = typename hashfun_t::input_t;
    using output_t = typename consumer_t::input_t; using hashvalue_t = typename hashfun_t::output_t;
    using joinpred_t = Tjoinpred; using concatfun_t = Tconcatfun;
    inline AlgHashJoinProbe(consumer_t* aConsumer, build_t* aBuildOperator)
      : AlgBase("AlgHashJoinProbe"), _consumer(aConsumer), _buildOperator(aBuildOperator), _outputTuple(), _numCmps() {
      hj3d::check_key_equality_predicate<joinpred_t, hashfun_t, typename build_t::hashfun_t>("AlgHashJoinProbe: Tjoinpred");
    }
    inline void init([[maybe_unused]] globstat_t* g) { reset(); _clk.ms = 0; _numCmps = 0; _in.clear(); _consumer->init(g); }
    inline void step(input_t* aTuple, [[maybe_unused]] globstat_t* g) { _in.push_copy(aTuple); }
    inline void step_bulk(input_t* first, size_t n, [[maybe_unused]] globstat_t* g) { _in.seq.push_bulk(first, n); }
    inline void fin(globstat_t* g) {
      using namespace hj3d;
      hj3d_ctx* c = Runtime::instance().ctx();
      const auto& table = _buildOperator->hashtable();
      const uint64_t n = _in.seq.size();
      const void* dprobe = _in.upload();
      hj3d_counters cnt{};
      bool need_tuples = true;
      if constexpr (detail::counts_only<consumer_t>) need_tuples = _consumer->wants_tuples();
      check(hj3d_probe_chaining(c, table.handle(), dprobe, n, _in.ks, nullptr, IsBuildKeyUnique ? 1 : 0, 0, nullptr, 0, &cnt)); _clk.tick();
      if (!need_tuples) {
        if constexpr (detail::counts_only<consumer_t>) _consumer->add_count(cnt.matches);
      } else {
        const uint64_t m = cnt.out_tuples;
        auto* dout = (uint32_t*)_dout.ensure(m * 8 + 8);
        check(hj3d_probe_chaining(c, table.handle(), dprobe, n, _in.ks, nullptr, IsBuildKeyUnique ? 1 : 0, 0, dout, m, &cnt)); _clk.tick();
        std::vector<uint32_t> pairs(2 * m);
        check(hj3d_memcpy_d2h(c, pairs.data(), dout, m * 8));
        // probe order like the tuple-at-a-time reference; within one probe tuple the reference emits its matches in
        // chain-walk order = the bucket's oldest tuple first, then newest to second oldest (ht_chaining.hh:185-194).
        // Matches of one probe are a sub-sequence of that chain, i.e. ordered by build row id: (the chain head if it
        // matches), then descending.  The device's output order is unspecified, so it is re-established here from
        // row ids alone; whether the smallest matching row is the chain head is decided by the bucket's minimum row.
        std::vector<uint64_t> order(m);
        for (uint64_t i = 0; i < m; ++i) order[i] = ((uint64_t)pairs[2 * i] << 32) | pairs[2 * i + 1];
        std::sort(order.begin(), order.end(), [](uint64_t a, uint64_t b) {
          if ((a >> 32) != (b >> 32)) return (a >> 32) < (b >> 32);
          return (uint32_t)a > (uint32_t)b;                                    // descending build row id
        });
        if constexpr (!IsBuildKeyUnique) {
          const auto head_row = table.bucket_min_rows();                         // bucket -> oldest row id of the bucket
          for (uint64_t lo = 0; lo < m;) {
            uint64_t hi = lo + 1;
            while (hi < m && (order[hi] >> 32) == (order[lo] >> 32)) ++hi;
            // the last element (smallest row id) moves to the front iff it is the bucket's first inserted tuple
            const uint32_t smallest = (uint32_t)order[hi - 1];
            if (hi - lo > 1 && head_row(smallest)) std::rotate(order.begin() + lo, order.begin() + hi - 1, order.begin() + hi);
            lo = hi;
          }
        }
        for (uint64_t i = 0; i < m; ++i) {
          _outputTuple = concatfun_t::eval(_in.seq.at((uint32_t)(order[i] >> 32)), table.row((uint32_t)order[i]));
          _consumer->step(&_outputTuple, g);
        }
      }
      inc(cnt.matches);
      _numCmps += cnt.num_cmps;
      _consumer->fin(g);
      setDeviceTime(_clk.ms); stopTimer();
    }
    inline const consumer_t* consumer() const { return _consumer; }
    inline uint64_t numCmps() const { return _numCmps; }
  private:
    hj3d::detail::DevClock _clk;
    consumer_t* _consumer;
    build_t*    _buildOperator;
    output_t    _outputTuple;
    uint64_t    _numCmps;
    hj3d::detail::ProbeInput<hashfun_t> _in;
    hj3d::DevBuf _dout;
};
