// Drop-in counterpart of the reference's main_experiment1 (key/foreign-key join, chaining vs nested table):
// same command line (main_experiment1.cc:1389-1422), same plans (scr scs Csr CsrUU Crs Nsr Nrs NrsNU, :532-1285),
// same CSV schema (:1289-1347, including the `scr`/`scs` rows that carry no `reps` field), same generated
// relations (hj3d/datagen.hh), same repeat loop semantics (util/measure_helpers.hh:15-41) -- with the join
// operators running on the GPU through hj3d/algebra.hh.
#include <chrono>
#include <filesystem>
#include <fstream>
#include <functional>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "hj3d/algebra.hh"
#include "hj3d/datagen.hh"

namespace {

struct tuple_uint32_3_t { uint32_t k, a, b; };
std::ostream& operator<<(std::ostream& os, const tuple_uint32_3_t& t) { return os << "[" << t.k << "," << t.a << "]"; }
using base_tuple_t = tuple_uint32_3_t;
using hashvalue_t = uint32_t;
struct GlobStat {};

inline uint32_t murmur32(uint32_t x) { x ^= x >> 16; x *= 0x85ebca6b; x ^= x >> 13; x *= 0xc2b2ae35; x ^= x >> 16; return x; }

// functor structs of the experiment (shapes of main_experiment1.cc:287-410)
struct HashfunOnR { using input_t = base_tuple_t; using output_t = hashvalue_t; static output_t eval(const input_t* t) { return murmur32(t->k); } };
struct HashfunOnS { using input_t = base_tuple_t; using output_t = hashvalue_t; static output_t eval(const input_t* t) { return murmur32(t->a); } };
struct EqfunBuildR { using left_t = base_tuple_t; using right_t = base_tuple_t; static bool eval(const left_t* l, const right_t* r) { return l->k == r->k; } };
struct EqfunBuildS { using left_t = base_tuple_t; using right_t = base_tuple_t; static bool eval(const left_t* l, const right_t* r) { return l->a == r->a; } };
struct EqfunJoinpredRS { using left_t = base_tuple_t; using right_t = base_tuple_t; static bool eval(const left_t* l, const right_t* r) { return l->k == r->a; } };
struct EqfunJoinpredSR { using left_t = base_tuple_t; using right_t = base_tuple_t; static bool eval(const left_t* l, const right_t* r) { return l->a == r->k; } };
using NestedOnS = HtNested1<base_tuple_t, HashfunOnS, EqfunBuildS>;
using NestedOnR = HtNested1<base_tuple_t, HashfunOnR, EqfunBuildR>;
struct nested_tuple_RS_t { base_tuple_t* _left; const NestedOnS::MainNode* _right; };
struct nested_tuple_SR_t { base_tuple_t* _left; const NestedOnR::MainNode* _right; };
struct result_tuple_t { const base_tuple_t* _left; const base_tuple_t* _right; };
std::ostream& operator<<(std::ostream& os, const result_tuple_t& t) { return os << "[" << *t._left << "," << *t._right << "]"; }
struct ConcatfunChaining { using left_t = base_tuple_t; using right_t = base_tuple_t; using output_t = result_tuple_t;
  static output_t eval(left_t* l, const right_t* r) { return {l, r}; } };
struct ConcatfunNestedRS { using left_t = base_tuple_t; using right_t = NestedOnS::MainNode; using output_t = nested_tuple_RS_t;
  static output_t eval(left_t* l, const right_t* r) { return {l, r}; } };
struct ConcatfunNestedSR { using left_t = base_tuple_t; using right_t = NestedOnR::MainNode; using output_t = nested_tuple_SR_t;
  static output_t eval(left_t* l, const right_t* r) { return {l, r}; } };
template <class Nested, class Ht> struct UnnestFunT {
  using input_t = Nested; using output_t = result_tuple_t; using MainNode = typename Ht::MainNode; using data_t = typename Ht::data_t;
  static const MainNode* getMainNode(input_t* t) { return t->_right; }
  static void eval_left(output_t* out, input_t* in) { out->_left = in->_left; }
  static void eval_right(output_t* out, input_t*, const data_t* data) { out->_right = data; }
};
using UnnestFunRS = UnnestFunT<nested_tuple_RS_t, NestedOnS>;
using UnnestFunSR = UnnestFunT<nested_tuple_SR_t, NestedOnR>;

using clk = std::chrono::steady_clock;
using ns_t = std::chrono::nanoseconds;

// util/measure_helpers.hh:15-41: start with minRepeat runs, double while the total stays below minTime;
// teardown after every run but the last
size_t repeat_mintime(ns_t minTime, const std::function<void()>& f, const std::function<void()>& teardown, size_t minRepeat) {
  size_t n = minRepeat;
  ns_t total{0};
  for (size_t i = 0; i < n; ++i) {
    auto t0 = clk::now(); f(); auto t1 = clk::now();
    total += (t1 - t0);
    if (i == n - 1 && total < minTime) n *= 2;
    if (i != n - 1) teardown();
  }
  return n;
}

class Csv {   // util/csv_writer.hh:35-54: ';' separated fields
  public:
    explicit Csv(const std::string& file) : _os(file, std::ofstream::trunc) {
      std::filesystem::path p{file};
      if (!std::filesystem::exists(p.remove_filename())) throw std::runtime_error("Directory " + p.string() + " does not exist");
    }
    template <class T> Csv& f(const T& v) { if (_col++) _os << ";"; _os << v; return *this; }
    Csv& nl() { _os << '\n'; _col = 0; return *this; }
  private:
    std::ofstream _os; uint32_t _col = 0;
};

struct Experiment1 {
  uint32_t log2R, log2S; bool skew; uint32_t t, b;
  std::chrono::milliseconds minRuntime{300}; size_t minRepeat{8};
  const uint32_t log2RsvChunk = 10;
  RelationRS<base_tuple_t> R, S;
  size_t numDvSa = 0;
  Csv csv;
  Experiment1(uint32_t r, uint32_t s, bool sk, uint32_t t_, uint32_t b_, const std::string& file)
    : log2R(r), log2S(s), skew(sk), t(t_), b(b_), csv(file) {}
  size_t cardR() const { return 1u << log2R; }
  size_t cardS() const { return 1u << log2S; }
  uint32_t fkMax() const { return 1u << (log2R - t); }

  void init() {
    auto d = hj3d::gen::experiment1(log2R, log2S, skew, t);
    R._tuples.resize(cardR()); S._tuples.resize(cardS());
    for (size_t i = 0; i < cardR(); ++i) R._tuples[i].k = d.Rk[i];
    for (size_t i = 0; i < cardS(); ++i) { S._tuples[i].k = d.Sk[i]; S._tuples[i].a = d.Sa[i]; }
    numDvSa = d.numDvSa;
  }
  void header() {
    for (const char* h : {"mintime", "minreps", "log2CardR", "log2CardS", "skew", "t", "fkMax", "numDvSa", "b", "plan", "ht_impl",
                          "build", "probe", "ht_buckets", "ht_fracEmpty", "cc0_avg", "cc0_min", "cc0_max", "cc1_avg", "cc1_min",
                          "cc1_max", "reps", "t_total", "t_buildStr", "t_probeStr", "t_top", "c_scanBuild", "c_selBuild", "c_htBuild",
                          "c_scanProbe", "c_selProbe", "c_htProbe", "c_htProbeCmp", "c_unnest", "c_top"}) csv.f(h);
    csv.nl();
  }
  void params() {
    csv.f(std::to_string(minRuntime.count()) + "ms").f(minRepeat).f(log2R).f(log2S).f(skew).f(t).f(fkMax()).f(numDvSa).f(b);
  }
  void scan_only(const char* plan, RelationRS<base_tuple_t>& rel) {
    GlobStat gs;
    using top_t = AlgTop<base_tuple_t, GlobStat>;
    top_t top(std::cout, false, [](const base_tuple_t* t, std::ostream& os) { os << t; });
    AlgScan<top_t> scan(&top, &rel);
    auto t0 = clk::now(); scan.run(&gs); auto t1 = clk::now();
    params();
    csv.f(plan);
    for (int i = 0; i < 11; ++i) csv.f("NA");
    csv.f((t1 - t0).count()).f("NA").f("NA").f(get_runtime_excl(&top).count()).f(scan.count());
    for (int i = 0; i < 7; ++i) csv.f("NA");
    csv.f(top.count()).nl();
  }

  // one join plan: build strand on `brel`, probe strand on `prel`; Tunnest = void for plans without unnest
  template <class build_t, class top_t, class probe_t, class unnest_t>
  void join_plan(const char* plan, const char* impl, const char* bname, const char* pname, RelationRS<base_tuple_t>& brel,
                 RelationRS<base_tuple_t>& prel, build_t& build, top_t& top, probe_t& probe, unnest_t* unnest) {
    GlobStat gs;
    AlgScan<build_t> scanB(&build, &brel);
    AlgScan<probe_t> scanP(&probe, &prel);
    ns_t dB{0}, dP{0}, dT{0};
    size_t it = repeat_mintime(minRuntime, [&] {
      auto t0 = clk::now(); scanB.run(&gs); auto t1 = clk::now(); scanP.run(&gs); auto t2 = clk::now();
      dB += (t1 - t0); dP += (t2 - t1); dT += (t2 - t0);
    }, [&] { build.clear_ht(); }, minRepeat);
    dB /= it; dP /= it; dT /= it;
    std::cout << "Plan " << plan << "\n  Build Strand\n"; print_strand(&scanB, 2);
    std::cout << "  Probe Strand\n"; print_strand(&scanP, 2);
    const HtStatistics st = build.hashtable().makeStatistics();
    params();
    csv.f(plan).f(impl).f(bname).f(pname).f(build.hashtable().numBuckets()).f(st.fracEmptyBuckets())
       .f(st._collisionChainLen.avg()).f(st._collisionChainLen.min()).f(st._collisionChainLen.max())
       .f(st._collisionChainLenNonempty.avg()).f(st._collisionChainLenNonempty.min()).f(st._collisionChainLenNonempty.max())
       .f(it).f(dT.count()).f(dB.count()).f(dP.count()).f(get_runtime_excl(&top).count())
       .f(scanB.count()).f("NA").f(build.count()).f(scanP.count()).f("NA").f(probe.count()).f(probe.numCmps());
    if constexpr (std::is_same_v<unnest_t, void>) csv.f("NA"); else csv.f(unnest->count());
    csv.f(top.count()).nl();
  }

  void run(const std::vector<std::string>& plans) {
    auto want = [&](const std::string& p) {
      for (auto& x : plans) if (x == p || x == "all" || x == "ALL") return true;
      return false;
    };
    header();
    if (want("scr")) scan_only("scr", R);
    if (want("scs")) scan_only("scs", S);
    const uint32_t bucketsR = (uint32_t)std::max<size_t>(cardR() / b, 1), bucketsS = (uint32_t)std::max<size_t>(numDvSa / b, 1);
    auto printer = [](const result_tuple_t* t, std::ostream& os) { os << "[" << *t->_left << "," << *t->_right << "]"; };
    using top_t = AlgTop<result_tuple_t, GlobStat>;
    if (want("Csr")) {        // chaining, build R, probe S, key property of R.k known (:624-744)
      using build_t = AlgHashJoinBuild<HashfunOnR, EqfunBuildR, GlobStat>;
      using probe_t = AlgHashJoinProbe<top_t, build_t, HashfunOnS, EqfunJoinpredSR, ConcatfunChaining, true>;
      build_t build(bucketsR, log2RsvChunk); top_t top(std::cout, false, printer); probe_t probe(&top, &build);
      join_plan<build_t, top_t, probe_t, void>("Csr", "chaining", "R", "S", R, S, build, top, probe, nullptr);
      std::cout << "  sizeof(Node): " << sizeof(build_t::hashtable_t::Node) << "\n";
    }
    if (want("CsrUU")) {      // same, key property unknown (:746-848)
      using build_t = AlgHashJoinBuild<HashfunOnR, EqfunBuildR, GlobStat>;
      using probe_t = AlgHashJoinProbe<top_t, build_t, HashfunOnS, EqfunJoinpredSR, ConcatfunChaining, false>;
      build_t build(bucketsR, log2RsvChunk); top_t top(std::cout, false, printer); probe_t probe(&top, &build);
      join_plan<build_t, top_t, probe_t, void>("CsrUU", "chaining", "R", "S", R, S, build, top, probe, nullptr);
    }
    if (want("Crs")) {        // chaining, build S, probe R (:850-967)
      using build_t = AlgHashJoinBuild<HashfunOnS, EqfunBuildS, GlobStat>;
      using probe_t = AlgHashJoinProbe<top_t, build_t, HashfunOnR, EqfunJoinpredRS, ConcatfunChaining>;
      build_t build(bucketsS, log2RsvChunk); top_t top(std::cout, false, printer); probe_t probe(&top, &build);
      join_plan<build_t, top_t, probe_t, void>("Crs", "chaining", "S", "R", S, R, build, top, probe, nullptr);
    }
    if (want("Nsr")) {        // nested, build R, probe S, unnest (:1078-1185)
      using build_t = AlgNestJoinBuild<HashfunOnR, EqfunBuildR, GlobStat>;
      using unnest_t = AlgUnnestHt<top_t, UnnestFunSR, build_t::hashtable_t>;
      using probe_t = AlgNestJoinProbe<unnest_t, build_t, HashfunOnS, EqfunJoinpredSR, ConcatfunNestedSR>;
      build_t build(bucketsR, log2RsvChunk, log2RsvChunk); top_t top(std::cout, false, printer); unnest_t un(&top); probe_t probe(&un, &build);
      join_plan<build_t, top_t, probe_t, unnest_t>("Nsr", "nested", "R", "S", R, S, build, top, probe, &un);
    }
    if (want("Nrs")) {        // nested, build S, probe R, unnest (:969-1076)
      using build_t = AlgNestJoinBuild<HashfunOnS, EqfunBuildS, GlobStat>;
      using unnest_t = AlgUnnestHt<top_t, UnnestFunRS, build_t::hashtable_t>;
      using probe_t = AlgNestJoinProbe<unnest_t, build_t, HashfunOnR, EqfunJoinpredRS, ConcatfunNestedRS>;
      build_t build(bucketsS, log2RsvChunk, log2RsvChunk); top_t top(std::cout, false, printer); unnest_t un(&top); probe_t probe(&un, &build);
      join_plan<build_t, top_t, probe_t, unnest_t>("Nrs", "nested", "S", "R", S, R, build, top, probe, &un);
      std::cout << "  sizeof(MainNode): " << sizeof(build_t::hashtable_t::MainNode) << "\n  sizeof(SubNode):  "
                << sizeof(build_t::hashtable_t::SubNode) << "\n";
    }
    if (want("NrsNU")) {      // nested, build S, probe R, no unnest (:1187-1285)
      using build_t = AlgNestJoinBuild<HashfunOnS, EqfunBuildS, GlobStat>;
      using ntop_t = AlgTop<nested_tuple_RS_t, GlobStat>;
      using probe_t = AlgNestJoinProbe<ntop_t, build_t, HashfunOnR, EqfunJoinpredRS, ConcatfunNestedRS>;
      build_t build(bucketsS, log2RsvChunk, log2RsvChunk);
      ntop_t top(std::cout, false, [](const nested_tuple_RS_t*, std::ostream&) {}); probe_t probe(&top, &build);
      join_plan<build_t, ntop_t, probe_t, void>("NrsNU", "nested", "S", "R", S, R, build, top, probe, nullptr);
    }
  }
};

[[noreturn]] void usage(const char* msg) {
  std::cerr << msg << "\nusage: main_experiment1.out -R <log2> -S <log2> (--skew|--no-skew) -t <0..9> [-b <1..4>] "
               "--measure-file <csv> [-p plan,plan,..] [--print-timers] [--print-relations]\n";
  std::exit(EXIT_FAILURE);
}

}  // namespace

int main(int argc, char** argv) {
  hj3d::Runtime::instance().cache_uploads(true);   // the relations do not change between the repetitions of a plan
  long R = -1, S = -1, t = -1, b = 1; int skew = -1; std::string file; std::vector<std::string> plans = {"all"};
  bool printRelations = false;
  auto need = [&](int& i) -> std::string { if (i + 1 >= argc) usage("missing value"); return argv[++i]; };
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i], v;
    auto eq = a.find('=');
    bool has = false;
    if (a.rfind("--", 0) == 0 && eq != std::string::npos) { v = a.substr(eq + 1); a = a.substr(0, eq); has = true; }
    auto val = [&]() { return has ? v : need(i); };
    if (a == "-R" || a == "--card-R") R = std::stol(val());
    else if (a == "-S" || a == "--card-S") S = std::stol(val());
    else if (a == "-t" || a == "--param-t") t = std::stol(val());
    else if (a == "-b" || a == "--param-b") b = std::stol(val());
    else if (a == "--skew") skew = 1;
    else if (a == "--no-skew") skew = 0;
    else if (a == "--measure-file") file = val();
    else if (a == "-p" || a == "--plans") {
      plans.clear();
      std::stringstream ss(val()); std::string item;
      while (std::getline(ss, item, ',')) if (!item.empty()) plans.push_back(item);
    }
    else if (a == "--print-timers" || a == "--no-print-timers") {}
    else if (a == "--print-relations") printRelations = true;
    else if (a == "--no-print-relations") printRelations = false;
    else usage(("unknown option " + a).c_str());
  }
  if (R < 0 || R > 30) usage("--card-R is required (0..30)");
  if (S < 0 || S > 30) usage("--card-S is required (0..30)");
  if (skew < 0) usage("--skew or --no-skew is required");
  if (t < 0 || t > 9) usage("--param-t is required (0..9)");
  if (b < 1 || b > 4) usage("--param-b must be in 1..4");
  if (file.empty()) usage("--measure-file is required");
  if (t > R) { std::cerr << "--param-t must not be greater than --card-R\n"; return EXIT_FAILURE; }
  std::cout << "Running Experiment 1 with the following config:\n  --card-R " << R << "\n  --card-S " << S << "\n  --skew "
            << std::boolalpha << (skew == 1) << "\n  --param-t " << t << "\n  --param-b " << b << "\n  --measure-file \"" << file
            << "\"\n  --plans ";
  for (auto& p : plans) std::cout << p << ",";
  std::cout << "\n";
  try {
    Experiment1 e((uint32_t)R, (uint32_t)S, skew == 1, (uint32_t)t, (uint32_t)b, file);
    e.init();
    if (printRelations) {
      std::cout << "-- R --\n"; for (auto& x : e.R._tuples) std::cout << x.k << "|" << x.a << "|" << x.b << "\n";
      std::cout << "-- S --\n"; for (auto& x : e.S._tuples) std::cout << x.k << "|" << x.a << "|" << x.b << "\n";
    }
    e.run(plans);
  } catch (const std::exception& ex) {
    std::cerr << "error: " << ex.what() << "\n";
    return EXIT_FAILURE;
  }
  std::cout << "----" << std::endl;
  return EXIT_SUCCESS;
}
