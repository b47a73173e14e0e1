// Drop-in counterpart of the reference's main_experiment4 (two joins on an inverted star, deferred unnesting):
// same command line (main_experiment4.cc:1065-1099), plans Ndu (:831-941) and Chj (:943-1043), CSV schema
// (:770-827) and generated relations (hj3d/datagen.hh), with the join / unnest operators on the GPU.
// Plan "Nnu" is accepted and, like in the reference, never run (:577-582).
#include <chrono>
#include <filesystem>
#include <fstream>
#include <functional>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "hj3d/algebra.hh"
#include "hj3d/datagen.hh"

namespace {

struct tuple_uint32_2_t { uint32_t k, a; };
std::ostream& operator<<(std::ostream& os, const tuple_uint32_2_t& t) { return os << "[" << t.k << "," << t.a << "]"; }
using base_tuple_t = tuple_uint32_2_t;
struct GlobStat {};
inline uint32_t murmur32(uint32_t x) { x ^= x >> 16; x *= 0x85ebca6b; x ^= x >> 13; x *= 0xc2b2ae35; x ^= x >> 16; return x; }

// functor structs (shapes of main_experiment4.cc:330-491)
struct HashfunR { using input_t = base_tuple_t; using output_t = uint32_t; static output_t eval(const input_t* t) { return murmur32(t->k); } };
struct HashfunFkRel { using input_t = base_tuple_t; using output_t = uint32_t; static output_t eval(const input_t* t) { return murmur32(t->a); } };
struct EqfunBuildFkRel { using left_t = base_tuple_t; using right_t = base_tuple_t; static bool eval(const left_t* l, const right_t* r) { return l->a == r->a; } };
struct JoinpredRS { using left_t = base_tuple_t; using right_t = base_tuple_t; static bool eval(const left_t* l, const right_t* r) { return l->k == r->a; } };
using NestedFk = HtNested1<base_tuple_t, HashfunFkRel, EqfunBuildFkRel>;
using MainNode = NestedFk::MainNode;
struct result_tuple_t { const base_tuple_t* _r; const base_tuple_t* _s; const base_tuple_t* _t; };
struct result_tuple_RS_t { const base_tuple_t* _r; const base_tuple_t* _s; };
struct nested_tuple_RS_t { base_tuple_t* _r; const MainNode* _s; };
struct nested_tuple_RST_t { base_tuple_t* _r; const MainNode* _s; const MainNode* _t; };
struct tuple_R_nS_xT_t { base_tuple_t* _r; const MainNode* _s; const base_tuple_t* _t; };
struct ConcatfunNested_RS { using left_t = base_tuple_t; using right_t = const MainNode; using output_t = nested_tuple_RS_t;
  static output_t eval(left_t* l, const right_t* r) { return {l, r}; } };
struct JoinpredRTnested { using left_t = nested_tuple_RS_t; using right_t = base_tuple_t;
  static bool eval(const left_t* l, const right_t* r) { return l->_r->k == r->a; } };
// the second probe hashes R.k reached through the nested tuple (main_experiment4.cc:413-419): tell the device
// shim which base functor that is and how to get to the base tuple
struct HashfunNestedRS { using input_t = nested_tuple_RS_t; using output_t = uint32_t;
  static output_t eval(const input_t* t) { return murmur32(t->_r->k); }
  using hj3d_base = HashfunR; static const base_tuple_t* hj3d_deref(const input_t* t) { return t->_r; } };
struct ConcatfunNested_RST { using left_t = nested_tuple_RS_t; using right_t = const MainNode; using output_t = nested_tuple_RST_t;
  static output_t eval(left_t* l, const right_t* r) { return {l->_r, l->_s, r}; } };
struct Unnestfun_R_nS_xT { using input_t = nested_tuple_RST_t; using output_t = tuple_R_nS_xT_t; using MainNode = ::MainNode; using data_t = base_tuple_t;
  static const MainNode* getMainNode(input_t* t) { return t->_t; }
  static void eval_left(output_t* o, input_t* i) { o->_r = i->_r; o->_s = i->_s; }
  static void eval_right(output_t* o, input_t*, const data_t* d) { o->_t = d; } };
struct Unnestfun_R_xS_xT { using input_t = tuple_R_nS_xT_t; using output_t = result_tuple_t; using MainNode = ::MainNode; using data_t = base_tuple_t;
  static const MainNode* getMainNode(input_t* t) { return t->_s; }
  static void eval_left(output_t* o, input_t* i) { o->_r = i->_r; o->_t = i->_t; }
  static void eval_right(output_t* o, input_t*, const data_t* d) { o->_s = d; } };
struct HashfunRS { using input_t = result_tuple_RS_t; using output_t = uint32_t;
  static output_t eval(const input_t* t) { return murmur32(t->_r->k); }
  using hj3d_base = HashfunR; static const base_tuple_t* hj3d_deref(const input_t* t) { return t->_r; } };
struct Joinpred_RS_T { using left_t = result_tuple_RS_t; using right_t = base_tuple_t;
  static bool eval(const left_t* l, const right_t* r) { return l->_r->k == r->a; } };
struct ConcatfunChaining_RS { using left_t = base_tuple_t; using right_t = base_tuple_t; using output_t = result_tuple_RS_t;
  static output_t eval(left_t* l, const right_t* r) { return {l, r}; } };
struct ConcatfunChaining_RS_T { using left_t = result_tuple_RS_t; using right_t = base_tuple_t; using output_t = result_tuple_t;
  static output_t eval(left_t* l, const right_t* r) { return {l->_r, l->_s, r}; } };
std::ostream& operator<<(std::ostream& os, const result_tuple_t& t) { return os << "[" << *t._r << "," << *t._s << "," << *t._t << "]"; }

using clk = std::chrono::steady_clock;
using ns_t = std::chrono::nanoseconds;
size_t repeat_mintime(ns_t minTime, const std::function<void()>& f, const std::function<void()>& teardown, size_t minRepeat) {
  size_t n = minRepeat; ns_t total{0};
  for (size_t i = 0; i < n; ++i) {
    auto t0 = clk::now(); f(); auto t1 = clk::now();
    total += (t1 - t0);
    if (i == n - 1 && total < minTime) n *= 2;
    if (i != n - 1) teardown();
  }
  return n;
}
class Csv {
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

struct Experiment4 {
  uint32_t log2R, alpha, beta, mA, mB;
  std::chrono::milliseconds minRuntime{300}; size_t minRepeat{8};
  RelationRS<base_tuple_t> R, S, T;
  Csv csv;
  Experiment4(uint32_t r, uint32_t a, uint32_t ma, uint32_t b, uint32_t mb, const std::string& file)
    : log2R(r), alpha(a), beta(b), mA(ma), mB(mb), csv(file) {}
  size_t cardR() const { return 1u << log2R; }
  size_t numFkCommon() const { return cardR() >> alpha; }
  size_t numFkExclusive() const { return cardR() >> beta; }
  size_t cardFk() const { return numFkCommon() * mA + numFkExclusive() * mB; }
  void init() {
    auto d = hj3d::gen::experiment4(log2R, alpha, mA, beta, mB);
    R._tuples.resize(cardR());
    for (size_t i = 0; i < cardR(); ++i) R._tuples[i] = {d.Rk[i], 0};
    S._tuples.resize(cardFk()); T._tuples.resize(cardFk());
    for (size_t i = 0; i < cardFk(); ++i) { S._tuples[i] = {(uint32_t)i, d.Sa[i]}; T._tuples[i] = {(uint32_t)i, d.Ta[i]}; }
  }
  void header() {
    for (const char* h : {"mintime", "minreps", "log2CardR", "a", "aM", "b", "bM", "cardR", "cardS", "cardT", "plan", "ht_impl", "reps",
                          "t_total", "t_build_S", "t_build_T", "t_probe_R", "c_sc_R", "c_sc_S", "c_sc_T", "c_build_S", "c_build_T",
                          "c_probe_RS", "c_probe_RS_cmp", "c_probe_RT", "c_probe_RT_cmp", "c_unnest_S", "c_unnest_T", "c_top"}) csv.f(h);
    csv.nl();
  }
  void params() {
    csv.f(std::to_string(minRuntime.count()) + "ms").f(minRepeat).f(log2R).f(alpha).f(mA).f(beta).f(mB).f(cardR()).f(cardFk()).f(cardFk());
  }
  template <class BS, class BT, class ScanR, class PRS, class PRT, class Top>
  void run_plan(const char* plan, const char* impl, BS& bS, BT& bT, ScanR& scR, PRS& pRS, PRT& pRT, Top& top,
                const std::function<void(Csv&)>& unnest_fields) {
    GlobStat gs;
    AlgScan<BS> scS(&bS, &S); AlgScan<BT> scT(&bT, &T);
    ns_t dS{0}, dT_{0}, dP{0}, dTot{0};
    size_t it = repeat_mintime(minRuntime, [&] {
      auto t0 = clk::now(); scS.run(&gs); auto t1 = clk::now(); scT.run(&gs); auto t2 = clk::now(); scR.run(&gs); auto t3 = clk::now();
      dS += (t1 - t0); dT_ += (t2 - t1); dP += (t3 - t2); dTot += (t3 - t0);
      top.printResult(false);
    }, [&] { bS.clear_ht(); bT.clear_ht(); }, minRepeat);
    dS /= it; dT_ /= it; dP /= it; dTot /= it;
    params();
    csv.f(plan).f(impl).f(it).f(dTot.count()).f(dS.count()).f(dT_.count()).f(dP.count())
       .f(scR.count()).f(scS.count()).f(scT.count()).f(bS.count()).f(bT.count())
       .f(pRS.count()).f(pRS.numCmps()).f(pRT.count()).f(pRT.numCmps());
    unnest_fields(csv);
    csv.f(top.count()).nl();
  }
  void run(const std::vector<std::string>& plans) {
    auto want = [&](const std::string& p) { for (auto& x : plans) if (x == p || x == "all" || x == "ALL") return true; return false; };
    header();
    const size_t D = numFkCommon() + numFkExclusive();
    auto printer = [](const result_tuple_t* t, std::ostream& os) { os << *t; };
    using top_t = AlgTop<result_tuple_t, GlobStat>;
    if (want("Ndu")) {
      std::cout << "void Experiment4::runNdu()" << std::endl;
      using build_t = AlgNestJoinBuild<HashfunFkRel, EqfunBuildFkRel, GlobStat>;
      using unnest_2_t = AlgUnnestHt<top_t, Unnestfun_R_xS_xT, build_t::hashtable_t>;
      using unnest_1_t = AlgUnnestHt<unnest_2_t, Unnestfun_R_nS_xT, build_t::hashtable_t>;
      using probe_RT_t = AlgNestJoinProbe<unnest_1_t, build_t, HashfunNestedRS, JoinpredRTnested, ConcatfunNested_RST>;
      using probe_RS_t = AlgNestJoinProbe<probe_RT_t, build_t, HashfunR, JoinpredRS, ConcatfunNested_RS>;
      build_t bS(D, 10, 10), bT(D, 10, 10);
      top_t top(std::cout, false, printer);
      unnest_2_t u2(&top); unnest_1_t u1(&u2); probe_RT_t pRT(&u1, &bT); probe_RS_t pRS(&pRT, &bS);
      AlgScan<probe_RS_t> scR(&pRS, &R);
      run_plan("Ndu", "nested", bS, bT, scR, pRS, pRT, top, [&](Csv& c) { c.f(u1.count()).f(u2.count()); });
      std::cout << "Plan Ndu\n  S: sizeof(MainNode): " << sizeof(MainNode) << "\n  S: sizeof(SubNode):  " << sizeof(NestedFk::SubNode) << "\n";
    }
    if (want("Chj")) {
      std::cout << "void Experiment4::runChj()" << std::endl;
      using build_t = AlgHashJoinBuild<HashfunFkRel, EqfunBuildFkRel, GlobStat>;
      using probe_RT_t = AlgHashJoinProbe<top_t, build_t, HashfunRS, Joinpred_RS_T, ConcatfunChaining_RS_T>;
      using probe_RS_t = AlgHashJoinProbe<probe_RT_t, build_t, HashfunR, JoinpredRS, ConcatfunChaining_RS>;
      build_t bS(D, 10), bT(D, 10);
      top_t top(std::cout, false, printer);
      probe_RT_t pRT(&top, &bT); probe_RS_t pRS(&pRT, &bS);
      AlgScan<probe_RS_t> scR(&pRS, &R);
      run_plan("Chj", "chaining", bS, bT, scR, pRS, pRT, top, [&](Csv& c) { c.f("NA").f("NA"); });
      std::cout << "Plan Chj\n  S: sizeof(Node): " << sizeof(build_t::hashtable_t::Node) << "\n";
    }
  }
};

[[noreturn]] void usage(const char* msg) {
  std::cerr << msg << "\nusage: main_experiment4.out -R <log2> -a <alpha> -A <mult> -b <beta> -B <mult> --measure-file <csv> [-p plans]\n";
  std::exit(EXIT_FAILURE);
}
}  // namespace

int main(int argc, char** argv) {
  hj3d::Runtime::instance().cache_uploads(true);   // the relations do not change between the repetitions of a plan
  long R = -1, a = -1, b = -1, A = -1, B = -1; std::string file; std::vector<std::string> plans = {"all"};
  for (int i = 1; i < argc; ++i) {
    std::string k = argv[i];
    auto val = [&]() -> std::string { if (i + 1 >= argc) usage("missing value"); return argv[++i]; };
    if (k == "-R" || k == "--card-R") R = std::stol(val());
    else if (k == "-a" || k == "--alpha") a = std::stol(val());
    else if (k == "-b" || k == "--beta") b = std::stol(val());
    else if (k == "-A" || k == "--alpha-mult") A = std::stol(val());
    else if (k == "-B" || k == "--beta-mult") B = std::stol(val());
    else if (k == "--measure-file") file = val();
    else if (k == "-p" || k == "--plans") { plans.clear(); std::stringstream ss(val()); std::string it; while (std::getline(ss, it, ',')) if (!it.empty()) plans.push_back(it); }
    else if (k.rfind("--print", 0) == 0 || k.rfind("--no-print", 0) == 0 || k == "--run" || k == "--no-run") {}
    else usage(("unknown option " + k).c_str());
  }
  if (R < 0 || R > 30 || a < 0 || b < 0 || A < 1 || B < 1 || file.empty()) usage("required: -R -a -A -b -B --measure-file");
  try {
    Experiment4 e((uint32_t)R, (uint32_t)a, (uint32_t)A, (uint32_t)b, (uint32_t)B, file);
    if (e.cardR() < e.numFkCommon() + 2 * e.numFkExclusive()) usage("cardR must be >= #common + 2 * #exclusive foreign keys");
    e.init();
    std::cout << "cardR " << e.cardR() << " cardS " << e.cardFk() << " cardT " << e.cardFk() << " expected |RST| "
              << e.numFkCommon() * A * A << "\n";
    e.run(plans);
  } catch (const std::exception& ex) { std::cerr << "error: " << ex.what() << "\n"; return EXIT_FAILURE; }
  std::cout << "----" << std::endl;
  return EXIT_SUCCESS;
}
