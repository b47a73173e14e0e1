// Drop-in counterpart of the reference's main_algebra_example.cc (algebra_test0..3, :147-435): the same four
// plans on the same literal relations, written against hj3d/algebra.hh instead of algebra.hh.  The functor
// structs are the shapes the reference driver uses (:31-145); nothing in them is device specific.
#include <iostream>

#include "hj3d/algebra.hh"

using attrval_t = int;
struct tuple_L_t { attrval_t a, b; };
struct tuple_R_t { attrval_t c, d; };
std::ostream& operator<<(std::ostream& os, const tuple_L_t t) { return os << "(" << t.a << "," << t.b << ")"; }
std::ostream& operator<<(std::ostream& os, const tuple_R_t t) { return os << "(" << t.c << "," << t.d << ")"; }
std::ostream& operator<<(std::ostream& os, const tuple_L_t* t) { return os << "(" << t->a << "," << t->b << ")"; }
std::ostream& operator<<(std::ostream& os, const tuple_R_t* t) { return os << "(" << t->c << "," << t->d << ")"; }

static inline uint64_t murmur64(uint64_t x) {
  x ^= (x >> 33); x *= 0xFF51AFD7ED558CCD; x ^= (x >> 33); x *= 0xC4CEB9FE1A95EC63; x ^= (x >> 33);
  return x;
}

struct SelectionL { using input_t = tuple_L_t; static bool eval(const input_t* t) { return t->b < 40; } };
struct DynSelectionL { using input_t = tuple_L_t; bool operator()(const input_t* t) { return t->b < 40; } };
struct HashfunBuild { using input_t = tuple_R_t; using output_t = uint64_t; static output_t eval(const input_t* t) { return murmur64(t->c); } };
struct HashfunProbe { using input_t = tuple_L_t; using output_t = uint64_t; static output_t eval(const input_t* t) { return murmur64(t->a); } };
struct EqFunBuild { using left_t = tuple_R_t; using right_t = tuple_R_t; static bool eval(const left_t* l, const right_t* r) { return l->c == r->c; } };
struct EqFunProbe { using left_t = tuple_L_t; using right_t = tuple_R_t; static bool eval(const left_t* l, const right_t* r) { return l->a == r->c; } };

using nested_ht_t = HtNested1<tuple_R_t, HashfunBuild, EqFunBuild>;
struct tuple_nested_t { tuple_L_t* _left; const nested_ht_t::MainNode* _right; };
struct tuple_LR_t { const tuple_L_t* _left; const tuple_R_t* _right; };
std::ostream& operator<<(std::ostream& os, const tuple_LR_t t) {
  return os << "(" << t._left->a << "," << t._left->b << "," << t._right->c << "," << t._right->d << ")";
}
struct ConcatFunNested { using left_t = tuple_L_t; using right_t = nested_ht_t::MainNode; using output_t = tuple_nested_t;
  static output_t eval(left_t* l, const right_t* r) { return tuple_nested_t{l, r}; } };
struct ConcatFunChaining { using left_t = tuple_L_t; using right_t = tuple_R_t; using output_t = tuple_LR_t;
  static output_t eval(left_t* l, const right_t* r) { return {l, r}; } };
struct UnnestFun {
  using input_t = tuple_nested_t; using output_t = tuple_LR_t; using MainNode = nested_ht_t::MainNode; using data_t = nested_ht_t::data_t;
  static const MainNode* getMainNode(input_t* t) { return t->_right; }
  static void eval_left(output_t* out, input_t* in) { out->_left = in->_left; }
  static void eval_right(output_t* out, input_t*, const data_t* data) { out->_right = data; }
};

struct GlobStat {};

int main() {
  RelationRS<tuple_L_t> L; L._tuples = {{1, 11}, {2, 21}, {3, 31}, {4, 41}};
  RelationRS<tuple_R_t> R; R._tuples = {{1, -1}, {1, -2}, {1, -3}, {2, -1}, {2, -2}, {3, -1}};
  GlobStat gs;
  {  // algebra_test0: scan -> selection -> top
    std::cout << "test0\n";
    using top_t = AlgTop<tuple_L_t, GlobStat>; using sel_t = AlgSelection<top_t, SelectionL>; using scan_t = AlgScan<sel_t>;
    top_t top(std::cout, true, [](const tuple_L_t* t, std::ostream& os) { os << t; });
    sel_t sel(&top); scan_t scan(&sel, &L);
    scan.run(&gs);
    print_strand(&scan, 1);
  }
  {  // algebra_test1: nested join, result not unnested
    std::cout << "test1\n";
    using build_t = AlgNestJoinBuild<HashfunBuild, EqFunBuild, GlobStat>; using scan_R_t = AlgScan<build_t>;
    using top_t = AlgTop<tuple_nested_t, GlobStat>;
    using probe_t = AlgNestJoinProbe<top_t, build_t, HashfunProbe, EqFunProbe, ConcatFunNested>;
    using sel_t = AlgDynSelection<probe_t, DynSelectionL>; using scan_L_t = AlgScan<sel_t>;
    build_t build(5, 4, 4); scan_R_t scanR(&build, &R);
    top_t top(std::cout, true, [](const tuple_nested_t* t, std::ostream& os) {
      os << "(" << t->_left->a << "," << t->_left->b << "," << t->_right->data()->c << "," << t->_right->data()->d << ")"; });
    probe_t probe(&top, &build); sel_t sel(&probe); scan_L_t scanL(&sel, &L);
    scanR.run(&gs); scanL.run(&gs);
    print_strand(&scanR, 1); print_strand(&scanL, 1);
    build.hashtable().makeStatistics().print();
  }
  {  // algebra_test2: nested join + unnest
    std::cout << "test2\n";
    using build_t = AlgNestJoinBuild<HashfunBuild, EqFunBuild, GlobStat>; using scan_R_t = AlgScan<build_t>;
    using top_t = AlgTop<tuple_LR_t, GlobStat>; using unnest_t = AlgUnnestHt<top_t, UnnestFun, build_t::hashtable_t>;
    using probe_t = AlgNestJoinProbe<unnest_t, build_t, HashfunProbe, EqFunProbe, ConcatFunNested>;
    using sel_t = AlgSelection<probe_t, SelectionL>; using scan_L_t = AlgScan<sel_t>;
    build_t build(5, 4, 4); scan_R_t scanR(&build, &R);
    top_t top(std::cout, true, [](const tuple_LR_t* t, std::ostream& os) { os << *t; });
    unnest_t unnest(&top); probe_t probe(&unnest, &build); sel_t sel(&probe); scan_L_t scanL(&sel, &L);
    scanR.run(&gs); scanL.run(&gs);
    print_strand(&scanR, 1); print_strand(&scanL, 1);
  }
  {  // algebra_test3: chaining join
    std::cout << "test3\n";
    using build_t = AlgHashJoinBuild<HashfunBuild, EqFunBuild, GlobStat>; using scan_R_t = AlgScan<build_t>;
    using top_t = AlgTop<tuple_LR_t, GlobStat>;
    using probe_t = AlgHashJoinProbe<top_t, build_t, HashfunProbe, EqFunProbe, ConcatFunChaining>;
    using sel_t = AlgSelection<probe_t, SelectionL>; using scan_L_t = AlgScan<sel_t>;
    build_t build(5, 4); scan_R_t scanR(&build, &R);
    top_t top(std::cout, true, [](const tuple_LR_t* t, std::ostream& os) { os << *t; });
    probe_t probe(&top, &build); sel_t sel(&probe); scan_L_t scanL(&sel, &L);
    scanR.run(&gs); scanL.run(&gs);
    print_strand(&scanR, 1); print_strand(&scanL, 1);
    build.hashtable().makeStatistics().print();
    std::cout << "sizeof(Node) " << sizeof(build_t::hashtable_t::Node) << "\n";
  }
  return 0;
}
