// Host-only check of hj3d::check_key_equality_predicate (hostcpp/hj3d/runtime.hh): the shims accept a join predicate /
// content-equality functor only when it is equality of the attributes the hash functors name.  No device call is made.
#include <cstdio>
#include <cstdint>
#include "hj3d/runtime.hh"

struct T3 { uint32_t k, a, b; };
inline uint32_t mm(uint32_t x) { return hj3d::ref_murmur32(x); }
struct HashK { using input_t = const T3; using output_t = uint32_t; static output_t eval(const input_t* t) { return mm(t->k); } };
struct HashA { using input_t = const T3; using output_t = uint32_t; static output_t eval(const input_t* t) { return mm(t->a); } };
struct PredKA    { using left_t = T3; using right_t = T3; static bool eval(const left_t* l, const right_t* r) { return l->k == r->a; } };
struct PredWrong { using left_t = T3; using right_t = T3; static bool eval(const left_t* l, const right_t* r) { return l->k == r->k; } };
struct PredExtra { using left_t = T3; using right_t = T3; static bool eval(const left_t* l, const right_t* r) { return l->k == r->a && l->b == r->b; } };

template <class P> int rejected() {
  try { hj3d::check_key_equality_predicate<P, HashK, HashA>("test"); } catch (const hj3d::Error&) { return 1; }
  return 0;
}
int main() {
  const int a = rejected<PredKA>(), b = rejected<PredWrong>(), c = rejected<PredExtra>();
  std::printf("%d %d %d\n", a, b, c);
  return (a == 0 && b == 1 && c == 1) ? 0 : 1;
}
