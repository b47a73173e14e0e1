// hj3d/tuple_types.hh -- the row-store tuple layouts of the reference (tuple_types.hh:5-30).
#pragma once

#include <cstdint>
#include <iostream>

template <typename Tattr> struct tuple_2_t { Tattr k, a; };
template <typename Tattr> struct tuple_3_t { Tattr k, a, b; };

using tuple_uint32_2_t = tuple_2_t<uint32_t>;
using tuple_uint64_2_t = tuple_2_t<uint64_t>;
using tuple_uint32_3_t = tuple_3_t<uint32_t>;
using tuple_uint64_3_t = tuple_3_t<uint64_t>;

template <typename Tattr> inline std::ostream& operator<<(std::ostream& os, const tuple_2_t<Tattr>& t) {
  return os << "[" << t.k << "," << t.a << "]";
}
template <typename Tattr> inline std::ostream& operator<<(std::ostream& os, const tuple_3_t<Tattr>& t) {
  return os << "[" << t.k << "," << t.a << "]";
}
