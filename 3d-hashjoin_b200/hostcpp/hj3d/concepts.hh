// hj3d/concepts.hh -- the functor contracts of the operator templates (same names and requirements as the
// reference's concepts.hh:16-86, so the drivers' functor structs satisfy them unchanged).
#pragma once

#include <concepts>
#include <ostream>

template <typename T>
concept Printable = requires(std::ostream& os, T a) { os << a; };

// hash function: input_t, output_t, static output_t eval(const input_t*)            (concepts.hh:22-28)
template <typename T>
concept alg_hashfun_c = requires {
  typename T::input_t;
  typename T::output_t;
  { T::eval(static_cast<const typename T::input_t*>(nullptr)) } -> std::same_as<typename T::output_t>;
};

// static unary predicate: static bool eval(const input_t*)                            (concepts.hh:31-36)
template <typename T>
concept alg_predicate_c = requires {
  typename T::input_t;
  { T::eval(static_cast<const typename T::input_t*>(nullptr)) } -> std::same_as<bool>;
};

// runtime unary predicate: bool operator()(const input_t*)                            (concepts.hh:39-45)
template <typename T>
concept alg_dyn_predicate_c = requires(T t) {
  typename T::input_t;
  { t(static_cast<const typename T::input_t*>(nullptr)) } -> std::same_as<bool>;
};

// key equality / join predicate: static bool eval(const left_t*, const right_t*)      (concepts.hh:49-56)
template <typename T>
concept alg_binary_predicate_c = requires {
  typename T::left_t;
  typename T::right_t;
  { T::eval(static_cast<const typename T::left_t*>(nullptr), static_cast<const typename T::right_t*>(nullptr)) }
      -> std::same_as<bool>;
};

// concatenation: static output_t eval(left_t*, right_t*)                              (concepts.hh:60-68)
template <typename T>
concept alg_concatfun_c = requires {
  typename T::left_t;
  typename T::right_t;
  typename T::output_t;
  { T::eval(static_cast<typename T::left_t*>(nullptr), static_cast<typename T::right_t*>(nullptr)) }
      -> std::same_as<typename T::output_t>;
};

// unnest function of the 3D hash join                                                 (concepts.hh:71-86)
template <typename T>
concept alg_unnestfun_c = requires {
  typename T::input_t;
  typename T::output_t;
  typename T::MainNode;
  typename T::data_t;
  { T::eval_left(static_cast<typename T::output_t*>(nullptr), static_cast<typename T::input_t*>(nullptr)) }
      -> std::same_as<void>;
  { T::eval_right(static_cast<typename T::output_t*>(nullptr), static_cast<typename T::input_t*>(nullptr),
                  static_cast<const typename T::data_t*>(nullptr)) } -> std::same_as<void>;
  { T::getMainNode(static_cast<typename T::input_t*>(nullptr)) } -> std::same_as<const typename T::MainNode*>;
};
