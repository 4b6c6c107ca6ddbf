// oracle/shim/boost/bind.hpp -- TEST INFRASTRUCTURE ONLY: the forms of boost::bind the reference's path uses:
// bind(f, _1, _2) on a free/static function, bind(&pair::second, _N) on a data member, and the relational
// composition `bind(...) < bind(...)` / `>` that Boost.Bind provides (list::sort comparators, map.cpp:140,
// reprojector.cpp:76,153).
#pragma once
#include <functional>
#include <tuple>
#include <utility>
namespace boost {
template <int N> struct arg {};
namespace bind_detail {
template <class F, class... B>
struct bound {
  F f;
  std::tuple<B...> b;
  template <class T, class Tup> static T&& pick(T&& v, Tup&) { return std::forward<T>(v); }
  template <int N, class Tup> static auto pick(arg<N>, Tup& t) -> decltype(std::get<N - 1>(t)) { return std::get<N - 1>(t); }
  template <class... A, size_t... I>
  decltype(auto) call(std::tuple<A&...> t, std::index_sequence<I...>) const {
    return std::invoke(f, pick(std::get<I>(b), t)...);
  }
  template <class... A> decltype(auto) operator()(A&&... a) const {
    std::tuple<A&...> t(a...);
    return call(t, std::index_sequence_for<B...>());
  }
};
template <class L, class R, class Op>
struct rel {
  L l; R r;
  template <class... A> bool operator()(A&&... a) const { return Op()(l(a...), r(a...)); }
};
}  // namespace bind_detail
template <class F, class... B> bind_detail::bound<F, B...> bind(F f, B... b) { return {f, std::tuple<B...>(b...)}; }
template <class F1, class... B1, class F2, class... B2>
bind_detail::rel<bind_detail::bound<F1, B1...>, bind_detail::bound<F2, B2...>, std::less<>> operator<(
    bind_detail::bound<F1, B1...> l, bind_detail::bound<F2, B2...> r) { return {l, r}; }
template <class F1, class... B1, class F2, class... B2>
bind_detail::rel<bind_detail::bound<F1, B1...>, bind_detail::bound<F2, B2...>, std::greater<>> operator>(
    bind_detail::bound<F1, B1...> l, bind_detail::bound<F2, B2...> r) { return {l, r}; }
}  // namespace boost
static boost::arg<1> _1;
static boost::arg<2> _2;
