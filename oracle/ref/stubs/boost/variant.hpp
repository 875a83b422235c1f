#pragma once
// Stand-in for <boost/variant.hpp> over std::variant: the subset libcore/properties.cpp uses.
#include <variant>
namespace boost {
template <class... T> using variant = std::variant<T...>;
template <class T, class... Ts> T *get(std::variant<Ts...> *v) { return std::get_if<T>(v); }
template <class T, class... Ts> const T *get(const std::variant<Ts...> *v) { return std::get_if<T>(v); }
template <class R> struct static_visitor { typedef R result_type; };
template <class V, class Var> typename V::result_type apply_visitor(V &vis, Var &&v) { return std::visit(vis, v); }
template <class V, class Var> typename V::result_type apply_visitor(const V &vis, Var &&v) { return std::visit(vis, v); }
}
