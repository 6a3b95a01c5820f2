// Stand-in for <boost/variant.hpp> (see README.md). Attribute values (attr_t) are not on the graph-exploration path.
#pragma once
namespace boost {
template <typename... T>
class variant {
public:
    variant() {}
    template <typename U> variant(const U &) {}
    bool operator<(const variant &) const { return false; }
    bool operator==(const variant &) const { return true; }
    int which() const { return 0; }
};
template <typename R>
class static_visitor { public: typedef R result_type; };
template <typename Visitor, typename V>
typename Visitor::result_type apply_visitor(const Visitor &, const V &) { return typename Visitor::result_type(); }
template <typename U, typename... T>
U get(const variant<T...> &) { return U(); }
}  // namespace boost
