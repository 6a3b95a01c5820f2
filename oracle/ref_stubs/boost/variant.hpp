// Stand-in for <boost/variant.hpp>, used ONLY to compile the reference's utils/variant.hpp when
// oracle/ref_layout_shim.cpp pulls core/store/vertex.hpp (Boost is not in this image). Attribute
// values (attr_t) are not on the graph-exploration path and are never constructed by the shim.
#pragma once
namespace boost {
template <typename... T>
class variant {
public:
    variant() {}
    template <typename U> variant(const U &) {}
    bool operator<(const variant &) const { return false; }
    bool operator==(const variant &) const { return true; }
};
template <typename R>
class static_visitor {};
}  // namespace boost
