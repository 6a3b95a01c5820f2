#pragma once
namespace boost { namespace serialization {
class access {};
struct object_serializable {};
struct track_never {};
template <class A, class T> void split_free(A &, T &, const unsigned int) {}
template <class A, class T> void split_member(A &, T &, const unsigned int) {}
template <class B, class D> B &base_object(D &d) { return d; }
} }
#define BOOST_SERIALIZATION_SPLIT_FREE(T)
#define BOOST_SERIALIZATION_SPLIT_MEMBER()
#define BOOST_CLASS_IMPLEMENTATION(T, L)
#define BOOST_CLASS_TRACKING(T, L)
