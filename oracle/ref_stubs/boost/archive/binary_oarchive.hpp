// no-op archives (see ../../README.md): nothing is ever serialised by the oracle/_ref shims
#pragma once
#include <istream>
#include <ostream>
namespace boost { namespace archive {
class binary_oarchive {
public:
    explicit binary_oarchive(std::ostream &) {}
    template <class T> binary_oarchive &operator<<(const T &) { return *this; }
    template <class T> binary_oarchive &operator&(const T &) { return *this; }
};
class binary_iarchive {
public:
    explicit binary_iarchive(std::istream &) {}
    template <class T> binary_iarchive &operator>>(T &) { return *this; }
    template <class T> binary_iarchive &operator&(T &) { return *this; }
};
} }
