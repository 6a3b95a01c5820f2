// declarations only (see README.md): the oracle/_ref shims never open a socket
#pragma once
#include <cstddef>
#include <cstring>
#include <stdexcept>
#define ZMQ_PULL 7
#define ZMQ_PUSH 8
#define ZMQ_NOBLOCK 1
#define ZMQ_DONTWAIT 1
namespace zmq {
class context_t { public: context_t(int = 1) {} };
class message_t {
    char *d = nullptr; size_t n = 0;
public:
    message_t() {}
    explicit message_t(size_t sz) : d(new char[sz]), n(sz) {}
    ~message_t() { delete[] d; }
    void *data() { return d; }
    size_t size() const { return n; }
};
class socket_t {
public:
    socket_t(context_t &, int) {}
    void bind(const char *) { throw std::runtime_error("zmq stub"); }
    void connect(const char *) { throw std::runtime_error("zmq stub"); }
    bool send(message_t &, int = 0) { throw std::runtime_error("zmq stub"); }
    bool recv(message_t *, int = 0) { throw std::runtime_error("zmq stub"); }
};
}
