// Shadow of core/comm/adaptor.hpp (TCP / RDMA transports): a single-server shim never sends.
#pragma once
#include <stdexcept>
#include <string>
#include "query.hpp"
class Adaptor {
public:
    int tid;
    Adaptor(int tid) : tid(tid) {}
    bool send(int, int, const std::string &) { throw std::runtime_error("oracle/_ref: no transport"); }
    bool send(int, int, const Bundle &) { throw std::runtime_error("oracle/_ref: no transport"); }
    bool send_dev2host(int, int, char *, uint64_t) { throw std::runtime_error("oracle/_ref: no transport"); }
    Bundle recv() { throw std::runtime_error("oracle/_ref: no transport"); }
    std::string recv(int) { throw std::runtime_error("oracle/_ref: no transport"); }
    bool tryrecv(std::string &) { return false; }
    bool tryrecv(Bundle &) { return false; }
};
