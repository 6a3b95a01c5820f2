// Shadow of core/string_server.hpp (id <-> string maps loaded with Boost.MPI / HDFS): only named by ORDER BY and FILTER,
// which are outside this path.
#pragma once
#include <string>
#include "type.hpp"
class StringServer {
public:
    bool exist(sid_t) { return false; }
    bool exist(const std::string &) { return false; }
    std::string id2str(sid_t) { return std::string(); }
    sid_t str2id(const std::string &) { return 0; }
};
