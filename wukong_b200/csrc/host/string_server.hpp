// String <-> id mapping loaded from an ID-triple dataset directory (reference core/string_server.hpp:
// str_index = predicates and types, str_normal = entities; "string \t id" per line, datagen/README.md).
#pragma once
#include <fstream>
#include <string>
#include <unordered_map>

#include "type.hpp"

namespace wukong {

class StringServer {
public:
    std::unordered_map<std::string, sid_t> str2id_map;
    std::unordered_map<sid_t, std::string> id2str_map;
    int num_index_lines = 0;   // lines of str_index; the loader derives num_normal_preds = lines - 1

    StringServer() {}
    explicit StringServer(std::string dname) { load(dname); }

    bool load(std::string dname) {
        if (!dname.empty() && dname.back() != '/') dname += '/';
        num_index_lines = load_file(dname + "str_index", true);
        load_file(dname + "str_normal", false);
        return num_index_lines > 0;
    }
    bool exist(const std::string &s) const { return str2id_map.count(s) != 0; }
    bool exist(sid_t id) const { return id2str_map.count(id) != 0; }
    sid_t str2id(const std::string &s) const { return str2id_map.at(s); }
    const std::string &id2str(sid_t id) const { return id2str_map.at(id); }
    void add(const std::string &s, sid_t id) { str2id_map[s] = id; id2str_map[id] = s; }

private:
    int load_file(const std::string &fname, bool) {
        std::ifstream f(fname.c_str());
        if (!f) return 0;
        std::string line;
        int n = 0;
        while (std::getline(f, line)) {
            size_t tab = line.find_last_of(" \t");
            if (tab == std::string::npos) continue;
            std::string s = line.substr(0, line.find_last_not_of(" \t", tab) + 1);
            sid_t id = (sid_t)strtoul(line.c_str() + tab + 1, nullptr, 10);
            add(s, id);
            n++;
        }
        return n;
    }
};

}  // namespace wukong
