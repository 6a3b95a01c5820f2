// the handful of string helpers the reference's planner / loaders call (see README.md), over <string>
#pragma once
#include <algorithm>
#include <cctype>
#include <string>
#include <vector>
namespace boost {
inline void trim(std::string &s) {
    size_t a = 0, b = s.size();
    while (a < b && std::isspace((unsigned char)s[a])) a++;
    while (b > a && std::isspace((unsigned char)s[b - 1])) b--;
    s = s.substr(a, b - a);
}
inline bool starts_with(const std::string &s, const std::string &p) { return s.size() >= p.size() && s.compare(0, p.size(), p) == 0; }
inline bool ends_with(const std::string &s, const std::string &p) { return s.size() >= p.size() && s.compare(s.size() - p.size(), p.size(), p) == 0; }
inline std::string to_lower_copy(std::string s) { for (auto &c : s) c = (char)std::tolower((unsigned char)c); return s; }
struct is_any_of { std::string set; explicit is_any_of(const std::string &s) : set(s) {} bool operator()(char c) const { return set.find(c) != std::string::npos; } };
enum token_compress_mode_type { token_compress_on, token_compress_off };
template <class Seq, class Pred>
Seq &split(Seq &out, const std::string &in, Pred pred, token_compress_mode_type m = token_compress_off) {
    out.clear();
    std::string cur;
    for (size_t i = 0; i < in.size(); i++) {
        if (pred(in[i])) { if (!(m == token_compress_on && cur.empty() && !out.empty())) out.push_back(cur); cur.clear(); }
        else cur.push_back(in[i]);
    }
    out.push_back(cur);
    return out;
}
}  // namespace boost
