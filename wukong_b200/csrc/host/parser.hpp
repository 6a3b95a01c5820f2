// Minimal SPARQL reader for basic graph patterns: PREFIX*, SELECT ?v..., WHERE { s p o . ... }.
// It covers what scripts/sparql_query/lubm/basic/lubm_q1..q7 use (the reference's full front-end,
// core/SPARQLParser.hpp + core/parser.hpp, is out of scope).  Variable numbering follows the
// reference: -1, -2, ... in order of first appearance, the SELECT list first
// (SPARQLParser.hpp:226-235, 1108-1135; parser.hpp:186-197).
#pragma once
#include <cstdlib>
#include <istream>
#include <map>
#include <sstream>
#include <string>

#include "query.hpp"
#include "string_server.hpp"

namespace wukong {

class Parser {
    StringServer *str_server;
    std::map<std::string, ssid_t> vars;
    std::map<std::string, std::string> prefixes;

    ssid_t var_id(const std::string &name) {
        auto it = vars.find(name);
        if (it != vars.end()) return it->second;
        ssid_t id = -(ssid_t)(vars.size() + 1);
        vars[name] = id;
        return id;
    }
    bool term(const std::string &tok, ssid_t &out) {
        if (tok.empty()) return false;
        if (tok[0] == '?') { out = var_id(tok.substr(1)); return true; }
        std::string iri;
        if (tok == "__PREDICATE__") iri = tok;
        else if (tok[0] == '<') iri = tok;
        else {
            size_t c = tok.find(':');
            if (c == std::string::npos) { strerror = "bad token " + tok; return false; }
            auto it = prefixes.find(tok.substr(0, c));
            if (it == prefixes.end()) { strerror = "unknown prefix in " + tok; return false; }
            iri = "<" + it->second + tok.substr(c + 1) + ">";
        }
        if (!str_server->exist(iri)) { strerror = "unknown IRI " + iri; return false; }
        out = (ssid_t)str_server->str2id(iri);
        return true;
    }

public:
    std::string strerror;
    explicit Parser(StringServer *ss) : str_server(ss) {}

    bool parse(std::istream &is, SPARQLQuery &sq) {
        vars.clear();
        prefixes.clear();
        strerror.clear();
        std::vector<std::string> toks;
        std::string t;
        while (is >> t) toks.push_back(t);
        size_t i = 0;
        auto upper = [](std::string s) { for (auto &c : s) c = (char)toupper(c); return s; };
        while (i + 2 < toks.size() && upper(toks[i]) == "PREFIX") {
            std::string name = toks[i + 1], iri = toks[i + 2];
            if (name.empty() || name.back() != ':' || iri.size() < 2 || iri[0] != '<') { strerror = "bad PREFIX"; return false; }
            prefixes[name.substr(0, name.size() - 1)] = iri.substr(1, iri.size() - 2);
            i += 3;
        }
        if (i >= toks.size() || upper(toks[i]) != "SELECT") { strerror = "SELECT expected"; return false; }
        i++;
        bool distinct = false;
        if (i < toks.size() && (upper(toks[i]) == "DISTINCT" || upper(toks[i]) == "REDUCED")) { distinct = true; i++; }   // SPARQLParser.hpp: both set `distinct`
        std::vector<ssid_t> required;
        while (i < toks.size() && toks[i][0] == '?') required.push_back(var_id(toks[i++].substr(1)));
        if (i >= toks.size() || upper(toks[i]) != "WHERE") { strerror = "WHERE expected"; return false; }
        i++;
        if (i >= toks.size() || toks[i] != "{") { strerror = "{ expected"; return false; }
        i++;
        SPARQLQuery::PatternGroup pg;
        while (i < toks.size() && toks[i] != "}") {
            if (i + 2 >= toks.size()) { strerror = "incomplete triple pattern"; return false; }
            ssid_t s, p, o;
            std::string ot = toks[i + 2];
            bool dot_glued = ot.size() > 1 && ot.back() == '.' && ot[ot.size() - 2] != '>' ? false : false;
            (void)dot_glued;
            if (!term(toks[i], s) || !term(toks[i + 1], p) || !term(ot, o)) return false;
            pg.patterns.push_back(SPARQLQuery::Pattern(s, p, OUT, o));
            i += 3;
            if (i < toks.size() && toks[i] == ".") i++;
        }
        if (i >= toks.size()) { strerror = "} expected"; return false; }
        if (pg.patterns.empty()) { strerror = "empty group"; return false; }
        i++;
        // solution modifiers: LIMIT n / OFFSET n in either order (ORDER BY needs string comparison on the proxy: rejected)
        int limit = -1;
        unsigned offset = 0;
        while (i < toks.size()) {
            const std::string kw = upper(toks[i]);
            if ((kw == "LIMIT" || kw == "OFFSET") && i + 1 < toks.size()) {
                char *end = nullptr;
                const long v = strtol(toks[i + 1].c_str(), &end, 10);
                if (*end != 0 || v < 0) { strerror = "bad " + kw; return false; }
                if (kw == "LIMIT") limit = (int)v; else offset = (unsigned)v;
                i += 2;
            } else {
                strerror = "unsupported solution modifier: " + toks[i];
                return false;
            }
        }
        sq = SPARQLQuery(pg, (int)vars.size(), required);
        sq.distinct = distinct;
        sq.limit = limit;
        sq.offset = offset;
        return true;
    }
};

}  // namespace wukong
