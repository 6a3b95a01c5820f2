// User-defined query plans: the .fmt reader of the reference planner (core/planner.hpp:1647-1754).
// Each non-comment line is "<pattern number, 1-based> <direction>":
//   >  keep (s, p, OUT, o)          <  swap subject/object, direction IN
//   << (p, PREDICATE_ID, IN, s)     >> (p, PREDICATE_ID, OUT, o)      -- seed from the predicate index
// The cost-based optimiser itself is out of scope: generate_plan() reports "no plan".
#pragma once
#include <istream>
#include <sstream>
#include <string>

#include "query.hpp"

namespace wukong {

class Planner {
    static std::string trim(const std::string &s) {
        size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
        return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
    }
    void set_direction(SPARQLQuery::PatternGroup &group, const std::vector<int> &orders, const std::vector<std::string> &dirs) {
        std::vector<SPARQLQuery::Pattern> patterns;
        for (size_t i = 0; i < orders.size(); i++) {
            SPARQLQuery::Pattern pt = group.patterns.at(orders[i] - 1);
            if (dirs[i] == "<") { pt.direction = IN; std::swap(pt.subject, pt.object); }
            else if (dirs[i] == ">") { pt.direction = OUT; }
            else if (dirs[i] == "<<") { pt.direction = IN; pt.object = pt.subject; pt.subject = pt.predicate; pt.predicate = PREDICATE_ID; }
            else if (dirs[i] == ">>") { pt.direction = OUT; pt.subject = pt.predicate; pt.predicate = PREDICATE_ID; }
            patterns.push_back(pt);
        }
        group.patterns = patterns;
    }

public:
    bool generate_plan(SPARQLQuery &) { return false; }   // optimiser: out of scope

    // @return false if no plan is set (wrong format: fewer steps than patterns, bad pattern number)
    bool set_plan(SPARQLQuery::PatternGroup &group, std::istream &fmt_stream) {
        if (!fmt_stream.good()) return false;
        std::vector<int> orders;
        std::vector<std::string> dirs;
        std::string line;
        while (std::getline(fmt_stream, line)) {
            line = trim(line);
            if (line.empty() || line[0] == '#' || line == "{") continue;
            if (line == "}") break;
            std::istringstream iss(line);
            int order = 0;
            std::string dir = ">";
            iss >> order >> dir;
            if (order < 1 || order > (int)group.patterns.size()) return false;
            orders.push_back(order);
            dirs.push_back(dir);
        }
        if (orders.size() < group.patterns.size()) return false;
        set_direction(group, orders, dirs);
        return true;
    }
};

}  // namespace wukong
