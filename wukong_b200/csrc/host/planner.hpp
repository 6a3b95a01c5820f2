// User-defined query plans: the .fmt reader of the reference planner (core/planner.hpp:1647-1754).
// Each non-comment line is "<pattern number, 1-based> <direction>":
//   >  keep (s, p, OUT, o)          <  swap subject/object, direction IN
//   << (p, PREDICATE_ID, IN, s)     >> (p, PREDICATE_ID, OUT, o)      -- seed from the predicate index
// "UNION {" ... "}" and "OPTIONAL {" ... "}" blocks plan the group's union / optional sub-groups in order (:1722-1738).
// The cost-based optimiser itself is out of scope: generate_plan() reports "no plan".
#pragma once
#include <cctype>
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

    // @return false if no plan is set (wrong format: fewer steps than patterns, bad pattern number, a UNION / OPTIONAL block
    // the group does not have).  Like the reference (planner.hpp:1700-1754) a "union ..." / "optional ..." line hands the
    // stream to the next union / optional sub-group, which reads up to its closing brace; the lines of a group itself may
    // come before, between and after its blocks.
    bool set_plan(SPARQLQuery::PatternGroup &group, std::istream &fmt_stream) {
        if (!fmt_stream.good()) return false;
        std::vector<int> orders;
        std::vector<std::string> dirs;
        std::string line;
        size_t nunions = 0, noptionals = 0;
        auto lower_starts = [](const std::string &l, const char *w) {
            size_t i = 0;
            for (; w[i]; i++)
                if (i >= l.size() || (char)tolower((unsigned char)l[i]) != w[i]) return false;
            return true;
        };
        while (std::getline(fmt_stream, line)) {
            line = trim(line);
            if (line.empty() || line[0] == '#' || line == "{") continue;
            if (line == "}") break;
            if (lower_starts(line, "union")) {
                if (nunions >= group.unions.size()) return false;
                set_plan(group.unions[nunions++], fmt_stream);     // like the reference, a refused sub-plan leaves that sub-group as parsed
                continue;
            }
            if (lower_starts(line, "optional")) {
                if (noptionals >= group.optional.size()) return false;
                set_plan(group.optional[noptionals++], fmt_stream);
                continue;
            }
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
