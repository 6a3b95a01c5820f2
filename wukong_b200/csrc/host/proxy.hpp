// Client stub with the surface of the reference's Proxy::run_single_query (core/proxy.hpp:298-385) and
// Monitor latency accounting (core/monitor.hpp:99-101): parse -> user-defined plan -> run `cnt` times,
// only the last run takes results back (unless global_silent).  In this single-process build the
// request goes straight to the engine object instead of through an Adaptor.
#pragma once
#include <chrono>
#include <istream>

#include "global.hpp"
#include "gpu_engine.hpp"
#include "parser.hpp"
#include "planner.hpp"
#include "string_server.hpp"

namespace wukong {

class Monitor {
    std::chrono::steady_clock::time_point t0, t1;
    int cnt = 1;
public:
    void init(int c = 1) { cnt = c < 1 ? 1 : c; t0 = std::chrono::steady_clock::now(); }
    void finish() { t1 = std::chrono::steady_clock::now(); }
    double latency_usec() const { return std::chrono::duration<double, std::micro>(t1 - t0).count() / cnt; }
};

class Proxy {
public:
    int sid, tid;
    StringServer *str_server;
    GPUEngine *engine;
    Global *global;
    Parser parser;
    Planner planner;

    Proxy(int sid, int tid, StringServer *ss, GPUEngine *eng, Global *g)
        : sid(sid), tid(tid), str_server(ss), engine(eng), global(g), parser(ss) {}

    // parse + plan only (what run_single_query does before sending)
    int prepare(std::istream &is, std::istream &fmt_stream, SPARQLQuery &request) {
        if (!parser.parse(is, request)) return SYNTAX_ERROR;
        if (global->enable_planner) return SETTING_ERROR;          // the optimiser is out of scope: plans come from .fmt files
        if (!planner.set_plan(request.pattern_group, fmt_stream)) return SYNTAX_ERROR;
        return SUCCESS;
    }

    // Run a single query `cnt` times (console "sparql -f <query> -p <plan> -n <cnt> -g").
    int run_single_query(std::istream &is, std::istream &fmt_stream, int mt_factor, int cnt, bool per_pattern,
                         SPARQLQuery &reply, Monitor &monitor) {
        SPARQLQuery request;
        int rc = prepare(is, fmt_stream, request);
        if (rc != SUCCESS) { reply.result.set_status_code(rc); return rc; }
        request.mt_factor = 1;   // one GPU engine scans the whole index slice (mt_factor only splits CPU engines)
        (void)mt_factor;
        request.dev_type = SPARQLQuery::GPU;
        monitor.init(cnt);
        for (int i = 0; i < cnt; i++) {
            SPARQLQuery q = request;
            q.result.blind = i < (cnt - 1) ? true : global->silent;   // only the last request takes results back
            if (per_pattern) engine->execute_patterns(q); else engine->execute_sparql_query(q);
            reply = q;
        }
        monitor.finish();
        return reply.result.status_code;
    }
};

}  // namespace wukong
