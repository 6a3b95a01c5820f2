// Seeded LUBM-shaped ID-triple generator (synthetic stand-in for the external Java UBA tool).
//
// The reference ships no LUBM generator; it consumes ID-triples produced
// by UBA + datagen/generate_data.cpp (reference datagen/README.md, generate_data.cpp:113-125:
// id 0 = __PREDICATE__, id 1 = rdf:type, index ids from 2, normal vertex ids from 1<<17).
// This generator emits triples with the same id conventions and the UBA profile the LUBM
// queries Q1-Q7 depend on.  Every university is generated from
// its own RNG stream (seed, univ), so any range of universities can be produced in parallel
// and independently, and vertex ids are closed-form per university block.
//
// Exposed as a plain C API (used through ctypes and by the C++ host code).
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <omp.h>

#include "lubm_vocab.h"

namespace {

struct Rng {  // splitmix64
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed) {}
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    // uniform integer in [lo, hi]
    uint32_t range(uint32_t lo, uint32_t hi) { return lo + (uint32_t)(next() % (uint64_t)(hi - lo + 1)); }
    uint32_t below(uint32_t n) { return (uint32_t)(next() % n); }
};

struct Sink {
    uint32_t *out;      // may be null: count only
    uint64_t cap;       // capacity in triples
    uint64_t n = 0;
    // optional string dictionary for the written-to-disk small datasets
    std::vector<std::pair<uint32_t, std::string>> *names = nullptr;
    inline void emit(uint32_t s, uint32_t p, uint32_t o) {
        if (out && n < cap) {
            out[3 * n + 0] = s;
            out[3 * n + 1] = p;
            out[3 * n + 2] = o;
        }
        n++;
    }
    inline void name(uint32_t id, const std::string &str) {
        if (names) names->emplace_back(id, str);
    }
};

struct Faculty {
    uint32_t id;
    int kind;                       // 0 full, 1 assoc, 2 asst, 3 lecturer
    std::vector<uint32_t> pubs;
    std::vector<uint32_t> courses;  // undergraduate courses taught
};

static const char *FAC_NAME[4] = {"FullProfessor", "AssociateProfessor", "AssistantProfessor", "Lecturer"};
static const uint32_t FAC_TYPE[4] = {T_FULLPROF, T_ASSOCPROF, T_ASSTPROF, T_LECTURER};
static const int FAC_NAMEPOOL[4] = {NP_FULLPROF, NP_ASSOCPROF, NP_ASSTPROF, NP_LECTURER};

static std::string univ_iri(uint32_t u) { return "<http://www.University" + std::to_string(u) + ".edu>"; }
static std::string dept_iri(uint32_t u, uint32_t d) {
    return "<http://www.Department" + std::to_string(d) + ".University" + std::to_string(u) + ".edu>";
}
static std::string ent_iri(uint32_t u, uint32_t d, const std::string &what) {
    return "<http://www.Department" + std::to_string(d) + ".University" + std::to_string(u) + ".edu/" + what + ">";
}

static void gen_university(uint32_t u, uint32_t total_univs, uint64_t seed, Sink &sk) {
    Rng rng(seed * 0x100000001B3ull + 0xC0FFEEull * (uint64_t)(u + 1));
    const uint32_t ubase = lubm_univ_id(u);
    uint32_t next_local = LUBM_LOCAL_FIRST_FREE;
    auto fresh = [&]() -> uint32_t { return ubase + next_local++; };
    const bool want_names = sk.names != nullptr;

    // university
    sk.emit(ubase, P_TYPE, T_UNIVERSITY);
    sk.emit(ubase, P_NAME, ubase + LUBM_LOCAL_UNIV_NAME);
    if (want_names) {
        sk.name(ubase, univ_iri(u));
        sk.name(ubase + LUBM_LOCAL_UNIV_NAME, "\"University" + std::to_string(u) + "\"");
    }

    auto degree_from = [&](uint32_t who, uint32_t pred) {
        uint32_t v = rng.below(total_univs);
        uint32_t vid = lubm_univ_id(v);
        sk.emit(who, pred, vid);
        // UBA re-declares the referenced university in the referencing file; the loader's
        // dedup (reference base_loader.hpp:81-95) removes the duplicates.
        sk.emit(vid, P_TYPE, T_UNIVERSITY);
        if (want_names) sk.name(vid, univ_iri(v));
    };
    auto person_common = [&](uint32_t id, int namepool, uint32_t idx) {
        sk.emit(id, P_NAME, lubm_name_literal(namepool, idx));
        sk.emit(id, P_EMAIL, fresh());
        sk.emit(id, P_TELEPHONE, LUBM_LIT_TELEPHONE);
    };

    const uint32_t ndepts = rng.range(15, 25);
    for (uint32_t d = 0; d < ndepts; d++) {
        const uint32_t dept = lubm_dept_id(u, d);
        sk.emit(dept, P_TYPE, T_DEPARTMENT);
        sk.emit(dept, P_NAME, lubm_name_literal(NP_DEPT, d));
        sk.emit(dept, P_SUBORG, ubase);
        if (want_names) sk.name(dept, dept_iri(u, d));

        // ---- faculty -------------------------------------------------------------------
        const uint32_t nfac[4] = {rng.range(7, 10), rng.range(10, 14), rng.range(8, 11), rng.range(5, 7)};
        std::vector<Faculty> fac;
        std::vector<uint32_t> ug_courses, gr_courses;
        uint32_t course_idx = 0, gcourse_idx = 0;
        for (int k = 0; k < 4; k++) {
            for (uint32_t i = 0; i < nfac[k]; i++) {
                Faculty f;
                f.id = fresh();
                f.kind = k;
                if (want_names) sk.name(f.id, ent_iri(u, d, FAC_NAME[k] + std::to_string(i)));
                sk.emit(f.id, P_TYPE, FAC_TYPE[k]);
                person_common(f.id, FAC_NAMEPOOL[k], i);
                sk.emit(f.id, P_WORKSFOR, dept);
                degree_from(f.id, P_UGDEGREE);
                degree_from(f.id, P_MSDEGREE);
                degree_from(f.id, P_DRDEGREE);
                sk.emit(f.id, P_RESEARCHINT, lubm_research_literal(rng.below(LUBM_NUM_RESEARCH)));
                // courses taught
                uint32_t nc = rng.range(1, 2);
                for (uint32_t c = 0; c < nc; c++) {
                    uint32_t cid = fresh();
                    if (want_names) sk.name(cid, ent_iri(u, d, "Course" + std::to_string(course_idx)));
                    sk.emit(cid, P_TYPE, T_COURSE);
                    sk.emit(cid, P_NAME, lubm_name_literal(NP_COURSE, course_idx));
                    sk.emit(f.id, P_TEACHEROF, cid);
                    ug_courses.push_back(cid);
                    f.courses.push_back(cid);
                    course_idx++;
                }
                uint32_t ngc = rng.range(1, 2);
                for (uint32_t c = 0; c < ngc; c++) {
                    uint32_t cid = fresh();
                    if (want_names) sk.name(cid, ent_iri(u, d, "GraduateCourse" + std::to_string(gcourse_idx)));
                    sk.emit(cid, P_TYPE, T_GRADCOURSE);
                    sk.emit(cid, P_NAME, lubm_name_literal(NP_GRADCOURSE, gcourse_idx));
                    sk.emit(f.id, P_TEACHEROF, cid);
                    gr_courses.push_back(cid);
                    gcourse_idx++;
                }
                // publications
                static const uint32_t PUB_LO[4] = {15, 10, 5, 0}, PUB_HI[4] = {20, 18, 10, 5};
                uint32_t np = rng.range(PUB_LO[k], PUB_HI[k]);
                for (uint32_t p = 0; p < np; p++) {
                    uint32_t pid = fresh();
                    if (want_names)
                        sk.name(pid, ent_iri(u, d, std::string(FAC_NAME[k]) + std::to_string(i) + "/Publication" + std::to_string(p)));
                    sk.emit(pid, P_TYPE, T_PUBLICATION);
                    sk.emit(pid, P_NAME, lubm_name_literal(NP_PUBLICATION, p));
                    sk.emit(pid, P_PUBAUTHOR, f.id);
                    f.pubs.push_back(pid);
                }
                fac.push_back(std::move(f));
            }
        }
        const uint32_t nfaculty = (uint32_t)fac.size();
        const uint32_t nprof = nfac[0] + nfac[1] + nfac[2];  // professors come first in `fac`
        // head of department: a full professor
        sk.emit(fac[rng.below(nfac[0])].id, P_HEADOF, dept);

        // ---- undergraduate students -------------------------------------------------------
        const uint32_t nug = nfaculty * rng.range(8, 14);
        for (uint32_t i = 0; i < nug; i++) {
            uint32_t sid = fresh();
            if (want_names) sk.name(sid, ent_iri(u, d, "UndergraduateStudent" + std::to_string(i)));
            sk.emit(sid, P_TYPE, T_UGSTUDENT);
            person_common(sid, NP_UGSTUDENT, i);
            sk.emit(sid, P_MEMBEROF, dept);
            // 2-4 distinct undergraduate courses
            uint32_t nc = rng.range(2, 4);
            uint32_t picked[4];
            for (uint32_t c = 0; c < nc; c++) {
                uint32_t cid;
                bool dup;
                do {
                    cid = ug_courses[rng.below((uint32_t)ug_courses.size())];
                    dup = false;
                    for (uint32_t q = 0; q < c; q++) dup |= (picked[q] == cid);
                } while (dup);
                picked[c] = cid;
                sk.emit(sid, P_TAKESCOURSE, cid);
            }
            // one in five undergraduates has an advisor (a professor)
            if (rng.below(5) == 0) sk.emit(sid, P_ADVISOR, fac[rng.below(nprof)].id);
            // NOTE: no undergraduateDegreeFrom for undergraduates => LUBM Q3 is empty by construction
        }

        // ---- graduate students ------------------------------------------------------------
        const uint32_t ngrad = nfaculty * rng.range(3, 4);
        for (uint32_t i = 0; i < ngrad; i++) {
            uint32_t sid = fresh();
            if (want_names) sk.name(sid, ent_iri(u, d, "GraduateStudent" + std::to_string(i)));
            sk.emit(sid, P_TYPE, T_GRADSTUDENT);
            person_common(sid, NP_GRADSTUDENT, i);
            sk.emit(sid, P_MEMBEROF, dept);
            degree_from(sid, P_UGDEGREE);
            uint32_t nc = rng.range(1, 3);
            uint32_t picked[3];
            for (uint32_t c = 0; c < nc; c++) {
                uint32_t cid;
                bool dup;
                do {
                    cid = gr_courses[rng.below((uint32_t)gr_courses.size())];
                    dup = false;
                    for (uint32_t q = 0; q < c; q++) dup |= (picked[q] == cid);
                } while (dup);
                picked[c] = cid;
                sk.emit(sid, P_TAKESCOURSE, cid);
            }
            Faculty &adv = fac[rng.below(nprof)];
            sk.emit(sid, P_ADVISOR, adv.id);
            // co-author 0-5 of the advisor's publications
            uint32_t nco = rng.range(0, 5);
            if (nco > adv.pubs.size()) nco = (uint32_t)adv.pubs.size();
            uint32_t first = adv.pubs.empty() ? 0 : rng.below((uint32_t)adv.pubs.size());
            for (uint32_t c = 0; c < nco; c++)
                sk.emit(adv.pubs[(first + c) % adv.pubs.size()], P_PUBAUTHOR, sid);
            if (rng.below(5) == 0) {  // teaching assistant
                sk.emit(sid, P_TYPE, T_TEACHASSIST);
                sk.emit(sid, P_TAOF, ug_courses[rng.below((uint32_t)ug_courses.size())]);
            }
            if (rng.below(4) == 0) sk.emit(sid, P_TYPE, T_RESEARCHASSIST);
        }

        // ---- research groups --------------------------------------------------------------
        const uint32_t nrg = rng.range(10, 20);
        for (uint32_t i = 0; i < nrg; i++) {
            uint32_t gid = fresh();
            if (want_names) sk.name(gid, ent_iri(u, d, "ResearchGroup" + std::to_string(i)));
            sk.emit(gid, P_TYPE, T_RESEARCHGROUP);
            sk.emit(gid, P_SUBORG, dept);
        }
    }
    if (next_local > LUBM_UNIV_BLOCK) {
        fprintf(stderr, "lubm_gen: university %u overflowed its id block (%u)\n", u, next_local);
        abort();
    }
}

}  // namespace

extern "C" {

// Generate universities [u_begin, u_end) of a `total_univs`-university dataset.
// Writes up to `cap` triples (3 x uint32 each) into `out` (may be NULL to only count).
// Returns the number of triples generated (which may exceed cap: call again with more room).
uint64_t wkgen_lubm(uint32_t u_begin, uint32_t u_end, uint32_t total_univs, uint64_t seed,
                    uint32_t *out, uint64_t cap) {
    if (u_end <= u_begin) return 0;
    const uint32_t nu = u_end - u_begin;
    std::vector<uint64_t> counts(nu + 1, 0);
#pragma omp parallel for schedule(dynamic, 4)
    for (uint32_t i = 0; i < nu; i++) {
        Sink sk{nullptr, 0};
        gen_university(u_begin + i, total_univs, seed, sk);
        counts[i + 1] = sk.n;
    }
    for (uint32_t i = 0; i < nu; i++) counts[i + 1] += counts[i];
    const uint64_t total = counts[nu];
    if (!out || cap < total) return total;
#pragma omp parallel for schedule(dynamic, 4)
    for (uint32_t i = 0; i < nu; i++) {
        Sink sk{out + 3 * counts[i], counts[i + 1] - counts[i]};
        gen_university(u_begin + i, total_univs, seed, sk);
    }
    return total;
}

// Write a dataset directory in the reference's ID-triple format (datagen/README.md):
//   id_uni<u>.nt   "s\tp\to" decimal lines     str_index   "string\tid"     str_normal  "string\tid"
// Intended for small scales (tests of the directory reader).  Returns #triples or 0 on error.
uint64_t wkgen_lubm_write_dir(const char *dir, uint32_t total_univs, uint64_t seed) {
    std::string d(dir);
    if (!d.empty() && d.back() != '/') d += '/';
    {
        FILE *f = fopen((d + "str_index").c_str(), "w");
        if (!f) return 0;
        for (int i = 0; i < LUBM_NUM_INDEX_IDS; i++) fprintf(f, "%s\t%d\n", lubm_index_string(i), i);
        fclose(f);
    }
    std::vector<std::pair<uint32_t, std::string>> names;
    // shared literal pool
    names.emplace_back(LUBM_LIT_TELEPHONE, "\"xxx-xxx-xxxx\"");
    for (uint32_t i = 0; i < LUBM_NUM_RESEARCH; i++)
        names.emplace_back(lubm_research_literal(i), "\"Research" + std::to_string(i) + "\"");
    uint64_t total = 0;
    for (uint32_t u = 0; u < total_univs; u++) {
        Sink cnt{nullptr, 0};
        gen_university(u, total_univs, seed, cnt);
        std::vector<uint32_t> buf(3 * cnt.n);
        Sink sk{buf.data(), cnt.n};
        sk.names = &names;
        gen_university(u, total_univs, seed, sk);
        FILE *f = fopen((d + "id_uni" + std::to_string(u) + ".nt").c_str(), "w");
        if (!f) return 0;
        for (uint64_t i = 0; i < sk.n; i++) fprintf(f, "%u\t%u\t%u\n", buf[3 * i], buf[3 * i + 1], buf[3 * i + 2]);
        fclose(f);
        total += sk.n;
    }
    FILE *f = fopen((d + "str_normal").c_str(), "w");
    if (!f) return 0;
    std::sort(names.begin(), names.end());
    uint32_t prev = 0;
    bool first = true;
    for (auto &kv : names) {
        if (!first && kv.first == prev) continue;
        fprintf(f, "%s\t%u\n", kv.second.c_str(), kv.first);
        prev = kv.first;
        first = false;
    }
    fclose(f);
    return total;
}

uint32_t wkgen_lubm_univ_id(uint32_t u) { return lubm_univ_id(u); }
uint32_t wkgen_lubm_dept_id(uint32_t u, uint32_t d) { return lubm_dept_id(u, d); }
int wkgen_lubm_num_index_ids() { return LUBM_NUM_INDEX_IDS; }
const char *wkgen_lubm_index_string(int id) { return lubm_index_string(id); }

}  // extern "C"
