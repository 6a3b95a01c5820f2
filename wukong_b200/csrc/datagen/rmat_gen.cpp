// Seeded R-MAT power-law graph generator for the roofline scan (BASELINE.json config 5):
// V = 2^scale vertices, E directed edges under one predicate (pid 2), plus one type (pid 1 -> type id 3)
// for every vertex that occurs.  Quadrant probabilities (a, b, c, d) as in Graph500 (0.57, 0.19, 0.19, 0.05).
// Vertex ids are offset by 2^17 (normal-vertex id space of the reference, generate_data.cpp:122-125).
// Duplicate edges are left in: the store builder's dedup removes them (base_loader.hpp:81-95).
#include <cstdint>
#include <omp.h>

namespace {
struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed) {}
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
};
}  // namespace

extern "C" {

enum { RMAT_PRED = 2, RMAT_TYPE = 3, RMAT_VID_BASE = 1 << 17 };

// Writes `nedges` (s, 2, o) triples into out (3 x uint32 each).  Returns nedges.
uint64_t wkgen_rmat_edges(int scale, uint64_t nedges, uint64_t seed, double a, double b, double c, uint32_t *out) {
    const uint64_t A = (uint64_t)(a * 4294967296.0), AB = A + (uint64_t)(b * 4294967296.0), ABC = AB + (uint64_t)(c * 4294967296.0);
#pragma omp parallel
    {
        const int nt = omp_get_num_threads(), t = omp_get_thread_num();
        const uint64_t lo = nedges * t / nt, hi = nedges * (t + 1) / nt;
        Rng rng(seed * 0x100000001B3ull + 0xABCDEFull * (uint64_t)(t + 1) + 17 * (uint64_t)nt);
        for (uint64_t e = lo; e < hi; e++) {
            uint32_t u = 0, v = 0;
            for (int l = 0; l < scale; l += 2) {   // two levels per 64-bit draw
                const uint64_t r = rng.next();
                const uint64_t r0 = r & 0xFFFFFFFFull, r1 = r >> 32;
                const int q0 = r0 < A ? 0 : (r0 < AB ? 1 : (r0 < ABC ? 2 : 3));
                u = (u << 1) | (uint32_t)(q0 >> 1);
                v = (v << 1) | (uint32_t)(q0 & 1);
                if (l + 1 < scale) {
                    const int q1 = r1 < A ? 0 : (r1 < AB ? 1 : (r1 < ABC ? 2 : 3));
                    u = (u << 1) | (uint32_t)(q1 >> 1);
                    v = (v << 1) | (uint32_t)(q1 & 1);
                }
            }
            out[3 * e + 0] = RMAT_VID_BASE + u;
            out[3 * e + 1] = RMAT_PRED;
            out[3 * e + 2] = RMAT_VID_BASE + v;
        }
    }
    return nedges;
}

}  // extern "C"
