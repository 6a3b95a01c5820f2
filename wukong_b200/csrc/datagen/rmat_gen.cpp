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

// Bijection on [0, 2^scale): Graph500 relabels the vertices the same way ("scramble"), because raw R-MAT ids carry the
// degree in their bit pattern -- ids whose low bits are zero are the hubs, so a vid % n sharding of the raw ids would give
// shard 0 several times the edges of the others.
static inline uint32_t rmat_scramble(uint32_t x, int scale) {
    const uint32_t mask = scale >= 32 ? 0xFFFFFFFFu : ((1u << scale) - 1u);
    const int h = scale > 1 ? scale / 2 : 1;
    x ^= x >> h;
    x = (x * 0x9E3779B1u) & mask;
    x ^= x >> h;
    x = (x * 0x85EBCA6Bu) & mask;
    x ^= x >> h;
    return x;
}

// Writes `nedges` (s, 2, o) triples into out (3 x uint32 each).  Returns nedges.  The edge list depends on (scale, nedges,
// seed, a, b, c, scramble) only, not on the number of threads: every block of 2^16 edges has its own stream.
uint64_t wkgen_rmat_edges(int scale, uint64_t nedges, uint64_t seed, double a, double b, double c, int scramble, uint32_t *out) {
    const uint64_t A = (uint64_t)(a * 4294967296.0), AB = A + (uint64_t)(b * 4294967296.0), ABC = AB + (uint64_t)(c * 4294967296.0);
    const uint64_t BLOCK = 1ull << 16;
    const int64_t nblocks = (int64_t)((nedges + BLOCK - 1) / BLOCK);
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t blk = 0; blk < nblocks; blk++) {
        const uint64_t lo = (uint64_t)blk * BLOCK, hi = lo + BLOCK < nedges ? lo + BLOCK : nedges;
        Rng rng(seed * 0x100000001B3ull + 0xABCDEFull * (uint64_t)(blk + 1));
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
            if (scramble) { u = rmat_scramble(u, scale); v = rmat_scramble(v, scale); }
            out[3 * e + 0] = RMAT_VID_BASE + u;
            out[3 * e + 1] = RMAT_PRED;
            out[3 * e + 2] = RMAT_VID_BASE + v;
        }
    }
    return nedges;
}

}  // extern "C"
