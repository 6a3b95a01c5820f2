/* Plain-C use of the C ABI (include/wukong_b200.h): build a store on the GPU from id triples, run one planned query,
 * print the answer.  No C++, no Python.
 *
 *   gcc -std=c99 -I include examples/query_c_abi.c -L wukong_b200 -l:libwukong_b200.so -Wl,-rpath,$PWD/wukong_b200 -o query_c_abi
 *   ./query_c_abi triples.bin 31        # triples.bin: n x (s, p, o) little-endian uint32; 31 = ids in str_index - 1
 *
 * The query below is LUBM Q5 under its OSDI16 plan: the members of Department0.University0 that are research groups
 * (ids from the LUBM-shaped generator of this repo: wukong_b200/datagen.py). */
#include <inttypes.h>
#include <stdio.h>
#include <stdlib.h>

#include "wukong_b200.h"

static int fail(const char *what, int rc) {
    fprintf(stderr, "%s: status %d (%s)\n", what, rc, wk_strerror(rc));
    return 1;
}

int main(int argc, char **argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s triples.bin num_normal_preds\n", argv[0]); return 2; }
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    fseek(f, 0, SEEK_END);
    const uint64_t n = (uint64_t)ftell(f) / 12;
    fseek(f, 0, SEEK_SET);
    wk_sid_t *triples = (wk_sid_t *)malloc(n * 12 + 12);
    if (fread(triples, 12, n, f) != n) { fprintf(stderr, "short read\n"); return 2; }
    fclose(f);

    int ndev = 0, rc = wk_device_count(&ndev);
    if (rc) return fail("wk_device_count", rc);            /* no CUDA device: there is no CPU fallback */

    wk_build_opts_t opts = {0};
    opts.num_servers = 1;
    opts.num_normal_preds = atoi(argv[2]);
    wk_store_t *store = NULL;
    wk_build_stats_t st;
    rc = wk_store_build(0, triples, n, &opts, &store, &st);
    if (rc) return fail("wk_store_build", rc);
    printf("store: %" PRIu64 " keys, %" PRIu64 " slots, built in %.1f ms\n", st.num_keys, st.num_slots, st.ms_total);

    wk_engine_t *engine = NULL;
    rc = wk_engine_create(store, 256u << 20, &engine);     /* two result buffers of 256 MB */
    if (rc) return fail("wk_engine_create", rc);

    /* ?X subOrganizationOf <Department0.University0> . ?X rdf:type ub:ResearchGroup   (planned: "1 <", "2 >") */
    const wk_pattern_t plan[2] = {{262146, 7, WK_DIR_IN, -1}, {-1, WK_TYPE_ID, WK_DIR_OUT, 28}};
    const int32_t required[1] = {-1};
    wk_sid_t table[4096];
    uint64_t rows = 0;
    int cols = 0;
    rc = wk_query_execute(engine, plan, 2, /*nvars*/ 1, required, 1, /*mt_tid*/ 0, /*mt_factor*/ 1, /*blind*/ 0,
                          table, 4096, &rows, &cols);
    if (rc) return fail("wk_query_execute", rc);
    printf("%" PRIu64 " rows x %d cols\n", rows, cols);
    for (uint64_t i = 0; i < rows && i < 16; i++) printf("  ?X = %u\n", table[i * (uint64_t)cols]);

    wk_engine_destroy(engine);
    wk_store_destroy(store);
    free(triples);
    return 0;
}
