/* Plain-C consumer of include/fsgpu.h: proves the boundary is a C ABI (no C++/torch types) and shows
 * the call sequence a non-Python host makes.  Build (needs libfsgpu.so and a B200 to RUN):
 *   gcc -std=c99 -I include examples/c_abi_smoke.c -L frankensearch_b200 -lfsgpu -o c_abi_smoke
 * tests/test_abi.py compiles it with -fsyntax-only on every CPU run. */
#include <stdio.h>
#include <stdlib.h>

#include "fsgpu.h"

int main(void) {
    enum { N = 1000, D = 128, K = 5 };
    float* rows = (float*)malloc(sizeof(float) * N * D);
    float query[D];
    fsgpu_hit hits[K];
    uint32_t count = 0;
    fsgpu_index* ix = NULL;
    fsgpu_index_options opts;
    int i, rc;
    for (i = 0; i < N * D; ++i) rows[i] = (float)((i * 2654435761u) >> 8 & 0xFFFF) / 65536.0f - 0.5f;
    for (i = 0; i < D; ++i) query[i] = rows[7 * D + i];
    fsgpu_index_options_default(&opts);
    rc = fsgpu_index_create_f32(rows, N, D, NULL, &opts, &ix);
    if (rc != FSGPU_OK) {
        fprintf(stderr, "create failed (%d): %s\n", rc, fsgpu_last_error());
        return 1;
    }
    rc = fsgpu_search_top_k(ix, query, 1, K, D, hits, &count);
    if (rc != FSGPU_OK) {
        fprintf(stderr, "search failed (%d): %s\n", rc, fsgpu_last_error());
        return 1;
    }
    for (i = 0; i < (int)count; ++i) printf("%u %.6f\n", hits[i].row, hits[i].score);
    fsgpu_index_destroy(ix);
    free(rows);
    return count == K && hits[0].row == 7 ? 0 : 2;
}
