/* abi_smoke.c — plain C (gcc, no C++/CUDA headers) against include/vega_b200.h: the same calls a cgo / JNI /
 * Rust `extern "C"` binding would make.  count_by_value of tests/test_pair_rdd.rs:84-109 → prints (1,2) (2,3) (3,4).
 * Needs a GPU to run; building it only proves the header is a valid C ABI and the library links from C. */
#include <stdio.h>
#include <stdlib.h>

#include "vega_b200.h"

#define CHECK(x) do { int rc_ = (x); if (rc_ != VB_OK) { fprintf(stderr, "%s -> %d: %s\n", #x, rc_, vb_last_error()); return 2; } } while (0)

int main(void)
{
    const uint64_t keys[9] = {1, 2, 1, 3, 2, 3, 3, 2, 3};
    uint64_t starts[6];
    const uint64_t n_map = vb_slice(9, 4, starts);                  /* ParallelCollection::slice */
    vb_ctx *ctx = NULL;
    vb_shuf *s = NULL;
    CHECK(vb_ctx_create(0, &ctx));
    CHECK(vb_shuffle_create(ctx, 0, (uint32_t)n_map, 4, VB_I64, VB_U64, VB_AGG_COUNT, VB_PART_HASH_METRO64, &s));
    CHECK(vb_shuffle_set_key_width(s, 4));                          /* i32 keys hash as 4 bytes */
    for (uint64_t m = 0; m < n_map; ++m)                            /* ShuffleMapTask::run per partition */
        CHECK(vb_shuffle_map_soa(s, (uint32_t)m, keys + starts[m], NULL, starts[m + 1] - starts[m], VB_HOST));
    CHECK(vb_shuffle_seal(s));                                      /* register_map_outputs */
    uint64_t total = 0;
    for (uint32_t r = 0; r < 4; ++r) {                              /* ShuffledRdd::compute per partition */
        uint64_t nk = 0, nv = 0;
        CHECK(vb_shuffle_reduce_size(s, r, &nk, &nv));
        uint64_t *k = (uint64_t *)malloc((nk + 1) * 8), *c = (uint64_t *)malloc((nk + 1) * 8);
        CHECK(vb_shuffle_reduce(s, r, k, c, NULL, NULL, VB_HOST));
        for (uint64_t i = 0; i < nk; ++i) {
            printf("(%llu, %llu) in partition %u (partitioner says %u)\n", (unsigned long long)k[i], (unsigned long long)c[i], r,
                   vb_get_partition(k[i], 4, 4));
            total += c[i];
        }
        free(k); free(c);
    }
    CHECK(vb_shuffle_free(s));
    CHECK(vb_ctx_destroy(ctx));
    return total == 9 ? 0 : 1;
}
