/*
 * vega_oracle.c — CPU restatement of rajasekarv/vega's shuffle + aggregation path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (vega_b200/, include/)
 * may link, import or call this file.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it, as the checker and as the
 * timed CPU baseline ("port": vega itself is Rust and cannot be built in this image).
 *
 * Parity pinning: the algorithm below is checked against every golden vector the
 * reference's own tests hold for this path (tests/test_pair_rdd.rs:8-135,
 * tests/test_rdd.rs:285-322,387-456,484-521,675-699) in tests/test_oracle_golden.py.
 * The *partition placement* hash (MetroHash64_1, third-party crate fasthash 0.4.0,
 * Cargo.toml:20, source not under /root/reference) is restated from the published
 * algorithm and pinned to MetroHash's published known-answer vector (63-byte test key,
 * seeds 0 and 1: 658F044F5C730E40 / AE49EBB0A856537B — every tail branch is exercised).
 * What stays "parity unpinned" is the binding fasthash::MetroHasher == metrohash64_1(seed 0)
 * over the bytes Rust's Hash impl writes: the reference's only test at that boundary
 * (src/partitioner.rs:63-82) asserts nothing about hash values.  No observable
 * result (collected multiset of (K,C), ordered value lists) depends on it.
 *
 * Each function cites the reference file:line it follows (paths relative to
 * /root/reference).
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
/* MetroHash64_1 (fasthash 0.4.0 `MetroHasher` = metro::Hasher64_1, seed 0).  */
/* Called from src/partitioner.rs:21-25 (`hash`) and :54-57 (`get_partition`). */
/* ------------------------------------------------------------------------- */
static inline uint64_t rotr64(uint64_t v, unsigned k) { return (v >> k) | (v << (64 - k)); }
static inline uint64_t rd64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }
static inline uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint16_t rd16(const uint8_t *p) { uint16_t v; memcpy(&v, p, 2); return v; }

uint64_t vo_metrohash64_1(const void *key, uint64_t len, uint32_t seed)
{
    static const uint64_t k0 = 0xC83A91E1ull, k1 = 0x8648DBDBull, k2 = 0x7BDEC03Bull, k3 = 0x2F5870A5ull;
    const uint8_t *ptr = (const uint8_t *)key;
    const uint8_t *const end = ptr + len;
    uint64_t hash = (((uint64_t)seed + k2) * k0) + len;

    if (len >= 32) {
        uint64_t v[4] = {hash, hash, hash, hash};
        do {
            v[0] += rd64(ptr) * k0; ptr += 8; v[0] = rotr64(v[0], 29) + v[2];
            v[1] += rd64(ptr) * k1; ptr += 8; v[1] = rotr64(v[1], 29) + v[3];
            v[2] += rd64(ptr) * k2; ptr += 8; v[2] = rotr64(v[2], 29) + v[0];
            v[3] += rd64(ptr) * k3; ptr += 8; v[3] = rotr64(v[3], 29) + v[1];
        } while (ptr <= (end - 32));
        v[2] ^= rotr64(((v[0] + v[3]) * k0) + v[1], 33) * k1;
        v[3] ^= rotr64(((v[1] + v[2]) * k1) + v[0], 33) * k0;
        v[0] ^= rotr64(((v[0] + v[2]) * k0) + v[3], 33) * k1;
        v[1] ^= rotr64(((v[1] + v[3]) * k1) + v[2], 33) * k0;
        hash += v[0] ^ v[1];
    }
    if ((end - ptr) >= 16) {
        uint64_t v0 = hash + (rd64(ptr) * k0); ptr += 8; v0 = rotr64(v0, 33) * k1;
        uint64_t v1 = hash + (rd64(ptr) * k1); ptr += 8; v1 = rotr64(v1, 33) * k2;
        v0 ^= rotr64(v0 * k0, 35) + v1;
        v1 ^= rotr64(v1 * k3, 35) + v0;
        hash += v1;
    }
    if ((end - ptr) >= 8) { hash += rd64(ptr) * k3; ptr += 8; hash ^= rotr64(hash, 33) * k1; }
    if ((end - ptr) >= 4) { hash += (uint64_t)rd32(ptr) * k3; ptr += 4; hash ^= rotr64(hash, 15) * k1; }
    if ((end - ptr) >= 2) { hash += (uint64_t)rd16(ptr) * k3; ptr += 2; hash ^= rotr64(hash, 13) * k1; }
    if ((end - ptr) >= 1) { hash += (uint64_t)ptr[0] * k3; hash ^= rotr64(hash, 25) * k1; }
    hash ^= rotr64(hash, 33);
    hash *= k0;
    hash ^= rotr64(hash, 33);
    return hash;
}

/* Rust `impl Hash for u64` → Hasher::write_u64 → 8 native-endian (LE) bytes;
 * `impl Hash for i32/u32` → 4 LE bytes.  key_width is 8 or 4.                */
uint64_t vo_hash_key(uint64_t key, uint32_t key_width)
{
    if (key_width == 4) { uint32_t k = (uint32_t)key; return vo_metrohash64_1(&k, 4, 0); }
    return vo_metrohash64_1(&key, 8, 0);
}

/* src/partitioner.rs:54-57: `hash(key) as usize % self.partitions` */
uint32_t vo_get_partition(uint64_t key, uint32_t key_width, uint32_t n_reduce)
{
    return (uint32_t)(vo_hash_key(key, key_width) % (uint64_t)n_reduce);
}

/* ------------------------------------------------------------------------- */
/* ParallelCollection::slice — src/rdd/parallel_collection_rdd.rs:116-145.    */
/* Writes slice start offsets into starts[0..n_slices] (starts[n_slices]=n)   */
/* and returns n_slices.  `starts` must hold min(n, num_slices)+2 entries.    */
/* The loop is restated literally, including the quirk that n < num_slices    */
/* yields an empty leading slice followed by singletons (n+1 slices).         */
/* ------------------------------------------------------------------------- */
uint64_t vo_slice(uint64_t n, uint64_t num_slices, uint64_t *starts)
{
    uint64_t slice_count = 0, iter_count = 0, n_out = 0;
    uint64_t end = ((slice_count + 1) * n) / num_slices;
    uint64_t tmp_start = 0;                      /* first element of `tmp` */
    for (uint64_t i = 0; i < n; i++) {
        if (iter_count < end) {
            iter_count++;
        } else {
            slice_count++;
            end = ((slice_count + 1) * n) / num_slices;
            starts[n_out++] = tmp_start;         /* output.push(tmp.drain(..)) */
            tmp_start = i;
            iter_count++;
        }
    }
    starts[n_out++] = tmp_start;                 /* final output.push */
    starts[n_out] = n;
    return n_out;
}

/* ------------------------------------------------------------------------- */
/* Insertion-ordered hash map u64 -> entry index.  The reference uses          */
/* std::collections::HashMap (SipHash, per-process random iteration order);    */
/* the oracle iterates in first-insertion order so its output is deterministic */
/* (results are compared as sorted multisets / ordered value lists).           */
/* ------------------------------------------------------------------------- */
typedef struct {
    uint64_t *keys;     /* entries, insertion order */
    uint64_t *acc;      /* reduce ops: combiner; group ops: count */
    uint64_t *head;     /* group ops: first node (row index) or UINT64_MAX */
    uint64_t *tail;
    uint64_t n, cap;
    uint64_t *idx;      /* open addressing: entry index + 1, 0 = empty */
    uint64_t idx_mask;
    int want_lists;
} omap;

static inline uint64_t mix64(uint64_t x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}

static void omap_init(omap *m, uint64_t hint, int want_lists)
{
    memset(m, 0, sizeof(*m));
    m->want_lists = want_lists;
    m->cap = hint < 16 ? 16 : hint;
    m->keys = (uint64_t *)malloc(m->cap * 8);
    m->acc = (uint64_t *)malloc(m->cap * 8);
    if (want_lists) { m->head = (uint64_t *)malloc(m->cap * 8); m->tail = (uint64_t *)malloc(m->cap * 8); }
    uint64_t ic = 32; while (ic < m->cap * 2) ic <<= 1;
    m->idx = (uint64_t *)calloc(ic, 8);
    m->idx_mask = ic - 1;
}

static void omap_free(omap *m)
{
    free(m->keys); free(m->acc); free(m->head); free(m->tail); free(m->idx);
    memset(m, 0, sizeof(*m));
}

static void omap_grow(omap *m)
{
    m->cap *= 2;
    m->keys = (uint64_t *)realloc(m->keys, m->cap * 8);
    m->acc = (uint64_t *)realloc(m->acc, m->cap * 8);
    if (m->want_lists) { m->head = (uint64_t *)realloc(m->head, m->cap * 8); m->tail = (uint64_t *)realloc(m->tail, m->cap * 8); }
    if (m->cap * 2 > m->idx_mask + 1) {
        uint64_t ic = (m->idx_mask + 1) * 2;
        free(m->idx);
        m->idx = (uint64_t *)calloc(ic, 8);
        m->idx_mask = ic - 1;
        for (uint64_t e = 0; e < m->n; e++) {
            uint64_t s = mix64(m->keys[e]) & m->idx_mask;
            while (m->idx[s]) s = (s + 1) & m->idx_mask;
            m->idx[s] = e + 1;
        }
    }
}

/* returns entry index; *is_new set when the key was inserted by this call */
static inline uint64_t omap_find_or_insert(omap *m, uint64_t key, int *is_new)
{
    uint64_t s = mix64(key) & m->idx_mask;
    for (;;) {
        uint64_t e = m->idx[s];
        if (!e) break;
        if (m->keys[e - 1] == key) { *is_new = 0; return e - 1; }
        s = (s + 1) & m->idx_mask;
    }
    if (m->n == m->cap) {
        omap_grow(m);
        s = mix64(key) & m->idx_mask;
        while (m->idx[s]) s = (s + 1) & m->idx_mask;
    }
    uint64_t e = m->n++;
    m->keys[e] = key;
    m->idx[s] = e + 1;
    *is_new = 1;
    return e;
}

static inline int64_t omap_find(const omap *m, uint64_t key)
{
    uint64_t s = mix64(key) & m->idx_mask;
    for (;;) {
        uint64_t e = m->idx[s];
        if (!e) return -1;
        if (m->keys[e - 1] == key) return (int64_t)(e - 1);
        s = (s + 1) & m->idx_mask;
    }
}

/* ------------------------------------------------------------------------- */
/* Aggregators.  src/aggregator.rs:33-52 (default = Vec append, group_by_key), */
/* src/rdd/pair_rdd.rs:74-78 (reduce_by_key: create = id, merge_* = f).        */
/* Named ops replace the serde_closure `f`: SUM/MIN/MAX over u64/i64/f64 and   */
/* COUNT (= rdd.rs:450-459 count_by_value: map(x -> (x,1u64)).reduce_by_key(+)).*/
/* ------------------------------------------------------------------------- */
enum { VO_U64 = 0, VO_I64 = 1, VO_F64 = 2 };
enum { VO_GROUP = 0, VO_SUM = 1, VO_MIN = 2, VO_MAX = 3, VO_COUNT = 4 };

static inline uint64_t apply_op(int op, int vdt, uint64_t a, uint64_t b)
{
    if (op == VO_SUM || op == VO_COUNT) {
        if (vdt == VO_F64 && op == VO_SUM) {
            double x, y; memcpy(&x, &a, 8); memcpy(&y, &b, 8); x = x + y; memcpy(&a, &x, 8); return a;
        }
        return a + b;       /* u64/i64: wrapping add (release-build semantics) */
    }
    int take_b;
    if (vdt == VO_U64) take_b = (op == VO_MIN) ? (b < a) : (b > a);
    else if (vdt == VO_I64) take_b = (op == VO_MIN) ? ((int64_t)b < (int64_t)a) : ((int64_t)b > (int64_t)a);
    else { double x, y; memcpy(&x, &a, 8); memcpy(&y, &b, 8); take_b = (op == VO_MIN) ? (y < x) : (y > x); }
    return take_b ? b : a;
}

/* ------------------------------------------------------------------------- */
/* A whole shuffle job: M map tasks then R reduce tasks.                       */
/* ------------------------------------------------------------------------- */
typedef struct {
    /* config */
    int op, vdt;
    uint32_t key_width;
    uint64_t n, M, R;
    const uint64_t *keys, *vals;     /* SoA input, n rows */
    uint64_t *starts;                /* M+1 slice starts (vo_slice) */
    /* map outputs: buckets[m*R + r] (== SHUFFLE_CACHE[(sid, m, r)], src/env.rs:27) */
    omap *buckets;
    uint64_t **next;                 /* group: per map split linked list over rows */
    /* reduce outputs, per r */
    uint64_t *out_nkeys, *out_nvals;
    uint64_t **out_keys, **out_comb, **out_offs, **out_vals;
} vo_job;

/* ShuffleDependency::do_shuffle_task — src/dependency.rs:164-229.
 * For one map split: R buckets; per row bucket_id = get_partition(k) (:201);
 * present → merge_value(old, v) (:203-206), else create_combiner(v) (:208).   */
static void map_task(vo_job *j, uint64_t m)
{
    uint64_t lo = j->starts[m], hi = j->starts[m + 1], R = j->R;
    omap *b = &j->buckets[m * R];
    uint64_t hint = (hi - lo) / R / 4 + 16;
    for (uint64_t r = 0; r < R; r++) omap_init(&b[r], hint, j->op == VO_GROUP);
    uint64_t *next = NULL;
    if (j->op == VO_GROUP) { next = (uint64_t *)malloc((hi - lo + 1) * 8); j->next[m] = next; }
    for (uint64_t i = lo; i < hi; i++) {
        uint64_t k = j->keys[i];
        uint64_t v = (j->op == VO_COUNT) ? 1ull : j->vals[i];
        uint32_t bid = vo_get_partition(k, j->key_width, (uint32_t)R);
        omap *bk = &b[bid];
        int is_new;
        uint64_t e = omap_find_or_insert(bk, k, &is_new);
        if (j->op == VO_GROUP) {
            uint64_t node = i - lo;
            next[node] = UINT64_MAX;
            if (is_new) { bk->head[e] = node; bk->tail[e] = node; bk->acc[e] = 1; }   /* vec![v] */
            else { next[bk->tail[e]] = node; bk->tail[e] = node; bk->acc[e]++; }        /* buf.push(v) */
        } else {
            if (is_new) bk->acc[e] = v;                                   /* create_combiner = id */
            else bk->acc[e] = apply_op(j->op, j->vdt, bk->acc[e], v);     /* merge_value = f */
        }
    }
}

/* ShuffledRdd::compute — src/rdd/shuffled_rdd.rs:149-170, fed by ShuffleFetcher::fetch
 * (src/shuffle/shuffle_fetcher.rs:34-39,63-97: local mode has one server URI, so chunks
 * are concatenated in map-id order).  merge_combiners for reduce ops = f; for the default
 * aggregator = b1.append(b2) (src/aggregator.rs:41-45).                                   */
static void reduce_task(vo_job *j, uint64_t r)
{
    uint64_t M = j->M, R = j->R;
    omap comb;
    uint64_t hint = 16;
    for (uint64_t m = 0; m < M; m++) if (j->buckets[m * R + r].n > hint) hint = j->buckets[m * R + r].n;
    omap_init(&comb, hint * 2, 0);
    for (uint64_t m = 0; m < M; m++) {
        omap *bk = &j->buckets[m * R + r];
        for (uint64_t e = 0; e < bk->n; e++) {
            int is_new;
            uint64_t c = omap_find_or_insert(&comb, bk->keys[e], &is_new);
            if (j->op == VO_GROUP) comb.acc[c] = (is_new ? 0 : comb.acc[c]) + bk->acc[e];
            else comb.acc[c] = is_new ? bk->acc[e] : apply_op(j->op, j->vdt, comb.acc[c], bk->acc[e]);
        }
    }
    uint64_t nk = comb.n;
    j->out_nkeys[r] = nk;
    j->out_keys[r] = (uint64_t *)malloc((nk + 1) * 8);
    memcpy(j->out_keys[r], comb.keys, nk * 8);
    if (j->op != VO_GROUP) {
        j->out_comb[r] = (uint64_t *)malloc((nk + 1) * 8);
        memcpy(j->out_comb[r], comb.acc, nk * 8);
        j->out_nvals[r] = 0;
    } else {
        uint64_t *offs = (uint64_t *)malloc((nk + 1) * 8);
        uint64_t tot = 0;
        for (uint64_t c = 0; c < nk; c++) { offs[c] = tot; tot += comb.acc[c]; }
        offs[nk] = tot;
        uint64_t *vals = (uint64_t *)malloc((tot + 1) * 8);
        uint64_t *cur = (uint64_t *)malloc((nk + 1) * 8);
        memcpy(cur, offs, (nk + 1) * 8);
        for (uint64_t m = 0; m < M; m++) {            /* map-id order ⇒ input order (F5) */
            omap *bk = &j->buckets[m * R + r];
            const uint64_t *next = j->next[m];
            uint64_t lo = j->starts[m];
            for (uint64_t e = 0; e < bk->n; e++) {
                int64_t c = omap_find(&comb, bk->keys[e]);
                for (uint64_t node = bk->head[e]; node != UINT64_MAX; node = next[node])
                    vals[cur[c]++] = j->vals[lo + node];
            }
        }
        free(cur);
        j->out_offs[r] = offs; j->out_vals[r] = vals; j->out_nvals[r] = tot;
    }
    omap_free(&comb);
}

typedef struct { vo_job *j; uint64_t tid, nthreads; int phase; } worker_arg;

static void *worker(void *p)
{
    worker_arg *a = (worker_arg *)p;
    vo_job *j = a->j;
    uint64_t cnt = a->phase == 0 ? j->M : j->R;
    for (uint64_t t = a->tid; t < cnt; t += a->nthreads) {
        if (a->phase == 0) map_task(j, t); else reduce_task(j, t);
    }
    return NULL;
}

static void run_phase(vo_job *j, int phase, uint64_t threads)
{
    uint64_t cnt = phase == 0 ? j->M : j->R;
    if (threads > cnt) threads = cnt;
    if (threads <= 1) { worker_arg a = {j, 0, 1, phase}; worker(&a); return; }
    pthread_t *th = (pthread_t *)malloc(threads * sizeof(pthread_t));
    worker_arg *args = (worker_arg *)malloc(threads * sizeof(worker_arg));
    for (uint64_t t = 0; t < threads; t++) {
        args[t].j = j; args[t].tid = t; args[t].nthreads = threads; args[t].phase = phase;
        pthread_create(&th[t], NULL, worker, &args[t]);
    }
    for (uint64_t t = 0; t < threads; t++) pthread_join(th[t], NULL);
    free(th); free(args);
}

/* combine_by_key end to end (src/rdd/pair_rdd.rs:20-80): slice the input into
 * `num_slices` map partitions (parallel_collection_rdd.rs:116-145), run the map stage
 * then the reduce stage (scheduler order, base_scheduler.rs:377-455), one thread per
 * task up to `threads`.  Returns an opaque job handle or NULL.                        */
vo_job *vo_shuffle_run(int op, int vdt, uint32_t key_width, const uint64_t *keys, const uint64_t *vals,
                       uint64_t n, uint64_t num_slices, uint64_t n_reduce, uint64_t threads)
{
    if (num_slices < 1 || n_reduce < 1) return NULL;
    vo_job *j = (vo_job *)calloc(1, sizeof(vo_job));
    j->op = op; j->vdt = vdt; j->key_width = key_width ? key_width : 8;
    j->n = n; j->R = n_reduce; j->keys = keys; j->vals = vals;
    uint64_t cap = (n < num_slices ? n : num_slices) + 2;
    j->starts = (uint64_t *)malloc(cap * 8);
    j->M = vo_slice(n, num_slices, j->starts);
    j->buckets = (omap *)calloc(j->M * j->R, sizeof(omap));
    j->next = (uint64_t **)calloc(j->M, sizeof(uint64_t *));
    j->out_nkeys = (uint64_t *)calloc(j->R, 8);
    j->out_nvals = (uint64_t *)calloc(j->R, 8);
    j->out_keys = (uint64_t **)calloc(j->R, sizeof(void *));
    j->out_comb = (uint64_t **)calloc(j->R, sizeof(void *));
    j->out_offs = (uint64_t **)calloc(j->R, sizeof(void *));
    j->out_vals = (uint64_t **)calloc(j->R, sizeof(void *));
    run_phase(j, 0, threads);
    run_phase(j, 1, threads);
    return j;
}

uint64_t vo_job_n_map(const vo_job *j) { return j->M; }
uint64_t vo_job_n_reduce(const vo_job *j) { return j->R; }
/* number of rows map task m put into bucket r (after map-side combine) */
uint64_t vo_job_bucket_rows(const vo_job *j, uint64_t m, uint64_t r) { return j->buckets[m * j->R + r].n; }
void vo_job_part_sizes(const vo_job *j, uint64_t r, uint64_t *n_keys, uint64_t *n_vals)
{
    *n_keys = j->out_nkeys[r]; *n_vals = j->out_nvals[r];
}
void vo_job_part_copy(const vo_job *j, uint64_t r, uint64_t *keys, uint64_t *comb, uint64_t *offs, uint64_t *vals)
{
    uint64_t nk = j->out_nkeys[r];
    if (keys) memcpy(keys, j->out_keys[r], nk * 8);
    if (comb && j->out_comb[r]) memcpy(comb, j->out_comb[r], nk * 8);
    if (offs && j->out_offs[r]) memcpy(offs, j->out_offs[r], (nk + 1) * 8);
    if (vals && j->out_vals[r]) memcpy(vals, j->out_vals[r], j->out_nvals[r] * 8);
}

void vo_job_free(vo_job *j)
{
    if (!j) return;
    for (uint64_t i = 0; i < j->M * j->R; i++) omap_free(&j->buckets[i]);
    for (uint64_t m = 0; m < j->M; m++) free(j->next[m]);
    for (uint64_t r = 0; r < j->R; r++) { free(j->out_keys[r]); free(j->out_comb[r]); free(j->out_offs[r]); free(j->out_vals[r]); }
    free(j->buckets); free(j->next); free(j->starts);
    free(j->out_nkeys); free(j->out_nvals); free(j->out_keys); free(j->out_comb); free(j->out_offs); free(j->out_vals);
    free(j);
}

/* ------------------------------------------------------------------------- */
/* join — src/rdd/pair_rdd.rs:104-121 over cogroup (:123-155) over              */
/* CoGroupedRdd::compute (src/rdd/co_grouped_rdd.rs:206-249).  Each side is     */
/* shuffled with the Vec-append aggregator (co_grouped_rdd.rs:78-124); compute  */
/* extends agg[k][dep_num] per dep in order; the join emits, per key,           */
/* `for v in vs { for w in ws }` (pair_rdd.rs:109-115) — inner join.            */
/* ------------------------------------------------------------------------- */
typedef struct {
    uint64_t R;
    uint64_t *n_out;
    uint64_t **k, **v, **w;
} vo_join_res;

vo_join_res *vo_join_run(uint32_t key_width,
                         const uint64_t *ka, const uint64_t *va, uint64_t na, uint64_t slices_a,
                         const uint64_t *kb, const uint64_t *vb, uint64_t nb, uint64_t slices_b,
                         uint64_t n_reduce, uint64_t threads)
{
    vo_job *A = vo_shuffle_run(VO_GROUP, VO_U64, key_width, ka, va, na, slices_a, n_reduce, threads);
    vo_job *B = vo_shuffle_run(VO_GROUP, VO_U64, key_width, kb, vb, nb, slices_b, n_reduce, threads);
    if (!A || !B) { vo_job_free(A); vo_job_free(B); return NULL; }
    vo_join_res *res = (vo_join_res *)calloc(1, sizeof(*res));
    res->R = n_reduce;
    res->n_out = (uint64_t *)calloc(n_reduce, 8);
    res->k = (uint64_t **)calloc(n_reduce, sizeof(void *));
    res->v = (uint64_t **)calloc(n_reduce, sizeof(void *));
    res->w = (uint64_t **)calloc(n_reduce, sizeof(void *));
    for (uint64_t r = 0; r < n_reduce; r++) {
        /* agg: keys of dep 0 first (insertion order), then unseen keys of dep 1 */
        omap idxB; omap_init(&idxB, B->out_nkeys[r] * 2 + 16, 0);
        for (uint64_t e = 0; e < B->out_nkeys[r]; e++) { int nw; uint64_t c = omap_find_or_insert(&idxB, B->out_keys[r][e], &nw); idxB.acc[c] = e; }
        uint64_t tot = 0;
        for (uint64_t e = 0; e < A->out_nkeys[r]; e++) {
            int64_t c = omap_find(&idxB, A->out_keys[r][e]);
            if (c < 0) continue;
            uint64_t eb = idxB.acc[c];
            tot += (A->out_offs[r][e + 1] - A->out_offs[r][e]) * (B->out_offs[r][eb + 1] - B->out_offs[r][eb]);
        }
        res->n_out[r] = tot;
        res->k[r] = (uint64_t *)malloc((tot + 1) * 8);
        res->v[r] = (uint64_t *)malloc((tot + 1) * 8);
        res->w[r] = (uint64_t *)malloc((tot + 1) * 8);
        uint64_t o = 0;
        for (uint64_t e = 0; e < A->out_nkeys[r]; e++) {
            int64_t c = omap_find(&idxB, A->out_keys[r][e]);
            if (c < 0) continue;                       /* empty side ⇒ no rows (inner join) */
            uint64_t eb = idxB.acc[c];
            for (uint64_t x = A->out_offs[r][e]; x < A->out_offs[r][e + 1]; x++)
                for (uint64_t y = B->out_offs[r][eb]; y < B->out_offs[r][eb + 1]; y++) {
                    res->k[r][o] = A->out_keys[r][e]; res->v[r][o] = A->out_vals[r][x]; res->w[r][o] = B->out_vals[r][y]; o++;
                }
        }
        omap_free(&idxB);
    }
    vo_job_free(A); vo_job_free(B);
    return res;
}

uint64_t vo_join_part_size(const vo_join_res *j, uint64_t r) { return j->n_out[r]; }
void vo_join_part_copy(const vo_join_res *j, uint64_t r, uint64_t *k, uint64_t *v, uint64_t *w)
{
    memcpy(k, j->k[r], j->n_out[r] * 8); memcpy(v, j->v[r], j->n_out[r] * 8); memcpy(w, j->w[r], j->n_out[r] * 8);
}
void vo_join_free(vo_join_res *j)
{
    if (!j) return;
    for (uint64_t r = 0; r < j->R; r++) { free(j->k[r]); free(j->v[r]); free(j->w[r]); }
    free(j->n_out); free(j->k); free(j->v); free(j->w); free(j);
}

/* ------------------------------------------------------------------------- */
/* sort_by_key — NOT in the reference: parity unpinned.  Oracle =  */
/* stable sort by key of the whole input; output partition i holds the i-th    */
/* contiguous key range.  Partition boundaries: cut points floor(i*n/R) moved   */
/* forward past any run of equal keys (a range partitioner never splits a key). */
/* ------------------------------------------------------------------------- */
static int key_less(int kdt, uint64_t a, uint64_t b)
{
    if (kdt == VO_I64) return (int64_t)a < (int64_t)b;
    if (kdt == VO_F64) { double x, y; memcpy(&x, &a, 8); memcpy(&y, &b, 8); return x < y; }
    return a < b;
}

static void msort(int kdt, uint64_t *k, uint64_t *v, uint64_t *tk, uint64_t *tv, uint64_t n)
{
    if (n < 2) return;
    uint64_t h = n / 2;
    msort(kdt, k, v, tk, tv, h);
    msort(kdt, k + h, v ? v + h : NULL, tk, tv, n - h);
    uint64_t i = 0, j = h, o = 0;
    while (i < h && j < n) {
        if (key_less(kdt, k[j], k[i])) { tk[o] = k[j]; if (v) tv[o] = v[j]; j++; }
        else { tk[o] = k[i]; if (v) tv[o] = v[i]; i++; }
        o++;
    }
    while (i < h) { tk[o] = k[i]; if (v) tv[o] = v[i]; i++; o++; }
    while (j < n) { tk[o] = k[j]; if (v) tv[o] = v[j]; j++; o++; }
    memcpy(k, tk, n * 8);
    if (v) memcpy(v, tv, n * 8);
}

/* out_keys/out_vals: n entries (sorted); part_starts: n_parts+1 entries */
void vo_sort_by_key(int kdt, const uint64_t *keys, const uint64_t *vals, uint64_t n, uint64_t n_parts,
                    uint64_t *out_keys, uint64_t *out_vals, uint64_t *part_starts)
{
    memcpy(out_keys, keys, n * 8);
    if (vals && out_vals) memcpy(out_vals, vals, n * 8);
    uint64_t *tk = (uint64_t *)malloc((n + 1) * 8), *tv = (uint64_t *)malloc((n + 1) * 8);
    msort(kdt, out_keys, (vals && out_vals) ? out_vals : NULL, tk, tv, n);
    free(tk); free(tv);
    part_starts[0] = 0;
    for (uint64_t p = 1; p < n_parts; p++) {
        uint64_t c = (p * n) / n_parts;
        if (c < part_starts[p - 1]) c = part_starts[p - 1];
        while (c > 0 && c < n && out_keys[c] == out_keys[c - 1]) c++;
        part_starts[p] = c;
    }
    part_starts[n_parts] = n;
}

/* ------------------------------------------------------------------------- */
/* Deterministic generator shared with the CUDA side.            */
/* ------------------------------------------------------------------------- */
static inline uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
uint64_t vo_splitmix64(uint64_t x) { return splitmix64(x); }

/* key_i = splitmix64(rank_i ^ 0xA5A5..), rank_i = splitmix64(seed_k + i) % D;
 * val_i = splitmix64(seed_v + i) & 0xFFFFF.  i = first .. first+n-1.              */
void vo_gen_uniform(uint64_t *keys, uint64_t *vals, uint64_t first, uint64_t n, uint64_t D, uint64_t seed_k, uint64_t seed_v)
{
    for (uint64_t t = 0; t < n; t++) {
        uint64_t i = first + t;
        uint64_t rank = splitmix64(seed_k + i) % D;
        keys[t] = splitmix64(rank ^ 0xA5A5A5A5A5A5A5A5ull);
        vals[t] = splitmix64(seed_v + i) & 0xFFFFFull;
    }
}
