// micro.cu — microbenchmarks that guided the design of hash_agg_kernel.
// Standalone: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o micro micro.cu ; ./micro [rows] [distinct]
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../vega_b200/csrc/kernels.cuh"
using namespace vb;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

enum { V_STREAM = 0, V_HASH, V_PROBE, V_RED, V_FULL, V_RED8, V_ATOM, V_PROBE_RED_NOCAS, V_N };
static const char *names[] = {"stream_read", "hash_only", "probe_only", "red_only", "full(hash_agg)", "red_only_8B_slots", "atom_ret_only", "probe+red(no cas)"};

template <int V, int ROWS>
__global__ void __launch_bounds__(256) micro_kernel(const u64 *__restrict__ rows, u64 n, Slot *tab, u64 *tab8, u32 log_cap, u64 *sink)
{
    const u64 pol = policy_evict_first();
    const u64 mask = (1ull << log_cap) - 1;
    const u32 shift = 64 - log_cap;
    constexpr int TILE = 256 * ROWS;
    const u64 n_tiles = (n + TILE - 1) / TILE;
    u64 acc = 0;
    for (u64 tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        u64 k[ROWS], v[ROWS];
#pragma unroll
        for (int j = 0; j < ROWS; ++j) {
            u64 idx = tile * TILE + (u64)j * 256 + threadIdx.x;
            k[j] = 0; v[j] = 0;
            if (idx < n) { ulonglong2 r = ld_stream_u64x2(rows + 2 * idx, pol); k[j] = r.x; v[j] = r.y; }
        }
        if (V == V_PROBE || V == V_PROBE_RED_NOCAS) {     // all probes issued before use (MLP)
            u64 s[ROWS], kk[ROWS];
#pragma unroll
            for (int j = 0; j < ROWS; ++j) { s[j] = slot_hash(k[j]) >> shift; kk[j] = ld_cg_u64(&tab[s[j]].key); }
#pragma unroll
            for (int j = 0; j < ROWS; ++j) {
                u32 guard = 0;
                while (kk[j] != k[j] && guard++ < 64) { s[j] = (s[j] + 1) & mask; kk[j] = ld_cg_u64(&tab[s[j]].key); }
                if (V == V_PROBE) acc += s[j];
                else atomicAdd((unsigned long long *)&tab[s[j]].acc, (unsigned long long)v[j]);
            }
            continue;
        }
#pragma unroll
        for (int j = 0; j < ROWS; ++j) {
            if (V == V_STREAM) acc += k[j] ^ v[j];
            else if (V == V_HASH) acc += slot_hash(k[j]) >> shift;
            else if (V == V_RED) atomicAdd((unsigned long long *)&tab[slot_hash(k[j]) >> shift].acc, (unsigned long long)v[j]);
            else if (V == V_RED8) atomicAdd((unsigned long long *)&tab8[slot_hash(k[j]) >> shift], (unsigned long long)v[j]);
            else if (V == V_ATOM) acc += atomicAdd((unsigned long long *)&tab[slot_hash(k[j]) >> shift].acc, (unsigned long long)v[j]);
            else if (V == V_FULL) { u32 slot, ins = 0; table_upsert<OPK_ADD_U64>(tab, mask, shift, k[j], v[j], slot, ins); acc += ins; }
        }
    }
    if (acc == 0x123456789ull) *sink = acc;
}

template <int V, int ROWS>
static float run(const u64 *rows, u64 n, Slot *tab, u64 *tab8, u32 log_cap, u64 *sink, int grid, int reps)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    micro_kernel<V, ROWS><<<grid, 256>>>(rows, n, tab, tab8, log_cap, sink);   // warm-up
    CK(cudaDeviceSynchronize());
    cudaEventRecord(a);
    for (int r = 0; r < reps; ++r) micro_kernel<V, ROWS><<<grid, 256>>>(rows, n, tab, tab8, log_cap, sink);
    cudaEventRecord(b);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

int main(int argc, char **argv)
{
    u64 n = argc > 1 ? (u64)atof(argv[1]) : 250000000ull;
    u64 D = argc > 2 ? (u64)atof(argv[2]) : 1000000ull;
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s, %d SMs; rows %.3g distinct %.3g\n", p.name, p.multiProcessorCount, (double)n, (double)D);
    u64 *rows; CK(cudaMalloc(&rows, n * 16));
    gen_pairs_kernel<<<p.multiProcessorCount * 16, 256>>>(rows, nullptr, nullptr, 0, n, GEN_UNIFORM, D, 0, 1, 2, nullptr);
    CK(cudaDeviceSynchronize());
    u64 *sink; CK(cudaMalloc(&sink, 8));
    for (u32 log_cap : {21u, 22u, 24u}) {
        const u64 cap = 1ull << log_cap;
        Slot *tab; CK(cudaMalloc(&tab, (cap + 1) * 16));
        u64 *tab8; CK(cudaMalloc(&tab8, cap * 8)); CK(cudaMemset(tab8, 0, cap * 8));
        TableCtl *ctl; CK(cudaMalloc(&ctl, sizeof(TableCtl))); CK(cudaMemset(ctl, 0, sizeof(TableCtl)));
        table_init_kernel<<<1024, 256>>>(tab, cap, 0);
        // pre-fill the table with every key (so V_PROBE finds them and V_FULL never CASes after warm-up)
        hash_agg_kernel<IN_AOS, OPK_ADD_U64, TX_NONE><<<p.multiProcessorCount * 8, 256>>>(rows, nullptr, n, tab, log_cap, ctl, ~0ull, nullptr);
        CK(cudaDeviceSynchronize());
        printf("--- table 2^%u slots (%.0f MB of 16-B slots), load %.2f\n", log_cap, cap * 16.0 / 1e6, (double)D / cap);
        for (int occ : {2, 4, 8}) {
            int grid = p.multiProcessorCount * occ;
            float ms[V_N];
            ms[V_STREAM] = run<V_STREAM, 4>(rows, n, tab, tab8, log_cap, sink, grid, 3);
            ms[V_HASH] = run<V_HASH, 4>(rows, n, tab, tab8, log_cap, sink, grid, 3);
            ms[V_PROBE] = run<V_PROBE, 4>(rows, n, tab, tab8, log_cap, sink, grid, 3);
            ms[V_RED] = run<V_RED, 4>(rows, n, tab, tab8, log_cap, sink, grid, 3);
            ms[V_FULL] = run<V_FULL, 4>(rows, n, tab, tab8, log_cap, sink, grid, 3);
            ms[V_RED8] = run<V_RED8, 4>(rows, n, tab, tab8, log_cap, sink, grid, 3);
            ms[V_ATOM] = run<V_ATOM, 4>(rows, n, tab, tab8, log_cap, sink, grid, 3);
            ms[V_PROBE_RED_NOCAS] = run<V_PROBE_RED_NOCAS, 4>(rows, n, tab, tab8, log_cap, sink, grid, 3);
            for (int v = 0; v < V_N; ++v)
                printf("  ctas/SM %d  %-22s %8.3f ms  %7.1f Grows/s  %7.1f GB/s(16B/row)\n", occ, names[v], ms[v], n / ms[v] / 1e6, n * 16.0 / ms[v] / 1e6);
        }
        // rows per thread = 8 variants at 8 CTAs/SM
        {
            int grid = p.multiProcessorCount * 8;
            printf("  rows/thread 8: stream %.3f ms, probe %.3f ms, red %.3f ms, probe+red %.3f ms\n",
                   run<V_STREAM, 8>(rows, n, tab, tab8, log_cap, sink, grid, 3), run<V_PROBE, 8>(rows, n, tab, tab8, log_cap, sink, grid, 3),
                   run<V_RED, 8>(rows, n, tab, tab8, log_cap, sink, grid, 3), run<V_PROBE_RED_NOCAS, 8>(rows, n, tab, tab8, log_cap, sink, grid, 3));
        }
        cudaFree(tab); cudaFree(tab8); cudaFree(ctl);
    }
    return 0;
}
