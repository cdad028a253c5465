// micro_r2.cu — round-2 micro-benchmarks behind DESIGN.md §3: where is the roof of an exact streaming
// aggregation into an L2-resident table?
//   1. request-rate ceilings with no input stream at all: random 32-byte loads (the probe), 64-/32-bit REDs,
//      returning 32-bit atomics, probe+RED pairs — each lane one L2 request, addresses from a counter hash;
//   2. the production kernels (register-staged hash_agg_kernel vs the bulk-staged hash_agg_bulk_kernel
//      compiled with this binary's -DVB_HB_ROWS/-DVB_HB_CTAS/-DVB_HB_STAGES) on a warm table.
//   nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -o micro_r2 micro_r2.cu
//   ./micro_r2 [rows] [distinct] [only_kernels(0/1)]
// Run under `ncu --metrics l1tex__m_l1tex2xbar_req_cycles_active...` to see which unit each one saturates.
#include <cstdio>
#include <cstdlib>
#include "../vega_b200/csrc/kernels.cuh"
using namespace vb;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

enum { C_LOAD32 = 0, C_RED64 = 1, C_RED32 = 2, C_ATOM32 = 3, C_LOAD_RED64 = 4, C_LOAD_RED64_SAME = 5, C_LOAD8 = 6 };

// n "rows" without any input: row i probes / updates slot hash(i).  ILP 4 like the production kernel.
template <int V>
__global__ void __launch_bounds__(256) ceiling_kernel(u64 n, Table t, u64 *sink)
{
    const u64 nb_shift = 64 - (t.log_cap - 2);
    u64 acc = 0;
    const u64 stride = (u64)gridDim.x * 256 * 4;
    for (u64 i0 = ((u64)blockIdx.x * 256 + threadIdx.x) * 4; i0 < n; i0 += stride) {
        u64 b[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = slot_hash(splitmix64(i0 + j)) >> nb_shift;
        if (V == C_LOAD32 || V == C_LOAD_RED64 || V == C_LOAD_RED64_SAME) {
            Bucket4 k[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) k[j] = ld_bucket(&t.keys[4 * b[j]]);
#pragma unroll
            for (int j = 0; j < 4; ++j) acc += k[j].k0 ^ k[j].k3;
        }
        if (V == C_LOAD8) {
#pragma unroll
            for (int j = 0; j < 4; ++j) acc += ld_cg_u64(&t.keys[4 * b[j]]);
        }
        if (V == C_RED64 || V == C_LOAD_RED64) {
#pragma unroll
            for (int j = 0; j < 4; ++j) atomicAdd((unsigned long long *)&t.accs[4 * b[j] + (acc & 1)], 1ull + (acc & 2));
        }
        if (V == C_LOAD_RED64_SAME) {   // RED into the sector that was just probed (interleaved key+acc layout)
#pragma unroll
            for (int j = 0; j < 4; ++j) atomicAdd((unsigned long long *)&t.keys[4 * b[j] + 2 + (acc & 1)], 1ull);
        }
        if (V == C_RED32) {
#pragma unroll
            for (int j = 0; j < 4; ++j) atomicAdd((unsigned int *)&t.accs[4 * b[j]], 1u);
        }
        if (V == C_ATOM32) {
#pragma unroll
            for (int j = 0; j < 4; ++j) acc += atomicAdd((unsigned int *)&t.accs[4 * b[j]], 1u);
        }
    }
    if (acc == 0x123456789ull) *sink = acc;
}

template <typename F>
static float timeit(F f, int reps = 3)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); CK(cudaDeviceSynchronize());
    cudaEventRecord(a);
    for (int r = 0; r < reps; ++r) f();
    cudaEventRecord(b);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

int main(int argc, char **argv)
{
    u64 n = argc > 1 ? (u64)atof(argv[1]) : 250000000ull;
    u64 D = argc > 2 ? (u64)atof(argv[2]) : 1000000ull;
    const int only_kernels = argc > 3 ? atoi(argv[3]) : 0;
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    printf("device %s, %d SMs; rows %.3g distinct %.3g; bulk variant rows/thread %d ctas/SM %d stages %d\n", p.name, sms, (double)n, (double)D,
           HB_ROWS, VB_HB_CTAS, HB_STAGES);
    u64 *rows; CK(cudaMalloc(&rows, n * 16));
    gen_pairs_kernel<<<sms * 16, 256>>>(rows, nullptr, nullptr, 0, n, GEN_UNIFORM, D, 0, 1, 2, nullptr);
    u64 *sink; CK(cudaMalloc(&sink, 8));
    TableCtl *ctl; CK(cudaMalloc(&ctl, sizeof(TableCtl)));
    for (u32 log_cap : {21u, 22u}) {
        void *base; CK(cudaMalloc(&base, table_bytes(log_cap)));
        Table t = table_at(base, log_cap);
        printf("--- table 2^%u slots (%zu MB), load %.2f\n", log_cap, table_bytes(log_cap) >> 20, (double)D / (1ull << log_cap));
        auto reset = [&] { table_init_kernel<<<1024, 256>>>(t, 0); cudaMemset(ctl, 0, sizeof(TableCtl)); };
        reset();
        if (!only_kernels) {
            for (int occ : {4, 8}) {
                const int grid = sms * occ;
#define CEIL(V, name, reqs)                                                                              \
    {                                                                                                    \
        float ms = timeit([&] { ceiling_kernel<V><<<grid, 256>>>(n, t, sink); });                        \
        printf("  ceiling %-18s ctas/SM %d: %7.3f ms  %6.1f Grows/s  %6.1f G L2 requests/s\n", name, occ, ms, n / ms / 1e6, reqs * n / ms / 1e6); \
    }
                CEIL(C_LOAD32, "load 32B", 1.0)
                CEIL(C_LOAD8, "load 8B", 1.0)
                CEIL(C_RED64, "red.u64", 1.0)
                CEIL(C_RED32, "red.u32", 1.0)
                CEIL(C_ATOM32, "atom.u32", 1.0)
                CEIL(C_LOAD_RED64, "load32B+red.u64", 2.0)
                CEIL(C_LOAD_RED64_SAME, "same-sector ld+red", 2.0)
            }
            reset();
        }
        // production kernels, warm table (all keys inserted by the first call)
        {
            auto kb = hash_agg_bulk_kernel<IN_AOS, OPK_ADD_U64, TX_NONE>;
            const size_t smem = hb_smem_bytes(true);
            CK(cudaFuncSetAttribute(kb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int occ_b = 0, occ_a = 0;
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_b, kb, HB_THREADS, smem));
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_a, hash_agg_kernel<IN_AOS, OPK_ADD_U64, TX_NONE>, HA_THREADS, 0));
            const int ga = sms * occ_a, gb = sms * occ_b;
            float a_ms = timeit([&] { hash_agg_kernel<IN_AOS, OPK_ADD_U64, TX_NONE><<<ga, HA_THREADS>>>(rows, nullptr, n, t, ctl, ~0ull, nullptr); });
            float b_ms = timeit([&] { kb<<<gb, HB_THREADS, smem>>>(rows, nullptr, n, t, ctl, ~0ull, nullptr); });
            printf("  hash_agg_kernel      (ctas/SM %d): %7.3f ms  %6.1f Grows/s  %6.0f GB/s\n", occ_a, a_ms, n / a_ms / 1e6, n * 16.0 / a_ms / 1e6);
            printf("  hash_agg_bulk_kernel (ctas/SM %d): %7.3f ms  %6.1f Grows/s  %6.0f GB/s\n", occ_b, b_ms, n / b_ms / 1e6, n * 16.0 / b_ms / 1e6);
        }
        cudaFree(base);
    }
    return 0;
}
