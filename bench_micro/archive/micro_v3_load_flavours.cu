// micro3.cu — flavours of the random 8-byte key read: plain/volatile/no_allocate loads vs "atomic reads"
// (atom.add 0 / atom.or 0 / atom.max 0 with return), alone and followed by the RED.
#include <cstdio>
#include <cstdlib>
#include "../vega_b200/csrc/kernels.cuh"
using namespace vb;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

enum { F_CG = 0, F_NOALLOC, F_VOLATILE, F_CV, F_ATOM_ADD0, F_ATOM_OR0, F_ATOM_MAX0, F_LDG128, F_ATOM_ADD_ACC, F_N };
static const char *fn[] = {"ld.cg", "ld.L1::no_allocate", "ld.volatile", "ld.cv", "atom.add(key,0)", "atom.or(key,0)", "atom.max(key,0)", "ld.cg.v2 (16B)", "atom.add(acc,v) ret"};

template <int F>
VB_D u64 probe(Slot *tab, u64 s, u64 v)
{
    u64 r;
    const u64 *p = &tab[s].key;
    if (F == F_CG) asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(r) : "l"(p));
    else if (F == F_NOALLOC) asm volatile("ld.global.L1::no_allocate.u64 %0, [%1];" : "=l"(r) : "l"(p));
    else if (F == F_VOLATILE) asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(r) : "l"(p));
    else if (F == F_CV) asm volatile("ld.global.cv.u64 %0, [%1];" : "=l"(r) : "l"(p));
    else if (F == F_ATOM_ADD0) r = atomicAdd((unsigned long long *)p, 0ull);
    else if (F == F_ATOM_OR0) r = atomicOr((unsigned long long *)p, 0ull);
    else if (F == F_ATOM_MAX0) r = atomicMax((unsigned long long *)p, 0ull);
    else if (F == F_LDG128) { u64 a; asm volatile("ld.global.cg.v2.u64 {%0, %1}, [%2];" : "=l"(r), "=l"(a) : "l"(p)); r ^= (a & 0); }
    else r = atomicAdd((unsigned long long *)&tab[s].acc, (unsigned long long)v);
    return r;
}

template <int F, int ROWS, int MODE>   // MODE 0: first probe only; 1: full probe loop + RED
__global__ void __launch_bounds__(256) k(const u64 *__restrict__ rows, u64 n, Slot *tab, u32 log_cap, u64 *sink)
{
    const u64 pol = policy_evict_first();
    const u64 mask = (1ull << log_cap) - 1;
    const u32 shift = 64 - log_cap;
    constexpr int TILE = 256 * ROWS;
    const u64 n_tiles = (n + TILE - 1) / TILE;
    u64 acc = 0;
    for (u64 tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        u64 kk[ROWS], v[ROWS], s[ROWS], got[ROWS];
#pragma unroll
        for (int j = 0; j < ROWS; ++j) {
            u64 idx = tile * TILE + (u64)j * 256 + threadIdx.x;
            if (idx >= n) idx = n - 1;
            ulonglong2 r = ld_stream_u64x2(rows + 2 * idx, pol); kk[j] = r.x; v[j] = r.y;
            s[j] = slot_hash(kk[j]) >> shift;
        }
#pragma unroll
        for (int j = 0; j < ROWS; ++j) got[j] = probe<F>(tab, s[j], v[j]);
#pragma unroll
        for (int j = 0; j < ROWS; ++j) {
            if (MODE == 0) { acc += got[j]; continue; }
            u32 guard = 0;
            while (got[j] != kk[j] && guard++ < 64) { s[j] = (s[j] + 1) & mask; got[j] = probe<F>(tab, s[j], v[j]); }
            atomicAdd((unsigned long long *)&tab[s[j]].acc, (unsigned long long)v[j]);
        }
    }
    if (acc == 0x123456789ull) *sink = acc;
}

template <int F, int ROWS, int MODE>
static float run(const u64 *rows, u64 n, Slot *tab, u32 log_cap, u64 *sink, int grid)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<F, ROWS, MODE><<<grid, 256>>>(rows, n, tab, log_cap, sink);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(a);
    for (int r = 0; r < 3; ++r) k<F, ROWS, MODE><<<grid, 256>>>(rows, n, tab, log_cap, sink);
    cudaEventRecord(b);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / 3;
}

template <int F>
static void sweep(const u64 *rows, u64 n, Slot *tab, u32 log_cap, u64 *sink, int sms)
{
    float a = run<F, 4, 0>(rows, n, tab, log_cap, sink, sms * 8), a2 = run<F, 8, 0>(rows, n, tab, log_cap, sink, sms * 8);
    float a3 = run<F, 4, 0>(rows, n, tab, log_cap, sink, sms * 4);
    float b = (F == F_ATOM_ADD_ACC) ? 0.f : run<F, 4, 1>(rows, n, tab, log_cap, sink, sms * 8);
    printf("  %-22s first-probe-only: %7.3f ms (%6.1f G/s) [8 rows/thr %7.3f; 4 CTAs/SM %7.3f]   probe-loop+RED: %7.3f ms (%6.1f G/s)\n", fn[F], a,
           n / a / 1e6, a2, a3, b, b > 0 ? n / b / 1e6 : 0.0);
}

int main(int argc, char **argv)
{
    u64 n = argc > 1 ? (u64)atof(argv[1]) : 250000000ull;
    u64 D = argc > 2 ? (u64)atof(argv[2]) : 1000000ull;
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    u64 *rows; CK(cudaMalloc(&rows, n * 16));
    gen_pairs_kernel<<<p.multiProcessorCount * 16, 256>>>(rows, nullptr, nullptr, 0, n, GEN_UNIFORM, D, 0, 1, 2, nullptr);
    u64 *sink; CK(cudaMalloc(&sink, 8));
    for (u32 log_cap : {21u, 22u}) {
        const u64 cap = 1ull << log_cap;
        Slot *tab; CK(cudaMalloc(&tab, (cap + 1) * 16));
        TableCtl *ctl; CK(cudaMalloc(&ctl, sizeof(TableCtl))); CK(cudaMemset(ctl, 0, sizeof(TableCtl)));
        table_init_kernel<<<1024, 256>>>(tab, cap, 0);
        hash_agg_kernel<IN_AOS, OPK_ADD_U64, TX_NONE><<<p.multiProcessorCount * 8, 256>>>(rows, nullptr, n, tab, log_cap, ctl, ~0ull, nullptr);
        CK(cudaDeviceSynchronize());
        printf("--- table 2^%u, load %.2f, rows %.3g\n", log_cap, (double)D / cap, (double)n);
        sweep<F_CG>(rows, n, tab, log_cap, sink, p.multiProcessorCount);
        sweep<F_NOALLOC>(rows, n, tab, log_cap, sink, p.multiProcessorCount);
        sweep<F_VOLATILE>(rows, n, tab, log_cap, sink, p.multiProcessorCount);
        sweep<F_CV>(rows, n, tab, log_cap, sink, p.multiProcessorCount);
        sweep<F_LDG128>(rows, n, tab, log_cap, sink, p.multiProcessorCount);
        sweep<F_ATOM_ADD0>(rows, n, tab, log_cap, sink, p.multiProcessorCount);
        sweep<F_ATOM_OR0>(rows, n, tab, log_cap, sink, p.multiProcessorCount);
        sweep<F_ATOM_MAX0>(rows, n, tab, log_cap, sink, p.multiProcessorCount);
        sweep<F_ATOM_ADD_ACC>(rows, n, tab, log_cap, sink, p.multiProcessorCount);
        cudaFree(tab); cudaFree(ctl);
    }
    return 0;
}
