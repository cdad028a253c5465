// micro2.cu — which memory path gives the fastest random key probe into an L2-resident table?
//   LDG (ld.cg)  |  LDG.nc  |  ATOM.CAS with return  |  cp.async (LDGSTS) gather into smem
// each measured alone and followed by the RED that adds the value.
#include <cstdio>
#include <cstdlib>
#include "../vega_b200/csrc/kernels.cuh"
using namespace vb;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

enum { P_LDG = 0, P_NC, P_CAS, P_LDGSTS16, P_LDGSTS8, P_MIX, P_N };
static const char *pn[] = {"ld.cg", "ld.nc", "atom.cas", "cp.async16", "cp.async8", "mix ldg/cas"};

VB_D u64 ld_nc_u64(const u64 *p) { u64 r; asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(r) : "l"(p)); return r; }
VB_D void cp_async16(void *smem, const void *g) { u32 s = (u32)__cvta_generic_to_shared(smem); asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(g)); }
VB_D void cp_async8(void *smem, const void *g) { u32 s = (u32)__cvta_generic_to_shared(smem); asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(g)); }
VB_D void cp_async_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }

template <int P, int ROWS, bool RED>
__global__ void __launch_bounds__(256) probe_kernel(const u64 *__restrict__ rows, u64 n, Slot *tab, u32 log_cap, u64 *sink)
{
    __shared__ __align__(16) Slot stage[(P == P_LDGSTS16 || P == P_LDGSTS8) ? ROWS * 256 : 1];
    const u64 pol = policy_evict_first();
    const u64 mask = (1ull << log_cap) - 1;
    const u32 shift = 64 - log_cap;
    constexpr int TILE = 256 * ROWS;
    const u64 n_tiles = (n + TILE - 1) / TILE;
    u64 acc = 0;
    for (u64 tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        u64 k[ROWS], v[ROWS], s[ROWS], kk[ROWS];
#pragma unroll
        for (int j = 0; j < ROWS; ++j) {
            u64 idx = tile * TILE + (u64)j * 256 + threadIdx.x;
            if (idx >= n) idx = n - 1;
            ulonglong2 r = ld_stream_u64x2(rows + 2 * idx, pol); k[j] = r.x; v[j] = r.y;
            s[j] = slot_hash(k[j]) >> shift;
        }
#pragma unroll
        for (int j = 0; j < ROWS; ++j) {
            if (P == P_LDG) kk[j] = ld_cg_u64(&tab[s[j]].key);
            else if (P == P_NC) kk[j] = ld_nc_u64(&tab[s[j]].key);
            else if (P == P_CAS) kk[j] = atomicCAS((unsigned long long *)&tab[s[j]].key, (unsigned long long)EMPTY_KEY, (unsigned long long)k[j]);
            else if (P == P_MIX) kk[j] = (j & 1) ? atomicCAS((unsigned long long *)&tab[s[j]].key, (unsigned long long)EMPTY_KEY, (unsigned long long)k[j]) : ld_cg_u64(&tab[s[j]].key);
            else if (P == P_LDGSTS16) cp_async16(&stage[j * 256 + threadIdx.x], &tab[s[j]]);
            else if (P == P_LDGSTS8) cp_async8(&stage[j * 256 + threadIdx.x].key, &tab[s[j]].key);
        }
        if (P == P_LDGSTS16 || P == P_LDGSTS8) {
            cp_async_wait_all();
#pragma unroll
            for (int j = 0; j < ROWS; ++j) kk[j] = stage[j * 256 + threadIdx.x].key;
        }
#pragma unroll
        for (int j = 0; j < ROWS; ++j) {
            u32 guard = 0;
            while (kk[j] != k[j] && guard++ < 64) { s[j] = (s[j] + 1) & mask; kk[j] = ld_cg_u64(&tab[s[j]].key); }
            if (RED) atomicAdd((unsigned long long *)&tab[s[j]].acc, (unsigned long long)v[j]);
            else acc += s[j];
        }
    }
    if (acc == 0x123456789ull) *sink = acc;
}

template <int P, int ROWS, bool RED>
static float run(const u64 *rows, u64 n, Slot *tab, u32 log_cap, u64 *sink, int grid)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    probe_kernel<P, ROWS, RED><<<grid, 256>>>(rows, n, tab, log_cap, sink);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(a);
    for (int r = 0; r < 3; ++r) probe_kernel<P, ROWS, RED><<<grid, 256>>>(rows, n, tab, log_cap, sink);
    cudaEventRecord(b);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / 3;
}

template <int P>
static void sweep(const u64 *rows, u64 n, Slot *tab, u32 log_cap, u64 *sink, int sms)
{
    for (int occ : {4, 8}) {
        float a4 = run<P, 4, false>(rows, n, tab, log_cap, sink, sms * occ), b4 = run<P, 4, true>(rows, n, tab, log_cap, sink, sms * occ);
        float a8 = run<P, 8, false>(rows, n, tab, log_cap, sink, sms * occ), b8 = run<P, 8, true>(rows, n, tab, log_cap, sink, sms * occ);
        printf("  %-12s ctas/SM %d | rows/thr 4: probe %7.3f ms (%6.1f G/s)  probe+red %7.3f ms (%6.1f G/s) | rows/thr 8: probe %7.3f  probe+red %7.3f\n",
               pn[P], occ, a4, n / a4 / 1e6, b4, n / b4 / 1e6, a8, b8);
    }
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
        sweep<P_LDG>(rows, n, tab, log_cap, sink, p.multiProcessorCount);
        sweep<P_NC>(rows, n, tab, log_cap, sink, p.multiProcessorCount);
        sweep<P_CAS>(rows, n, tab, log_cap, sink, p.multiProcessorCount);
        sweep<P_LDGSTS16>(rows, n, tab, log_cap, sink, p.multiProcessorCount);
        sweep<P_LDGSTS8>(rows, n, tab, log_cap, sink, p.multiProcessorCount);
        sweep<P_MIX>(rows, n, tab, log_cap, sink, p.multiProcessorCount);
        cudaFree(tab); cudaFree(ctl);
    }
    return 0;
}
