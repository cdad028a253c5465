// micro.cu — times the production hash_agg_kernel (map-side combine) in isolation for a sweep of
// table sizes / resident CTAs, next to a pure streaming read and a RED-only floor.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o micro micro.cu ; ./micro [rows] [distinct]
#include <cstdio>
#include <cstdlib>
#include "../vega_b200/csrc/kernels.cuh"
using namespace vb;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

template <int V>   // 0 stream read, 1 RED only (no key check)
__global__ void __launch_bounds__(256) floor_kernel(const u64 *__restrict__ rows, u64 n, Table t, u64 *sink)
{
    const u64 pol = policy_evict_first();
    const u64 n_tiles = (n + HA_TILE - 1) / HA_TILE;
    u64 acc = 0;
    for (u64 tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
#pragma unroll
        for (int j = 0; j < HA_ROWS; ++j) {
            u64 idx = tile * HA_TILE + (u64)j * 256 + threadIdx.x;
            if (idx >= n) continue;
            ulonglong2 r = ld_stream_u64x2(rows + 2 * idx, pol);
            if (V == 0) acc += r.x ^ r.y;
            else atomicAdd((unsigned long long *)&t.accs[4 * home_bucket(r.x, t.log_cap)], (unsigned long long)r.y);
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
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    printf("device %s, %d SMs; rows %.3g distinct %.3g\n", p.name, sms, (double)n, (double)D);
    u64 *rows; CK(cudaMalloc(&rows, n * 16));
    gen_pairs_kernel<<<sms * 16, 256>>>(rows, nullptr, nullptr, 0, n, GEN_UNIFORM, D, 0, 1, 2, nullptr);
    u64 *sink; CK(cudaMalloc(&sink, 8));
    u32 *slots; CK(cudaMalloc(&slots, n * 4));
    TableCtl *ctl; CK(cudaMalloc(&ctl, sizeof(TableCtl)));
    for (u32 log_cap : {21u, 22u, 23u}) {
        void *base; CK(cudaMalloc(&base, table_bytes(log_cap)));
        Table t = table_at(base, log_cap);
        printf("--- table 2^%u slots, load %.2f\n", log_cap, (double)D / (1ull << log_cap));
        for (int occ : {2, 3, 4, 6, 8}) {
            const int grid = sms * occ;
            auto reset = [&] { table_init_kernel<<<1024, 256>>>(t, 0); cudaMemset(ctl, 0, sizeof(TableCtl)); };
            reset();
            float cold = timeit([&] { reset(); hash_agg_kernel<IN_AOS, OPK_ADD_U64, TX_NONE><<<grid, 256>>>(rows, nullptr, n, t, ctl, ~0ull, nullptr); }, 2);
            float init = timeit([&] { reset(); }, 2);
            float warm = timeit([&] { hash_agg_kernel<IN_AOS, OPK_ADD_U64, TX_NONE><<<grid, 256>>>(rows, nullptr, n, t, ctl, ~0ull, nullptr); });
            float cnt = timeit([&] { hash_agg_kernel<IN_AOS, OPK_COUNT, TX_NONE><<<grid, 256>>>(rows, nullptr, n, t, ctl, ~0ull, nullptr); });
            float mx = timeit([&] { hash_agg_kernel<IN_AOS, OPK_MAX_U64, TX_NONE><<<grid, 256>>>(rows, nullptr, n, t, ctl, ~0ull, nullptr); });
            float dict = timeit([&] { hash_agg_kernel<IN_AOS, OPK_DICT, TX_NONE><<<grid, 256>>>(rows, nullptr, n, t, ctl, ~0ull, slots); });
            float st = timeit([&] { floor_kernel<0><<<grid, 256>>>(rows, n, t, sink); });
            float rd = timeit([&] { floor_kernel<1><<<grid, 256>>>(rows, n, t, sink); });
            printf("  ctas/SM %d: hash_agg sum %7.3f ms (%6.1f Grows/s, %6.0f GB/s) [incl. inserts: %7.3f]  count %7.3f  max %7.3f  dict %7.3f | stream %6.3f  red-only %6.3f\n",
                   occ, warm, n / warm / 1e6, n * 16.0 / warm / 1e6, cold - init, cnt, mx, dict, st, rd);
        }
        cudaFree(base);
    }
    return 0;
}
