// cub_yardstick.cu — YARDSTICK ONLY (never part of libvega_b200): cub::DeviceRadixSort on the same
// inputs as tools/bench_ops.py, to place our sweep pass (vega_b200/csrc/sweep.cuh) on the same box.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o cub_yardstick cub_yardstick.cu ; ./cub_yardstick [n]
#include <cstdio>
#include <cstdlib>
#include <cub/device/device_radix_sort.cuh>
#include "../vega_b200/csrc/common.cuh"
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__global__ void gen(u64 *k, u64 *v, u32 *k32, u64 n, u64 D)
{
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u64 st = (u64)gridDim.x * blockDim.x;
    for (; i < n; i += st) {
        const u64 x = vb::splitmix64(3 + i);
        if (k) k[i] = x;
        if (v) v[i] = vb::splitmix64(2 + i) & 0xFFFFF;
        if (k32) k32[i] = (u32)(vb::splitmix64(1 + i) % D);
    }
}

template <typename F> static float timeit(F f, int reps = 3)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(a); f(); cudaEventRecord(b); CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, a, b); best = ms < best ? ms : best;
    }
    return best;
}

int main(int argc, char **argv)
{
    const u64 n = argc > 1 ? (u64)atof(argv[1]) : 1000000000ull;
    u64 *k, *k2, *v, *v2; u32 *i32, *i32b;
    CK(cudaMalloc(&k, n * 8)); CK(cudaMalloc(&k2, n * 8)); CK(cudaMalloc(&v, n * 8)); CK(cudaMalloc(&v2, n * 8));
    CK(cudaMalloc(&i32, n * 4)); CK(cudaMalloc(&i32b, n * 4));
    gen<<<148 * 16, 256>>>(k, v, i32, n, 1000000);
    CK(cudaDeviceSynchronize());
    void *tmp = nullptr; size_t tb = 0, t2 = 0, t3 = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, tb, k, k2, (long long)n);
    cub::DeviceRadixSort::SortPairs(nullptr, t2, k, k2, v, v2, (long long)n);
    cub::DeviceRadixSort::SortPairs(nullptr, t3, i32, i32b, v, v2, (long long)n, 0, 20);
    tb = tb > t2 ? tb : t2; tb = tb > t3 ? tb : t3;
    CK(cudaMalloc(&tmp, tb));
    float a = timeit([&] { cub::DeviceRadixSort::SortKeys(tmp, tb, k, k2, (long long)n); });
    printf("{\"yardstick\": \"cub::DeviceRadixSort::SortKeys u64\", \"rows\": %llu, \"ms\": %.3f, \"Gkeys_per_s\": %.2f}\n", (unsigned long long)n, a, n / a / 1e6);
    float b = timeit([&] { cub::DeviceRadixSort::SortPairs(tmp, tb, k, k2, v, v2, (long long)n); });
    printf("{\"yardstick\": \"cub::DeviceRadixSort::SortPairs (u64,u64)\", \"rows\": %llu, \"ms\": %.3f, \"Grows_per_s\": %.2f}\n", (unsigned long long)n, b, n / b / 1e6);
    float c = timeit([&] { cub::DeviceRadixSort::SortPairs(tmp, tb, i32, i32b, v, v2, (long long)n, 0, 20); });
    printf("{\"yardstick\": \"cub::DeviceRadixSort::SortPairs (u32 ids < 1e6, u64), 20 bits = the group_by_key sort\", \"rows\": %llu, \"ms\": %.3f, \"Grows_per_s\": %.2f}\n", (unsigned long long)n, c, n / c / 1e6);
    return 0;
}
