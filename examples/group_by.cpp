// Mirror of the reference's examples/group_by.rs (same data; String keys "x","y" become the
// u64 ids 120,121 because rows crossing the C ABI are POD).  Expected output (any outer order):
//   result: [(120, [1, 2, 3, 4, 5, 6, 7]), (121, [1, 2, 3, 4, 5, 6, 7, 8])]
#include <algorithm>
#include <cstdio>

#include "vega_b200.hpp"

int main()
{
    try {
        auto sc = vega::Context::create();
        std::vector<std::pair<uint64_t, int64_t>> vec;
        for (int i = 1; i <= 7; ++i) vec.emplace_back('x', i);
        for (int i = 1; i <= 8; ++i) vec.emplace_back('y', i);
        auto r = sc->make_rdd(vec, 4);
        auto res = r.group_by_key(4);
        std::sort(res.begin(), res.end());
        std::printf("result: [");
        for (size_t i = 0; i < res.size(); ++i) {
            std::printf("%s(%llu, [", i ? ", " : "", (unsigned long long)res[i].first);
            for (size_t j = 0; j < res[i].second.size(); ++j) std::printf("%s%lld", j ? ", " : "", (long long)res[i].second[j]);
            std::printf("])");
        }
        std::printf("]\n");
        bool ok = res.size() == 2 && res[0].second == std::vector<int64_t>{1, 2, 3, 4, 5, 6, 7} &&
                  res[1].second == std::vector<int64_t>{1, 2, 3, 4, 5, 6, 7, 8};
        return ok ? 0 : 1;
    } catch (const vega::Error &e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 2;
    }
}
