// set_ops.cpp — tests/test_rdd.rs:484-521 (intersection → [3,4,5,13]) and :675-699 (subtract → [0,1,2,10,12,19])
// through the C++ host mirror (include/vega_b200.hpp) over the C ABI.  Needs a GPU to run.
#include <algorithm>
#include <cstdio>
#include "vega_b200.hpp"

int main()
{
    auto sc = vega::Context::create(0);
    const std::vector<int32_t> col1{1, 2, 3, 4, 5, 10, 12, 13, 19, 0}, col2{3, 4, 5, 6, 7, 8, 11, 13};
    auto first = sc->parallelize(col1, 2), second = sc->parallelize(col2, 4);
    auto inter = first.intersection(second, 3);
    std::sort(inter.begin(), inter.end());
    auto sub = sc->parallelize(col1, 4).subtract(sc->parallelize(col2, 4));
    std::sort(sub.begin(), sub.end());
    const bool ok = inter == std::vector<int32_t>{3, 4, 5, 13} && sub == std::vector<int32_t>{0, 1, 2, 10, 12, 19};
    std::printf("intersection:");
    for (auto x : inter) std::printf(" %d", x);
    std::printf("\nsubtract:");
    for (auto x : sub) std::printf(" %d", x);
    std::printf("\n%s\n", ok ? "ok" : "MISMATCH");
    return ok ? 0 : 1;
}
