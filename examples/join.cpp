// Mirror of the reference's examples/join.rs: col2.join(col1, 4) with the String payloads
// replaced by indices into the tables below.  Expected 6 rows: keys 1,1,2,2,3,3 (key 4 has no
// partner: inner join).
#include <algorithm>
#include <cstdio>

#include "vega_b200.hpp"

int main()
{
    try {
        auto sc = vega::Context::create();
        const char *t1[] = {"(A,B)", "(C,D)", "(E,F)", "(G,H)"};
        const char *t2[] = {"A1", "A2", "B1", "B2", "C1", "C2"};
        std::vector<std::pair<int32_t, uint64_t>> col1 = {{1, 0}, {2, 1}, {3, 2}, {4, 3}};
        std::vector<std::pair<int32_t, uint64_t>> col2 = {{1, 0}, {1, 1}, {2, 2}, {2, 3}, {3, 4}, {3, 5}};
        auto r1 = sc->parallelize(col1, 4);
        auto r2 = sc->parallelize(col2, 4);
        auto res = r2.join(r1, 4);
        std::sort(res.begin(), res.end());
        std::printf("result: [");
        for (size_t i = 0; i < res.size(); ++i)
            std::printf("%s(%d, (%s, %s))", i ? ", " : "", res[i].first, t2[res[i].second.first], t1[res[i].second.second]);
        std::printf("]\n");
        return res.size() == 6 ? 0 : 1;
    } catch (const vega::Error &e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 2;
    }
}
