// Test infrastructure (built with nvcc by tests/test_bvh_device_partition.py, runs on the CPU): the device builder's index
// introsort (csrc/bvh_device.cuh, DbIntroSort) against the host builder's item introsort (csrc/bvh_build.hpp, NetIntroSort,
// the restatement of System.Array.Sort): same PERMUTATION on keys with many ties, NaNs, equal keys, sorted and reversed runs.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../yetanotherconsolegameengine_b200/csrc/bvh_build.hpp"
#include "../yetanotherconsolegameengine_b200/csrc/bvh_device.cuh"
int main() {
    unsigned seed = 12345;
    auto rnd = [&]() { seed = seed * 1664525u + 1013904223u; return seed >> 8; };
    long cases = 0;
    for (int rep = 0; rep < 400; rep++) {
        const int n = 1 + (int)(rnd() % (rep < 200 ? 64 : 5000)), kind = rep % 6, start = (int)(rnd() % 7);
        std::vector<float> key(n + start + 3);
        for (int i = 0; i < (int)key.size(); i++) {
            switch (kind) {
                case 0: key[i] = (float)(rnd() % 1000) * 0.25f; break;           // random
                case 1: key[i] = (float)(rnd() % 4); break;                      // many ties
                case 2: key[i] = 1.5f; break;                                    // all equal (the builder's fallback case)
                case 3: key[i] = (float)i; break;                                // sorted
                case 4: key[i] = (float)-i; break;                               // reversed
                default: key[i] = (rnd() % 9 == 0) ? NAN : (float)(rnd() % 50);  // NaNs
            }
        }
        std::vector<ycge::BuildItem> items(key.size());
        std::vector<int> idx(key.size());
        for (int i = 0; i < (int)key.size(); i++) { items[i].index = i; items[i].c[0] = items[i].c[1] = items[i].c[2] = key[i]; idx[i] = i; }
        ycge::detail::NetIntroSort(items.data(), 1).run(start, n);
        ycge::DbIntroSort s; s.a_ = idx.data(); s.key_ = key.data();
        s.run(start, n);
        for (int i = 0; i < (int)key.size(); i++)
            if (items[i].index != idx[i]) { printf("FAIL: rep %d kind %d n %d: position %d holds %d vs %d\n", rep, kind, n, i, idx[i], items[i].index); return 1; }
        cases++;
    }
    printf("%ld cases: same permutation -> ok\n", cases);
    return 0;
}
