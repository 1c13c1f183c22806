// TEST INFRASTRUCTURE (oracle) -- never linked into the product library.
//
// Exposes the permutation that libstdc++'s std::sort produces for the
// comparator of the reference's SetFeatureExtractorMask
// (vins/estimator/feature_tracker.cpp:228-231: sort by track count,
// descending, comparator looks at .first only, unstable).  The tie order of
// that sort decides which of two close features survives the min_dist mask and
// the order of ids_, so the oracle has to use the real std::sort.
//
// Build: g++ -O2 -shared -fPIC -o oracle/_build/libvf_stdsort.so oracle/stdsort_helper.cpp
#include <algorithm>
#include <utility>
#include <vector>

extern "C" void vf_oracle_stdsort_perm(const int* track_count, int n, int* perm_out) {
    // Same element shape as the reference: pair<count, payload>; payload = original index.
    std::vector<std::pair<int, std::pair<float, int>>> v;
    v.reserve(n);
    for (int i = 0; i < n; ++i) v.emplace_back(track_count[i], std::make_pair(0.0f, i));
    std::sort(v.begin(), v.end(),
              [](const std::pair<int, std::pair<float, int>>& a,
                 const std::pair<int, std::pair<float, int>>& b) { return a.first > b.first; });
    for (int i = 0; i < n; ++i) perm_out[i] = v[i].second.second;
}
