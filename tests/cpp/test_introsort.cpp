// CPU check of vins_fast_b200/csrc/introsort_emul.h against the real libstdc++ std::sort with the
// reference's comparator (feature_tracker.cpp:228-231).  Exit code 0 = every case identical.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <utility>
#include <vector>
#include "../../vins_fast_b200/csrc/introsort_emul.h"

static std::vector<int> ref_perm(const std::vector<int>& keys) {
    std::vector<std::pair<int, std::pair<float, int>>> v;
    for (int i = 0; i < (int)keys.size(); ++i) v.emplace_back(keys[i], std::make_pair(0.f, i));
    std::sort(v.begin(), v.end(), [](const std::pair<int, std::pair<float, int>>& a,
                                     const std::pair<int, std::pair<float, int>>& b) { return a.first > b.first; });
    std::vector<int> p;
    for (auto& e : v) p.push_back(e.second.second);
    return p;
}

static std::vector<int> emu_perm(const std::vector<int>& keys) {
    int n = (int)keys.size();
    std::vector<vf::SortItem> v(n);
    for (int i = 0; i < n; ++i) { v[i].key = keys[i]; v[i].payload = i; }
    std::vector<int> stack(3 * (2 * vf::floor_log2(n > 0 ? n : 1) + 4));
    vf::vf_introsort_loop(v.data(), n, stack.data());
    // final insertion sort == stable sort by key (descending) of the loop's output: parallel rank on the GPU
    std::vector<int> out(n);
    for (int i = 0; i < n; ++i) {
        int r = 0;
        for (int j = 0; j < n; ++j)
            if (v[j].key > v[i].key || (v[j].key == v[i].key && j < i)) ++r;
        out[r] = v[i].payload;
    }
    return out;
}

int main() {
    std::mt19937 rng(12345);
    long cases = 0, heap_cases = 0;
    auto check = [&](const std::vector<int>& k) {
        ++cases;
        if (ref_perm(k) != emu_perm(k)) { std::printf("MISMATCH n=%zu case=%ld\n", k.size(), cases); std::exit(1); }
    };
    for (int n = 0; n <= 600; ++n)
        for (int rep = 0; rep < 6; ++rep) {
            std::vector<int> k(n);
            int span = 1 + (int)(rng() % (rep < 2 ? 3 : rep < 4 ? 40 : 100000));
            for (auto& x : k) x = 1 + (int)(rng() % span);
            check(k);
        }
    for (int n : {150, 200, 400, 1000, 2000, 4096}) {       // tracker-like: few distinct counts, sorted runs
        for (int rep = 0; rep < 50; ++rep) {
            std::vector<int> k(n);
            int c = 60;
            for (int i = 0; i < n; ++i) { if (rng() % 9 == 0 && c > 2) --c; k[i] = c; }
            for (int i = n - (int)(rng() % 12) - 1; i < n; ++i) if (i >= 0) k[i] = 2;
            check(k);
            std::vector<int> s(k); std::sort(s.begin(), s.end()); check(s);
            std::reverse(s.begin(), s.end()); check(s);
        }
    }
    // median-of-3 killer (Musser): drives the depth limit to 0 -> heap-sort fallback path
    for (int n : {64, 128, 256, 1024, 2048}) {
        std::vector<int> k(n);
        int half = n / 2;
        for (int i = 0; i < half; ++i) {
            if (i % 2 == 0) k[i] = i + 1; else k[i] = half + i + (half % 2 == 0 ? 0 : 1);
            k[half + i] = 2 * (i + 1);
        }
        for (auto& x : k) x = n * 3 - x;     // descending comparator -> mirror the keys
        check(k); ++heap_cases;
        std::vector<int> a(n);               // organ pipe, all equal, ascending/descending keys
        for (int i = 0; i < n; ++i) a[i] = std::min(i, n - i); check(a);
        for (int i = 0; i < n; ++i) a[i] = 7; check(a);
        for (int i = 0; i < n; ++i) a[i] = i; check(a);
        for (int i = 0; i < n; ++i) a[i] = n - i; check(a);
    }
    std::printf("OK %ld cases\n", cases);
    return 0;
}
