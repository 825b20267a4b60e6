// tests/cpp/warp_partition_model.cpp — the rule behind csrc/ransac.cu: warp_unguarded_partition, checked on the host against the real
// libstdc++ std::__unguarded_partition (bits/stl_algo.h) and against csrc/replay.h's sequential restatement of it.
//
// The kernel partitions the first (long) ranges of the std::sort replay (Odometry/ransac.cpp:199 sorts the good matches with
// DMatch::operator<, distance only, and the order of equal distances feeds the sample ids — quirk Q6) with a warp instead of a thread.
// Rule: the sequential loop exchanges the k-th element from the left that is NOT LESS than the pivot with the k-th element from the
// right that is NOT GREATER than the pivot for as long as the two positions have not crossed; neither scan looks at an element the other
// has moved before they cross.  So the two stop lists follow from flags over the untouched range, K = the number of leading k with
// L[k] < R[k], and the cut = (K == 0) ? L[0] : min(R[K-1], L[K] if it exists).  This file runs that rule the way the kernel does — 32
// lanes, ballots as masks, prefix popcounts, the same loop bounds and the same early exit — on ranges prepared like introsort prepares
// them (median of three moved to `first`), and compares the cut AND the whole array afterwards.
//
// It is a model of the device function (which cannot run without a GPU); the device function itself is covered by the GPU tests
// (tests/test_gpu_ransac.py: sorted match lists / sample ids / inlier lists against the oracle's real std::sort).
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../../adaptive-rgbd-localization-mappig_b200/csrc/replay.h"

typedef unsigned long long u64;

struct KeyLess {                        // the kernel's order: the high word (distance bits) only; the low word carries the match index
    bool operator()(u64 a, u64 b) const { return (uint32_t)(a >> 32) < (uint32_t)(b >> 32); }
};

// 32 lanes in lockstep: every "instruction" of the device function is a loop over the lanes
static int warp_partition_model(std::vector<u64>& keys, std::vector<uint16_t>& scratch, int first, int last)
{
    const uint32_t pv = (uint32_t)(keys[first] >> 32);
    const int lo0 = first + 1, n = last - lo0;
    uint16_t* Lpos = scratch.data();
    uint16_t* Rpos = Lpos + n;
    int nL = 0, nR = 0;
    for (int b = 0; b < n; b += 32) {
        uint32_t ml = 0, mr = 0;
        bool fl[32], fr[32];
        for (int lane = 0; lane < 32; ++lane) {
            const int i = lo0 + b + lane, j = last - 1 - (b + lane);
            const bool in = b + lane < n;
            fl[lane] = in && !((uint32_t)(keys[i] >> 32) < pv);
            fr[lane] = in && !(pv < (uint32_t)(keys[j] >> 32));
            ml |= (uint32_t)fl[lane] << lane; mr |= (uint32_t)fr[lane] << lane;
        }
        for (int lane = 0; lane < 32; ++lane) {
            const uint32_t below = (1u << lane) - 1;
            const int i = lo0 + b + lane, j = last - 1 - (b + lane);
            if (fl[lane]) Lpos[nL + __builtin_popcount(ml & below)] = (uint16_t)i;
            if (fr[lane]) Rpos[nR + __builtin_popcount(mr & below)] = (uint16_t)j;
        }
        nL += __builtin_popcount(ml); nR += __builtin_popcount(mr);
    }
    const int nmin = std::min(nL, nR);
    int K = 0;
    for (int b = 0; b < nmin; b += 32) {
        uint32_t m = 0;
        for (int lane = 0; lane < 32; ++lane) {
            const int k = b + lane;
            m |= (uint32_t)(k < nmin && Lpos[k] < Rpos[k]) << lane;
        }
        K += __builtin_popcount(m);
        if (m != 0xffffffffu) break;
    }
    for (int k = 0; k < K; ++k) std::swap(keys[Lpos[k]], keys[Rpos[k]]);     // disjoint positions: any lane order
    int cut;
    if (K == 0) cut = Lpos[0];
    else { cut = Rpos[K - 1]; if (K < nL && (int)Lpos[K] < cut) cut = Lpos[K]; }
    return cut;
}

int main(int argc, char** argv)
{
    const int rounds = argc > 1 ? atoi(argv[1]) : 20000;
    std::mt19937 rng(12345);
    long checked = 0;
    for (int t = 0; t < rounds; ++t) {
        // a range inside a longer array (the kernel partitions [first, last) of the pair's key list), length > 16 like introsort's loop
        const int len = 17 + (int)(rng() % (t % 7 == 0 ? 1500 : 200));
        const int first = (int)(rng() % 40), tail = (int)(rng() % 40), last = first + len;
        std::vector<u64> a(first + len + tail);
        const int mode = t % 6;
        const uint32_t span = mode == 0 ? 2 : mode == 1 ? 3 : mode == 2 ? 64 : mode == 3 ? 256 : 1u << 30;
        for (size_t i = 0; i < a.size(); ++i) a[i] = ((u64)(rng() % span) << 32) | (uint32_t)i;
        if (mode == 5) {                                   // all equal / sorted / reversed ranges
            const int sub = (t / 6) % 3;
            for (int i = first; i < last; ++i) {
                const uint32_t v = sub == 0 ? 7u : sub == 1 ? (uint32_t)i : (uint32_t)(last - i);
                a[i] = ((u64)v << 32) | (uint32_t)i;
            }
        }
        // introsort's preparation of the range: median of (first + 1, mid, last - 1) moved to first (real libstdc++)
        const int mid = first + (last - first) / 2;
        std::__move_median_to_first(a.begin() + first, a.begin() + first + 1, a.begin() + mid, a.begin() + last - 1,
            __gnu_cxx::__ops::__iter_comp_iter(KeyLess()));
        std::vector<u64> lib = a, seq = a, par = a;
        // (1) the real library function
        const int cutLib = (int)(std::__unguarded_partition(lib.begin() + first + 1, lib.begin() + last, lib.begin() + first,
            __gnu_cxx::__ops::__iter_comp_iter(KeyLess())) - lib.begin());
        // (2) csrc/replay.h's sequential restatement (what the kernel's thread-per-range levels run)
        replay::IntroSort<u64, KeyLess> S{ seq.data(), KeyLess() };
        const int cutSeq = S.unguarded_partition(first + 1, last, first);
        // (3) the warp rule
        std::vector<uint16_t> scratch(2 * (size_t)len + 4);
        const int cutPar = warp_partition_model(par, scratch, first, last);
        if (cutLib != cutSeq || cutLib != cutPar || lib != seq || lib != par) {
            fprintf(stderr, "MISMATCH round %d mode %d len %d: cut lib %d seq %d warp %d, arrays %s / %s\n", t, mode, len, cutLib, cutSeq,
                cutPar, lib == seq ? "seq ok" : "seq differs", lib == par ? "warp ok" : "warp differs");
            return 1;
        }
        ++checked;
    }
    printf("warp partition model: %ld ranges identical to std::__unguarded_partition (cut and contents)\n", checked);
    return 0;
}
