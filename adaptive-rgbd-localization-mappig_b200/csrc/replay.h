// csrc/replay.h — host+device restatements of two library algorithms the reference's RANSAC depends on:
//
//  * libstdc++ std::sort (introsort: median-of-3 to first, unguarded Hoare partition, threshold 16,
//    heapsort fallback at depth 2*floor(log2 n), final insertion sort).  Ransac::Iterate sorts the good
//    matches with std::sort on DMatch::operator< (distance only; Odometry/ransac.cpp:199), which is NOT stable
//    (quirk Q6): the resulting permutation of equal-distance matches feeds both the sample ids and the f32
//    accumulation order, so the device replays the exact algorithm instead of "a" sort.
//  * glibc rand()/srand() (TYPE_3 additive feedback generator, degree 31, separation 3), which
//    Ransac::SampleMatches draws from (ransac.cpp:269-293).
//
// Both are pinned against the real std::sort / libc rand() in tests (oracle calls the library versions).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define ORBF_HD __host__ __device__
#else
#define ORBF_HD
#endif

namespace replay {

// ------------------------------- std::sort (libstdc++) ------------------------------------------------
template <typename T, typename Less>
struct IntroSort {
    T* a;
    Less less;
    ORBF_HD void swp(int i, int j) { T t = a[i]; a[i] = a[j]; a[j] = t; }

    ORBF_HD void move_median_to_first(int result, int x, int y, int z)
    {
        if (less(a[x], a[y])) {
            if (less(a[y], a[z])) swp(result, y);
            else if (less(a[x], a[z])) swp(result, z);
            else swp(result, x);
        } else if (less(a[x], a[z])) swp(result, x);
        else if (less(a[y], a[z])) swp(result, z);
        else swp(result, y);
    }
    ORBF_HD int unguarded_partition(int first, int last, int pivot)
    {
        while (true) {
            while (less(a[first], a[pivot])) ++first;
            --last;
            while (less(a[pivot], a[last])) --last;
            if (!(first < last)) return first;
            swp(first, last);
            ++first;
        }
    }
    // ---- heap helpers (std::__adjust_heap / __push_heap with operator<) ----
    ORBF_HD void push_heap(int first, int hole, int top, T value)
    {
        int parent = (hole - 1) / 2;
        while (hole > top && less(a[first + parent], value)) {
            a[first + hole] = a[first + parent];
            hole = parent;
            parent = (hole - 1) / 2;
        }
        a[first + hole] = value;
    }
    ORBF_HD void adjust_heap(int first, int hole, int len, T value)
    {
        const int top = hole;
        int child = hole;
        while (child < (len - 1) / 2) {
            child = 2 * (child + 1);
            if (less(a[first + child], a[first + (child - 1)])) child--;
            a[first + hole] = a[first + child];
            hole = child;
        }
        if ((len & 1) == 0 && child == (len - 2) / 2) {
            child = 2 * (child + 1);
            a[first + hole] = a[first + (child - 1)];
            hole = child - 1;
        }
        push_heap(first, hole, top, value);
    }
    ORBF_HD void heap_sort(int first, int last)   // std::__partial_sort(first, last, last)
    {
        const int len = last - first;
        if (len >= 2) {
            int parent = (len - 2) / 2;
            while (true) {
                T v = a[first + parent];
                adjust_heap(first, parent, len, v);
                if (parent == 0) break;
                parent--;
            }
        }
        for (int l = last; l - first > 1;) {
            --l;
            T v = a[l];
            a[l] = a[first];
            adjust_heap(first, 0, l - first, v);
        }
    }
    ORBF_HD void unguarded_linear_insert(int last)
    {
        T val = a[last];
        int next = last - 1;
        while (less(val, a[next])) { a[last] = a[next]; last = next; --next; }
        a[last] = val;
    }
    ORBF_HD void insertion_sort(int first, int last)
    {
        if (first == last) return;
        for (int i = first + 1; i != last; ++i) {
            if (less(a[i], a[first])) {
                T val = a[i];
                for (int j = i; j > first; --j) a[j] = a[j - 1];
                a[first] = val;
            } else unguarded_linear_insert(i);
        }
    }
    ORBF_HD void sort(int n)
    {
        if (n <= 1) return;
        // __introsort_loop with an explicit stack (recursion on the right part, iteration on the left)
        int stFirst[64], stLast[64], stDepth[64];
        int sp = 0;
        int lg = 0;
        for (int t = n; t > 1; t >>= 1) ++lg;
        stFirst[0] = 0; stLast[0] = n; stDepth[0] = 2 * lg; sp = 1;
        while (sp > 0) {
            --sp;
            int first = stFirst[sp], last = stLast[sp], depth = stDepth[sp];
            while (last - first > 16) {
                if (depth == 0) { heap_sort(first, last); break; }
                --depth;
                const int mid = first + (last - first) / 2;
                move_median_to_first(first, first + 1, mid, last - 1);
                const int cut = unguarded_partition(first + 1, last, first);
                // the library recurses into [cut, last) first, then loops on [first, cut); the two ranges are
                // disjoint, so processing order does not change the outcome — push the right part for later
                if (sp < 64) { stFirst[sp] = cut; stLast[sp] = last; stDepth[sp] = depth; ++sp; }
                last = cut;
            }
        }
        // __final_insertion_sort
        if (n > 16) {
            insertion_sort(0, 16);
            for (int i = 16; i != n; ++i) unguarded_linear_insert(i);
        } else insertion_sort(0, n);
    }
};

// ------------------------------- glibc srand / rand ------------------------------------------------
struct GlibcRand {
    int32_t r[31];
    int f, b;
    ORBF_HD void seed(uint32_t s)
    {
        if (s == 0) s = 1;
        r[0] = (int32_t)s;
        int32_t word = (int32_t)s;
        for (int i = 1; i < 31; ++i) {
            const int32_t hi = word / 127773, lo = word % 127773;
            word = 16807 * lo - 2836 * hi;
            if (word < 0) word += 2147483647;
            r[i] = word;
        }
        f = 3; b = 0;
        for (int i = 0; i < 310; ++i) next();
    }
    ORBF_HD int32_t next()
    {
        const uint32_t v = (uint32_t)r[f] + (uint32_t)r[b];
        r[f] = (int32_t)v;
        const int32_t out = (int32_t)(v >> 1);
        if (++f >= 31) f = 0;
        if (++b >= 31) b = 0;
        return out;
    }
};

// Ransac::SampleMatches (ransac.cpp:269-293): until S unique ids: id = min(rand() % M, rand() % M); ids returned
// ascending (std::set order), -1 padded if the 10000-draw safety net fires first.
ORBF_HD inline void sample_row(GlibcRand& g, int M, int S, int* row)
{
    int ids[8];
    int cnt = 0, safety = 0;
    while (cnt < S && M >= S) {
        int id1 = g.next() % M;
        const int id2 = g.next() % M;
        if (id1 > id2) id1 = id2;
        int pos = 0;
        bool dup = false;
        while (pos < cnt && ids[pos] <= id1) { if (ids[pos] == id1) dup = true; ++pos; }
        if (!dup) {
            for (int j = cnt; j > pos; --j) ids[j] = ids[j - 1];
            ids[pos] = id1;
            ++cnt;
        }
        if (++safety > 10000) break;
    }
    for (int k = 0; k < S; ++k) row[k] = (k < cnt) ? ids[k] : -1;
}

}  // namespace replay
