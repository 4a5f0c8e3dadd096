// sort_canon.h — TEST INFRASTRUCTURE, force-included (-include) ONLY into the "canon" build of the reference's
// src/MOVExtractor.cc (oracle/Makefile: _ref/libmovref_canon.so).
// The reference orders prev->mvVF with std::sort and a comparator that leaves ties (equal age and equal descriptor
// popcount) in an implementation-defined order (MOVExtractor.cc:249-252). The CUDA path and the oracle canonicalise that
// order to a STABLE sort. To compare whole multi-frame runs against the reference's own code, this header
// makes the unqualified call `sort(begin(mvVF), end(mvVF), cmp)` inside namespace MOV_SLAM resolve to a stable merge sort
// (a more specialised overload found before std::sort). The plain build (_ref/libmovref.so) keeps std::sort and is compared
// on tie-free tables. The reference source itself is not modified in either build.
#pragma once
#include <vector>
#include "Frame.h"
namespace MOV_SLAM {
template <class Cmp>
inline void sort(std::vector<VideoFeature>::iterator first, std::vector<VideoFeature>::iterator last, Cmp cmp) {
    const size_t n = (size_t)(last - first);
    std::vector<size_t> idx(n), tmp(n);
    for (size_t i = 0; i < n; i++) idx[i] = i;
    for (size_t w = 1; w < n; w *= 2) {  // bottom-up merge: take from the right run only when it is strictly before
        for (size_t lo = 0; lo < n; lo += 2 * w) {
            const size_t mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
            size_t a = lo, b = mid, o = lo;
            while (a < mid && b < hi) tmp[o++] = cmp(first[idx[b]], first[idx[a]]) ? idx[b++] : idx[a++];
            while (a < mid) tmp[o++] = idx[a++];
            while (b < hi) tmp[o++] = idx[b++];
        }
        idx.swap(tmp);
    }
    std::vector<VideoFeature> sorted;
    sorted.reserve(n);
    for (size_t i = 0; i < n; i++) sorted.push_back(first[idx[i]]);
    for (size_t i = 0; i < n; i++) first[i] = sorted[i];
}
}  // namespace MOV_SLAM
