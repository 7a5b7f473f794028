// msvc_sort.h -- the order in which the Microsoft STL's std::sort (VS2015-2019 <algorithm>: introsort, Tukey-ninther
// median guess, three-way partition around the pivot's equal range, insertion sort below 33 elements, 1.5 log2 N depth
// budget) leaves EQUAL elements.  Restated from the published algorithm.
//
// Why it exists: pcl::VoxelGrid sorts (voxel index, point index) pairs by voxel index only and then sums the points of a
// voxel in the sorted order, in float.  std::sort is not stable, so that order -- and with it the last bit of a few
// centroids -- depends on the standard library.  The reference's recorded results (results/4DPCReg) were produced by
// its Windows build; with this order the pre-processing reproduces that build's clouds bit for bit and ALL recorded
// 4x4 matrices are reproduced within print precision (profiles/r01m_refdata_oracle_msvc_*.txt), with the input order
// of a stable sort 3 of 19 differ by up to 9e-4 rad.
#pragma once
#include <algorithm>
#include <utility>
#include <vector>
namespace msvc {
template <class It, class Pr> void med3(It a, It b, It c, Pr lt) {
    if (lt(*b, *a)) std::iter_swap(b, a);
    if (lt(*c, *b)) { std::iter_swap(c, b); if (lt(*b, *a)) std::iter_swap(b, a); }
}
template <class It, class Pr> void guess_median(It first, It mid, It last, Pr lt) {
    const auto count = last - first;
    if (40 < count) {
        const auto step = (count + 1) >> 3, two = step << 1;
        msvc::med3(first, first + step, first + two, lt);
        msvc::med3(mid - step, mid, mid + step, lt);
        msvc::med3(last - two, last - step, last, lt);
        msvc::med3(first + step, mid, last - step, lt);
    } else msvc::med3(first, mid, last, lt);
}
template <class It, class Pr> std::pair<It, It> part3(It first, It last, Pr lt) {
    It mid = first + ((last - first) >> 1);
    msvc::guess_median(first, mid, last - 1, lt);
    It pfirst = mid, plast = pfirst + 1;
    while (first < pfirst && !lt(*(pfirst - 1), *pfirst) && !lt(*pfirst, *(pfirst - 1))) --pfirst;
    while (plast < last && !lt(*plast, *pfirst) && !lt(*pfirst, *plast)) ++plast;
    It gfirst = plast, glast = pfirst;
    for (;;) {
        for (; gfirst < last; ++gfirst) {
            if (lt(*pfirst, *gfirst)) continue;
            else if (lt(*gfirst, *pfirst)) break;
            else if (plast != gfirst) { std::iter_swap(plast, gfirst); ++plast; }
            else ++plast;
        }
        for (; first < glast; --glast) {
            if (lt(*(glast - 1), *pfirst)) continue;
            else if (lt(*pfirst, *(glast - 1))) break;
            else if (--pfirst != glast - 1) std::iter_swap(pfirst, glast - 1);
        }
        if (glast == first && gfirst == last) return {pfirst, plast};
        if (glast == first) {
            if (plast != gfirst) std::iter_swap(pfirst, plast);
            ++plast;
            std::iter_swap(pfirst, gfirst);
            ++pfirst; ++gfirst;
        } else if (gfirst == last) {
            if (--glast != --pfirst) std::iter_swap(glast, pfirst);
            std::iter_swap(pfirst, --plast);
        } else {
            std::iter_swap(gfirst, --glast);
            ++gfirst;
        }
    }
}
template <class It, class Pr> void insertion(It first, It last, Pr lt) {
    if (first == last) return;
    for (It next = first; ++next != last;) {
        It next1 = next;
        auto val = std::move(*next);
        if (lt(val, *first)) { std::move_backward(first, next, ++next1); *first = std::move(val); }
        else {
            for (It first1 = next1; lt(val, *--first1); next1 = first1) *next1 = std::move(*first1);
            *next1 = std::move(val);
        }
    }
}
template <class It, class Pr> void sort_rec(It first, It last, long long ideal, Pr lt) {
    for (;;) {
        if (last - first <= 32) { msvc::insertion(first, last, lt); return; }
        if (ideal <= 0) { std::make_heap(first, last, lt); std::sort_heap(first, last, lt); return; }
        auto mid = msvc::part3(first, last, lt);
        ideal = (ideal >> 1) + (ideal >> 2);
        if (mid.first - first < last - mid.second) { msvc::sort_rec(first, mid.first, ideal, lt); first = mid.second; }
        else { msvc::sort_rec(mid.second, last, ideal, lt); last = mid.first; }
    }
}
template <class It, class Pr> void sort(It first, It last, Pr lt) { msvc::sort_rec(first, last, (long long)(last - first), lt); }
}  // namespace msvc
