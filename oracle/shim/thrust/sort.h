// oracle/shim/thrust/sort.h — TEST INFRASTRUCTURE.  bvh.h:4,87-91 calls thrust::sort from one
// device thread; on the host a stable merge sort stands in (std::sort may not be used: the
// reference comparator returns true on equality, bvh.h:65-69, which is not a strict weak order).
#pragma once
#include <algorithm>
namespace thrust {
template <class It, class Cmp>
inline void sort(It first, It last, Cmp cmp) {
    if (last > first) std::stable_sort(first, last, cmp);
}
} // namespace thrust
