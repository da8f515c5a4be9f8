// Sequential replay of ATen's CPU torch.topk tie rule.
//
// The reference's assignment / selection indices are "whatever torch.topk does on CPU": ATen builds
// (value, index) pairs per slice and runs libstdc++ with a comparator on the VALUE only:
//     k*64 <= n  ->  std::partial_sort(first, first+k, last)          (heap select + sort_heap)
//     otherwise  ->  std::nth_element(first, first+k-1, last) then std::sort(first, first+k-1)
// Those algorithms are deterministic, so replaying the same element moves on the device reproduces the same
// winners *and* the same output order on exact ties.  One thread runs one slice; slices are short
// (bags: 25..125, assignment stage 2: num_pre) or streamed (assignment stage 1 keeps only the k-heap).
#pragma once
#include <math.h>

namespace ptb {

struct VI { float v; int i; };

// value-only comparator; NaN sorts as the largest value like ATen's lambdas
template <bool LARGEST>
__device__ __forceinline__ bool tk_comp(const VI& x, const VI& y) {
  if (LARGEST) return (isnan(x.v) && !isnan(y.v)) || (x.v > y.v);
  return (!isnan(x.v) && isnan(y.v)) || (x.v < y.v);
}

struct PairArray {
  float* v; int* i;
  __device__ __forceinline__ VI get(int p) const { VI r; r.v = v[p]; r.i = i[p]; return r; }
  __device__ __forceinline__ void set(int p, const VI& x) const { v[p] = x.v; i[p] = x.i; }
  __device__ __forceinline__ void swap(int a, int b) const { VI t = get(a); set(a, get(b)); set(b, t); }
};

// ---- libstdc++ heap primitives (bits/stl_heap.h), offsets relative to `first`
template <bool L>
__device__ inline void tk_push_heap(const PairArray& a, int first, int hole, int top, VI value) {
  int parent = (hole - 1) / 2;
  while (hole > top && tk_comp<L>(a.get(first + parent), value)) {
    a.set(first + hole, a.get(first + parent));
    hole = parent;
    parent = (hole - 1) / 2;
  }
  a.set(first + hole, value);
}

template <bool L>
__device__ inline void tk_adjust_heap(const PairArray& a, int first, int hole, int len, VI value) {
  const int top = hole;
  int child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (tk_comp<L>(a.get(first + child), a.get(first + child - 1))) child--;
    a.set(first + hole, a.get(first + child));
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    a.set(first + hole, a.get(first + child - 1));
    hole = child - 1;
  }
  tk_push_heap<L>(a, first, hole, top, value);
}

template <bool L>
__device__ inline void tk_make_heap(const PairArray& a, int first, int last) {
  const int len = last - first;
  if (len < 2) return;
  int parent = (len - 2) / 2;
  while (true) {
    VI value = a.get(first + parent);
    tk_adjust_heap<L>(a, first, parent, len, value);
    if (parent == 0) return;
    parent--;
  }
}

// __pop_heap(first, last, result): *result <- *first, re-heapify [first, last) with the old *result
template <bool L>
__device__ inline void tk_pop_heap(const PairArray& a, int first, int last, int result) {
  VI value = a.get(result);
  a.set(result, a.get(first));
  tk_adjust_heap<L>(a, first, 0, last - first, value);
}

template <bool L>
__device__ inline void tk_sort_heap(const PairArray& a, int first, int last) {
  while (last - first > 1) {
    --last;
    tk_pop_heap<L>(a, first, last, last);
  }
}

// ---- insertion sort (bits/stl_algo.h __insertion_sort / __unguarded_linear_insert)
template <bool L>
__device__ inline void tk_unguarded_linear_insert(const PairArray& a, int last) {
  VI val = a.get(last);
  int next = last - 1;
  while (tk_comp<L>(val, a.get(next))) {
    a.set(last, a.get(next));
    last = next;
    --next;
  }
  a.set(last, val);
}

template <bool L>
__device__ inline void tk_insertion_sort(const PairArray& a, int first, int last) {
  if (first == last) return;
  for (int i = first + 1; i != last; ++i) {
    if (tk_comp<L>(a.get(i), a.get(first))) {
      VI val = a.get(i);
      for (int j = i; j > first; --j) a.set(j, a.get(j - 1));  // move_backward(first, i, i+1)
      a.set(first, val);
    } else {
      tk_unguarded_linear_insert<L>(a, i);
    }
  }
}

// ---- introselect (std::nth_element)
template <bool L>
__device__ inline void tk_move_median_to_first(const PairArray& p, int result, int a, int b, int c) {
  if (tk_comp<L>(p.get(a), p.get(b))) {
    if (tk_comp<L>(p.get(b), p.get(c))) p.swap(result, b);
    else if (tk_comp<L>(p.get(a), p.get(c))) p.swap(result, c);
    else p.swap(result, a);
  } else if (tk_comp<L>(p.get(a), p.get(c))) p.swap(result, a);
  else if (tk_comp<L>(p.get(b), p.get(c))) p.swap(result, c);
  else p.swap(result, b);
}

template <bool L>
__device__ inline int tk_unguarded_partition(const PairArray& p, int first, int last, int pivot) {
  while (true) {
    while (tk_comp<L>(p.get(first), p.get(pivot))) ++first;
    --last;
    while (tk_comp<L>(p.get(pivot), p.get(last))) --last;
    if (!(first < last)) return first;
    p.swap(first, last);
    ++first;
  }
}

template <bool L>
__device__ inline void tk_heap_select(const PairArray& a, int first, int middle, int last) {
  tk_make_heap<L>(a, first, middle);
  for (int i = middle; i < last; ++i)
    if (tk_comp<L>(a.get(i), a.get(first))) tk_pop_heap<L>(a, first, middle, i);
}

template <bool L>
__device__ inline void tk_nth_element(const PairArray& a, int first, int nth, int last) {
  if (first == last || nth == last) return;
  int n = last - first, depth = 0;
  while (n > 1) { n >>= 1; depth++; }   // __lg(last - first)
  depth *= 2;
  while (last - first > 3) {
    if (depth == 0) {
      tk_heap_select<L>(a, first, nth + 1, last);
      a.swap(first, nth);
      return;
    }
    --depth;
    const int mid = first + (last - first) / 2;
    tk_move_median_to_first<L>(a, first, first + 1, mid, last - 1);
    const int cut = tk_unguarded_partition<L>(a, first + 1, last, first);
    if (cut <= nth) first = cut; else last = cut;
  }
  tk_insertion_sort<L>(a, first, last);
}

template <bool L>
__device__ inline void tk_topk(const PairArray& a, int n, int k) {
  if (k * 64 <= n) {
    tk_heap_select<L>(a, 0, k, n);     // std::partial_sort
    tk_sort_heap<L>(a, 0, k);
  } else {
    tk_nth_element<L>(a, 0, k - 1, n);
    tk_insertion_sort<L>(a, 0, k - 1);  // std::sort on <= 16 elements is a plain insertion sort
  }
}

// vals / idx: n (value, index) pairs in slice order; on return the first k entries are torch.topk's
// (values, indices) in its output order.  k <= 17.
__device__ inline void cpu_topk_replay(float* vals, int* idx, int n, int k, bool largest) {
  PairArray a{vals, idx};
  if (largest) tk_topk<true>(a, n, k); else tk_topk<false>(a, n, k);
}

}  // namespace ptb
