// Structured tree-causal mask, shared by the attention kernels.
#pragma once
#include "sq_common.cuh"

namespace sq {

struct RowMask {
  int lim;          // keys c <= lim are visible (causal part)
  int node;         // tree node id (>= 1) or -1
};
__device__ __forceinline__ RowMask row_mask(int slot, int P) {
  RowMask r;
  r.lim = min(slot, P - 1);
  r.node = (slot >= P) ? (slot - (P - 1)) : -1;
  return r;
}
// visibility bits of key columns [c0, c0+32) for one row; bits = that row's packed ancestor words (may be smem)
__device__ __forceinline__ uint32_t vis_word(const RowMask& rm, int c0, int P, int kv_len, const uint32_t* bits,
                                             int tree_words) {
  uint32_t v;
  if (rm.lim >= c0 + 31) v = 0xFFFFFFFFu;
  else if (rm.lim < c0) v = 0u;
  else v = (1u << (rm.lim - c0 + 1)) - 1u;
  if (rm.node >= 0) {
    const int j0 = c0 - (P - 1);                 // tree column of key c0
    const int w0 = j0 >> 5;                      // arithmetic shift: floor
    const int sh = j0 & 31;
    const uint32_t lo = (w0 >= 0 && w0 < tree_words) ? bits[w0] : 0u;
    const uint32_t hi = (w0 + 1 >= 0 && w0 + 1 < tree_words) ? bits[w0 + 1] : 0u;
    v |= __funnelshift_r(lo, hi, sh);
  }
  const int rem = kv_len - c0;
  if (rem <= 0) v = 0u;
  else if (rem < 32) v &= (1u << rem) - 1u;
  return v;
}


}  // namespace sq
