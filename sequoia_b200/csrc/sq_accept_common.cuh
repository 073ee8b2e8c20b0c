// Shared by the single-CTA and the cluster accept kernels.
#pragma once
#include "sq_common.cuh"

namespace sq {

int launch_accept_cluster(const sq_half* target_logits, int64_t ld_t, const sq_half* draft_logits, int64_t ld_d,
                          const sq_half* r, const sq_half* noise, const int32_t* succ_off, const int32_t* succ,
                          const int32_t* depth, int S, int V, float T, int64_t* tokens, int64_t* position_ids,
                          int32_t* accept_idx, int32_t* state, int max_target_seq, int policy, void* stream);

// policy bits: SQ_ACCEPT_GE / SQ_ACCEPT_KEEP_Q from include/sequoia_b200.h

// Post-processing shared by both walks.  Runs with the whole block; thread 0 does the (short, ordered) serial part.
// sh_acc[0..n_new) = accepted absolute slots; publishes state[].
// bonus_first: SpecTree writes the bonus token at slot a BEFORE gathering tokens[accept_list] (SpecTree.py:222-224), so
// an accepted node that happens to live at slot a is returned as the bonus token -- reproduced here; GreedyTree
// gathers first (GreedyTree.py:205-207).
__device__ __forceinline__ void finish_verify(const int32_t* sh_acc, int n_new, int P, bool terminal, bool nan_flag, int64_t bonus,
                              bool bonus_first, const int32_t* __restrict__ depth, int S, int64_t* __restrict__ tokens,
                              int64_t* __restrict__ position_ids, int32_t* __restrict__ accept_idx,
                              int32_t* __restrict__ state, int max_target_seq) {
  const int a = P + n_new;
  // Buffer bound (the reference raises an IndexError / shape error at the equivalent slice assignments): tokens[a] needs
  // a < M, the re-laid tree positions need a + S - 1 < M.  M travels in the state word (set by the Tree constructor);
  // 0 = unknown -> fall back to max_target_seq, which callers size as the buffer length.
  const int M = state[ST_M] > 0 ? state[ST_M] : max_target_seq;
  const bool prepare = !terminal && (a + 1 <= max_target_seq) && (a + S <= M);
  const bool bonus_ok = !terminal && a < M;
  if (threadIdx.x == 0) {
    if (bonus_ok && bonus_first) tokens[a] = bonus;         // SpecTree.py:222
    for (int j = 0; j < n_new; ++j) {                       // tokens[:a] = tokens[accept_list]  (SpecTree.py:224)
      const int src = sh_acc[j];
      accept_idx[j] = src;
      tokens[P + j] = tokens[src];
    }
    if (bonus_ok && !bonus_first) tokens[a] = bonus;        // GreedyTree.py:207
    if (prepare) {                                          // prepare_for_next_iter (SpecTree.py:261-271)
      for (int j = 0; j < n_new; ++j) position_ids[P + j] = position_ids[sh_acc[j]];
      position_ids[a] = a;
    }
    state[ST_ACCEPT_LEN] = a;
    state[ST_TERMINAL] = terminal ? 1 : 0;
    state[ST_N_NEW] = n_new;
    state[ST_P_OLD] = P;
    state[ST_BONUS] = terminal ? -1 : (int32_t)bonus;
    state[ST_NAN] = nan_flag ? 1 : 0;
    state[ST_SKIPPED] = (!terminal && !prepare) ? 1 : 0;
    if (prepare) state[ST_P] = a + 1;
  }
  __syncthreads();   // the gather above reads old tree positions that the re-lay below overwrites
  if (prepare) {
    for (int k = 1 + threadIdx.x; k < S; k += blockDim.x) position_ids[a + k] = (int64_t)depth[k] + a;
  }
}

}  // namespace sq
