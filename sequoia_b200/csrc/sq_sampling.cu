// Draft-side sampling kernels (utils.py:5-32): temperature softmax, exponential-race sampling without
// replacement / top-k, residual distribution, row argmax.  One 1024-thread CTA per vocabulary row; the row lives in
// registers (16-byte striped loads, fully coalesced); reductions are warp-shuffle / redux + one shared-memory hop.
// HBM-bound by bytes (V*2 logits + V*2 rand per row) but in practice SFU/latency-bound: few rows, one SM each.
//
// fp16 rounding chain of the reference, reproduced step by step:
//   xt = fp16(x * (1/T))        torch's CUDA div-by-scalar multiplies by the fp32 reciprocal
//   q  = fp16(exp(xt - max) / sum)            softmax computes in fp32, rounds once
//   sc = fp16(fp16(log(u)) / q)               rand.log() and the division are separate fp16 ops
// exp uses ex2.approx (rel. error ~1e-6, far below the fp16 rounding that follows); log stays full precision
// because log(u) for u close to 1 decides the top ranks.
#include "sq_common.cuh"

namespace sq {

constexpr int NT = 1024;
constexpr int NW = NT / 32;
constexpr int CH = 4;          // 16-byte chunks per thread: V <= NT*CH*8 = 32768

// fp16 bits -> uint16 whose unsigned order equals the float order (-inf lowest, +inf highest)
__device__ __forceinline__ uint32_t ord16(__half h) {
  const uint32_t b = __half_as_ushort(h);
  return (b & 0x8000u) ? (~b & 0xFFFFu) : (b | 0x8000u);
}

__device__ __forceinline__ uint32_t block_max_u32(uint32_t v, uint32_t* red) {
  v = __reduce_max_sync(0xffffffffu, v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  return __reduce_max_sync(0xffffffffu, red[l]);   // NW == 32
}

// Loads the row striped: chunk c = i*NT + tid holds elements [8c, 8c+8).  Out-of-range chunks read as -inf.
__device__ __forceinline__ void load_row(const __half* __restrict__ row, int V, Pack8 (&x)[CH]) {
  const int nvec = V / 8;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    const int c = i * NT + threadIdx.x;
    if (c < nvec) x[i].u = reinterpret_cast<const uint4*>(row)[c];
    else x[i].u = make_uint4(0xFC00FC00u, 0xFC00FC00u, 0xFC00FC00u, 0xFC00FC00u);
  }
}

// in place: x <- fp16(x * inv_T); returns the row max and sum(exp(xt - max)) (fp32)
__device__ __forceinline__ void scale_and_stats(Pack8 (&x)[CH], float inv_T, float* red, float& mx, float& sum) {
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < CH; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      x[i].h[j] = f2h(h2f(x[i].h[j]) * inv_T);
      m = fmaxf(m, h2f(x[i].h[j]));
    }
  mx = block_max<NW>(m, red);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < CH; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) s += __expf(h2f(x[i].h[j]) - mx);
  sum = block_sum<NW>(s, red);
}

__device__ __forceinline__ __half softmax_val(__half xt, float mx, float inv_sum_unused, float sum) {
  return f2h(__fdividef(__expf(h2f(xt) - mx), sum));
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) softmax_T_kernel(const __half* __restrict__ logits, int64_t ld_in,
                                                        __half* __restrict__ out, int64_t ld_out, int V, float inv_T) {
  __shared__ float red[NW];
  Pack8 x[CH];
  load_row(logits + blockIdx.x * ld_in, V, x);
  float mx, sum;
  scale_and_stats(x, inv_T, red, mx, sum);
  const int nvec = V / 8;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    const int c = i * NT + threadIdx.x;
    if (c < nvec) {
      Pack8 o;
#pragma unroll
      for (int j = 0; j < 8; ++j) o.h[j] = softmax_val(x[i].h[j], mx, 0.f, sum);
      reinterpret_cast<uint4*>(out + blockIdx.x * ld_out)[c] = o.u;
    }
  }
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) sample_level_kernel(
    const __half* __restrict__ logits, int64_t ld_logits, const __half* __restrict__ rand, int64_t ld_rand,
    const int32_t* __restrict__ parent_rows, const int32_t* __restrict__ child_first,
    const int32_t* __restrict__ n_branch, int k_max, int V, float inv_T, int mode, int64_t* __restrict__ positions,
    int64_t* __restrict__ tokens, const int32_t* __restrict__ state) {
  __shared__ float red[NW];
  __shared__ uint32_t redu[NW];
  const int j = blockIdx.x;
  const int nb = n_branch ? n_branch[j] : 0;
  const int k_need = positions ? k_max : min(nb, k_max);   // children this row actually needs (block-uniform)
  if (k_need == 0) return;
  const int prow = parent_rows ? parent_rows[j] : j;
  const int nvec = V / 8;
  uint32_t key[CH * 8];
  {
    Pack8 x[CH];
    load_row(logits + prow * ld_logits, V, x);
    if (mode == 0) {
      float mx, sum;
      scale_and_stats(x, inv_T, red, mx, sum);
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        const int c = i * NT + threadIdx.x;
        Pack8 u;
        if (c < nvec) u.u = reinterpret_cast<const uint4*>(rand + prow * ld_rand)[c];
        else u.u = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float q = h2f(softmax_val(x[i].h[e], mx, 0.f, sum));
          const float lg = rnd16(logf(h2f(u.h[e])));                 // rand.log() in fp16
          const __half sc = f2h(__fdividef(lg, q));                  // / sampling_q in fp16
          key[i * 8 + e] = (c < nvec) ? ((ord16(sc) << 16) | (0xFFFFu - (uint32_t)(c * 8 + e))) : 0u;
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        const int c = i * NT + threadIdx.x;
#pragma unroll
        for (int e = 0; e < 8; ++e)
          key[i * 8 + e] = (c < nvec) ? ((ord16(x[i].h[e]) << 16) | (0xFFFFu - (uint32_t)(c * 8 + e))) : 0u;
      }
    }
  }
  // top-k over unique (score, index) keys; ties on the score resolve to the lower index
  uint32_t best = 0u;
#pragma unroll
  for (int e = 0; e < CH * 8; ++e) best = max(best, key[e]);
  const int base = tokens ? row_base(state, child_first[j]) : 0;
  // Fast path (k <= 32): the k-th largest of the 32 warp maxima is a lower bound of the k-th largest key (at least k keys
  // reach it), and usually only a few more do: collect the keys >= that threshold in shared memory and rank them directly
  // (rank = number of larger candidates) -- three block barriers instead of two per selected child.
  if (k_need <= 32) {
    __shared__ uint32_t cand[NT];
    __shared__ int cnt;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t wmax = __reduce_max_sync(0xffffffffu, best);
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();                       // (also orders the last read of `redu`/`red` by scale_and_stats)
    if (lane == 0) redu[warp] = wmax;
    __syncthreads();
    uint32_t v = redu[lane], thr = 0u;     // NW == 32
    for (int r = 0; r < k_need; ++r) {
      thr = __reduce_max_sync(0xffffffffu, v);
      if (v == thr) v = 0u;
    }
    if (best >= thr && best != 0u) {       // (valid keys are never 0; thr == 0 only when few warps hold valid keys)
#pragma unroll
      for (int e = 0; e < CH * 8; ++e)
        if (key[e] >= thr && key[e] != 0u) {
          const int p = atomicAdd(&cnt, 1);
          if (p < NT) cand[p] = key[e];
        }
    }
    __syncthreads();
    const int C = cnt;
    if (C <= NT) {
      if ((int)threadIdx.x < C) {
        const uint32_t mine = cand[threadIdx.x];
        int rank = 0;
        for (int i = 0; i < C; ++i) rank += (cand[i] > mine) ? 1 : 0;
        if (rank < k_need) {
          const int64_t idx = (int64_t)(0xFFFFu - (mine & 0xFFFFu));
          if (positions) positions[(int64_t)j * k_max + rank] = idx;
          if (tokens && rank < nb) tokens[base + rank] = idx;
        }
      }
      return;
    }
    // (more than NT keys reach the threshold -- e.g. an index-sorted row in top-k mode: fall through to the round loop)
  }
  for (int rnd = 0; rnd < k_need; ++rnd) {
    const uint32_t top = block_max_u32(best, redu);
    if (best == top) {                    // unique owner (keys embed the index)
      const int64_t idx = (int64_t)(0xFFFFu - (top & 0xFFFFu));
      if (positions) positions[(int64_t)j * k_max + rnd] = idx;
      if (tokens && rnd < nb) tokens[base + rnd] = idx;
      best = 0u;
#pragma unroll
      for (int e = 0; e < CH * 8; ++e) {
        if (key[e] == top) key[e] = 0u;
        best = max(best, key[e]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Sampling WITH replacement (SpecInferTree.py:100-105: softmax(l/T).multinomial(k, replacement=True)) by exact integer
// inverse-CDF.  Every fp16 probability is an integer multiple of 2^-24, so w = q * 2^24 is an exact uint32 weight, the
// prefix sums are exact (order-independent) and draw c is the first index whose inclusive prefix exceeds
// (word_c * total) >> 32.  The oracle's `multinomial_words` does the same arithmetic: given equal q rows the draws agree
// bit for bit.  The striped row layout (chunk c = i*NT + tid = elements [8c, 8c+8)) is already in index order, so the
// CDF needs one 4-wide block scan of the per-chunk sums.
__global__ void __launch_bounds__(NT) sample_replace_kernel(
    const __half* __restrict__ logits, int64_t ld_logits, const int64_t* __restrict__ words,
    const int32_t* __restrict__ parent_rows, const int32_t* __restrict__ child_first,
    const int32_t* __restrict__ n_branch, int k_max, int V, float inv_T, int64_t* __restrict__ positions,
    int64_t* __restrict__ tokens, const int32_t* __restrict__ state) {
  __shared__ float red[NW];
  __shared__ uint32_t wtot[CH][NW];
  const int j = blockIdx.x;
  const int prow = parent_rows ? parent_rows[j] : j;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t w[CH * 8];
  uint32_t cs[CH], excl[CH];
  {
    Pack8 x[CH];
    load_row(logits + prow * ld_logits, V, x);
    float mx, sum;
    scale_and_stats(x, inv_T, red, mx, sum);
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      cs[i] = 0u;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float q = h2f(softmax_val(x[i].h[e], mx, 0.f, sum));        // fp16 q; padding chunks are exp(-inf) = 0
        w[i * 8 + e] = (uint32_t)(q * 16777216.f);                        // exact
        cs[i] += w[i * 8 + e];
      }
    }
  }
  // inclusive warp scans of the 4 chunk sums, then the warp totals
  uint32_t inc[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    uint32_t v = cs[i];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    inc[i] = v;
    if (lane == 31) wtot[i][warp] = v;
  }
  __syncthreads();
  uint32_t total = 0u;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    uint32_t v = wtot[i][lane];                   // NW == 32
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    const uint32_t before = __shfl_sync(0xffffffffu, v, (warp + 31) & 31);   // inclusive total of warps < warp
    excl[i] = total + (warp ? before : 0u) + inc[i] - cs[i];
    total += __shfl_sync(0xffffffffu, v, 31);
  }
  const int count = n_branch ? n_branch[j] : k_max;
  const int wbase = child_first ? child_first[j] : j * k_max;
  const int tbase = tokens ? row_base(state, child_first[j]) : 0;
  for (int c = 0; c < count; ++c) {
    const uint32_t word = (uint32_t)words[wbase + c];
    const uint32_t t = (uint32_t)(((uint64_t)word * (uint64_t)total) >> 32);
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      if (t >= excl[i] && t - excl[i] < cs[i]) {  // unique owner chunk
        uint32_t run = excl[i];
        int idx = -1;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          run += w[i * 8 + e];
          if (idx < 0 && run > t) idx = (i * NT + (int)threadIdx.x) * 8 + e;
        }
        if (positions) positions[(int64_t)j * k_max + c] = idx;
        if (tokens) tokens[tbase + c] = idx;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) residual_kernel(const __half* __restrict__ p, const __half* __restrict__ q,
                                                       __half* __restrict__ out, int V) {
  __shared__ float red[NW];
  Pack8 a[CH], b[CH];
  load_row(p, V, a);
  load_row(q, V, b);
  const int nvec = V / 8;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < CH; ++i)
    if (i * NT + threadIdx.x < nvec) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float d = rnd16(h2f(a[i].h[e]) - h2f(b[i].h[e]));
        d = (d < 0.f) ? 0.f : d;                                   // relu_ (NaN propagates like torch)
        a[i].h[e] = f2h(d);
        s += d;
      }
    }
  const float tot = rnd16(block_sum<NW>(s, red));                  // .sum() of an fp16 tensor -> fp16
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    const int c = i * NT + threadIdx.x;
    if (c < nvec) {
      Pack8 o;
#pragma unroll
      for (int e = 0; e < 8; ++e) o.h[e] = f2h(h2f(a[i].h[e]) / tot);
      reinterpret_cast<uint4*>(out)[c] = o.u;
    }
  }
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) argmax_rows_kernel(const __half* __restrict__ logits, int64_t ld, int V,
                                                          int64_t* __restrict__ out) {
  __shared__ uint32_t redu[NW];
  Pack8 x[CH];
  load_row(logits + blockIdx.x * ld, V, x);
  const int nvec = V / 8;
  uint32_t best = 0u;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    const int c = i * NT + threadIdx.x;
    if (c < nvec) {
#pragma unroll
      for (int e = 0; e < 8; ++e) best = max(best, (ord16(x[i].h[e]) << 16) | (0xFFFFu - (uint32_t)(c * 8 + e)));
    }
  }
  const uint32_t top = block_max_u32(best, redu);
  if (threadIdx.x == 0) out[blockIdx.x] = (int64_t)(0xFFFFu - (top & 0xFFFFu));
}


// ---------------------------------------------------------------------------------------------
// get_sampling_logits (utils.py:65-77): nucleus (top-p) filter, in place.  The reference sorts the row, takes
// cumsum(softmax(sorted / T)) in fp16 and removes every token whose PREDECESSOR's cumulative probability exceeds top_p
// (the first sorted token is always kept).  No sort here: the fp16 probabilities are integer multiples of 2^-24, so
// "mass of all tokens ranked before token i" is an exact integer S(i) that a two-level histogram over the 16-bit
// order-preserving key of fp16(logit / T) yields directly (high byte, then low byte inside the boundary bin; per-warp
// private histograms, integer atomics => order-independent, deterministic).  Token i is removed iff
// fp16(S(i) * 2^-24) > fp16(top_p) -- the comparison torch performs.  Tokens with the SAME fp16 logit tie: they are
// ranked by ascending index (a stable descending sort), resolved with a block scan over the boundary key only.
// (torch's cumsum adds the same fp16 terms in fp32 in scan order; the exact sum differs from it by < 2^-20 relative, far
// below the fp16 rounding of the comparison.)
__device__ __forceinline__ bool topp_pred(uint32_t S, float tp) { return h2f(f2h((float)S * (1.0f / 16777216.f))) > tp; }

__global__ void __launch_bounds__(NT) top_p_filter_kernel(__half* __restrict__ logits, int64_t ld, int V, float inv_T,
                                                           float tp) {
  __shared__ float red[NW];
  __shared__ uint32_t whist[NW][256];
  __shared__ uint32_t hist[256];
  __shared__ int sel[2];            // boundary bin, mass ranked before it
  __shared__ uint32_t wtot[CH][NW];
  __half* row = logits + blockIdx.x * ld;
  const int nvec = V / 8;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  Pack8 xt[CH];
  load_row(row, V, xt);
  float mx, sum;
  scale_and_stats(xt, inv_T, red, mx, sum);
  uint32_t before = 0u;             // mass ranked before the current boundary bin
  int hi = -1, key_b = -1;
  for (int level = 0; level < 2; ++level) {
    for (int i = threadIdx.x; i < NW * 256; i += NT) (&whist[0][0])[i] = 0u;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < CH; ++i)
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const uint32_t k = ord16(xt[i].h[e]);
        const uint32_t w = (uint32_t)(h2f(softmax_val(xt[i].h[e], mx, 0.f, sum)) * 16777216.f);
        if (w != 0u && (level == 0 || (int)(k >> 8) == hi)) atomicAdd(&whist[warp][level == 0 ? (k >> 8) : (k & 255u)], w);
      }
    __syncthreads();
    if (threadIdx.x < 256) {
      uint32_t t = 0u;
#pragma unroll 8
      for (int w = 0; w < NW; ++w) t += whist[w][threadIdx.x];
      hist[threadIdx.x] = t;
    }
    if (threadIdx.x == 0) sel[0] = -1;
    __syncthreads();
    if (warp == 0) {
      // lane l owns bins 255-8l .. 248-8l (descending order of value); exclusive scan over the lanes
      uint32_t m[8], tot = 0u;
#pragma unroll
      for (int b = 0; b < 8; ++b) { m[b] = hist[255 - 8 * lane - b]; tot += m[b]; }
      uint32_t inc = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      uint32_t a = before + inc - tot;
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        if (!topp_pred(a, tp) && topp_pred(a + m[b], tp)) { sel[0] = 255 - 8 * lane - b; sel[1] = (int)a; }   // unique
        a += m[b];
      }
    }
    __syncthreads();
    const int bb = sel[0];
    if (bb < 0) return;             // the whole row stays below top_p: nothing to remove (block-uniform)
    before = (uint32_t)sel[1];
    if (level == 0) hi = bb; else key_b = (hi << 8) | bb;
    __syncthreads();
  }
  // ties on the boundary key: the first t_keep of them (ascending index) stay
  uint32_t w_b = 0u;
  int cs[CH], excl[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    cs[i] = 0;
#pragma unroll
    for (int e = 0; e < 8; ++e)
      if ((int)ord16(xt[i].h[e]) == key_b) {
        ++cs[i];
        w_b = (uint32_t)(h2f(softmax_val(xt[i].h[e], mx, 0.f, sum)) * 16777216.f);
      }
  }
  w_b = __reduce_max_sync(0xffffffffu, w_b);
  if (lane == 0) red[warp] = __uint_as_float(w_b);
  int inc[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    int v = cs[i];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    inc[i] = v;
    if (lane == 31) wtot[i][warp] = (uint32_t)v;
  }
  __syncthreads();
  w_b = __reduce_max_sync(0xffffffffu, __float_as_uint(red[lane]));          // NW == 32: every warp gets the bin's weight
  int total = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    int v = (int)wtot[i][lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    const int prev = __shfl_sync(0xffffffffu, v, (warp + 31) & 31);
    excl[i] = total + (warp ? prev : 0) + inc[i] - cs[i];
    total += __shfl_sync(0xffffffffu, v, 31);
  }
  // smallest t with pred(before + t * w_b): tie ranks >= t are removed (pred(before) is false, pred at t = total true)
  int lo = 0, hi_t = total;
  while (lo < hi_t) {
    const int mid = (lo + hi_t) >> 1;
    if (topp_pred(before + (uint32_t)mid * w_b, tp)) hi_t = mid; else lo = mid + 1;
  }
  const int t_keep = lo;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    const int c = i * NT + threadIdx.x;
    if (c >= nvec) continue;
    bool any = false;
    bool rm[8];
    int rank = excl[i];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = (int)ord16(xt[i].h[e]);
      rm[e] = (k < key_b) || (k == key_b && rank >= t_keep);
      if (k == key_b) ++rank;
      any |= rm[e];
    }
    if (any) {
      Pack8 x;
      x.u = reinterpret_cast<const uint4*>(row)[c];
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (rm[e]) x.h[e] = __ushort_as_half((unsigned short)0xFC00u);     // -inf
      reinterpret_cast<uint4*>(row)[c] = x.u;
    }
  }
}

}  // namespace sq

using namespace sq;

#define SQ_CHECK_V(V) \
  SQ_CHECK_ARG((V) % 8 == 0 && (V) > 0 && (V) <= NT * CH * 8, "V=%d must be a multiple of 8, <= 32768", (V))

extern "C" int sq_softmax_T(const sq_half* logits, int64_t ld_in, sq_half* out, int64_t ld_out, int n, int V, float T,
                            void* stream) {
  SQ_CHECK_V(V);
  if (n == 0) return SQ_OK;
  softmax_T_kernel<<<n, NT, 0, (cudaStream_t)stream>>>((const __half*)logits, ld_in, (__half*)out, ld_out, V, 1.0f / T);
  SQ_CHECK_LAUNCH("sq_softmax_T");
  return SQ_OK;
}

extern "C" int sq_sample_level(const sq_half* logits, int64_t ld_logits, const sq_half* rand, int64_t ld_rand,
                               const int32_t* parent_rows, const int32_t* child_first, const int32_t* n_branch,
                               int n_parents, int k_max, int V, float T, int mode, int64_t* positions, int64_t* tokens,
                               const int32_t* state, void* stream) {
  SQ_CHECK_V(V);
  if (n_parents == 0 || k_max == 0) return SQ_OK;
  SQ_CHECK_ARG(mode == 1 || rand != nullptr, "sq_sample_level: rand required for mode 0");
  SQ_CHECK_ARG(tokens == nullptr || (child_first && n_branch), "sq_sample_level: tokens needs child_first/n_branch");
  SQ_CHECK_ARG(k_max <= V, "sq_sample_level: k_max > V");
  sample_level_kernel<<<n_parents, NT, 0, (cudaStream_t)stream>>>((const __half*)logits, ld_logits, (const __half*)rand,
                                                                 ld_rand, parent_rows, child_first, n_branch, k_max, V,
                                                                 1.0f / T, mode, positions, tokens, state);
  SQ_CHECK_LAUNCH("sq_sample_level");
  return SQ_OK;
}

extern "C" int sq_sample_replace(const sq_half* logits, int64_t ld_logits, const int64_t* words,
                                 const int32_t* parent_rows, const int32_t* child_first, const int32_t* n_branch,
                                 int n_parents, int k_max, int V, float T, int64_t* positions, int64_t* tokens,
                                 const int32_t* state, void* stream) {
  SQ_CHECK_V(V);
  if (n_parents == 0 || k_max == 0) return SQ_OK;
  SQ_CHECK_ARG(words != nullptr, "sq_sample_replace: words required");
  SQ_CHECK_ARG(tokens == nullptr || (child_first && n_branch), "sq_sample_replace: tokens needs child_first/n_branch");
  SQ_CHECK_ARG(positions != nullptr || tokens != nullptr, "sq_sample_replace: no output");
  sample_replace_kernel<<<n_parents, NT, 0, (cudaStream_t)stream>>>((const __half*)logits, ld_logits, words, parent_rows,
                                                                   child_first, n_branch, k_max, V, 1.0f / T, positions,
                                                                   tokens, state);
  SQ_CHECK_LAUNCH("sq_sample_replace");
  return SQ_OK;
}

extern "C" int sq_residual(const sq_half* p, const sq_half* q, sq_half* out, int V, void* stream) {
  SQ_CHECK_V(V);
  residual_kernel<<<1, NT, 0, (cudaStream_t)stream>>>((const __half*)p, (const __half*)q, (__half*)out, V);
  SQ_CHECK_LAUNCH("sq_residual");
  return SQ_OK;
}

extern "C" int sq_argmax_rows(const sq_half* logits, int64_t ld, int n, int V, int64_t* out, void* stream) {
  SQ_CHECK_V(V);
  if (n == 0) return SQ_OK;
  argmax_rows_kernel<<<n, NT, 0, (cudaStream_t)stream>>>((const __half*)logits, ld, V, out);
  SQ_CHECK_LAUNCH("sq_argmax_rows");
  return SQ_OK;
}

extern "C" int sq_top_p_filter(sq_half* logits, int64_t ld, int n, int V, float top_p, float T, void* stream) {
  SQ_CHECK_V(V);
  if (n == 0 || top_p >= 1.0f) return SQ_OK;                       // utils.py:68: only when top_p < 1
  const float tp = __half2float(__float2half_rn(top_p));           // torch compares in the tensor's dtype (fp16)
  top_p_filter_kernel<<<n, NT, 0, (cudaStream_t)stream>>>((__half*)logits, ld, V, 1.0f / T, tp);
  SQ_CHECK_LAUNCH("sq_top_p_filter");
  return SQ_OK;
}
