// Draft-side sampling kernels (utils.py:5-32): temperature softmax, exponential-race sampling without
// replacement / top-k, residual distribution, row argmax.  One 512-thread CTA per vocabulary row; the row lives
// in registers of 512 threads (16-byte striped loads, fully coalesced); reductions are warp-shuffle + one shared-memory hop.
// HBM-bound: algorithmic bytes per row = V*2 (logits) + V*2 (rand).  No tensor cores.
#include "sq_common.cuh"

namespace sq {

constexpr int NT = 512;
constexpr int NW = NT / 32;

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
  return __reduce_max_sync(0xffffffffu, l < NW ? red[l] : 0u);
}

// Loads the row striped: chunk c = i*NT + tid holds elements [8c, 8c+8).
template <int CH>
__device__ __forceinline__ void load_row(const __half* __restrict__ row, int V, Pack8 (&x)[CH]) {
  const int nvec = V / 8;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    const int c = i * NT + threadIdx.x;
    if (c < nvec) x[i].u = reinterpret_cast<const uint4*>(row)[c];
  }
}

// softmax statistics of fp16(x / T) : returns max and sum(exp(xt - max)) (fp32), xt kept by the caller via recompute
template <int CH>
__device__ __forceinline__ void softmax_stats(const Pack8 (&x)[CH], int V, float T, float* red, float& mx, float& sum) {
  const int nvec = V / 8;
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < CH; ++i)
    if (i * NT + threadIdx.x < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) m = fmaxf(m, rnd16(h2f(x[i].h[j]) / T));
    }
  mx = block_max<NW>(m, red);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < CH; ++i)
    if (i * NT + threadIdx.x < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) s += expf(rnd16(h2f(x[i].h[j]) / T) - mx);
    }
  sum = block_sum<NW>(s, red);
}

__device__ __forceinline__ __half softmax_val(__half x, float T, float mx, float sum) {
  return f2h(expf(rnd16(h2f(x) / T) - mx) / sum);
}

// ---------------------------------------------------------------------------------------------
template <int CH>
__global__ void __launch_bounds__(NT) softmax_T_kernel(const __half* __restrict__ logits, int64_t ld_in,
                                                        __half* __restrict__ out, int64_t ld_out, int V, float T) {
  __shared__ float red[NW];
  Pack8 x[CH];
  const __half* row = logits + blockIdx.x * ld_in;
  load_row<CH>(row, V, x);
  float mx, sum;
  softmax_stats<CH>(x, V, T, red, mx, sum);
  const int nvec = V / 8;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    const int c = i * NT + threadIdx.x;
    if (c < nvec) {
      Pack8 o;
#pragma unroll
      for (int j = 0; j < 8; ++j) o.h[j] = softmax_val(x[i].h[j], T, mx, sum);
      reinterpret_cast<uint4*>(out + blockIdx.x * ld_out)[c] = o.u;
    }
  }
}

// ---------------------------------------------------------------------------------------------
template <int CH>
__global__ void __launch_bounds__(NT) sample_level_kernel(
    const __half* __restrict__ logits, int64_t ld_logits, const __half* __restrict__ rand, int64_t ld_rand,
    const int32_t* __restrict__ parent_rows, const int32_t* __restrict__ child_first,
    const int32_t* __restrict__ n_branch, int k_max, int V, float T, int mode, int64_t* __restrict__ positions,
    int64_t* __restrict__ tokens, const int32_t* __restrict__ state) {
  __shared__ float red[NW];
  __shared__ uint32_t redu[NW];
  const int j = blockIdx.x;
  const int prow = parent_rows ? parent_rows[j] : j;
  const int nvec = V / 8;
  Pack8 x[CH];
  load_row<CH>(logits + prow * ld_logits, V, x);
  uint32_t key[CH * 8];
  if (mode == 0) {
    float mx, sum;
    softmax_stats<CH>(x, V, T, red, mx, sum);
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const int c = i * NT + threadIdx.x;
      if (c < nvec) {
        Pack8 u;
        u.u = reinterpret_cast<const uint4*>(rand + prow * ld_rand)[c];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float q = h2f(softmax_val(x[i].h[e], T, mx, sum));
          const float lg = rnd16(logf(h2f(u.h[e])));              // rand.log() in fp16
          const __half sc = f2h(lg / q);                          // / sampling_q in fp16
          key[i * 8 + e] = (ord16(sc) << 16) | (0xFFFFu - (uint32_t)(c * 8 + e));
        }
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) key[i * 8 + e] = 0u;
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
  // k rounds of block-wide arg-max over unique (score, index) keys
  uint32_t best = 0u;
#pragma unroll
  for (int e = 0; e < CH * 8; ++e) best = max(best, key[e]);
  const int nb = n_branch ? n_branch[j] : 0;
  const int base = tokens ? row_base(state, child_first[j]) : 0;
  for (int rnd = 0; rnd < k_max; ++rnd) {
    const uint32_t top = block_max_u32(best, redu);
    if (best == top) {                    // unique owner (keys embed the index); top==0 only if V exhausted
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
template <int CH>
__global__ void __launch_bounds__(NT) residual_kernel(const __half* __restrict__ p, const __half* __restrict__ q,
                                                       __half* __restrict__ out, int V) {
  __shared__ float red[NW];
  Pack8 a[CH], b[CH];
  load_row<CH>(p, V, a);
  load_row<CH>(q, V, b);
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
template <int CH>
__global__ void __launch_bounds__(NT) argmax_rows_kernel(const __half* __restrict__ logits, int64_t ld, int V,
                                                          int64_t* __restrict__ out) {
  __shared__ uint32_t redu[NW];
  Pack8 x[CH];
  load_row<CH>(logits + blockIdx.x * ld, V, x);
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

}  // namespace sq

using namespace sq;

#define SQ_DISPATCH_CH(V, CALL4, CALL8)                                                        \
  do {                                                                                         \
    SQ_CHECK_ARG((V) % 8 == 0 && (V) > 0 && (V) <= 32768, "V=%d must be a multiple of 8, <= 32768", (V)); \
    if ((V) <= 16384) { CALL4; } else { CALL8; }                                               \
  } while (0)

extern "C" int sq_softmax_T(const sq_half* logits, int64_t ld_in, sq_half* out, int64_t ld_out, int n, int V, float T,
                            void* stream) {
  if (n == 0) return SQ_OK;
  cudaStream_t st = (cudaStream_t)stream;
  SQ_DISPATCH_CH(V, (softmax_T_kernel<4><<<n, NT, 0, st>>>((const __half*)logits, ld_in, (__half*)out, ld_out, V, T)),
                 (softmax_T_kernel<8><<<n, NT, 0, st>>>((const __half*)logits, ld_in, (__half*)out, ld_out, V, T)));
  SQ_CHECK_LAUNCH("sq_softmax_T");
  return SQ_OK;
}

extern "C" int sq_sample_level(const sq_half* logits, int64_t ld_logits, const sq_half* rand, int64_t ld_rand,
                               const int32_t* parent_rows, const int32_t* child_first, const int32_t* n_branch,
                               int n_parents, int k_max, int V, float T, int mode, int64_t* positions, int64_t* tokens,
                               const int32_t* state, void* stream) {
  if (n_parents == 0 || k_max == 0) return SQ_OK;
  SQ_CHECK_ARG(mode == 1 || rand != nullptr, "sq_sample_level: rand required for mode 0");
  SQ_CHECK_ARG(tokens == nullptr || (child_first && n_branch), "sq_sample_level: tokens needs child_first/n_branch");
  SQ_CHECK_ARG(k_max <= V, "sq_sample_level: k_max > V");
  cudaStream_t st = (cudaStream_t)stream;
  SQ_DISPATCH_CH(V,
                 (sample_level_kernel<4><<<n_parents, NT, 0, st>>>((const __half*)logits, ld_logits, (const __half*)rand,
                                                                  ld_rand, parent_rows, child_first, n_branch, k_max, V,
                                                                  T, mode, positions, tokens, state)),
                 (sample_level_kernel<8><<<n_parents, NT, 0, st>>>((const __half*)logits, ld_logits, (const __half*)rand,
                                                                  ld_rand, parent_rows, child_first, n_branch, k_max, V,
                                                                  T, mode, positions, tokens, state)));
  SQ_CHECK_LAUNCH("sq_sample_level");
  return SQ_OK;
}

extern "C" int sq_residual(const sq_half* p, const sq_half* q, sq_half* out, int V, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SQ_DISPATCH_CH(V, (residual_kernel<4><<<1, NT, 0, st>>>((const __half*)p, (const __half*)q, (__half*)out, V)),
                 (residual_kernel<8><<<1, NT, 0, st>>>((const __half*)p, (const __half*)q, (__half*)out, V)));
  SQ_CHECK_LAUNCH("sq_residual");
  return SQ_OK;
}

extern "C" int sq_argmax_rows(const sq_half* logits, int64_t ld, int n, int V, int64_t* out, void* stream) {
  if (n == 0) return SQ_OK;
  cudaStream_t st = (cudaStream_t)stream;
  SQ_DISPATCH_CH(V, (argmax_rows_kernel<4><<<n, NT, 0, st>>>((const __half*)logits, ld, V, out)),
                 (argmax_rows_kernel<8><<<n, NT, 0, st>>>((const __half*)logits, ld, V, out)));
  SQ_CHECK_LAUNCH("sq_argmax_rows");
  return SQ_OK;
}
