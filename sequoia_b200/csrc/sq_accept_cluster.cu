// Cluster version of the stochastic accept/reject walk (see sq_accept.cu for the algorithm and the reference citations).
// A thread-block cluster of 8 CTAs x 512 threads splits the vocabulary (one 16-byte chunk per thread); block-level
// partials are exchanged through distributed shared memory with ONE cluster barrier per tested child:
// every CTA speculatively computes its share of the residual sum while the CTA owning the tested token evaluates the
// accept rule, and both travel in the same exchange.  Per child: ~8 exp + 8 div per thread + 1 cluster.sync.
#include <cooperative_groups.h>

#include "sq_common.cuh"
#include "sq_accept_common.cuh"

namespace cg = cooperative_groups;

namespace sq {

constexpr int CL = 8;
constexpr int CNT = 512;
constexpr int CNW = CNT / 32;

struct Xch {                       // double-buffered exchange slots, one per source CTA
  float f[2][CL][4];
  uint32_t u[2][CL][2];
};

__device__ __forceinline__ uint32_t c_ord16(__half h) {
  const uint32_t b = __half_as_ushort(h);
  return (b & 0x8000u) ? (~b & 0xFFFFu) : (b | 0x8000u);
}

struct ClusterCtx {
  cg::cluster_group cluster;
  Xch* x;
  int rank;
  int ph;
  // publish (f0..f3, u0, u1) of this CTA to every CTA, barrier, then read all slots (fixed order => identical results)
  __device__ __forceinline__ void exchange(float f0, float f1, float f2, float f3, uint32_t u0, uint32_t u1) {
    if (threadIdx.x < CL) {
      Xch* remote = cluster.map_shared_rank(x, threadIdx.x);
      remote->f[ph][rank][0] = f0; remote->f[ph][rank][1] = f1;
      remote->f[ph][rank][2] = f2; remote->f[ph][rank][3] = f3;
      remote->u[ph][rank][0] = u0; remote->u[ph][rank][1] = u1;
    }
    cluster.sync();
  }
  __device__ __forceinline__ float fsum(int k) const {
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < CL; ++r) s += x->f[ph][r][k];
    return s;
  }
  __device__ __forceinline__ float fmax_(int k) const {
    float s = -INFINITY;
#pragma unroll
    for (int r = 0; r < CL; ++r) s = fmaxf(s, x->f[ph][r][k]);
    return s;
  }
  __device__ __forceinline__ uint32_t uor(int k) const {
    uint32_t s = 0;
#pragma unroll
    for (int r = 0; r < CL; ++r) s |= x->u[ph][r][k];
    return s;
  }
  __device__ __forceinline__ uint32_t umax(int k) const {
    uint32_t s = 0;
#pragma unroll
    for (int r = 0; r < CL; ++r) s = max(s, x->u[ph][r][k]);
    return s;
  }
  __device__ __forceinline__ void next() { ph ^= 1; }
};

__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(CNT) accept_stochastic_cluster_kernel(
    const __half* __restrict__ target_logits, int64_t ld_t, const __half* __restrict__ draft_logits, int64_t ld_d,
    const __half* __restrict__ r, const __half* __restrict__ noise, const int32_t* __restrict__ succ_off,
    const int32_t* __restrict__ succ, const int32_t* __restrict__ depth, int S, int V, float inv_T,
    int64_t* __restrict__ tokens, int64_t* __restrict__ position_ids, int32_t* __restrict__ accept_idx,
    int32_t* __restrict__ state, int max_target_seq, int policy) {
  __shared__ Xch xch;
  __shared__ float red[CNW];
  __shared__ int32_t sh_acc[1024];
  __shared__ float sh_own[2];                    // owner thread -> block: {etok, flag bits as float}
  ClusterCtx cx{cg::this_cluster(), &xch, 0, 0};
  cx.rank = (int)cx.cluster.block_rank();
  const int P = state[ST_P];
  const int nvec = V / 8;
  const int cpb = (nvec + CL - 1) / CL;          // chunks per CTA (<= CNT)
  const int chunk = cx.rank * cpb + threadIdx.x;
  const bool active = threadIdx.x < cpb && chunk < nvec;
  const uint4 NEG_INF = make_uint4(0xFC00FC00u, 0xFC00FC00u, 0xFC00FC00u, 0xFC00FC00u);
  Pack8 p, xd;
  int cur = 0, n_new = 0;
  bool terminal = false;
  while (true) {
    const int c0 = succ_off[cur], c1 = succ_off[cur + 1];
    const bool leaf = (c0 == c1);
    // load + scale both rows; softmax statistics of both in two exchanges
    p.u = active ? reinterpret_cast<const uint4*>(target_logits + cur * ld_t)[chunk] : NEG_INF;
    xd.u = (active && !leaf) ? reinterpret_cast<const uint4*>(draft_logits + cur * ld_d)[chunk] : NEG_INF;
    float m1 = -INFINITY, m2 = -INFINITY;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      p.h[e] = f2h(h2f(p.h[e]) * inv_T);
      xd.h[e] = f2h(h2f(xd.h[e]) * inv_T);
      m1 = fmaxf(m1, h2f(p.h[e]));
      m2 = fmaxf(m2, h2f(xd.h[e]));
    }
    m1 = block_max<CNW>(m1, red);
    m2 = block_max<CNW>(m2, red);
    cx.exchange(m1, m2, 0.f, 0.f, 0u, 0u);
    const float mxt = cx.fmax_(0);
    float mxd = cx.fmax_(1);
    cx.next();
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      s1 += __expf(h2f(p.h[e]) - mxt);
      s2 += __expf(h2f(xd.h[e]) - mxd);
    }
    s1 = block_sum<CNW>(s1, red);
    s2 = block_sum<CNW>(s2, red);
    cx.exchange(s1, s2, 0.f, 0.f, 0u, 0u);
    const float sumt = cx.fsum(0);
    float sumd = cx.fsum(1);
    cx.next();
#pragma unroll
    for (int e = 0; e < 8; ++e) p.h[e] = f2h(__fdividef(__expf(h2f(p.h[e]) - mxt), sumt));   // p = softmax(target/T)
    if (leaf) break;                                         // residual = p   (SpecTree.py:143-144)
    int accepted = -1;
    for (int ci = c0; ci < c1; ++ci) {
      const int child = succ[ci];
      const int slot = P - 1 + child;
      const int tok = (int)tokens[slot];
      const int tc = tok >> 3, te = tok & 7;
      const bool owner = (tc / cpb == cx.rank) && (threadIdx.x == tc % cpb);
      if (threadIdx.x == 0) { sh_own[0] = 0.f; sh_own[1] = 0.f; }
      __syncthreads();
      // speculative residual share: d = relu(fp16(p - q)), partial sum
      Pack8 dtmp;
      float s = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float q = h2f(f2h(__fdividef(__expf(h2f(xd.h[e]) - mxd), sumd)));
        float d = rnd16(h2f(p.h[e]) - q);
        d = (d < 0.f) ? 0.f : d;                             // relu_; NaN propagates like torch
        dtmp.h[e] = f2h(d);
        s += d;
        if (owner && e == te) {
          const float etok = __expf(h2f(xd.h[e]) - mxd);
          const float thr = rnd16(h2f(r[slot]) * q);         // r * q[token] in fp16
          const float pv = h2f(p.h[e]);
          // strict > (SpecTree.py:152); the SpecInfer policy accepts on >= (SpecInferTree.py:158)
          const int acc = ((policy & SQ_ACCEPT_GE) ? (pv >= thr) : (pv > thr)) ? 1 : 0;
          sh_own[0] = etok;
          sh_own[1] = (float)(acc | ((!(policy & SQ_ACCEPT_KEEP_Q) && h2f(xd.h[e]) >= mxd) ? 2 : 0));
        }
      }
      s = block_sum<CNW>(s, red);                            // (contains the barriers that publish sh_own)
      cx.exchange(s, sh_own[0], 0.f, 0.f, (uint32_t)sh_own[1], 0u);
      const float tot = rnd16(cx.fsum(0));
      const float etok = cx.fsum(1);
      const uint32_t flag = cx.uor(0);
      cx.next();
      if (flag & 1u) { accepted = child; break; }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        p.h[e] = f2h(h2f(dtmp.h[e]) / tot);                  // get_residual (utils.py:5-8)
        if (owner && e == te && !(policy & SQ_ACCEPT_KEEP_Q))
          xd.h[e] = __ushort_as_half((unsigned short)0xFC00u);   // draft_logits[token] = min (SpecTree.py:156)
      }
      if (policy & SQ_ACCEPT_KEEP_Q) continue;               // SpecInfer: q stays softmax(draft/T) for every child
      if (flag & 2u) {                                       // rare: the masked token held the max -> new statistics
        float m = -INFINITY;
#pragma unroll
        for (int e = 0; e < 8; ++e) m = fmaxf(m, h2f(xd.h[e]));
        m = block_max<CNW>(m, red);
        cx.exchange(m, 0.f, 0.f, 0.f, 0u, 0u);
        mxd = cx.fmax_(0);
        cx.next();
        float ss = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) ss += __expf(h2f(xd.h[e]) - mxd);
        ss = block_sum<CNW>(ss, red);
        cx.exchange(ss, 0.f, 0.f, 0.f, 0u, 0u);
        sumd = cx.fsum(0);
        cx.next();
      } else {
        sumd -= etok;
      }
    }
    if (accepted < 0) break;                                 // residual = p   (:157)
    const int slot = P - 1 + accepted;
    if (threadIdx.x == 0) sh_acc[n_new] = slot;
    ++n_new;
    const int64_t t = tokens[slot];
    if (t == 0 || t == 2) { terminal = true; break; }        // (:208)
    cur = accepted;
  }
  bool nan_flag = false;
  int64_t bonus = -1;
  if (!terminal) {
    uint32_t has_nan = 0u, best = 0u;
    if (active) {
      Pack8 nz;
      nz.u = reinterpret_cast<const uint4*>(noise)[chunk];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        has_nan |= __hisnan(p.h[e]) ? 1u : 0u;
        const __half v = f2h(h2f(p.h[e]) / h2f(nz.h[e]));    // multinomial(1) = argmax(residual / Exp(1))  (:222)
        best = max(best, (c_ord16(v) << 16) | (0xFFFFu - (uint32_t)(chunk * 8 + e)));
      }
    }
    has_nan = __syncthreads_or((int)has_nan) ? 1u : 0u;
    best = __reduce_max_sync(0xffffffffu, best);
    __shared__ uint32_t redu[CNW];
    if ((threadIdx.x & 31) == 0) redu[threadIdx.x >> 5] = best;
    __syncthreads();
    best = __reduce_max_sync(0xffffffffu, (threadIdx.x & 31) < CNW ? redu[threadIdx.x & 31] : 0u);
    cx.exchange(0.f, 0.f, 0.f, 0.f, has_nan, best);
    nan_flag = cx.uor(0) != 0u;                              // torch.isnan(residual).any()  (:219)
    bonus = (int64_t)(0xFFFFu - (cx.umax(1) & 0xFFFFu));
    cx.next();
    if (nan_flag) terminal = true;
  }
  if (cx.rank != 0) return;
  __syncthreads();
  finish_verify(sh_acc, n_new, P, terminal, nan_flag, bonus, true, depth, S, tokens, position_ids, accept_idx, state,
                max_target_seq);
}

}  // namespace sq

using namespace sq;

int sq::launch_accept_cluster(const sq_half* target_logits, int64_t ld_t, const sq_half* draft_logits, int64_t ld_d,
                              const sq_half* r, const sq_half* noise, const int32_t* succ_off, const int32_t* succ,
                              const int32_t* depth, int S, int V, float T, int64_t* tokens, int64_t* position_ids,
                              int32_t* accept_idx, int32_t* state, int max_target_seq, int policy, void* stream) {
  SQ_CHECK_ARG(V % 8 == 0 && V > 0 && V <= CL * CNT * 8, "sq_accept_stochastic: V=%d unsupported", V);
  accept_stochastic_cluster_kernel<<<CL, CNT, 0, (cudaStream_t)stream>>>(
      (const __half*)target_logits, ld_t, (const __half*)draft_logits, ld_d, (const __half*)r, (const __half*)noise,
      succ_off, succ, depth, S, V, 1.0f / T, tokens, position_ids, accept_idx, state, max_target_seq, policy);
  SQ_CHECK_LAUNCH("sq_accept_stochastic(cluster)");
  return SQ_OK;
}
