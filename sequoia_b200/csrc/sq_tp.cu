// Tensor-parallel target forward: ONE kernel = one-shot all-reduce over NVLink peer memory + residual add + RMSNorm.
//
// Under target TP every row-parallel GEMM (o_proj, down_proj) is followed by a sum over the ranks and
// then by `hidden += x; normed = rmsnorm(hidden)`.  Instead of NCCL allreduce + sq_add_rmsnorm (two launches, ~20 us
// for a 1 MiB payload), each rank's kernel reads the partial GEMM outputs of ALL ranks straight from their HBM through
// NVLink/NVSwitch peer mappings (CUDA IPC), sums them in fp32 in a fixed rank order (bit-identical on every rank),
// and continues with the residual add and the norm.  Synchronisation is a per-launch epoch handshake through flag
// words in peer memory (st.release.sys / ld.acquire.sys); the epoch lives on the device, so the kernel is CUDA-graph
// replayable.  The partial buffers alternate (A for o_proj, B for down_proj): the handshake of the NEXT reduction
// proves that every peer has finished reading the previous contents of a buffer before it is overwritten.
#include <cstring>

#include "sq_common.cuh"

namespace sq {

struct TpArgs {
  const __half* proj[8];     // partial GEMM outputs of rank 0..N-1 (peer-mapped, same layout (n_max, hidden))
  uint32_t* flags[8];        // flags[r] = flag array living on rank r (N words; word s is written by rank s)
  uint32_t* epoch;           // local: [0] epoch of the last completed reduction, [1] CTA ticket, [2] error word
  int rank, N;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint4 ld_relaxed_sys_v4(const void* p) {
  uint4 v;
  asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p)
               : "memory");
  return v;
}

// One CTA (256 threads) per row, hidden <= 256*8*MAXV.
template <int MAXV>
__global__ void __launch_bounds__(256) tp_allreduce_add_rmsnorm_kernel(TpArgs t, __half* __restrict__ resid,
                                                                       const __half* __restrict__ w,
                                                                       __half* __restrict__ out, int hidden, float eps) {
  __shared__ float red[8];
  const int r = blockIdx.x, tid = threadIdx.x;
  pdl_wait();                                                // the row-parallel GEMM's partial is complete and visible
  pdl_trigger();
  const uint32_t e = t.epoch[0] + 1;                         // identical in every CTA: only bumped by the last CTA
  if (r == 0 && tid < t.N && tid != t.rank) st_release_sys(t.flags[tid] + t.rank, e);   // "my partial is ready"
  if (tid < t.N && tid != t.rank) {                          // wait until every peer's partial of this epoch is ready
    const uint32_t* f = t.flags[t.rank] + tid;
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys(f) - e) < 0) {
      if (clock64() - t0 > 4000000000LL) { atomicExch(t.epoch + 2, 1u); break; }      // ~2 s: never hang the box
    }
  }
  __syncthreads();
  const int nvec = hidden / 8;
  Pack8 v[MAXV];
  float ss = 0.f;
  // Peer loads cross NVLink (~2 us each): issue ALL of them (every chunk of this thread x every rank) before the first
  // use, instead of paying the latency once per rank and chunk.
  Pack8 part[MAXV][8];
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = tid + i * 256;
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      part[i][s].u = make_uint4(0, 0, 0, 0);
      if (c < nvec && s < t.N) {
        const uint4* src = reinterpret_cast<const uint4*>(t.proj[s] + (int64_t)r * hidden) + c;
        part[i][s].u = (s == t.rank) ? *src : ld_relaxed_sys_v4(src);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = tid + i * 256;
    if (c < nvec) {
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int s = 0; s < 8; ++s)                             // fixed rank order => every rank computes the same bits
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += h2f(part[i][s].h[j]);   // (+0 for s >= N)
      Pack8 a;
      a.u = reinterpret_cast<const uint4*>(resid + (int64_t)r * hidden)[c];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const __half x = f2h(acc[j]);                        // the all-reduced projection output (fp16 like the reference's)
        v[i].h[j] = f2h(h2f(a.h[j]) + h2f(x));               // residual + x  (fp16 add)
        const float f = h2f(v[i].h[j]);
        ss += f * f;
      }
      reinterpret_cast<uint4*>(resid + (int64_t)r * hidden)[c] = v[i].u;
    }
  }
  ss = block_sum<8>(ss, red);
  const float inv = rsqrtf(ss / (float)hidden + eps);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = tid + i * 256;
    if (c < nvec) {
      Pack8 wv, o;
      wv.u = reinterpret_cast<const uint4*>(w)[c];
#pragma unroll
      for (int j = 0; j < 8; ++j) o.h[j] = f2h(h2f(wv.h[j]) * h2f(f2h(h2f(v[i].h[j]) * inv)));
      reinterpret_cast<uint4*>(out + (int64_t)r * hidden)[c] = o.u;
    }
  }
  // the last CTA to finish publishes the new epoch (all CTAs have read the old one by then)
  __syncthreads();
  if (tid == 0) {
    const uint32_t ticket = atomicAdd(t.epoch + 1, 1u);
    if (ticket == gridDim.x - 1) {
      t.epoch[1] = 0u;
      __threadfence();
      t.epoch[0] = e;
    }
  }
}


// ------------------------------------------------------------------------------------------------------------------
// Two-shot variant (reduce-scatter + all-gather in ONE kernel) for N >= 4 and / or large payloads: the one-shot kernel
// above makes every rank pull (N-1) x the payload (7 x 12.6 MB per reduction at TP-8 on the 70B shapes); here row r is
// OWNED by rank r % N: the owner's CTA pulls the N partial rows, sums them in fp32 in rank order, rounds to fp16 and
// STORES the reduced row into the `red` buffer of every rank (st.relaxed.sys over NVLink), then raises a per-row flag on
// every rank (fence + st.release.sys).  Every rank's CTA for that row waits for the flag (owners skip the wait), reads the
// reduced row from its LOCAL memory and does residual add + RMSNorm.  Per rank and reduction: (N-1)/N of the payload
// pulled + (N-1)/N pushed, independent of N.  Buffers alternate A/B exactly like the partial buffers (the epoch handshake
// of the next reduction proves every rank finished the previous one), so a row flag can never be two epochs ahead of a
// reader.  Block b of rank k handles an OWNED row first (b < owned rows), so an owner never queues behind a waiter.
struct Tp2Args {
  const __half* proj[8];     // partial GEMM outputs of rank 0..N-1
  __half* red[8];            // reduced rows, (n_max, hidden) fp16 on every rank
  uint32_t* flags[8];        // epoch flags (as in TpArgs)
  uint32_t* rowflags[8];     // rowflags[r] = per-row epoch words living on rank r (n_max words)
  uint32_t* epoch;
  int rank, N;
};

__device__ __forceinline__ void st_relaxed_sys_v4(void* p, uint4 v) {
  asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <int MAXV>
__global__ void __launch_bounds__(256) tp_allreduce2_add_rmsnorm_kernel(Tp2Args t, __half* __restrict__ resid,
                                                                        const __half* __restrict__ w,
                                                                        __half* __restrict__ out, int n, int hidden, float eps) {
  __shared__ float red[8];
  const int tid = threadIdx.x, N = t.N, rank = t.rank;
  pdl_wait();
  pdl_trigger();
  // block -> row: owned rows (rank, rank+N, ...) first, then the others in order
  const int n_own = (n - rank + N - 1) / N;                  // rows r < n with r % N == rank   (n > rank assumed below)
  int r;
  bool own;
  {
    const int b = blockIdx.x;
    const int owned = (rank < n) ? n_own : 0;
    if (b < owned) { r = rank + b * N; own = true; }
    else {
      // the (b - owned)-th row whose owner is another rank
      const int j = b - owned;
      // rows not owned: for every full group of N rows, N-1 of them; walk arithmetically
      const int g = j / (N - 1), o = j % (N - 1);
      r = g * N + (o < rank ? o : o + 1);
      own = false;
    }
  }
  const uint32_t e = t.epoch[0] + 1;
  if (blockIdx.x == 0 && tid < N && tid != rank) st_release_sys(t.flags[tid] + rank, e);      // "my partial is ready"
  const int nvec = hidden / 8;
  Pack8 v[MAXV];
  float ss = 0.f;
  if (own) {
    if (tid < N && tid != rank) {                            // every peer's partial of this epoch must be ready
      const uint32_t* f = t.flags[rank] + tid;
      const long long t0 = clock64();
      while ((int32_t)(ld_acquire_sys(f) - e) < 0) {
        if (clock64() - t0 > 4000000000LL) { atomicExch(t.epoch + 2, 1u); break; }
      }
    }
    __syncthreads();
    Pack8 part[MAXV][8];
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = tid + i * 256;
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        part[i][s].u = make_uint4(0, 0, 0, 0);
        if (c < nvec && s < N) {
          const uint4* src = reinterpret_cast<const uint4*>(t.proj[s] + (int64_t)r * hidden) + c;
          part[i][s].u = (s == rank) ? *src : ld_relaxed_sys_v4(src);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = tid + i * 256;
      if (c < nvec) {
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int s = 0; s < 8; ++s)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] += h2f(part[i][s].h[j]);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i].h[j] = f2h(acc[j]);          // the all-reduced projection output (fp16)
#pragma unroll
        for (int s = 0; s < 8; ++s)
          if (s < N && s != rank) st_relaxed_sys_v4(reinterpret_cast<uint4*>(t.red[s] + (int64_t)r * hidden) + c, v[i].u);
      }
    }
    __syncthreads();                                          // every thread's remote stores are issued ...
    if (tid < N && tid != rank) {
      __threadfence_system();                                 // ... and ordered before the flag (cumulative release)
      st_release_sys(t.rowflags[tid] + r, e);
    }
  } else {
    if (tid == 0) {
      const uint32_t* f = t.rowflags[rank] + r;
      const long long t0 = clock64();
      while ((int32_t)(ld_acquire_sys(f) - e) < 0) {
        if (clock64() - t0 > 4000000000LL) { atomicExch(t.epoch + 2, 2u); break; }
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = tid + i * 256;
      if (c < nvec) v[i].u = ld_relaxed_sys_v4(reinterpret_cast<const uint4*>(t.red[rank] + (int64_t)r * hidden) + c);
    }
  }
  // residual add + RMSNorm of row r (identical arithmetic to the one-shot kernel / sq_add_rmsnorm)
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = tid + i * 256;
    if (c < nvec) {
      Pack8 a;
      a.u = reinterpret_cast<const uint4*>(resid + (int64_t)r * hidden)[c];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[i].h[j] = f2h(h2f(a.h[j]) + h2f(v[i].h[j]));
        const float f = h2f(v[i].h[j]);
        ss += f * f;
      }
      reinterpret_cast<uint4*>(resid + (int64_t)r * hidden)[c] = v[i].u;
    }
  }
  ss = block_sum<8>(ss, red);
  const float inv = rsqrtf(ss / (float)hidden + eps);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = tid + i * 256;
    if (c < nvec) {
      Pack8 wv, o;
      wv.u = reinterpret_cast<const uint4*>(w)[c];
#pragma unroll
      for (int j = 0; j < 8; ++j) o.h[j] = f2h(h2f(wv.h[j]) * h2f(f2h(h2f(v[i].h[j]) * inv)));
      reinterpret_cast<uint4*>(out + (int64_t)r * hidden)[c] = o.u;
    }
  }
  __syncthreads();
  if (tid == 0) {
    const uint32_t ticket = atomicAdd(t.epoch + 1, 1u);
    if (ticket == gridDim.x - 1) {
      t.epoch[1] = 0u;
      __threadfence();
      t.epoch[0] = e;
    }
  }
}


// ------------------------------------------------------------------------------------------------------------------
// One-shot PUSH variant for small payloads (the 128-row verify of a 7B: 1 MB): the pull kernels above need the partial to
// be announced (flag, one NVLink trip) and then fetched (a round trip); here every rank WRITES its partial row into a
// receive slot on every peer, fences once, raises a per-(source, row) flag, and then only reads LOCAL memory.  (In-graph at
// TP-8 the two-shot kernel takes 21.7 us per reduction, 39 % of the step: profiles/r02_timeline_c2_tp8.md.)  Parity-green at
// TP-2 but NOT faster (18.5 vs 12.7 us per reduction): `__threadfence_system` between the remote stores and the flag waits
// for the stores to be acknowledged, i.e. the same round trip.  Opt-in (SQ_TP_SHOT=3); an LL-style protocol (flag packed with
// the data, no fence) would be the next step.
// Receive slots alternate with the partial buffers (A / B), so a sender can be at most one reduction ahead of a reader.
struct Tp3Args {
  const __half* proj;        // this rank's partial GEMM output (local)
  __half* recv[8];           // recv[r] = receive area on rank r: (N sources, rows_max, hidden) fp16 for THIS buffer parity
  uint32_t* pflags[8];       // pflags[r] = per-(source, row) epoch words on rank r: (N, rows_max)
  uint32_t* epoch;
  int rank, N, rows_max;
};

template <int MAXV>
__global__ void __launch_bounds__(256) tp_allreduce3_add_rmsnorm_kernel(Tp3Args t, __half* __restrict__ resid,
                                                                        const __half* __restrict__ w,
                                                                        __half* __restrict__ out, int hidden, float eps) {
  __shared__ float red[8];
  const int r = blockIdx.x, tid = threadIdx.x, N = t.N, rank = t.rank;
  pdl_wait();                                                // the row-parallel GEMM's partial is complete and visible
  pdl_trigger();
  const uint32_t e = t.epoch[0] + 1;
  const int nvec = hidden / 8;
  Pack8 mine[MAXV];
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = tid + i * 256;
    mine[i].u = make_uint4(0, 0, 0, 0);
    if (c < nvec) mine[i].u = reinterpret_cast<const uint4*>(t.proj + (int64_t)r * hidden)[c];
  }
  // phase 1: push my row into slot [rank][r] of every peer
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = tid + i * 256;
    if (c < nvec) {
#pragma unroll
      for (int s = 0; s < 8; ++s)
        if (s < N && s != rank)
          st_relaxed_sys_v4(reinterpret_cast<uint4*>(t.recv[s] + ((int64_t)rank * t.rows_max + r) * hidden) + c, mine[i].u);
    }
  }
  __syncthreads();                                            // every thread's remote stores are issued ...
  if (tid < N && tid != rank) {
    __threadfence_system();                                   // ... and ordered before the flag (cumulative release)
    st_release_sys(t.pflags[tid] + rank * t.rows_max + r, e);
    // phase 2: wait for the same row from that peer
    const uint32_t* f = t.pflags[rank] + tid * t.rows_max + r;
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys(f) - e) < 0) {
      if (clock64() - t0 > 4000000000LL) { atomicExch(t.epoch + 2, 3u); break; }
    }
  }
  __syncthreads();
  Pack8 v[MAXV];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = tid + i * 256;
    if (c < nvec) {
      Pack8 part[8];
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        part[s].u = make_uint4(0, 0, 0, 0);
        if (s < N) part[s].u = (s == rank) ? mine[i].u
                                           : ld_relaxed_sys_v4(reinterpret_cast<const uint4*>(t.recv[rank] + ((int64_t)s * t.rows_max + r) * hidden) + c);
      }
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int s = 0; s < 8; ++s)                             // fixed rank order => every rank computes the same bits
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += h2f(part[s].h[j]);
      Pack8 a;
      a.u = reinterpret_cast<const uint4*>(resid + (int64_t)r * hidden)[c];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const __half x = f2h(acc[j]);
        v[i].h[j] = f2h(h2f(a.h[j]) + h2f(x));
        const float f = h2f(v[i].h[j]);
        ss += f * f;
      }
      reinterpret_cast<uint4*>(resid + (int64_t)r * hidden)[c] = v[i].u;
    }
  }
  ss = block_sum<8>(ss, red);
  const float inv = rsqrtf(ss / (float)hidden + eps);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = tid + i * 256;
    if (c < nvec) {
      Pack8 wv, o;
      wv.u = reinterpret_cast<const uint4*>(w)[c];
#pragma unroll
      for (int j = 0; j < 8; ++j) o.h[j] = f2h(h2f(wv.h[j]) * h2f(f2h(h2f(v[i].h[j]) * inv)));
      reinterpret_cast<uint4*>(out + (int64_t)r * hidden)[c] = o.u;
    }
  }
  __syncthreads();
  if (tid == 0) {
    const uint32_t ticket = atomicAdd(t.epoch + 1, 1u);
    if (ticket == gridDim.x - 1) {
      t.epoch[1] = 0u;
      __threadfence();
      t.epoch[0] = e;
    }
  }
}


// ------------------------------------------------------------------------------------------------------------------
// LL ("low latency") two-shot for small payloads: no flags and no fences -- every 8-byte word that crosses NVLink carries
// 4 bytes of data (two halfs) and the 4-byte epoch of this reduction, written with ONE store (8-byte stores are atomic), and
// the reader simply polls the words it needs until their epoch matches (the scheme of NCCL's LL protocol).  Row r is owned
// by rank r % N:   (1) every non-owner pushes its partial row to the owner's slot [src][r / N];   (2) the owner polls the
// N-1 slots, sums in rank order in fp32, rounds to fp16 and pushes the reduced row to every peer's slot [r];   (3) the others
// poll that slot.  Two one-way NVLink trips, no round trip, 2x the bytes (fine for the 128-row verify of a 7B: 3.6 MB per rank
// at TP-8).  Slots alternate with the partial buffers (A / B); a sender can be at most one reduction ahead of any reader
// because completing a reduction needs a word from every owner.
struct TpLLArgs {
  const __half* proj;        // this rank's partial GEMM output (local)
  uint4* ll1[8];             // ll1[r] = gather area on rank r: (N sources, own_max rows, hidden/4 pairs) for this parity
  uint4* ll2[8];             // ll2[r] = reduced-row area on rank r: (rows_max, hidden/4 pairs)
  uint32_t* epoch;
  int rank, N, rows_max, own_max;
  int oneshot;               // 1: every rank pushes its row to EVERY peer's gather slot [src][r] and reduces locally (one trip;
                             //    (N-1) x 2 x payload on the wire: N <= 3); own_max == rows_max then
};

__device__ __forceinline__ uint4 ld_volatile_v4(const uint4* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint2 ll_wait(const uint4* p, uint32_t e, uint32_t* err) {
  const long long t0 = clock64();
  uint4 v = ld_volatile_v4(p);
  while (v.y != e || v.w != e) {
    if (clock64() - t0 > 4000000000LL) { atomicExch(err, 4u); break; }      // ~2 s: never hang the box
    v = ld_volatile_v4(p);
  }
  return make_uint2(v.x, v.z);
}

template <int MAXP>   // 16-byte pairs (4 halfs of payload) per thread: hidden <= 256 * 4 * MAXP
__global__ void __launch_bounds__(256) tp_allreduce_ll_add_rmsnorm_kernel(TpLLArgs t, __half* __restrict__ resid,
                                                                          const __half* __restrict__ w,
                                                                          __half* __restrict__ out, int hidden, float eps) {
  __shared__ float red[8];
  const int r = blockIdx.x, tid = threadIdx.x, N = t.N, rank = t.rank;
  pdl_wait();                                                // the row-parallel GEMM's partial is complete and visible
  pdl_trigger();
  const uint32_t e = t.epoch[0] + 1;
  const int npair = hidden / 4;
  const int owner = t.oneshot ? rank : r % N, lrow = t.oneshot ? r : r / N;
  uint2 mine[MAXP], redv[MAXP];
#pragma unroll
  for (int i = 0; i < MAXP; ++i) {
    const int p = tid + i * 256;
    mine[i] = make_uint2(0u, 0u);
    if (p < npair) mine[i] = reinterpret_cast<const uint2*>(t.proj + (int64_t)r * hidden)[p];
  }
  if (t.oneshot) {
    // one-shot: my partial row goes to every peer's gather slot [rank][r]; the reduction below polls the peers' rows
#pragma unroll
    for (int i = 0; i < MAXP; ++i) {
      const int p = tid + i * 256;
      if (p < npair) {
        const uint4 word = make_uint4(mine[i].x, e, mine[i].y, e);
        for (int s = 0; s < N; ++s)
          if (s != rank) st_relaxed_sys_v4(t.ll1[s] + ((int64_t)rank * t.own_max + lrow) * npair + p, word);
      }
    }
  }
  if (owner != rank) {
    // (1) push my partial row to the owner
    uint4* dst = t.ll1[owner] + ((int64_t)rank * t.own_max + lrow) * npair;
#pragma unroll
    for (int i = 0; i < MAXP; ++i) {
      const int p = tid + i * 256;
      if (p < npair) st_relaxed_sys_v4(dst + p, make_uint4(mine[i].x, e, mine[i].y, e));
    }
    // (3) poll the reduced row
    const uint4* src = t.ll2[rank] + (int64_t)r * npair;
#pragma unroll
    for (int i = 0; i < MAXP; ++i) {
      const int p = tid + i * 256;
      redv[i] = make_uint2(0u, 0u);
      if (p < npair) redv[i] = ll_wait(src + p, e, t.epoch + 2);
    }
  } else {
    // (2) owner: gather, reduce in rank order, scatter
#pragma unroll
    for (int i = 0; i < MAXP; ++i) {
      const int p = tid + i * 256;
      redv[i] = make_uint2(0u, 0u);
      if (p < npair) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int s = 0; s < N; ++s) {
          const uint2 v = (s == rank) ? mine[i] : ll_wait(t.ll1[rank] + ((int64_t)s * t.own_max + lrow) * npair + p, e, t.epoch + 2);
          const __half2 a = *reinterpret_cast<const __half2*>(&v.x), b = *reinterpret_cast<const __half2*>(&v.y);
          acc[0] += __low2float(a); acc[1] += __high2float(a); acc[2] += __low2float(b); acc[3] += __high2float(b);
        }
        const __half2 lo = __floats2half2_rn(acc[0], acc[1]), hi = __floats2half2_rn(acc[2], acc[3]);
        redv[i].x = *reinterpret_cast<const uint32_t*>(&lo);
        redv[i].y = *reinterpret_cast<const uint32_t*>(&hi);
        if (!t.oneshot) {
          const uint4 word = make_uint4(redv[i].x, e, redv[i].y, e);
          for (int s = 0; s < N; ++s)
            if (s != rank) st_relaxed_sys_v4(t.ll2[s] + (int64_t)r * npair + p, word);
        }
      }
    }
  }
  // residual add + RMSNorm of row r (identical arithmetic to the other variants / sq_add_rmsnorm)
  float ss = 0.f;
  uint2 v[MAXP];
#pragma unroll
  for (int i = 0; i < MAXP; ++i) {
    const int p = tid + i * 256;
    if (p < npair) {
      const uint2 a = reinterpret_cast<const uint2*>(resid + (int64_t)r * hidden)[p];
      const __half2 a0 = *reinterpret_cast<const __half2*>(&a.x), a1 = *reinterpret_cast<const __half2*>(&a.y);
      const __half2 x0 = *reinterpret_cast<const __half2*>(&redv[i].x), x1 = *reinterpret_cast<const __half2*>(&redv[i].y);
      const __half2 s0 = __floats2half2_rn(__low2float(a0) + __low2float(x0), __high2float(a0) + __high2float(x0));
      const __half2 s1 = __floats2half2_rn(__low2float(a1) + __low2float(x1), __high2float(a1) + __high2float(x1));
      const float f0 = __low2float(s0), f1 = __high2float(s0), f2 = __low2float(s1), f3 = __high2float(s1);
      ss += f0 * f0 + f1 * f1 + f2 * f2 + f3 * f3;
      v[i].x = *reinterpret_cast<const uint32_t*>(&s0);
      v[i].y = *reinterpret_cast<const uint32_t*>(&s1);
      reinterpret_cast<uint2*>(resid + (int64_t)r * hidden)[p] = v[i];
    }
  }
  ss = block_sum<8>(ss, red);
  const float inv = rsqrtf(ss / (float)hidden + eps);
#pragma unroll
  for (int i = 0; i < MAXP; ++i) {
    const int p = tid + i * 256;
    if (p < npair) {
      const uint2 wv = reinterpret_cast<const uint2*>(w)[p];
      const __half2 w0 = *reinterpret_cast<const __half2*>(&wv.x), w1 = *reinterpret_cast<const __half2*>(&wv.y);
      const __half2 s0 = *reinterpret_cast<const __half2*>(&v[i].x), s1 = *reinterpret_cast<const __half2*>(&v[i].y);
      const __half2 o0 = __floats2half2_rn(__low2float(w0) * h2f(f2h(__low2float(s0) * inv)), __high2float(w0) * h2f(f2h(__high2float(s0) * inv)));
      const __half2 o1 = __floats2half2_rn(__low2float(w1) * h2f(f2h(__low2float(s1) * inv)), __high2float(w1) * h2f(f2h(__high2float(s1) * inv)));
      uint2 o;
      o.x = *reinterpret_cast<const uint32_t*>(&o0);
      o.y = *reinterpret_cast<const uint32_t*>(&o1);
      reinterpret_cast<uint2*>(out + (int64_t)r * hidden)[p] = o;
    }
  }
  __syncthreads();
  if (tid == 0) {
    const uint32_t ticket = atomicAdd(t.epoch + 1, 1u);
    if (ticket == gridDim.x - 1) {
      t.epoch[1] = 0u;
      __threadfence();
      t.epoch[0] = e;
    }
  }
}


// ------------------------------------------------------------------------------------------------------------------
// Driver -> follower messages (tokens / position ids / state word before the target forward, accept list / state after the
// walk) as LL words through peer memory instead of five NCCL broadcast kernels per step (11.5 us each inside the graph at
// TP-8, profiles/r02_timeline_c2_tp8.md): ONE CTA on the driver writes every 4-byte word of up to three segments, paired with
// this message's epoch, into each follower's mailbox; ONE CTA on each follower polls its mailbox and scatters the words into
// its local tensors.  Mailboxes are double-buffered by epoch parity; epochs are per channel and live on the device.
struct MsgSeg { const uint32_t* src; uint32_t* dst; int words; };

__global__ void __launch_bounds__(256) tp_ll_publish_kernel(uint2* const mbox0, uint2* const mbox1, uint2* const mbox2,
                                                            uint2* const mbox3, uint2* const mbox4, uint2* const mbox5,
                                                            uint2* const mbox6, int n_peers, int cap_words, uint32_t* epoch,
                                                            MsgSeg s0, MsgSeg s1, MsgSeg s2) {
  pdl_wait();
  pdl_trigger();
  uint2* const boxes[7] = {mbox0, mbox1, mbox2, mbox3, mbox4, mbox5, mbox6};
  const uint32_t e = epoch[0] + 1;
  const int half = (int)(e & 1u) * cap_words;
  const MsgSeg segs[3] = {s0, s1, s2};
  int off = 0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    for (int i = threadIdx.x; i < segs[k].words; i += 256) {
      const uint32_t v = segs[k].src[i];
      for (int p = 0; p < n_peers; ++p)
        asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(boxes[p] + half + off + i), "r"(v), "r"(e) : "memory");
    }
    off += segs[k].words;
  }
  __syncthreads();
  if (threadIdx.x == 0) epoch[0] = e;
}

__global__ void __launch_bounds__(256) tp_ll_consume_kernel(const uint2* mbox, int cap_words, uint32_t* epoch, uint32_t* err,
                                                            MsgSeg s0, MsgSeg s1, MsgSeg s2) {
  pdl_wait();
  pdl_trigger();
  const uint32_t e = epoch[0] + 1;
  const uint2* box = mbox + (int)(e & 1u) * cap_words;
  const MsgSeg segs[3] = {s0, s1, s2};
  int off = 0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    for (int i = threadIdx.x; i < segs[k].words; i += 256) {
      const long long t0 = clock64();
      uint2 v;
      do {
        asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(box + off + i) : "memory");
        if (v.y != e && clock64() - t0 > 4000000000LL) { atomicExch(err, 5u); break; }
      } while (v.y != e);
      segs[k].dst[i] = v.x;
    }
    off += segs[k].words;
  }
  __syncthreads();
  if (threadIdx.x == 0) epoch[0] = e;
}

}  // namespace sq

using namespace sq;

extern "C" int sq_tp_alloc(void** ptr, int64_t bytes) {
  cudaError_t e = cudaMalloc(ptr, (size_t)bytes);
  if (e == cudaSuccess) e = cudaMemset(*ptr, 0, (size_t)bytes);
  if (e != cudaSuccess) { set_error("sq_tp_alloc: %s", cudaGetErrorString(e)); return SQ_ERR_CUDA; }
  return SQ_OK;
}

extern "C" int sq_tp_free(void* ptr) {
  cudaFree(ptr);
  return SQ_OK;
}

extern "C" int sq_tp_ipc_export(void* ptr, uint8_t* handle64) {
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, ptr);
  if (e != cudaSuccess) { set_error("sq_tp_ipc_export: %s", cudaGetErrorString(e)); return SQ_ERR_CUDA; }
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &h, 64);
  return SQ_OK;
}

extern "C" int sq_tp_ipc_open(const uint8_t* handle64, void** ptr) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) { set_error("sq_tp_ipc_open: %s", cudaGetErrorString(e)); return SQ_ERR_CUDA; }
  return SQ_OK;
}

extern "C" int sq_tp_ipc_close(void* ptr) {
  cudaIpcCloseMemHandle(ptr);
  return SQ_OK;
}

extern "C" int sq_tp_allreduce_add_rmsnorm(sq_half* resid, const void* const* host_proj_ptrs,
                                           void* const* host_flag_ptrs, uint32_t* epoch, int rank, int N,
                                           const sq_half* weight, sq_half* out, int n, int hidden, float eps,
                                           void* stream) {
  SQ_CHECK_ARG(N >= 2 && N <= 8 && rank >= 0 && rank < N, "sq_tp_allreduce_add_rmsnorm: bad rank/N %d/%d", rank, N);
  SQ_CHECK_ARG(hidden % 8 == 0 && hidden <= 256 * 8 * 8, "sq_tp_allreduce_add_rmsnorm: hidden=%d unsupported", hidden);
  if (n == 0) return SQ_OK;
  TpArgs t;
  for (int i = 0; i < 8; ++i) {
    t.proj[i] = i < N ? (const __half*)host_proj_ptrs[i] : nullptr;
    t.flags[i] = i < N ? (uint32_t*)host_flag_ptrs[i] : nullptr;
  }
  t.epoch = epoch;
  t.rank = rank;
  t.N = N;
  cudaStream_t st = (cudaStream_t)stream;
  const int nvec = hidden / 8;
  if (nvec <= 256) launch_k(tp_allreduce_add_rmsnorm_kernel<1>, dim3(n), dim3(256), 0, st, t, (__half*)resid, (const __half*)weight, (__half*)out, hidden, eps);
  else if (nvec <= 512) launch_k(tp_allreduce_add_rmsnorm_kernel<2>, dim3(n), dim3(256), 0, st, t, (__half*)resid, (const __half*)weight, (__half*)out, hidden, eps);
  else if (nvec <= 1024) launch_k(tp_allreduce_add_rmsnorm_kernel<4>, dim3(n), dim3(256), 0, st, t, (__half*)resid, (const __half*)weight, (__half*)out, hidden, eps);
  else launch_k(tp_allreduce_add_rmsnorm_kernel<8>, dim3(n), dim3(256), 0, st, t, (__half*)resid, (const __half*)weight, (__half*)out, hidden, eps);
  SQ_CHECK_LAUNCH("sq_tp_allreduce_add_rmsnorm");
  return SQ_OK;
}

extern "C" int sq_tp_allreduce2_add_rmsnorm(sq_half* resid, const void* const* host_proj_ptrs, void* const* host_red_ptrs,
                                            void* const* host_flag_ptrs, void* const* host_rowflag_ptrs, uint32_t* epoch,
                                            int rank, int N, const sq_half* weight, sq_half* out, int n, int hidden,
                                            float eps, void* stream) {
  SQ_CHECK_ARG(N >= 2 && N <= 8 && rank >= 0 && rank < N, "sq_tp_allreduce2_add_rmsnorm: bad rank/N %d/%d", rank, N);
  SQ_CHECK_ARG(hidden % 8 == 0 && hidden <= 256 * 8 * 8, "sq_tp_allreduce2_add_rmsnorm: hidden=%d unsupported", hidden);
  if (n == 0) return SQ_OK;
  Tp2Args t;
  for (int i = 0; i < 8; ++i) {
    t.proj[i] = i < N ? (const __half*)host_proj_ptrs[i] : nullptr;
    t.red[i] = i < N ? (__half*)host_red_ptrs[i] : nullptr;
    t.flags[i] = i < N ? (uint32_t*)host_flag_ptrs[i] : nullptr;
    t.rowflags[i] = i < N ? (uint32_t*)host_rowflag_ptrs[i] : nullptr;
  }
  t.epoch = epoch;
  t.rank = rank;
  t.N = N;
  cudaStream_t st = (cudaStream_t)stream;
  const int nvec = hidden / 8;
  __half* r = (__half*)resid; const __half* w = (const __half*)weight; __half* o = (__half*)out;
  if (nvec <= 256) launch_k(tp_allreduce2_add_rmsnorm_kernel<1>, dim3(n), dim3(256), 0, st, t, r, w, o, n, hidden, eps);
  else if (nvec <= 512) launch_k(tp_allreduce2_add_rmsnorm_kernel<2>, dim3(n), dim3(256), 0, st, t, r, w, o, n, hidden, eps);
  else if (nvec <= 1024) launch_k(tp_allreduce2_add_rmsnorm_kernel<4>, dim3(n), dim3(256), 0, st, t, r, w, o, n, hidden, eps);
  else launch_k(tp_allreduce2_add_rmsnorm_kernel<8>, dim3(n), dim3(256), 0, st, t, r, w, o, n, hidden, eps);
  SQ_CHECK_LAUNCH("sq_tp_allreduce2_add_rmsnorm");
  return SQ_OK;
}

extern "C" int sq_tp_allreduce3_add_rmsnorm(sq_half* resid, const sq_half* proj_local, void* const* host_recv_ptrs,
                                            void* const* host_pflag_ptrs, uint32_t* epoch, int rank, int N, int rows_max,
                                            const sq_half* weight, sq_half* out, int n, int hidden, float eps, void* stream) {
  SQ_CHECK_ARG(N >= 2 && N <= 8 && rank >= 0 && rank < N, "sq_tp_allreduce3_add_rmsnorm: bad rank/N %d/%d", rank, N);
  SQ_CHECK_ARG(hidden % 8 == 0 && hidden <= 256 * 8 * 8, "sq_tp_allreduce3_add_rmsnorm: hidden=%d unsupported", hidden);
  SQ_CHECK_ARG(n <= rows_max, "sq_tp_allreduce3_add_rmsnorm: n=%d exceeds the receive area (%d rows)", n, rows_max);
  if (n == 0) return SQ_OK;
  Tp3Args t;
  t.proj = (const __half*)proj_local;
  for (int i = 0; i < 8; ++i) {
    t.recv[i] = i < N ? (__half*)host_recv_ptrs[i] : nullptr;
    t.pflags[i] = i < N ? (uint32_t*)host_pflag_ptrs[i] : nullptr;
  }
  t.epoch = epoch; t.rank = rank; t.N = N; t.rows_max = rows_max;
  cudaStream_t st = (cudaStream_t)stream;
  const int nvec = hidden / 8;
  __half* r = (__half*)resid; const __half* w = (const __half*)weight; __half* o = (__half*)out;
  if (nvec <= 256) launch_k(tp_allreduce3_add_rmsnorm_kernel<1>, dim3(n), dim3(256), 0, st, t, r, w, o, hidden, eps);
  else if (nvec <= 512) launch_k(tp_allreduce3_add_rmsnorm_kernel<2>, dim3(n), dim3(256), 0, st, t, r, w, o, hidden, eps);
  else if (nvec <= 1024) launch_k(tp_allreduce3_add_rmsnorm_kernel<4>, dim3(n), dim3(256), 0, st, t, r, w, o, hidden, eps);
  else launch_k(tp_allreduce3_add_rmsnorm_kernel<8>, dim3(n), dim3(256), 0, st, t, r, w, o, hidden, eps);
  SQ_CHECK_LAUNCH("sq_tp_allreduce3_add_rmsnorm");
  return SQ_OK;
}

extern "C" int sq_tp_allreduce_ll_add_rmsnorm(sq_half* resid, const sq_half* proj_local, void* const* host_ll1_ptrs,
                                              void* const* host_ll2_ptrs, uint32_t* epoch, int rank, int N, int rows_max,
                                              int own_max, const sq_half* weight, sq_half* out, int n, int hidden, float eps,
                                              void* stream) {
  // own_max == rows_max selects the one-shot form (gather slots for every row from every source)
  const int oneshot = own_max == rows_max ? 1 : 0;
  SQ_CHECK_ARG(N >= 2 && N <= 8 && rank >= 0 && rank < N, "sq_tp_allreduce_ll_add_rmsnorm: bad rank/N %d/%d", rank, N);
  SQ_CHECK_ARG(hidden % 4 == 0 && hidden <= 256 * 4 * 8, "sq_tp_allreduce_ll_add_rmsnorm: hidden=%d unsupported", hidden);
  SQ_CHECK_ARG(n <= rows_max && (oneshot || (n + N - 1) / N <= own_max), "sq_tp_allreduce_ll_add_rmsnorm: n=%d exceeds the LL areas", n);
  if (n == 0) return SQ_OK;
  TpLLArgs t;
  t.proj = (const __half*)proj_local;
  for (int i = 0; i < 8; ++i) {
    t.ll1[i] = i < N ? (uint4*)host_ll1_ptrs[i] : nullptr;
    t.ll2[i] = i < N ? (uint4*)host_ll2_ptrs[i] : nullptr;
  }
  t.epoch = epoch; t.rank = rank; t.N = N; t.rows_max = rows_max; t.own_max = own_max; t.oneshot = oneshot;
  cudaStream_t st = (cudaStream_t)stream;
  const int npair = hidden / 4;
  __half* r = (__half*)resid; const __half* w = (const __half*)weight; __half* o = (__half*)out;
  if (npair <= 256) launch_k(tp_allreduce_ll_add_rmsnorm_kernel<1>, dim3(n), dim3(256), 0, st, t, r, w, o, hidden, eps);
  else if (npair <= 512) launch_k(tp_allreduce_ll_add_rmsnorm_kernel<2>, dim3(n), dim3(256), 0, st, t, r, w, o, hidden, eps);
  else if (npair <= 1024) launch_k(tp_allreduce_ll_add_rmsnorm_kernel<4>, dim3(n), dim3(256), 0, st, t, r, w, o, hidden, eps);
  else launch_k(tp_allreduce_ll_add_rmsnorm_kernel<8>, dim3(n), dim3(256), 0, st, t, r, w, o, hidden, eps);
  SQ_CHECK_LAUNCH("sq_tp_allreduce_ll_add_rmsnorm");
  return SQ_OK;
}

extern "C" int sq_tp_ll_publish(void* const* host_mbox_ptrs, int n_peers, int cap_words, uint32_t* epoch, const void* src0,
                                int words0, const void* src1, int words1, const void* src2, int words2, void* stream) {
  SQ_CHECK_ARG(n_peers >= 1 && n_peers <= 7 && words0 >= 0 && words1 >= 0 && words2 >= 0 && words0 + words1 + words2 <= cap_words,
               "sq_tp_ll_publish: %d peers, %d words exceed the mailbox (%d)", n_peers, words0 + words1 + words2, cap_words);
  uint2* b[7];
  for (int i = 0; i < 7; ++i) b[i] = i < n_peers ? (uint2*)host_mbox_ptrs[i] : nullptr;
  MsgSeg s0{(const uint32_t*)src0, nullptr, words0}, s1{(const uint32_t*)src1, nullptr, words1}, s2{(const uint32_t*)src2, nullptr, words2};
  launch_k(tp_ll_publish_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream, b[0], b[1], b[2], b[3], b[4], b[5], b[6], n_peers,
           cap_words, epoch, s0, s1, s2);
  SQ_CHECK_LAUNCH("sq_tp_ll_publish");
  return SQ_OK;
}

extern "C" int sq_tp_ll_consume(const void* mbox_local, int cap_words, uint32_t* epoch, uint32_t* err, void* dst0, int words0,
                                void* dst1, int words1, void* dst2, int words2, void* stream) {
  SQ_CHECK_ARG(words0 >= 0 && words1 >= 0 && words2 >= 0 && words0 + words1 + words2 <= cap_words,
               "sq_tp_ll_consume: %d words exceed the mailbox (%d)", words0 + words1 + words2, cap_words);
  MsgSeg s0{nullptr, (uint32_t*)dst0, words0}, s1{nullptr, (uint32_t*)dst1, words1}, s2{nullptr, (uint32_t*)dst2, words2};
  launch_k(tp_ll_consume_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream, (const uint2*)mbox_local, cap_words, epoch, err, s0, s1, s2);
  SQ_CHECK_LAUNCH("sq_tp_ll_consume");
  return SQ_OK;
}
