// Accepted-path KV compaction (Engine/Llama_KV.py:50-68).  One CTA per (layer, kv-head, K|V) plane stages every
// source row on chip (registers, 16 B per thread per row-chunk) before writing, which gives the reference's
// gather-to-temp-then-copy semantics for arbitrary (also overlapping) index lists with no extra HBM traffic:
// algorithmic bytes = 2 (K,V) * 2 (read+write) * L * Hkv * n * D * 2.
#include "sq_common.cuh"

namespace sq {

// threads: ROWS_PER_PASS rows x (D/8) 16-byte lanes.  Dynamic shared memory holds all n rows when n > ROWS_PER_PASS.
__global__ void kv_gather_kernel(__half* __restrict__ k_cache, __half* __restrict__ v_cache, int M, int D,
                                 const int32_t* __restrict__ idx, int n_host, int offset_host,
                                 const int32_t* __restrict__ state, int max_n) {
  extern __shared__ uint4 stage[];
  __half* plane = (blockIdx.y == 0 ? k_cache : v_cache) + (int64_t)blockIdx.x * M * D;
  int n = n_host, offset = offset_host;
  if (state) { n = state[ST_N_NEW]; offset = state[ST_P_OLD]; }
  if (n > max_n) n = max_n;
  const int lanes = D / 8;
  const int total = n * lanes;
  // phase 1: gather every source row chunk into shared memory
  for (int t = threadIdx.x; t < total; t += blockDim.x) {
    const int j = t / lanes, c = t % lanes;
    stage[t] = reinterpret_cast<const uint4*>(plane + (int64_t)idx[j] * D)[c];
  }
  __syncthreads();
  // phase 2: write to the compacted destination rows
  for (int t = threadIdx.x; t < total; t += blockDim.x) {
    const int j = t / lanes, c = t % lanes;
    reinterpret_cast<uint4*>(plane + (int64_t)(offset + j) * D)[c] = stage[t];
  }
}

// zero rows >= offset + n of every plane (Llama_KV.py:65-66).  grid (planes, 2, chunks)
__global__ void kv_zero_tail_kernel(__half* __restrict__ k_cache, __half* __restrict__ v_cache, int M, int D,
                                    int n_host, int offset_host, const int32_t* __restrict__ state) {
  __half* plane = (blockIdx.y == 0 ? k_cache : v_cache) + (int64_t)blockIdx.x * M * D;
  int n = n_host, offset = offset_host;
  if (state) { n = state[ST_N_NEW]; offset = state[ST_P_OLD]; }
  const int64_t first = (int64_t)(offset + n) * D / 8;
  const int64_t last = (int64_t)M * D / 8;
  uint4* p = reinterpret_cast<uint4*>(plane);
  const uint4 z = make_uint4(0, 0, 0, 0);
  for (int64_t i = first + blockIdx.z * (int64_t)blockDim.x + threadIdx.x; i < last;
       i += (int64_t)gridDim.z * blockDim.x)
    p[i] = z;
}

// Index lists too long to stage on chip (reference API: gather_kv with a whole accept list of several hundred rows):
// gather into a global scratch (planes, n, D), then copy back -- literally the reference's temp-then-copy.
// grid (planes, 2, chunks); dir 0: cache[idx[j]] -> scratch[j], dir 1: scratch[j] -> cache[offset + j]
__global__ void kv_gather_scratch_kernel(__half* __restrict__ k_cache, __half* __restrict__ v_cache, int M, int D,
                                         const int32_t* __restrict__ idx, int n, int offset, uint4* __restrict__ scratch,
                                         int dir) {
  __half* plane = (blockIdx.y == 0 ? k_cache : v_cache) + (int64_t)blockIdx.x * M * D;
  const int lanes = D / 8;
  uint4* sp = scratch + ((int64_t)blockIdx.x * 2 + blockIdx.y) * n * lanes;
  for (int t = blockIdx.z * blockDim.x + threadIdx.x; t < n * lanes; t += gridDim.z * blockDim.x) {
    const int j = t / lanes, c = t % lanes;
    if (dir == 0) sp[t] = reinterpret_cast<const uint4*>(plane + (int64_t)idx[j] * D)[c];
    else reinterpret_cast<uint4*>(plane + (int64_t)(offset + j) * D)[c] = sp[t];
  }
}

}  // namespace sq

using namespace sq;

extern "C" int64_t sq_kv_gather_scratch_bytes(int L, int Hkv, int D, int n) { return (int64_t)L * Hkv * 2 * n * D * 2; }

extern "C" int sq_kv_gather_big(sq_half* k_cache, sq_half* v_cache, int L, int Hkv, int M, int D, const int32_t* idx,
                                int n, int offset, void* scratch, int64_t scratch_bytes, int zero_tail, void* stream) {
  SQ_CHECK_ARG(D % 8 == 0 && n >= 0, "sq_kv_gather_big: bad shape");
  SQ_CHECK_ARG(n == 0 || (scratch && scratch_bytes >= sq_kv_gather_scratch_bytes(L, Hkv, D, n)),
               "sq_kv_gather_big: scratch too small");
  SQ_CHECK_ARG(offset >= 0 && offset + n <= M, "sq_kv_gather_big: offset %d + n %d > M %d", offset, n, M);
  cudaStream_t st = (cudaStream_t)stream;
  const int planes = L * Hkv;
  if (n > 0) {
    const int chunks = (n * (D / 8) + 1023) / 1024;
    for (int dir = 0; dir < 2; ++dir) {
      kv_gather_scratch_kernel<<<dim3(planes, 2, chunks), 256, 0, st>>>((__half*)k_cache, (__half*)v_cache, M, D, idx, n,
                                                                       offset, (uint4*)scratch, dir);
      SQ_CHECK_LAUNCH("sq_kv_gather_big");
    }
  }
  if (zero_tail) {
    kv_zero_tail_kernel<<<dim3(planes, 2, 4), 256, 0, st>>>((__half*)k_cache, (__half*)v_cache, M, D, n, offset, nullptr);
    SQ_CHECK_LAUNCH("sq_kv_zero_tail");
  }
  return SQ_OK;
}

extern "C" int sq_kv_gather(sq_half* k_cache, sq_half* v_cache, int L, int Hkv, int M, int D, const int32_t* idx,
                            int n, int offset, const int32_t* state, int max_n, int zero_tail, void* stream) {
  SQ_CHECK_ARG(D % 8 == 0, "sq_kv_gather: D %% 8 != 0");
  if (!state) max_n = n;
  SQ_CHECK_ARG(max_n >= 0 && (int64_t)max_n * D * 2 <= 200 * 1024, "sq_kv_gather: max_n=%d rows do not fit on chip",
               max_n);
  cudaStream_t st = (cudaStream_t)stream;
  const int planes = L * Hkv;
  if (max_n > 0) {
    const size_t smem = (size_t)max_n * D * 2;
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(kv_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) { set_error("sq_kv_gather: smem attr: %s", cudaGetErrorString(e)); return SQ_ERR_CUDA; }
    }
    int threads = max_n * (D / 8);
    threads = threads < 64 ? 64 : (threads > 512 ? 512 : ((threads + 31) / 32) * 32);
    kv_gather_kernel<<<dim3(planes, 2), threads, smem, st>>>((__half*)k_cache, (__half*)v_cache, M, D, idx, n, offset,
                                                            state, max_n);
    SQ_CHECK_LAUNCH("sq_kv_gather");
  }
  if (zero_tail) {
    kv_zero_tail_kernel<<<dim3(planes, 2, 4), 256, 0, st>>>((__half*)k_cache, (__half*)v_cache, M, D, n, offset, state);
    SQ_CHECK_LAUNCH("sq_kv_zero_tail");
  }
  return SQ_OK;
}
