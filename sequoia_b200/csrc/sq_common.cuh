// Shared helpers for the sequoia_b200 kernels (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sequoia_b200.h"

namespace sq {

void set_error(const char* fmt, ...);
void count_launch(int n);

#define SQ_CHECK_ARG(cond, ...)                     \
  do {                                              \
    if (!(cond)) {                                  \
      sq::set_error(__VA_ARGS__);                   \
      return SQ_ERR_INVALID_ARG;                    \
    }                                               \
  } while (0)

#define SQ_CHECK_LAUNCH(name)                                                   \
  do {                                                                          \
    cudaError_t e__ = cudaGetLastError();                                       \
    if (e__ != cudaSuccess) {                                                   \
      sq::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));    \
      return SQ_ERR_CUDA;                                                       \
    }                                                                           \
    sq::count_launch(1);                                                        \
  } while (0)

// Programmatic dependent launch (SQ_PDL=1): kernel N+1 is scheduled while kernel N drains; every kernel launched this
// way executes pdl_wait() FIRST (before any global access and before any early return), so completion stays transitive
// along the stream (N+1 cannot finish before N has finished and flushed), then pdl_trigger() to let N+2 be scheduled.
bool pdl_enabled();
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// Device-side decode state (int32 words), owned by the Tree object, read by every kernel
// that needs the dynamic prefix length so that whole iterations are CUDA-graph static.
enum StateWord : int {
  ST_P = 0,          // ground_truth_len: committed tokens incl. the root (bonus) token
  ST_ACCEPT_LEN = 1, // a = len(accept_list) of the last verify
  ST_TERMINAL = 2,   // 1 if the last verify hit EOS/pad or a NaN residual
  ST_N_NEW = 3,      // number of tree nodes accepted by the last verify (a - P_old)
  ST_P_OLD = 4,      // ground_truth_len the last verify started from
  ST_BONUS = 5,      // bonus token (-1 when terminal)
  ST_NAN = 6,        // residual had a NaN
  ST_SKIPPED = 7,    // prepare_for_next_iter was skipped (a + 1 > max_target_seq, or the tree would overrun the buffers)
  ST_M = 8,          // length of the tokens / position_ids buffers (max_length), written by the host once per prompt
  ST_WORDS = 16
};

__device__ __forceinline__ int row_base(const int32_t* P_ptr, int n0) {
  return (P_ptr ? (P_ptr[ST_P] - 1) : 0) + n0;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide reductions for blockDim.x = 32 * NW threads; `red` is NW floats of shared memory.
// The result is identical in every thread (fixed combination order => deterministic).
template <int NW>
__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float r = (l < NW) ? red[l] : -INFINITY;
  return warp_max(r);
}
template <int NW>
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float r = (l < NW) ? red[l] : 0.f;
  return warp_sum(r);
}

// fp16 helpers that reproduce torch's "compute in fp32, round to fp16" element-wise semantics.
__device__ __forceinline__ float h2f(__half h) { return __half2float(h); }
__device__ __forceinline__ __half f2h(float f) { return __float2half_rn(f); }
__device__ __forceinline__ float rnd16(float f) { return __half2float(__float2half_rn(f)); }

union Pack8 {
  uint4 u;
  __half h[8];
  __half2 h2[4];
};

}  // namespace sq
