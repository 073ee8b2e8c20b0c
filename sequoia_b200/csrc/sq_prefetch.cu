// L2 prefetch of upcoming weight panels.  A decode step is a chain  GEMM -> small kernels -> GEMM ...: HBM is idle while
// the small (latency-bound) kernels of a layer run (~30 us per 7B layer), and the next GEMM then waits for HBM again.
// This kernel, launched on a forked stream next to those small kernels, issues fire-and-forget bulk prefetches of the
// first K-columns of the next weight matrix into the 126 MB L2, so the GEMM's first tiles hit L2 (LTS cap ~12 TB/s)
// instead of HBM.  A prefetch is only a hint: it cannot change any result.
#include "sq_common.cuh"

namespace sq {

// One thread per (row, <= 4 KB chunk) of the panel  base[r*pitch + off .. + seg).
__global__ void l2_prefetch_kernel(const char* __restrict__ base, int64_t pitch, int rows, int64_t off, int64_t seg,
                                   int chunks_per_row) {
  const int64_t total = (int64_t)rows * chunks_per_row;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / chunks_per_row), c = (int)(i % chunks_per_row);
    const int64_t b0 = (int64_t)c * 4096;
    const int64_t len = (seg - b0 < 4096) ? (seg - b0) : 4096;
    const char* p = base + (int64_t)r * pitch + off + b0;
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"((uint32_t)len) : "memory");
  }
}

}  // namespace sq

using namespace sq;

extern "C" int sq_l2_prefetch(const void* base, int64_t pitch_bytes, int rows, int64_t off_bytes, int64_t seg_bytes,
                              void* stream) {
  SQ_CHECK_ARG(base != nullptr && rows >= 0 && seg_bytes >= 0 && off_bytes >= 0, "sq_l2_prefetch: bad arguments");
  SQ_CHECK_ARG(off_bytes + seg_bytes <= pitch_bytes || rows <= 1, "sq_l2_prefetch: segment exceeds the row pitch");
  SQ_CHECK_ARG(((reinterpret_cast<uintptr_t>(base) | (uintptr_t)pitch_bytes | (uintptr_t)off_bytes |
                 (uintptr_t)seg_bytes) & 15u) == 0, "sq_l2_prefetch: base / pitch / offset / size must be multiples of 16 B");
  if (rows == 0 || seg_bytes == 0) return SQ_OK;
  const int cpr = (int)((seg_bytes + 4095) / 4096);
  const int64_t total = (int64_t)rows * cpr;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 592) blocks = 592;                       // 4 x 148
  l2_prefetch_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const char*)base, pitch_bytes, rows, off_bytes, seg_bytes,
                                                              cpr);
  SQ_CHECK_LAUNCH("sq_l2_prefetch");
  return SQ_OK;
}
