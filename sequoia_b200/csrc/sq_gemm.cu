// Weight-streaming GEMM for the tree-verify shapes:  C[n <= 128, N] = A[n, K] * W[N, K]^T   (fp16 in, fp32 accumulate
// in tensor memory, fp16 out) -- nn.Linear of the target / draft forward (Engine/Llama_modules.py:108-110,138,270-272,
// Llama_model.py:213) for at most 128 rows.  With 128 rows every weight byte is used once: the kernel is a pure HBM
// stream (roofline = weight bytes / HBM bandwidth), so the design goal is bytes in flight, not FLOPs:
//   * one CTA per 128- or 256-wide slice of N (one wave on 148 SMs), warp-specialised: 1 TMA producer lane, 1 MMA
//     issuer lane, 4 epilogue warps; a 4-6 stage mbarrier ring of (A 128x64, W BNx64) SWIZZLE_128B tiles keeps
//     128-192 KB per SM in flight; tcgen05.mma kind::f16 M=128 N=BN K=16, accumulator in TMEM;
//   * narrow outputs (o_proj / down_proj: N = hidden = 32 slices only) are split along K over a thread-block cluster
//     (1, SPLIT, 1): each CTA streams 1/SPLIT of K, pushes its fp32 partial rows into the shared memory of the row's
//     owner CTA (st.shared::cluster), one cluster barrier, owners add in a fixed order and store fp16.
#include <cstdio>
#include <cstdlib>

#include "sq_common.cuh"
#include "sq_ptx.cuh"

struct sq_gemm_plan {
  CUtensorMap tm_a, tm_w;
  __half* c;
  int ldc, n_max, N, K, bn, split, stages, mc, pdl, tiled, epi, n_out;
  int* err_flag;
};

namespace sq {

struct GemmArgs {
  __half* c;
  int ldc, n, N, K, kb_per_split;   // kb = 64-wide K blocks handled by one CTA
  int tiled, kb_total;              // weights pre-tiled: tile (n-tile, kb) is one contiguous BN x 64 block
  int m0;                           // first activation / output row of this launch (row tiles of 128 for n > 128)
  int epi, n_out;                   // epi 1: weight rows interleave 16 gate | 16 up rows -> out = silu(gate) * up, n_out columns
  int* err_flag;
};

constexpr int G_BK = 64;
constexpr int G_THREADS = 192;       // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2..5: epilogue

__host__ __device__ constexpr int tmem_cols_for(int bn) { return bn <= 32 ? 32 : bn <= 64 ? 64 : bn <= 128 ? 128 : 256; }

template <int BN, int STAGES, int SPLIT>
struct GemmSmem {
  static constexpr int A_BYTES = 128 * 128;            // 128 rows x 64 halfs
  static constexpr int W_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_BYTES + ((W_BYTES + 1023) / 1024) * 1024;   // stages stay 1024 B aligned (SW128)
  static constexpr int OFF_BAR = STAGES * STAGE_BYTES; // full[STAGES], empty[STAGES], tmem_full, tmem ptr
  static constexpr int OFF_RED = OFF_BAR + 256;        // split-K: SPLIT slots x (128/SPLIT rows) x (BN+4) floats
  static constexpr int R_STRIDE = BN + 4;
  static constexpr int RED_BYTES = SPLIT > 1 ? 128 * R_STRIDE * 4 : 0;
  static constexpr int TOTAL = OFF_RED + RED_BYTES;
};

__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5}], [%2], %3;" ::"r"(dst),
      "l"(tm), "r"(bar), "h"(mask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
               : "memory");
}

// MC = CTAs (along N) that share one activation K-slab: each loads 128/MC of its rows and multicasts them to all.
template <int BN, int STAGES, int SPLIT, int MC>
__global__ void __launch_bounds__(G_THREADS, 1)
    gemm_tn_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w, GemmArgs g) {
  // cluster (MC, SPLIT): rank = mcr + MC * ks.  CTAs with the same ks share the activation slab (multicast group); CTAs
  // with the same mcr hold the K-splits of one output tile (DSMEM reduction group).
  using SM = GemmSmem<BN, STAGES, SPLIT>;
  constexpr int TCOLS = tmem_cols_for(BN);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.x * BN;
  const int ks = blockIdx.y;                               // K split index == rank in the (1, SPLIT) cluster
  const int mcr = (MC > 1) ? (int)(blockIdx.x % MC) : 0;   // rank in the (MC, 1) cluster
  const int kb0 = ks * g.kb_per_split;
  const int nkb = g.kb_per_split;
  const uint32_t s_base = ptx::smem_u32(smem);
  const uint32_t bar_full = s_base + SM::OFF_BAR, bar_empty = bar_full + 8 * STAGES, bar_tmem = bar_empty + 8 * STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + SM::OFF_BAR + 16 * STAGES + 8);

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(bar_full + 8 * s, 1);
      ptx::mbar_init(bar_empty + 8 * s, MC);               // every CTA that received this slot's A rows must release it
    }
    ptx::mbar_init(bar_tmem, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(ptx::smem_u32(tmem_ptr_smem), TCOLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (MC > 1)   // the peers' barriers must be initialised before any multicast / remote arrive can target them
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_ptr_smem;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
      // Weights do not depend on the previous kernel: under programmatic dependent launch the first ring of weight
      // tiles streams in while that kernel is still running; only the activation loads wait for it.
      const int pre = nkb < STAGES ? nkb : STAGES;
      for (int kb = 0; kb < pre; ++kb) {
        const uint32_t sa = s_base + kb * SM::STAGE_BYTES, sw = sa + SM::A_BYTES;
        ptx::mbar_expect_tx(bar_full + 8 * kb, SM::A_BYTES + SM::W_BYTES);
        if (g.tiled) ptx::tma_load_2d(sw, &tm_w, bar_full + 8 * kb, 0, ((int)blockIdx.x * g.kb_total + kb0 + kb) * BN);
        else ptx::tma_load_2d(sw, &tm_w, bar_full + 8 * kb, (kb0 + kb) * G_BK, n0);
      }
      asm volatile("griddepcontrol.wait;" ::: "memory");
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        const uint32_t sa = s_base + stage * SM::STAGE_BYTES, sw = sa + SM::A_BYTES;
        if (kb >= pre) {
          ptx::mbar_wait_one(bar_empty + 8 * stage, phase ^ 1, g.err_flag, 11);   // slot free in EVERY CTA of the cluster
          ptx::mbar_expect_tx(bar_full + 8 * stage, SM::A_BYTES + SM::W_BYTES);
          if (g.tiled) ptx::tma_load_2d(sw, &tm_w, bar_full + 8 * stage, 0, ((int)blockIdx.x * g.kb_total + kb0 + kb) * BN);
          else ptx::tma_load_2d(sw, &tm_w, bar_full + 8 * stage, (kb0 + kb) * G_BK, n0);
        }
        if (MC == 1) ptx::tma_load_2d(sa, &tm_a, bar_full + 8 * stage, (kb0 + kb) * G_BK, g.m0);
        else tma_load_2d_mc(sa + mcr * (SM::A_BYTES / MC), &tm_a, bar_full + 8 * stage, (kb0 + kb) * G_BK, g.m0 + mcr * (128 / MC),
                            (uint16_t)(((1u << MC) - 1u) << (MC * ks)));
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(BN, false);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        ptx::mbar_wait_one(bar_full + 8 * stage, phase, g.err_flag, 12);           // TMA bytes have landed
        ptx::tc_fence_after();
        const uint32_t sa = s_base + stage * SM::STAGE_BYTES, sw = sa + SM::A_BYTES;
#pragma unroll
        for (int k = 0; k < G_BK / 16; ++k)
          ptx::mma_ss(tmem, umma_desc(sa + k * 32, 16, 1024), umma_desc(sw + k * 32, 16, 1024), idesc, (kb | k) != 0);
        // frees the slot when the MMAs retire -- in every CTA whose multicast writes into this CTA's slot
        if (MC == 1) ptx::tc_commit(bar_empty + 8 * stage);
        else tc_commit_mc(bar_empty + 8 * stage, (uint16_t)(((1u << MC) - 1u) << (MC * ks)));
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      ptx::tc_commit(bar_tmem);                                                    // accumulator complete
    }
  } else {
    // ===== epilogue: 4 warps, thread == output row == TMEM lane =====
    const int lg = warp & 3;                                // TMEM lane group this warp may access
    const int row = lg * 32 + lane;
    ptx::mbar_wait(bar_tmem, 0, g.err_flag, 13);
    ptx::tc_fence_after();
    const uint32_t lane_base = (uint32_t)(lg * 32) << 16;
    if (SPLIT == 1 && g.epi == 1) {
      // fused SwiGLU epilogue (Engine/Llama_modules.py:272: down(act(gate(x)) * up(x))): the weight rows interleave 16 gate
      // rows with their 16 up rows, so a 32-column accumulator chunk holds both halves of 16 outputs.  Rounding points of the
      // unfused path: gate, up -> fp16; silu in fp32 -> fp16; product -> fp16.
      __half* orow = g.c + (int64_t)(g.m0 + row) * g.ldc;
#pragma unroll 1
      for (int j = 0; j < BN / 32; ++j) {
        uint32_t r[32];
        ptx::tmem_ld32(tmem + lane_base + j * 32, r);
        const int oc = (n0 + j * 32) / 2;
        if (row < g.n && oc < g.n_out) {
          Pack8 o[2];
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const float x = h2f(f2h(__uint_as_float(r[e])));
            const __half act = f2h(x / (1.0f + expf(-x)));
            o[e / 8].h[e % 8] = f2h(h2f(act) * h2f(f2h(__uint_as_float(r[16 + e]))));
          }
          *reinterpret_cast<uint4*>(orow + oc) = o[0].u;
          *reinterpret_cast<uint4*>(orow + oc + 8) = o[1].u;
        }
      }
    } else if (SPLIT == 1) {
      __half* crow = g.c + (int64_t)(g.m0 + row) * g.ldc + n0;
#pragma unroll 1
      for (int j = 0; j < BN / 32; ++j) {
        uint32_t r[32];
        ptx::tmem_ld32(tmem + lane_base + j * 32, r);
        if (row < g.n && n0 + j * 32 < g.N) {               // (ragged last tile: N is a multiple of 32)
#pragma unroll
          for (int e = 0; e < 32; e += 8) {
            Pack8 o;
#pragma unroll
            for (int q = 0; q < 8; ++q) o.h[q] = f2h(__uint_as_float(r[e + q]));
            *reinterpret_cast<uint4*>(crow + j * 32 + e) = o.u;
          }
        }
      }
    } else {
      // push this K-split's fp32 partial row to the CTA that owns the row (rows dealt in blocks of 128/SPLIT)
      constexpr int RPC = 128 / SPLIT;
      const uint32_t owner = (uint32_t)(row / RPC);
      uint32_t dst;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(dst) : "r"(s_base + SM::OFF_RED), "r"((uint32_t)mcr + MC * owner));
      dst += (uint32_t)((ks * RPC + row % RPC) * SM::R_STRIDE * 4);
#pragma unroll 1
      for (int j = 0; j < BN / 32; ++j) {
        uint32_t r[32];
        ptx::tmem_ld32(tmem + lane_base + j * 32, r);
#pragma unroll
        for (int e = 0; e < 32; e += 4)
          asm volatile("st.shared::cluster.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (j * 32 + e) * 4), "r"(r[e]),
                       "r"(r[e + 1]), "r"(r[e + 2]), "r"(r[e + 3])
                       : "memory");
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem, TCOLS);

  // (MC > 1: no CTA may leave while a peer can still multicast into its slots / arrive on its barriers;
  //  SPLIT > 1: the partial rows of every K-split must have landed in the owners' shared memory)
  if (MC > 1 || SPLIT > 1)
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");

  if (SPLIT > 1) {
    // owner CTA ks reduces rows [ks*RPC, (ks+1)*RPC): sum of the SPLIT slots in split order, fp16 out
    constexpr int RPC = 128 / SPLIT;
    constexpr int CPR = BN / 4;
    const float* red = reinterpret_cast<const float*>(smem + SM::OFF_RED);
    for (int i = tid; i < RPC * CPR; i += G_THREADS) {
      const int lr = i / CPR, cc = i % CPR;
      const int row = ks * RPC + lr;
      if (row >= g.n) continue;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int s = 0; s < SPLIT; ++s) {
        const float4 o = *reinterpret_cast<const float4*>(red + (s * RPC + lr) * SM::R_STRIDE + cc * 4);
        acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
      }
      const __half2 lo = __floats2half2_rn(acc.x, acc.y), hi = __floats2half2_rn(acc.z, acc.w);
      uint2 pk;
      pk.x = *reinterpret_cast<const uint32_t*>(&lo);
      pk.y = *reinterpret_cast<const uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(g.c + (int64_t)(g.m0 + row) * g.ldc + n0 + cc * 4) = pk;
    }
  }
}

}  // namespace sq

using namespace sq;

typedef CUresult (*PFN_tmapEncodeTiledG)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                         const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                         CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int encode_2d(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes,
                     uint32_t box_inner, uint32_t box_outer, CUtensorMapL2promotion prom) {
  static PFN_tmapEncodeTiledG fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (PFN_tmapEncodeTiledG)p;
  }
  if (!fn) { set_error("cuTensorMapEncodeTiled unavailable"); return SQ_ERR_CUDA; }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, prom, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed: %d", (int)r); return SQ_ERR_CUDA; }
  return SQ_OK;
}

// Tile selection: put a CTA on (nearly) every SM with the widest N tile that allows -- per-SM ingest, not HBM, is what
// limits a 128-row weight stream, and the activation tile every CTA re-reads is pure overhead: a wider BN and a 2-CTA
// multicast of the activation slab both raise the weight share of each SM's ingest.
static void choose_tiles(sq_gemm_plan* p, bool allow_split) {
  const int N = p->N, kb = p->K / 64;
  const int cands[] = {256, 224, 192, 160, 128, 96, 64};
  int best_bn = 128, best_split = 1, best_mc = 1;
  double best_score = -1.0;
  for (int bn : cands) {
    const int tiles = (N + bn - 1) / bn;
    for (int split : {1, 2, 4}) {
      if (kb % split || (split > 1 && !allow_split)) continue;
      if (split > 1 && (bn > 128 || N % bn)) continue;       // DSMEM reduction buffers: BN <= 128, no ragged tile
      const int ctas = tiles * split;
      if (ctas > 148) continue;
      for (int mc : {1, 2}) {
        if (mc == 2 && (tiles % 2 || (split > 1 && bn == 96))) continue;       // (combinations in sq_gemm_run's table)
        // modelled weight bandwidth ~ CTAs x weight share of the per-SM ingest; split-K pays a reduction tail
        const double a_share = 128.0 / mc, share = bn / (bn + a_share);
        double score = ctas * share;
        if (split > 1) score *= 0.85;
        if (score > best_score) { best_score = score; best_bn = bn; best_split = split; best_mc = mc; }
      }
    }
  }
  p->bn = best_bn; p->split = best_split; p->mc = best_mc;
}

static void pick_tiles(sq_gemm_plan* p) {
  const int kb = p->K / 64;
  const bool allow_split = p->epi == 0;                 // a fused epilogue needs the whole K sum in one CTA
  choose_tiles(p, allow_split);
  const char* force = getenv("SQ_GEMM_FORCE");          // tuning / debugging: "bn,split,mc"
  if (force) {
    int fb = 0, fs = 0, fm = 1;
    const int nf = sscanf(force, "%d,%d,%d", &fb, &fs, &fm);
    const bool bn_ok = fb == 64 || fb == 96 || fb == 128 || fb == 160 || fb == 192 || fb == 224 || fb == 256;
    if (nf >= 2 && bn_ok && (fs == 1 || fs == 2 || fs == 4) && kb % fs == 0 && (fm == 1 || fm == 2) &&
        !(fs > 1 && (fb > 128 || p->N % fb || !allow_split)) && !(fm == 2 && ((p->N + fb - 1) / fb) % 2)) {
      p->bn = fb; p->split = fs; p->mc = fm;
    }
  }
}

/* The tile shape a plan for (N, K) will use: callers that pre-tile the weights need BN before they build the copy. */
extern "C" int sq_gemm_pick_tiles(int N, int K, int* bn, int* split, int* mc) { return sq_gemm_pick_tiles_ex(N, K, 0, bn, split, mc); }

extern "C" int sq_gemm_pick_tiles_ex(int N, int K, int flags, int* bn, int* split, int* mc) {
  SQ_CHECK_ARG(K % 64 == 0 && N % 32 == 0, "sq_gemm_pick_tiles: K %% 64, N %% 32");
  sq_gemm_plan p{};
  p.N = N; p.K = K; p.epi = (flags & SQ_GEMM_SWIGLU) ? 1 : 0;
  pick_tiles(&p);
  *bn = p.bn; *split = p.split; *mc = p.mc;
  return SQ_OK;
}

static int plan_create(sq_gemm_plan** plan, const sq_half* a, int lda, int n_max, const sq_half* w, int N, int K, sq_half* c,
                       int ldc, int* err_flag, int flags);

extern "C" int sq_gemm_plan_create(sq_gemm_plan** plan, const sq_half* a, int lda, int n_max, const sq_half* w, int N,
                                   int K, sq_half* c, int ldc, int* err_flag) {
  return plan_create(plan, a, lda, n_max, w, N, K, c, ldc, err_flag, 0);
}

/* Same, for weights stored PRE-TILED: (ceil(N/BN), K/64, BN, 64) fp16 contiguous (rows beyond N zero), BN from
 * sq_gemm_pick_tiles -- every TMA weight load is then ONE contiguous BN*128-byte block of HBM instead of BN separate
 * 128-byte segments K*2 bytes apart. */
extern "C" int sq_gemm_plan_create_tiled(sq_gemm_plan** plan, const sq_half* a, int lda, int n_max, const sq_half* w_tiled,
                                         int N, int K, sq_half* c, int ldc, int* err_flag) {
  return plan_create(plan, a, lda, n_max, w_tiled, N, K, c, ldc, err_flag, SQ_GEMM_TILED);
}

/* flags: SQ_GEMM_TILED (weights pre-tiled as above) | SQ_GEMM_SWIGLU (fused epilogue, see sq_gemm_plan_set_epilogue; the
 * tile choice then excludes split-K) */
extern "C" int sq_gemm_plan_create_ex(sq_gemm_plan** plan, const sq_half* a, int lda, int n_max, const sq_half* w, int N,
                                      int K, sq_half* c, int ldc, int* err_flag, int flags) {
  return plan_create(plan, a, lda, n_max, w, N, K, c, ldc, err_flag, flags);
}

static int plan_create(sq_gemm_plan** plan, const sq_half* a, int lda, int n_max, const sq_half* w, int N, int K, sq_half* c,
                       int ldc, int* err_flag, int flags) {
  const int tiled = (flags & SQ_GEMM_TILED) ? 1 : 0;
  SQ_CHECK_ARG(plan && a && w && c, "sq_gemm_plan_create: null pointer");
  SQ_CHECK_ARG(K % 64 == 0 && N % 32 == 0 && lda % 8 == 0 && ldc % 8 == 0, "sq_gemm_plan_create: K %% 64, N %% 32");
  sq_gemm_plan* p = new sq_gemm_plan();
  p->c = (__half*)c; p->ldc = ldc; p->n_max = n_max; p->N = N; p->K = K; p->err_flag = err_flag;
  p->tiled = tiled; p->epi = (flags & SQ_GEMM_SWIGLU) ? 1 : 0; p->n_out = p->epi ? N / 2 : 0;
  if (p->epi) SQ_CHECK_ARG(N % 32 == 0, "sq_gemm_plan_create: SwiGLU needs N %% 32 == 0");
  pick_tiles(p);
  {
    p->pdl = pdl_enabled() ? 1 : 0;
  }
  p->stages = p->split > 1 ? 4 : (p->bn >= 224 ? 4 : p->bn >= 160 ? 5 : p->bn >= 96 ? 6 : 8);   // == the dispatch table below
  int rc = encode_2d(&p->tm_a, a, (uint64_t)K, (uint64_t)n_max, (uint64_t)lda * 2, 64, (uint32_t)(128 / p->mc),
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
  if (!rc) {
    if (tiled) rc = encode_2d(&p->tm_w, w, 64, (uint64_t)((N + p->bn - 1) / p->bn) * (uint64_t)(K / 64) * (uint64_t)p->bn, 128, 64,
                              (uint32_t)p->bn, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
    else rc = encode_2d(&p->tm_w, w, (uint64_t)K, (uint64_t)N, (uint64_t)K * 2, 64, (uint32_t)p->bn, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
  }
  if (rc) { delete p; return rc; }
  *plan = p;
  return SQ_OK;
}

extern "C" int sq_gemm_plan_destroy(sq_gemm_plan* plan) {
  delete plan;
  return SQ_OK;
}

template <int BN, int STAGES, int SPLIT, int MC>
static int launch_gemm(sq_gemm_plan* p, GemmArgs& g, cudaStream_t st) {
  using SM = GemmSmem<BN, STAGES, SPLIT>;
  constexpr int smem = SM::TOTAL + 1024;
  static_assert(smem <= 227 * 1024, "shared memory budget");
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tn_kernel<BN, STAGES, SPLIT, MC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { set_error("sq_gemm: smem attr: %s", cudaGetErrorString(e)); return SQ_ERR_CUDA; }
    attr = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((p->N + BN - 1) / BN, SPLIT, 1);
  cfg.blockDim = dim3(G_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = MC;
  at[0].val.clusterDim.y = SPLIT;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  if (p->pdl) {
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 2;
  }
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_tn_kernel<BN, STAGES, SPLIT, MC>, p->tm_a, p->tm_w, g);
  if (e != cudaSuccess) { set_error("sq_gemm: launch failed: %s", cudaGetErrorString(e)); return SQ_ERR_CUDA; }
  SQ_CHECK_LAUNCH("sq_gemm");
  return SQ_OK;
}

static int run_tile(sq_gemm_plan* plan, GemmArgs& g, cudaStream_t st) {
  const int bn = plan->bn, sp = plan->split, mc = plan->mc;
#define SQ_G(BN_, ST_, SP_, MC_) if (bn == BN_ && sp == SP_ && mc == MC_) return launch_gemm<BN_, ST_, SP_, MC_>(plan, g, st)
  SQ_G(256, 4, 1, 1); SQ_G(256, 4, 1, 2);
  SQ_G(224, 4, 1, 1); SQ_G(224, 4, 1, 2);
  SQ_G(192, 5, 1, 1); SQ_G(192, 5, 1, 2);
  SQ_G(160, 5, 1, 1); SQ_G(160, 5, 1, 2);
  SQ_G(128, 6, 1, 1); SQ_G(128, 6, 1, 2); SQ_G(128, 4, 2, 1); SQ_G(128, 4, 4, 1); SQ_G(128, 4, 2, 2); SQ_G(128, 4, 4, 2);
  SQ_G(96, 6, 1, 1); SQ_G(96, 6, 1, 2); SQ_G(96, 4, 2, 1); SQ_G(96, 4, 4, 1);
  SQ_G(64, 8, 1, 1); SQ_G(64, 8, 1, 2); SQ_G(64, 4, 2, 1); SQ_G(64, 4, 4, 1); SQ_G(64, 4, 2, 2); SQ_G(64, 4, 4, 2);
#undef SQ_G
  set_error("sq_gemm_run: no kernel for bn=%d split=%d mc=%d", bn, sp, mc);
  return SQ_ERR_UNSUPPORTED;
}

extern "C" int sq_gemm_run(sq_gemm_plan* plan, int n, void* stream) { return sq_gemm_run_at(plan, n, 0, nullptr, 0, stream); }

/* rows [a_row0, a_row0 + n) of the plan's activation buffer -> rows [0, n) of `c` (pitch ldc halfs; NULL = the plan's own
 * output buffer, rows [a_row0, ...)). */
extern "C" int sq_gemm_run_at(sq_gemm_plan* plan, int n, int a_row0, sq_half* c, int ldc, void* stream) {
  SQ_CHECK_ARG(plan != nullptr, "sq_gemm_run: null plan");
  SQ_CHECK_ARG(n >= 0 && a_row0 >= 0 && a_row0 + n <= plan->n_max, "sq_gemm_run: rows [%d, %d) exceed the plan's n_max=%d", a_row0, a_row0 + n, plan->n_max);
  SQ_CHECK_ARG(c == nullptr || (ldc % 8 == 0 && ((uintptr_t)c % 16) == 0), "sq_gemm_run: output override must be 16 B aligned, pitch %% 8");
  // more than 128 rows (prefill): one launch per 128-row tile, each streaming the weights again (once per prompt)
  for (int m0 = 0; m0 < n; m0 += 128) {
    GemmArgs g;
    g.n = n - m0 < 128 ? n - m0 : 128; g.N = plan->N; g.K = plan->K;
    g.kb_per_split = plan->K / 64 / plan->split;
    g.tiled = plan->tiled; g.kb_total = plan->K / 64;
    g.m0 = a_row0 + m0; g.epi = plan->epi; g.n_out = plan->n_out;
    // the kernel addresses output row (g.m0 + row): bias the base so that activation row a_row0 lands on output row 0
    if (c) { g.ldc = ldc; g.c = (__half*)c - (int64_t)a_row0 * ldc; }
    else { g.ldc = plan->ldc; g.c = plan->c; }
    g.err_flag = plan->err_flag;
    const int rc = run_tile(plan, g, (cudaStream_t)stream);
    if (rc != SQ_OK) return rc;
  }
  return SQ_OK;
}

/* Fused epilogue.  kind 0: C = A W^T (default).  kind 1 (SwiGLU): the weight rows interleave 16 gate rows / 16 up rows
 * (row 32b + t = gate[16b + t], row 32b + 16 + t = up[16b + t]); C (n, n_out = N/2) = silu(gate) * up.  Split-K plans
 * cannot fuse it. */
extern "C" int sq_gemm_plan_set_epilogue(sq_gemm_plan* plan, int kind, int n_out) {
  SQ_CHECK_ARG(plan != nullptr && (kind == 0 || kind == 1), "sq_gemm_plan_set_epilogue: bad arguments");
  SQ_CHECK_ARG(kind == 0 || (plan->split == 1 && n_out % 16 == 0 && 2 * n_out <= plan->N + 31), "sq_gemm_plan_set_epilogue: SwiGLU needs split 1, n_out %% 16 == 0");
  plan->epi = kind; plan->n_out = n_out;
  return SQ_OK;
}

extern "C" int sq_gemm_plan_info(sq_gemm_plan* plan, int* bn, int* split, int* stages) {
  *bn = plan->bn; *split = plan->split; *stages = plan->stages + 100 * plan->mc;
  return SQ_OK;
}
