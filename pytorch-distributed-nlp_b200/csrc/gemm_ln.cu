// Dense + bias + dropout + residual + LayerNorm in ONE kernel: BertSelfOutput.forward / BertOutput.forward
// (SP/transformers/models/bert/modeling_bert.py:294-298, :352-356: `LayerNorm(dropout(dense(h)) + input)`).
//
// Why a kernel of its own.  The two N = hidden GEMMs of a layer are small (4.8 / 19 GFLOP at BERT-base, batch 32 x 128)
// and were followed by a separate LayerNorm launch each: per layer 2 x (~5 us kernel + ~2 us launch gap + a 12.6 MB
// round trip of the pre-LN tensor), and with 256-wide tiles an N = 768 problem occupies only 48 CTA pairs = 96 of
// the 148 SMs.  LayerNorm needs full rows, and a row (768 / 1024 fp32 accumulators) does not fit one CTA's 512 TMEM
// columns -- so the row is spread over a CLUSTER: 8 CTAs = 4 CTA pairs (tcgen05 cta_group::2), pair p owns the
// 256 x BN tile of columns [p BN, (p+1) BN) with BN = N / 4 (192 for BERT-base, 256 for BERT-large), all four pairs
// share the same 256 rows.  16 clusters x 8 = 128 SMs for batch 32 x 128.
//
// Per CTA (128 rows x BN columns, accumulator in TMEM):
//   warp 0   TMA producer (this CTA's 128 A rows + its half of the pair's B tile; completion credited to the pair
//            leader's mbarrier)          warp 1   MMA issuer (pair leader only)         warp 2   TMEM allocator
//   warps 4-19  epilogue, warp -> (TMEM lane quarter, column group of BN/4):
//     pass 1  z = bf16(acc + bias -> dropout -> + residual) into the warp's 4 KB staging tile (residual tile
//             prefetched by cp.async under the mainloop), stored to D with coalesced 128-byte row segments;
//             per-row partial statistics over the warp's BN/4 columns (mean, M2 = sum (z - mean)^2, of the ROUNDED z)
//             written into the `stats` pad of the four CTAs that hold the same rows -- distributed shared memory
//     cluster barrier (arrive.release / wait.acquire)
//     pass 2  every thread merges the 16 partials of its row (Chan's parallel-variance formula: exact two-pass
//             statistics, no E[x^2] - mean^2 cancellation), normalises its staged z in place, stores y; mean / rstd
//             (fp32, what the LayerNorm backward reads) are written by the first pair's first column group
// Numerics: identical inputs and the same fp32 operations as the unfused pair of kernels up to the summation order of
// the row statistics.
#include "common.cuh"
#include "gemm_epilogue.cuh"
#include "pair.cuh"
#include "../../include/b2_ddp_bert.h"

namespace b2 {

constexpr int kLnBM = 128, kLnBK = 64, kLnUmmaK = 16;
constexpr int kLnEW = 16;                        // epilogue warps
constexpr int kLnThreads = (4 + kLnEW) * 32;     // 640
// (a value-dependent argument, as in gemm.cu: nvcc rejects __launch_bounds__ next to __maxnreg__ only when the bound is
// a non-dependent constant)
template <int BN> constexpr int gemm_ln_threads() { return BN > 0 ? kLnThreads : 0; }
constexpr int kLnPairs = 4;                      // column tiles (CTA pairs) per cluster
constexpr int kLnCluster = 2 * kLnPairs;         // 8 CTAs

template <int BN>
struct GemmLnCfg {
  static constexpr int kABytes = kLnBM * kLnBK * 2;          // this CTA's 128 A rows
  static constexpr int kBBytes = (BN / 2) * kLnBK * 2;       // this CTA's half of the pair's B tile
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN == 256) ? 4 : 5;
  static constexpr int kTmemCols = 256;
  static constexpr int kPipeBytes = kStages * kStageBytes;
  static constexpr int kStatsBytes = 4 * kLnPairs * kLnBM * 8;   // [pair][column group][row] float2
  static constexpr int kSmemBytes = kPipeBytes + kLnEW * kEpiStageBytes + kStatsBytes + 1024 /*align*/ + 256;
};

struct GemmLnParams {
  int M, N, kblocks;
  __nv_bfloat16* D; long long ldd;                 // pre-LayerNorm sum z, bf16 [M, N] (kept for the backward)
  const __nv_bfloat16* bias;
  const __nv_bfloat16* resid; long long ld_resid;  // residual input, bf16 [M, N]
  float dropout_p; const unsigned long long* rng; unsigned rng_site;
  const __nv_bfloat16* gamma; const __nv_bfloat16* beta; float eps;
  __nv_bfloat16* Y; long long ldy;                 // LayerNorm output, bf16 [M, N]
  float* mean; float* rstd;                        // fp32 [M]
};

// [32 rows x 128 B] staging-tile copies that touch only the first USED 16-byte chunks of every row
template <int USED>
__device__ __forceinline__ void tile8_g2s_async(uint8_t* stage, const uint8_t* g, long long pitch, int lane,
                                                int rows_valid) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int r = 4 * k + (lane >> 3), c = lane & 7;
    if (c < USED) {
      uint4* dst = stage_ptr<8>(stage, r, c);
      if (r < rows_valid) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)),
                     "l"(g + (size_t)r * pitch + c * 16)
                     : "memory");
      } else {
        *dst = make_uint4(0u, 0u, 0u, 0u);
      }
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int USED>
__device__ __forceinline__ void tile8_s2g(uint8_t* stage, uint8_t* g, long long pitch, int lane, int rows_valid) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int r = 4 * k + (lane >> 3), c = lane & 7;
    if (c < USED && r < rows_valid)
      *reinterpret_cast<uint4*>(g + (size_t)r * pitch + c * 16) = *stage_ptr<8>(stage, r, c);
  }
}
__device__ __forceinline__ void unpack8(const uint4& t, float (&f)[8]) {
  f[0] = bf16_lo(t.x); f[1] = bf16_hi(t.x); f[2] = bf16_lo(t.y); f[3] = bf16_hi(t.y);
  f[4] = bf16_lo(t.z); f[5] = bf16_hi(t.z); f[6] = bf16_lo(t.w); f[7] = bf16_hi(t.w);
}

template <int BN>
__global__ void __cluster_dims__(kLnCluster, 1, 1) __launch_bounds__(gemm_ln_threads<BN>()) __maxnreg__(BN > 0 ? 96 : 64)
gemm_ln_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const GemmLnParams p) {
  using Cfg = GemmLnCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi_stage = smem + Cfg::kPipeBytes;
  float2* stats = reinterpret_cast<float2*>(epi_stage + kLnEW * kEpiStageBytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(stats) + Cfg::kStatsBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tmem_full = empty_bar + Cfg::kStages;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();            // 0..7
  const uint32_t parity = rank & 1u;                  // which 128 of the pair's 256 rows
  const uint32_t lead_rank = rank & ~1u;              // the pair's leader CTA
  const bool leader = parity == 0;
  const int pair = (int)(rank >> 1);                  // column tile of this pair
  const int row_block = blockIdx.x / kLnCluster;
  const int m0 = row_block * (2 * kLnBM) + (int)parity * kLnBM;
  const int n_tile = pair * BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc_2sm(tmem_holder, Cfg::kTmemCols);
  tc_fence_before();
  cluster_sync_all();   // barrier inits + TMEM allocation visible cluster-wide; every CTA of the cluster is running
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  // PDL: nothing above touches global data
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    // ------------------------------ TMA producer (every CTA) ------------------------------
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      const int nb = n_tile + (int)parity * (BN / 2);
      for (int kb = 0; kb < p.kblocks; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (leader) mbar_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
        const uint32_t bar = mapa_u32(smem_u32(&full_bar[stage]), lead_rank);
        uint8_t* sa = smem + stage * Cfg::kStageBytes;
        uint8_t* sb = sa + Cfg::kABytes;
        tma_load_2d_2sm(sa, &tmap_a, bar, kb * kLnBK, m0);     // box {64 k, 128 rows}
        tma_load_2d_2sm(sb, &tmap_b, bar, kb * kLnBK, nb);     // box {64 k, BN/2 rows}
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (pair leaders) ------------------------------
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * kLnBM, BN, false, false);
      const uint16_t pair_mask = (uint16_t)(3u << lead_rank);
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < p.kblocks; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
        const uint32_t sb = sa + Cfg::kABytes;
#pragma unroll
        for (int k = 0; k < kLnBK / kLnUmmaK; ++k)
          umma_bf16_2sm(tmem_base, make_smem_desc(sa + k * kLnUmmaK * 2, 16, 1024),
                        make_smem_desc(sb + k * kLnUmmaK * 2, 16, 1024), idesc, (kb > 0 || k > 0) ? 1u : 0u);
        umma_commit_2sm(&empty_bar[stage], pair_mask);
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
      umma_commit_2sm(tmem_full, pair_mask);
    }
  } else if (warp >= 4) {
    // ------------------------------ epilogue ------------------------------
    constexpr int CW = BN / 4;          // columns per warp: 48 / 64
    constexpr int NCH = CW / 8;         // 16-byte chunks per staged row: 6 / 8
    const DropCtx drop = make_drop_ctx(p.rng, p.rng_site, p.dropout_p);
    uint8_t* stage = epi_stage + (warp - 4) * kEpiStageBytes;
    const int q = warp & 3, cg = (warp - 4) >> 2;
    const int row_l = q * 32 + lane;                  // row inside this CTA's 128
    const int row0 = m0 + q * 32;
    const int rows_valid = p.M - row0;                // <= 0: nothing of this warp is stored
    const int m = row0 + lane;
    const int nw = n_tile + cg * CW;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cg * CW);

    // residual tile: independent of the accumulator, fetched under the mainloop
    tile8_g2s_async<NCH>(stage, reinterpret_cast<const uint8_t*>(p.resid + (size_t)row0 * p.ld_resid + nw),
                         p.ld_resid * 2, lane, rows_valid);
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    float sum = 0.f;
#pragma unroll 1
    for (int j = 0; j < CW / 16; ++j) {
      uint32_t v[16];
      tmem_ld16(taddr + j * 16, v);
      tmem_ld_wait();
      if (j == 0) {
        tile_async_wait();
        __syncwarp();
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int col = nw + j * 16 + c * 8;
        float f[8], r[8];
        unpack8(ldg16(p.bias + col), r);
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(v[c * 8 + i]) + r[i];
        uint4* sp = stage_ptr<8>(stage, lane, 2 * j + c);
        unpack8(*sp, r);                                   // residual, own row
        const uint32_t keep = dropout_keep8(drop, (unsigned long long)m * p.N + col);
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = (((keep >> i) & 1u) ? f[i] * drop.scale : 0.f) + r[i];
        uint4 o;
        o.x = pack_bf16_round(f[0], f[1]); o.y = pack_bf16_round(f[2], f[3]);
        o.z = pack_bf16_round(f[4], f[5]); o.w = pack_bf16_round(f[6], f[7]);
        *sp = o;                                           // z (bf16), in place
#pragma unroll
        for (int i = 0; i < 8; ++i) sum += f[i];
      }
    }
    // partial statistics of the ROUNDED z over this warp's CW columns: two passes over the staged row
    const float mean_l = sum * (1.0f / CW);
    float m2_l = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      float f[8];
      unpack8(*stage_ptr<8>(stage, lane, c), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float d = f[i] - mean_l;
        m2_l = fmaf(d, d, m2_l);
      }
    }
    {
      const uint32_t slot = smem_u32(&stats[(pair * 4 + cg) * kLnBM + row_l]);
#pragma unroll
      for (int dp = 0; dp < kLnPairs; ++dp) st_cluster_f32x2(mapa_u32(slot, (uint32_t)(2 * dp) + parity), mean_l, m2_l);
    }
    __syncwarp();
    tile8_s2g<NCH>(stage, reinterpret_cast<uint8_t*>(p.D + (size_t)row0 * p.ldd + nw), p.ldd * 2, lane, rows_valid);
    __syncwarp();
    cluster_arrive_release();
    cluster_wait_acquire();
    // merge the 16 partials of this row (equal counts CW): mean = avg(mean_i), M2 = sum M2_i + CW sum (mean_i - mean)^2
    float mean = 0.f;
#pragma unroll
    for (int i = 0; i < 4 * kLnPairs; ++i) mean += stats[i * kLnBM + row_l].x;
    mean *= 1.0f / (4 * kLnPairs);
    float m2 = 0.f;
#pragma unroll
    for (int i = 0; i < 4 * kLnPairs; ++i) {
      const float2 t = stats[i * kLnBM + row_l];
      const float d = t.x - mean;
      m2 += t.y + (float)CW * d * d;
    }
    const float rstd = rsqrtf(m2 / (float)p.N + p.eps);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int col = nw + c * 8;
      float f[8], g[8], b[8];
      uint4* sp = stage_ptr<8>(stage, lane, c);
      unpack8(*sp, f);
      unpack8(ldg16(p.gamma + col), g);
      unpack8(ldg16(p.beta + col), b);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = (f[i] - mean) * rstd * g[i] + b[i];
      uint4 o;
      o.x = pack_bf16(f[0], f[1]); o.y = pack_bf16(f[2], f[3]); o.z = pack_bf16(f[4], f[5]); o.w = pack_bf16(f[6], f[7]);
      *sp = o;
    }
    __syncwarp();
    tile8_s2g<NCH>(stage, reinterpret_cast<uint8_t*>(p.Y + (size_t)row0 * p.ldy + nw), p.ldy * 2, lane, rows_valid);
    if (pair == 0 && cg == 0 && lane < rows_valid) {
      p.mean[m] = mean;
      p.rstd[m] = rstd;
    }
  }
  if (warp < 4) {
    // the statistics barrier counts every thread of the cluster
    __syncwarp();
    cluster_arrive_release();
    cluster_wait_acquire();
  }

  // no CTA may exit (or free TMEM) while a partner can still multicast-commit into it or write its stats pad
  tc_fence_before();
  __syncwarp();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, Cfg::kTmemCols);
  }
}

template <int BN>
static int32_t gemm_ln_prepare() {
  static int state = 0;    // 0 = not yet, 1 = ready, -1 = failed
  if (state == 0) {
    auto kern = gemm_ln_kernel<BN>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmLnCfg<BN>::kSmemBytes) !=
        cudaSuccess) {
      (void)cudaGetLastError();
      state = -1;
    } else {
      state = 1;
    }
  }
  return state;
}

template <int BN>
static int32_t gemm_ln_max_clusters() {
  if (gemm_ln_prepare<BN>() != 1) return 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(kLnCluster);
  cfg.blockDim = dim3(kLnThreads);
  cfg.dynamicSmemBytes = GemmLnCfg<BN>::kSmemBytes;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kLnCluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, gemm_ln_kernel<BN>, &cfg) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  return n;
}

template <int BN>
static int32_t launch_gemm_ln(const b2_gemm_args_t& a, const void* gamma, const void* beta, float eps, void* y,
                              int64_t ldy, float* mean, float* rstd, cudaStream_t stream) {
  using Cfg = GemmLnCfg<BN>;
  B2_REQUIRE(gemm_ln_prepare<BN>() == 1, "b2_gemm_ln_fwd: cannot reserve %d bytes of shared memory", Cfg::kSmemBytes);
  CUtensorMap ta, tb;
  int32_t st = get_tensor_map_2d(&ta, a.A, (uint64_t)a.M, (uint64_t)a.K, (uint64_t)a.lda * 2, kLnBM, 64);
  if (st) return st;
  st = get_tensor_map_2d(&tb, a.B, (uint64_t)a.N, (uint64_t)a.K, (uint64_t)a.ldb * 2, BN / 2, 64);
  if (st) return st;
  GemmLnParams p;
  p.M = (int)a.M; p.N = (int)a.N; p.kblocks = (int)((a.K + kLnBK - 1) / kLnBK);
  p.D = (__nv_bfloat16*)a.D; p.ldd = a.ldd;
  p.bias = (const __nv_bfloat16*)a.bias;
  p.resid = (const __nv_bfloat16*)a.aux_in; p.ld_resid = a.ld_aux_in;
  p.dropout_p = a.dropout_p; p.rng = (const unsigned long long*)a.rng_state; p.rng_site = a.rng_site;
  p.gamma = (const __nv_bfloat16*)gamma; p.beta = (const __nv_bfloat16*)beta; p.eps = eps;
  p.Y = (__nv_bfloat16*)y; p.ldy = ldy; p.mean = mean; p.rstd = rstd;
  const int row_blocks = (int)((a.M + 2 * kLnBM - 1) / (2 * kLnBM));
  B2_LAUNCH(gemm_ln_kernel<BN>, kLnCluster * row_blocks, kLnThreads, Cfg::kSmemBytes, stream, ta, tb, p);
  B2_CUDA(cudaGetLastError());
  count_launches(1);
  return 0;
}

}  // namespace b2

using namespace b2;

extern "C" int32_t b2_gemm_ln_max_clusters(int64_t hidden) {
  if (hidden == 768) return gemm_ln_max_clusters<192>();
  if (hidden == 1024) return gemm_ln_max_clusters<256>();
  return 0;
}

extern "C" int32_t b2_gemm_ln_fwd(const b2_gemm_args_t* a, const void* gamma, const void* beta, float eps, void* y,
                                  int64_t ldy, float* mean, float* rstd, void* stream_) {
  B2_REQUIRE(a != nullptr, "b2_gemm_ln_fwd: null args");
  B2_REQUIRE(a->M > 0 && a->K > 0, "b2_gemm_ln_fwd: empty problem");
  B2_REQUIRE(a->N == 768 || a->N == 1024, "b2_gemm_ln_fwd: N=%lld: the row cluster covers hidden sizes 768 and 1024",
             (long long)a->N);
  B2_REQUIRE(a->a_major == B2_MAJOR_K && a->b_major == B2_MAJOR_K, "b2_gemm_ln_fwd: NT layout only (y = x W^T)");
  B2_REQUIRE(a->epilogue == B2_EPI_BIAS_DROPOUT_RESIDUAL, "b2_gemm_ln_fwd: epilogue must be BIAS_DROPOUT_RESIDUAL");
  B2_REQUIRE(a->A && a->B && a->D && a->bias && a->aux_in && gamma && beta && y && mean && rstd,
             "b2_gemm_ln_fwd: null pointer");
  B2_REQUIRE(a->K % 8 == 0 && a->lda % 8 == 0 && a->ldb % 8 == 0 && a->ldd % 8 == 0 && a->ld_aux_in % 8 == 0 &&
                 ldy % 8 == 0,
             "b2_gemm_ln_fwd: K and leading dimensions must be multiples of 8 elements (16 B)");
  B2_REQUIRE(((uintptr_t)a->A % 16 == 0) && ((uintptr_t)a->B % 16 == 0) && ((uintptr_t)a->D % 16 == 0) &&
                 ((uintptr_t)a->aux_in % 16 == 0) && ((uintptr_t)y % 16 == 0) && ((uintptr_t)a->bias % 16 == 0) &&
                 ((uintptr_t)gamma % 16 == 0) && ((uintptr_t)beta % 16 == 0),
             "b2_gemm_ln_fwd: operands must be 16-byte aligned");
  B2_REQUIRE(a->dropout_p >= 0.f && a->dropout_p < 1.f, "b2_gemm_ln_fwd: dropout_p out of range");
  B2_REQUIRE(!(a->dropout_p > 0.f) || a->rng_state != nullptr, "b2_gemm_ln_fwd: dropout needs rng_state");
  if (a->N == 768)
    return launch_gemm_ln<192>(*a, gamma, beta, eps, y, ldy, mean, rstd, (cudaStream_t)stream_);
  return launch_gemm_ln<256>(*a, gamma, beta, eps, y, ldy, mean, rstd, (cudaStream_t)stream_);
}
