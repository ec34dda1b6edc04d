// Row helpers shared by the LayerNorm and embedding kernels: a warp owns one row of H = VPL*256 bf16 values;
// lane l holds vectors (v*32 + l), v = 0..VPL-1, of 8 consecutive elements each (coalesced 16-byte accesses).
#pragma once
#include "common.cuh"

namespace b2 {

#ifdef __CUDACC__
template <int VPL>
__device__ __forceinline__ void load_row(const __nv_bfloat16* __restrict__ row, int lane, float (&v)[VPL * 8]) {
#pragma unroll
  for (int vv = 0; vv < VPL; ++vv) {
    const uint4 t = ldg16(row + (vv * 32 + lane) * 8);
    v[vv * 8 + 0] = bf16_lo(t.x); v[vv * 8 + 1] = bf16_hi(t.x);
    v[vv * 8 + 2] = bf16_lo(t.y); v[vv * 8 + 3] = bf16_hi(t.y);
    v[vv * 8 + 4] = bf16_lo(t.z); v[vv * 8 + 5] = bf16_hi(t.z);
    v[vv * 8 + 6] = bf16_lo(t.w); v[vv * 8 + 7] = bf16_hi(t.w);
  }
}
template <int VPL>
__device__ __forceinline__ void store_row(__nv_bfloat16* __restrict__ row, int lane, const float (&v)[VPL * 8]) {
#pragma unroll
  for (int vv = 0; vv < VPL; ++vv) {
    uint4 o;
    o.x = pack_bf16(v[vv * 8 + 0], v[vv * 8 + 1]); o.y = pack_bf16(v[vv * 8 + 2], v[vv * 8 + 3]);
    o.z = pack_bf16(v[vv * 8 + 4], v[vv * 8 + 5]); o.w = pack_bf16(v[vv * 8 + 6], v[vv * 8 + 7]);
    stg16(row + (vv * 32 + lane) * 8, o);
  }
}
template <int VPL>
__device__ __forceinline__ void load_row_f32(const float* __restrict__ row, int lane, float (&v)[VPL * 8]) {
#pragma unroll
  for (int vv = 0; vv < VPL; ++vv) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(row + (vv * 32 + lane) * 8));
    const float4 b = __ldg(reinterpret_cast<const float4*>(row + (vv * 32 + lane) * 8 + 4));
    v[vv * 8 + 0] = a.x; v[vv * 8 + 1] = a.y; v[vv * 8 + 2] = a.z; v[vv * 8 + 3] = a.w;
    v[vv * 8 + 4] = b.x; v[vv * 8 + 5] = b.y; v[vv * 8 + 6] = b.z; v[vv * 8 + 7] = b.w;
  }
}
template <int VPL>
__device__ __forceinline__ void store_row_f32(float* __restrict__ row, int lane, const float (&v)[VPL * 8]) {
#pragma unroll
  for (int vv = 0; vv < VPL; ++vv) {
    *reinterpret_cast<float4*>(row + (vv * 32 + lane) * 8) =
        make_float4(v[vv * 8 + 0], v[vv * 8 + 1], v[vv * 8 + 2], v[vv * 8 + 3]);
    *reinterpret_cast<float4*>(row + (vv * 32 + lane) * 8 + 4) =
        make_float4(v[vv * 8 + 4], v[vv * 8 + 5], v[vv * 8 + 6], v[vv * 8 + 7]);
  }
}
// two-pass mean / variance in registers (biased variance, like torch.nn.LayerNorm)
template <int VPL>
__device__ __forceinline__ void row_stats(const float (&v)[VPL * 8], float eps, float& mean, float& rstd) {
  constexpr float inv = 1.0f / (VPL * 256);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPL * 8; ++i) s += v[i];
  mean = warp_sum(s) * inv;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VPL * 8; ++i) {
    const float d = v[i] - mean;
    q += d * d;
  }
  rstd = rsqrtf(warp_sum(q) * inv + eps);
}
template <int VPL>
__device__ __forceinline__ void normalize(float (&v)[VPL * 8], float mean, float rstd,
                                          const __nv_bfloat16* __restrict__ gamma,
                                          const __nv_bfloat16* __restrict__ beta, int lane) {
  float g[VPL * 8], b[VPL * 8];
  load_row<VPL>(gamma, lane, g);
  load_row<VPL>(beta, lane, b);
#pragma unroll
  for (int i = 0; i < VPL * 8; ++i) v[i] = (v[i] - mean) * rstd * g[i] + b[i];
}
template <int VPL>
__device__ __forceinline__ void normalize_store(float (&v)[VPL * 8], float mean, float rstd,
                                                const __nv_bfloat16* __restrict__ gamma,
                                                const __nv_bfloat16* __restrict__ beta, int lane,
                                                __nv_bfloat16* __restrict__ out) {
  normalize<VPL>(v, mean, rstd, gamma, beta, lane);
  store_row<VPL>(out, lane, v);
}
#endif

// host-side launchers shared with embed.cu
int32_t launch_colsum(const void* x, int64_t rows, int64_t cols, int64_t ldx, const int* filter, int filter_value,
                      void* out, float* scratch, int64_t scratch_bytes, cudaStream_t stream);
int32_t launch_layernorm_bwd(const void* dy, const void* dy_add, const void* x, const float* mean, const float* rstd,
                             const void* gamma, int64_t rows, int64_t hidden, float dropout_p, const void* rng,
                             uint32_t site, int mode, int dy_f32, int dx_f32, void* dx, void* dx_drop, void* d_gamma,
                             void* d_beta, void* d_bias, float* scratch, int64_t scratch_bytes, cudaStream_t stream,
                             int32_t* deferred_nparts = nullptr, float* accum = nullptr);

}  // namespace b2
