// K2 -- fused RMSNorm + gather into expert-sorted order.  HBM-bound: 2 * H * 2 bytes per row.
//
// Restates RMSNorm.forward (modeling_cogvlm.py:36-41): fp32 statistics, w * x_hat in fp32, ONE
// rounding to bf16; fused with the hidden_states[padding_mask] gather (:307, :326).
// One warp per row, the whole row stays in registers (H/256 x 16-byte loads in flight per lane),
// no shared memory, no block barrier.
#include "common.cuh"

namespace vex {

constexpr int K2_WARPS = 8;

template <int NCHUNK>  // H = NCHUNK * 256
__global__ void __launch_bounds__(K2_WARPS * 32)
    k2_rmsnorm(const __nv_bfloat16* __restrict__ x, const void* __restrict__ weight, int weight_is_fp32, float eps,
               const int32_t* __restrict__ row_src, const int32_t* __restrict__ row_dst,
               const int32_t* __restrict__ n_rows_ptr, __nv_bfloat16* __restrict__ y, int rows_cap) {
  constexpr int H = NCHUNK * 256;
  const int lane = lane_id();
  const int n_rows = min(*n_rows_ptr, rows_cap);
  const int warps_total = gridDim.x * K2_WARPS;
  for (int r = blockIdx.x * K2_WARPS + (threadIdx.x >> 5); r < n_rows; r += warps_total) {
    const int src = row_src ? row_src[r] : r;
    const uint4* xp = reinterpret_cast<const uint4*>(x + static_cast<int64_t>(src) * H);
    uint4 v[NCHUNK];
#pragma unroll
    for (int i = 0; i < NCHUNK; ++i) v[i] = ld_stream(xp + i * 32 + lane);
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NCHUNK; ++i) {
      const uint32_t u[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float a = bf16_lo(u[j]), b = bf16_hi(u[j]);
        ss = fmaf(a, a, ss);
        ss = fmaf(b, b, ss);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float inv = rsqrtf(ss * (1.0f / H) + eps);
    const int dst = row_dst ? row_dst[r] : r;
    uint4* yp = reinterpret_cast<uint4*>(y + static_cast<int64_t>(dst) * H);
#pragma unroll
    for (int i = 0; i < NCHUNK; ++i) {
      const int col = (i * 32 + lane) * 8;
      float w[8];
      if (weight_is_fp32) {
        const float4* wp = reinterpret_cast<const float4*>(static_cast<const float*>(weight) + col);
        const float4 w0 = __ldg(wp), w1 = __ldg(wp + 1);
        w[0] = w0.x; w[1] = w0.y; w[2] = w0.z; w[3] = w0.w;
        w[4] = w1.x; w[5] = w1.y; w[6] = w1.z; w[7] = w1.w;
      } else {
        const uint4 wb = __ldg(reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(weight) + col));
        const uint32_t u[4] = {wb.x, wb.y, wb.z, wb.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          w[2 * j] = bf16_lo(u[j]);
          w[2 * j + 1] = bf16_hi(u[j]);
        }
      }
      const uint32_t u[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
      uint4 o;
      uint32_t* op = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        // reference order: (x * rsqrt(var + eps)) in fp32, then weight * that, then one cast
        const float a = w[2 * j] * (bf16_lo(u[j]) * inv);
        const float b = w[2 * j + 1] * (bf16_hi(u[j]) * inv);
        op[j] = pack_bf16(a, b);
      }
      st_stream(yp + i * 32 + lane, o);
    }
  }
}

}  // namespace vex

extern "C" int vex_rmsnorm_gather(const void* x, const void* weight, int weight_is_fp32, float eps,
                                  const int32_t* row_src, const int32_t* row_dst, const int32_t* n_rows, void* y,
                                  int rows_cap, int H, vexStream stream) {
  if (!x || !weight || !n_rows || !y || rows_cap <= 0) return VEX_E_INVALID;
  if (H % 256 != 0 || H <= 0 || H > 4096) return VEX_E_UNSUPPORTED;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int grid = vex::ceil_div(rows_cap, vex::K2_WARPS);
  auto xp = static_cast<const __nv_bfloat16*>(x);
  auto yp = static_cast<__nv_bfloat16*>(y);
#define VEX_K2_CASE(NC)                                                                                        \
  case NC:                                                                                                     \
    vex::k2_rmsnorm<NC><<<grid, vex::K2_WARPS * 32, 0, s>>>(xp, weight, weight_is_fp32, eps, row_src, row_dst, \
                                                             n_rows, yp, rows_cap);                                   \
    break;
  switch (H / 256) {
    VEX_K2_CASE(1) VEX_K2_CASE(2) VEX_K2_CASE(3) VEX_K2_CASE(4) VEX_K2_CASE(5) VEX_K2_CASE(6) VEX_K2_CASE(8)
    VEX_K2_CASE(16)
    default:
      return VEX_E_UNSUPPORTED;
  }
#undef VEX_K2_CASE
  VEX_LAUNCH_CHECK();
  return VEX_OK;
}
