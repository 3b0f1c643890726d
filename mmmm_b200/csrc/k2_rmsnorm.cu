// K2 -- fused RMSNorm + gather into expert-sorted order.  HBM-bound: 2 * H * 2 bytes per row.
//
// Restates RMSNorm.forward (modeling_cogvlm.py:36-41): fp32 statistics, w * x_hat in fp32, ONE
// rounding to bf16; fused with the hidden_states[padding_mask] gather (:307, :326).
// One warp per row, the whole row stays in registers (H/256 x 16-byte loads in flight per lane),
// no shared memory, no block barrier.
#include <algorithm>

#include "common.cuh"

namespace vex {

constexpr int K2_WARPS = 8;

template <int NCHUNK>  // H = NCHUNK * 256
__global__ void __launch_bounds__(K2_WARPS * 32, 2)
    k2_rmsnorm(const __nv_bfloat16* __restrict__ x, const void* __restrict__ weight, int weight_is_fp32, float eps,
               const int32_t* __restrict__ row_src, const int32_t* __restrict__ row_dst,
               const int32_t* __restrict__ n_rows_ptr, __nv_bfloat16* __restrict__ y, int rows_cap) {
  constexpr int H = NCHUNK * 256;
  // WPR warps share one row so that a lane holds at most 8 x 16 bytes of it (32 registers).  The NEXT row's slice is
  // loaded into a second register set before the current row is reduced and stored (software pipelining): a warp has
  // loads in flight during its shuffle / barrier / store phase too, and the dependent row-index -> row-data latency of
  // the gather is hidden behind the previous row (round 1: 0.70 of the HBM roofline at c2, loads only in flight during
  // ~40 % of a warp's iteration).
  constexpr int WPR = NCHUNK > 8 ? 2 : 1;   // warps per row
  constexpr int CPL = NCHUNK / WPR;         // 16-byte chunks per lane
  constexpr int ROWS = K2_WARPS / WPR;      // rows per CTA iteration
  static_assert(NCHUNK % WPR == 0, "row slices must be whole chunks");
  // weight vector staged once per CTA in shared memory (fp32): short-latency LDS in the per-row epilogue
  __shared__ __align__(16) float w_s[H];
  __shared__ float part[K2_WARPS];
  const int lane = lane_id();
  const int warp = threadIdx.x >> 5;
  const int slot = warp / WPR, half = warp % WPR;
  pdl_trigger();
  // Weight staging with every load issued up front (16-byte vectors, fully unrolled).  The scalar one-element-per-trip
  // loop this replaces was a chain of up to 16 dependent-latency round trips (~12 us: the whole duration of the
  // 8-row decode launch, and a quarter of the 42 us prefill launch at c2 -- ncu, profiles/r2_decode_ncu.md).  The weight
  // is a constant: its loads are issued BEFORE the dependency wait of a programmatic dependent launch.
  constexpr int NVF = (H / 4 + K2_WARPS * 32 - 1) / (K2_WARPS * 32);  // float4 per thread
  constexpr int NVB = (H / 8 + K2_WARPS * 32 - 1) / (K2_WARPS * 32);  // uint4 (8 bf16) per thread
  float4 wf[NVF];
  uint4 wb[NVB];
  if (weight_is_fp32) {
#pragma unroll
    for (int j = 0; j < NVF; ++j) {
      const int i = threadIdx.x + j * K2_WARPS * 32;
      if (i < H / 4) wf[j] = __ldg(static_cast<const float4*>(weight) + i);
    }
  } else {
#pragma unroll
    for (int j = 0; j < NVB; ++j) {
      const int i = threadIdx.x + j * K2_WARPS * 32;
      if (i < H / 8) wb[j] = __ldg(static_cast<const uint4*>(weight) + i);
    }
  }
  pdl_wait();  // rows, row maps and the row count come from earlier kernels
  const int n_rows = min(*n_rows_ptr, rows_cap);
  const int stride = gridDim.x * ROWS;

  auto load_row = [&](int r, uint4 (&dst)[CPL]) {
    const int src = row_src ? row_src[r] : r;
    const uint4* xp = reinterpret_cast<const uint4*>(x + static_cast<int64_t>(src) * H) + half * CPL * 32;
#pragma unroll
    for (int i = 0; i < CPL; ++i) dst[i] = ld_coherent_stream(xp + i * 32 + lane);
  };

  uint4 nx[CPL];
  {
    const int r = blockIdx.x * ROWS + slot;
    if (r < n_rows) load_row(r, nx);  // the first row is in flight while the weight vector is staged
  }
  if (weight_is_fp32) {
#pragma unroll
    for (int j = 0; j < NVF; ++j) {
      const int i = threadIdx.x + j * K2_WARPS * 32;
      if (i < H / 4) *reinterpret_cast<float4*>(&w_s[4 * i]) = wf[j];
    }
  } else {
#pragma unroll
    for (int j = 0; j < NVB; ++j) {
      const int i = threadIdx.x + j * K2_WARPS * 32;
      if (i < H / 8) {
        *reinterpret_cast<float4*>(&w_s[8 * i]) =
            make_float4(bf16_lo(wb[j].x), bf16_hi(wb[j].x), bf16_lo(wb[j].y), bf16_hi(wb[j].y));
        *reinterpret_cast<float4*>(&w_s[8 * i + 4]) =
            make_float4(bf16_lo(wb[j].z), bf16_hi(wb[j].z), bf16_lo(wb[j].w), bf16_hi(wb[j].w));
      }
    }
  }
  __syncthreads();
  // trip count uniform per warp pair (both warps of a row take the pair barrier below)
  for (int r0 = blockIdx.x * ROWS; r0 < n_rows; r0 += stride) {
    const int r = r0 + slot;
    const bool live = r < n_rows;
    uint4 v[CPL];
#pragma unroll
    for (int i = 0; i < CPL; ++i) v[i] = nx[i];
    if (r + stride < n_rows) load_row(r + stride, nx);  // prefetch: in flight during the reduction and the stores
    float ss = 0.f;
    if (live) {
#pragma unroll
      for (int i = 0; i < CPL; ++i) {
        const uint32_t u[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float a = bf16_lo(u[j]), b = bf16_hi(u[j]);
          ss = fmaf(a, a, ss);
          ss = fmaf(b, b, ss);
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if constexpr (WPR == 2) {
      // pair-scoped named barrier (id 1 + slot, 64 threads): the two warps of a row meet, the other rows of the
      // CTA keep streaming -- a CTA-wide barrier would put all 8 warps in lock-step load / store phases
      if (lane == 0) part[warp] = ss;
      asm volatile("bar.sync %0, 64;" ::"r"(slot + 1) : "memory");
      ss = part[slot * 2] + part[slot * 2 + 1];
      asm volatile("bar.sync %0, 64;" ::"r"(slot + 1) : "memory");  // part[] is rewritten in the next iteration
    }
    if (live) {
      const float inv = rsqrtf(ss * (1.0f / H) + eps);
      const int dst = row_dst ? row_dst[r] : r;
      uint4* yp = reinterpret_cast<uint4*>(y + static_cast<int64_t>(dst) * H) + half * CPL * 32;
#pragma unroll
      for (int i = 0; i < CPL; ++i) {
        const int col = ((half * CPL + i) * 32 + lane) * 8;
        const float4 w0 = *reinterpret_cast<const float4*>(&w_s[col]);
        const float4 w1 = *reinterpret_cast<const float4*>(&w_s[col + 4]);
        uint4 o;
        // reference order: (x * rsqrt(var + eps)) in fp32, then weight * that, then one cast
        o.x = pack_bf16(w0.x * (bf16_lo(v[i].x) * inv), w0.y * (bf16_hi(v[i].x) * inv));
        o.y = pack_bf16(w0.z * (bf16_lo(v[i].y) * inv), w0.w * (bf16_hi(v[i].y) * inv));
        o.z = pack_bf16(w1.x * (bf16_lo(v[i].z) * inv), w1.y * (bf16_hi(v[i].z) * inv));
        o.w = pack_bf16(w1.z * (bf16_lo(v[i].w) * inv), w1.w * (bf16_hi(v[i].w) * inv));
        st_stream(yp + i * 32 + lane, o);
      }
    }
  }
}

}  // namespace vex

extern "C" int vex_rmsnorm_gather(const void* x, const void* weight, int weight_is_fp32, float eps,
                                  const int32_t* row_src, const int32_t* row_dst, const int32_t* n_rows, void* y,
                                  int rows_cap, int H, vexStream stream) {
  if (!x || !weight || !n_rows || !y || rows_cap <= 0) return VEX_E_INVALID;
  if (reinterpret_cast<uintptr_t>(weight) & 15) return VEX_E_INVALID;  // staged with 16-byte loads
  if (H % 256 != 0 || H <= 0 || H > 4096) return VEX_E_UNSUPPORTED;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // persistent-ish grid: every CTA stages the weight vector once, then strides over rows
  const int grid = std::min(vex::ceil_div(rows_cap, vex::K2_WARPS / 2), 148 * 2);
  auto xp = static_cast<const __nv_bfloat16*>(x);
  auto yp = static_cast<__nv_bfloat16*>(y);
#define VEX_K2_CASE(NC)                                                                                        \
  case NC:                                                                                                     \
    VEX_CUDA_TRY(vex::launch_pdl(vex::k2_rmsnorm<NC>, dim3(grid), dim3(vex::K2_WARPS * 32), 0, s, xp, weight,    \
                                 weight_is_fp32, eps, row_src, row_dst, n_rows, yp, rows_cap));                       \
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
