// K6 -- standalone residual add + scatter back to the reference [B, L, H] layout, HBM-bound
// (3 * H * 2 bytes per row), plus the pass-through of padded rows.
// Restates out[mask] = expert(x[mask]) followed by residual + out (modeling_cogvlm.py:278-279 + :321,
// :96-97 + :330): one bf16 rounding of the sum, like the eager add of two bf16 tensors.
#include <algorithm>

#include "common.cuh"

namespace vex {

__global__ void __launch_bounds__(256)
    k6_residual_scatter(const uint4* __restrict__ y, const __nv_bfloat16* __restrict__ residual,
                        const int32_t* __restrict__ row_dst, const int32_t* __restrict__ n_rows_ptr,
                        __nv_bfloat16* __restrict__ out, int rows_cap, int H) {
  const int vec_per_row = H / 8;
  const int n_rows = min(*n_rows_ptr, rows_cap);
  // one warp walks a row: consecutive lanes touch consecutive 16-byte vectors of the same row
  const int warps_total = gridDim.x * (blockDim.x >> 5);
  const int lane = lane_id();
  for (int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n_rows; r += warps_total) {
    const int64_t dst = row_dst ? row_dst[r] : r;
    const uint4* yr = y + static_cast<int64_t>(r) * vec_per_row;
    const uint4* rr = reinterpret_cast<const uint4*>(residual + dst * H);
    uint4* orow = reinterpret_cast<uint4*>(out + dst * H);
    for (int c = lane; c < vec_per_row; c += 32) {
      const uint4 a = ld_stream(yr + c), b = ld_stream(rr + c);
      const uint32_t au[4] = {a.x, a.y, a.z, a.w}, bu[4] = {b.x, b.y, b.z, b.w};
      uint4 o;
      uint32_t* op = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        op[j] = pack_bf16(bf16_lo(au[j]) + bf16_lo(bu[j]), bf16_hi(au[j]) + bf16_hi(bu[j]));
      st_stream(orow + c, o);
    }
  }
}

__global__ void __launch_bounds__(256)
    k6_copy_padded_rows(const __nv_bfloat16* __restrict__ x, const int32_t* __restrict__ flat_to_sorted,
                        __nv_bfloat16* __restrict__ out, int n_flat, int H) {
  const int vec_per_row = H / 8;
  const int warps_total = gridDim.x * (blockDim.x >> 5);
  const int lane = lane_id();
  for (int f = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); f < n_flat; f += warps_total) {
    if (flat_to_sorted[f] >= 0) continue;  // valid rows are written by the GEMM epilogues
    const uint4* src = reinterpret_cast<const uint4*>(x + static_cast<int64_t>(f) * H);
    uint4* dst = reinterpret_cast<uint4*>(out + static_cast<int64_t>(f) * H);
    for (int c = lane; c < vec_per_row; c += 32) st_stream(dst + c, ld_stream(src + c));
  }
}

// KV cache rows of padded positions: k[b, :, l, :] = v[b, :, l, :] = 0 where padding_mask[b, l] == False (the
// reference multiplies the projected q / k / v by the mask, modeling_cogvlm.py:243, so its cache holds zeros there).
// One warp per (flat position, head); valid positions are written by the QKV epilogue and skipped here.
__global__ void __launch_bounds__(256)
    k6_kv_clear_padded(__nv_bfloat16* __restrict__ k, __nv_bfloat16* __restrict__ v,
                       const int32_t* __restrict__ flat_to_sorted, int B, int L, int heads, int cap) {
  const int warps_total = gridDim.x * (blockDim.x >> 5);
  const int lane = lane_id();
  const int64_t n = static_cast<int64_t>(B) * L * heads;
  for (int64_t i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += warps_total) {
    const int f = static_cast<int>(i / heads), h = static_cast<int>(i - static_cast<int64_t>(f) * heads);
    if (flat_to_sorted[f] >= 0) continue;
    const int b = f / L, l = f - b * L;
    const int64_t row = (static_cast<int64_t>(b) * heads + h) * cap + l;
    if (lane < 16) {
      reinterpret_cast<uint4*>(k + row * 128)[lane] = make_uint4(0, 0, 0, 0);
      reinterpret_cast<uint4*>(v + row * 128)[lane] = make_uint4(0, 0, 0, 0);
    }
  }
}

}  // namespace vex

extern "C" int vex_kv_clear_padded(void* k_cache, void* v_cache, const int32_t* flat_to_sorted, int B, int L, int heads,
                                   int kv_capacity, vexStream stream) {
  if (!k_cache || !v_cache || !flat_to_sorted || B <= 0 || L <= 0 || heads <= 0 || kv_capacity < L) return VEX_E_INVALID;
  const int64_t n = static_cast<int64_t>(B) * L * heads;
  const int grid = static_cast<int>(std::min<int64_t>((n + 7) / 8, 148 * 8));
  vex::k6_kv_clear_padded<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<__nv_bfloat16*>(k_cache), static_cast<__nv_bfloat16*>(v_cache), flat_to_sorted, B, L, heads,
      kv_capacity);
  VEX_LAUNCH_CHECK();
  return VEX_OK;
}

extern "C" int vex_residual_scatter(const void* y, const void* residual, const int32_t* row_dst,
                                    const int32_t* n_rows, void* out, int rows_cap, int H, vexStream stream) {
  if (!y || !residual || !n_rows || !out || rows_cap <= 0 || H <= 0) return VEX_E_INVALID;
  if (H % 8 != 0) return VEX_E_UNSUPPORTED;
  static const int resident = vex::resident_ctas(vex::k6_residual_scatter, 256);
  const int grid = std::min(vex::ceil_div(rows_cap, 8), resident);
  vex::k6_residual_scatter<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(y), static_cast<const __nv_bfloat16*>(residual), row_dst, n_rows,
      static_cast<__nv_bfloat16*>(out), rows_cap, H);
  VEX_LAUNCH_CHECK();
  return VEX_OK;
}

extern "C" int vex_copy_padded_rows(const void* x, const int32_t* flat_to_sorted, void* out, int n_flat, int H,
                                    vexStream stream) {
  if (!x || !flat_to_sorted || !out || n_flat <= 0 || H <= 0) return VEX_E_INVALID;
  if (H % 8 != 0) return VEX_E_UNSUPPORTED;
  const int grid = std::min(vex::ceil_div(n_flat, 8), 148 * 8);
  vex::k6_copy_padded_rows<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), flat_to_sorted, static_cast<__nv_bfloat16*>(out), n_flat, H);
  VEX_LAUNCH_CHECK();
  return VEX_OK;
}
