// K5 -- standalone SiLU gate, HBM-bound (3 * I * 2 bytes per row).
// Restates act_fn(gate_proj(x)) * up_proj(x) (modeling_cogvlm.py:55, act_fn = ACT2FN['silu'] :52) with the
// eager-bf16 rounding points: silu(gate) -> bf16, product -> bf16.  16-byte coalesced accesses,
// grid sized from the host-known row bound, live row count read from the device.
#include "common.cuh"

namespace vex {

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }

__global__ void __launch_bounds__(256) k5_silu_mul(const uint4* __restrict__ gate, const uint4* __restrict__ up,
                                                   uint4* __restrict__ out, const int32_t* __restrict__ n_rows_ptr,
                                                   int rows_cap, int vec_per_row) {
  const int64_t n = static_cast<int64_t>(min(*n_rows_ptr, rows_cap)) * vec_per_row;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint4 g = ld_stream(gate + i), u = ld_stream(up + i);
    const uint32_t gu[4] = {g.x, g.y, g.z, g.w}, uu[4] = {u.x, u.y, u.z, u.w};
    uint4 o;
    uint32_t* op = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a = bf16r(silu_f(bf16_lo(gu[j]))) * bf16_lo(uu[j]);
      const float b = bf16r(silu_f(bf16_hi(gu[j]))) * bf16_hi(uu[j]);
      op[j] = pack_bf16(a, b);
    }
    st_stream(out + i, o);
  }
}

}  // namespace vex

extern "C" int vex_silu_mul(const void* gate, const void* up, void* out, const int32_t* n_rows, int rows_cap, int I,
                            vexStream stream) {
  if (!gate || !up || !out || !n_rows || rows_cap <= 0 || I <= 0) return VEX_E_INVALID;
  if (I % 8 != 0) return VEX_E_UNSUPPORTED;
  const int64_t total = static_cast<int64_t>(rows_cap) * (I / 8);
  const int grid = static_cast<int>(std::min<int64_t>((total + 255) / 256, 148 * 16));  // see vex_silu_mul_backward
  vex::k5_silu_mul<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(gate), static_cast<const uint4*>(up), static_cast<uint4*>(out), n_rows, rows_cap,
      I / 8);
  VEX_LAUNCH_CHECK();
  return VEX_OK;
}
