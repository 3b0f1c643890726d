// K4 entry point: argument checks and implementation choice for the causal varlen attention.
#include "common.cuh"

namespace vex {
// the persistent two-tile tcgen05 kernel (k4_attention_tc3.cu).  The earlier kernels it superseded (one-tile tc1,
// non-persistent tc2, the mma.sync baseline) are kept for A/B measurements in csrc/baselines/ and are built into a
// separate libvex_baselines.so (vex_attention_baseline) that the product path never loads.
int launch_attention_tc3(const void* qkv, const int32_t* cu_seqlens, int B, int max_len_cap, int heads,
                         const int32_t* out_row_map, void* out, float scale, int rows_cap, float* lse, int causal,
                         cudaStream_t s);
}  // namespace vex

extern "C" int vex_attention(const void* qkv, const int32_t* cu_seqlens, int B, int max_len_cap, int heads,
                             const int32_t* out_row_map, void* out, float scale, vexStream stream) {
  return vex_attention_lse(qkv, cu_seqlens, B, max_len_cap, heads, out_row_map, out, scale, nullptr, stream);
}

extern "C" int vex_attention_lse(const void* qkv, const int32_t* cu_seqlens, int B, int max_len_cap, int heads,
                                 const int32_t* out_row_map, void* out, float scale, float* lse, vexStream stream) {
  if (!qkv || !cu_seqlens || !out || B <= 0 || max_len_cap <= 0 || heads <= 0) return VEX_E_INVALID;
  if (B > 65535 || heads > 65535) return VEX_E_UNSUPPORTED;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t rows_cap = static_cast<int64_t>(B) * max_len_cap;
  if (rows_cap > 0x7fffffff) return VEX_E_UNSUPPORTED;
  return vex::launch_attention_tc3(qkv, cu_seqlens, B, max_len_cap, heads, out_row_map, out, scale,
                                       static_cast<int>(rows_cap), lse, /*causal=*/1, s);
}

extern "C" int vex_attention_blockdiag(const void* qkv, const int32_t* cu_seqlens, int B, int max_len_cap, int heads,
                                       const int32_t* out_row_map, void* out, float scale, vexStream stream) {
  if (!qkv || !cu_seqlens || !out || B <= 0 || max_len_cap <= 0 || heads <= 0) return VEX_E_INVALID;
  if (B > 65535 || heads > 65535) return VEX_E_UNSUPPORTED;
  const int64_t rows_cap = static_cast<int64_t>(B) * max_len_cap;
  if (rows_cap > 0x7fffffff) return VEX_E_UNSUPPORTED;
  return vex::launch_attention_tc3(qkv, cu_seqlens, B, max_len_cap, heads, out_row_map, out, scale,
                                       static_cast<int>(rows_cap), nullptr, /*causal=*/0,
                                       static_cast<cudaStream_t>(stream));
}
