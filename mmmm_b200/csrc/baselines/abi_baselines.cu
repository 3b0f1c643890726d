// libvex_baselines.so -- the attention kernels that k4_attention_tc3.cu superseded, kept OUT of the product library for
// A/B measurements (tools/bench_kernels.py) and cross-checks (tests/test_kernels_gpu.py).  Same buffers and semantics
// as vex_attention_lse / vex_attention_blockdiag (include/vex.h); `impl` = "mma" (mma.sync baseline, causal forward
// only), "tc1" (one 128-query tile per CTA, tcgen05), "tc2" (two tiles per CTA, non-persistent; schedules behind
// VEX_ATTN_P=tmem|smem|token).
#include <cstring>

#include "../common.cuh"

namespace vex {
thread_local int g_last_cuda_error = 0;
int launch_attention_mma(const void* qkv, const int32_t* cu_seqlens, int B, int max_len_cap, int heads,
                         const int32_t* out_row_map, void* out, float scale, cudaStream_t s);
int launch_attention_tc(const void* qkv, const int32_t* cu_seqlens, int B, int max_len_cap, int heads,
                        const int32_t* out_row_map, void* out, float scale, int rows_cap, float* lse, int causal,
                        cudaStream_t s);
int launch_attention_tc2(const void* qkv, const int32_t* cu_seqlens, int B, int max_len_cap, int heads,
                         const int32_t* out_row_map, void* out, float scale, int rows_cap, float* lse, int causal,
                         cudaStream_t s);
}  // namespace vex

extern "C" int vex_baselines_last_cuda_error(void) { return vex::g_last_cuda_error; }

extern "C" int vex_attention_baseline(const char* impl, const void* qkv, const int32_t* cu_seqlens, int B,
                                      int max_len_cap, int heads, const int32_t* out_row_map, void* out, float scale,
                                      float* lse, int causal, vexStream stream) {
  if (!impl || !qkv || !cu_seqlens || !out || B <= 0 || max_len_cap <= 0 || heads <= 0) return VEX_E_INVALID;
  if (B > 65535 || heads > 65535) return VEX_E_UNSUPPORTED;
  const int64_t rows_cap = static_cast<int64_t>(B) * max_len_cap;
  if (rows_cap > 0x7fffffff) return VEX_E_UNSUPPORTED;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (std::strcmp(impl, "mma") == 0) {
    if (lse != nullptr || !causal) return VEX_E_UNSUPPORTED;
    return vex::launch_attention_mma(qkv, cu_seqlens, B, max_len_cap, heads, out_row_map, out, scale, s);
  }
  if (std::strcmp(impl, "tc1") == 0)
    return vex::launch_attention_tc(qkv, cu_seqlens, B, max_len_cap, heads, out_row_map, out, scale,
                                    static_cast<int>(rows_cap), lse, causal, s);
  if (std::strcmp(impl, "tc2") == 0)
    return vex::launch_attention_tc2(qkv, cu_seqlens, B, max_len_cap, heads, out_row_map, out, scale,
                                     static_cast<int>(rows_cap), lse, causal, s);
  return VEX_E_INVALID;
}
