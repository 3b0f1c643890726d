// K4 entry point: argument checks and implementation choice for the causal varlen attention.
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace vex {
int launch_attention_mma(const void* qkv, const int32_t* cu_seqlens, int B, int max_len_cap, int heads,
                         const int32_t* out_row_map, void* out, float scale, cudaStream_t s);
int launch_attention_tc(const void* qkv, const int32_t* cu_seqlens, int B, int max_len_cap, int heads,
                        const int32_t* out_row_map, void* out, float scale, int rows_cap, float* lse, int causal,
                        cudaStream_t s);
int launch_attention_tc2(const void* qkv, const int32_t* cu_seqlens, int B, int max_len_cap, int heads,
                         const int32_t* out_row_map, void* out, float scale, int rows_cap, float* lse, int causal,
                         cudaStream_t s);

int launch_attention_tc3(const void* qkv, const int32_t* cu_seqlens, int B, int max_len_cap, int heads,
                         const int32_t* out_row_map, void* out, float scale, int rows_cap, float* lse, int causal,
                         cudaStream_t s);

// Default: the persistent two-tile kernel (k4_attention_tc3.cu).  VEX_ATTN_IMPL=tc2 / tc1 select the non-persistent
// two-tile kernel (k4_attention_tc2.cu, schedules behind VEX_ATTN_P) and the one-tile kernel (k4_attention_tc.cu);
// all three are parity-tested (tests/test_kernels_gpu.py) and timed side by side by tools/bench_kernels.py.
static int launch_attention_tcgen05(const void* qkv, const int32_t* cu_seqlens, int B, int max_len_cap, int heads,
                                    const int32_t* out_row_map, void* out, float scale, int rows_cap, float* lse,
                                    int causal, cudaStream_t s) {
  const char* impl = std::getenv("VEX_ATTN_IMPL");
  if (impl && std::strcmp(impl, "tc1") == 0)
    return launch_attention_tc(qkv, cu_seqlens, B, max_len_cap, heads, out_row_map, out, scale, rows_cap, lse, causal, s);
  if (impl && std::strcmp(impl, "tc2") == 0)
    return launch_attention_tc2(qkv, cu_seqlens, B, max_len_cap, heads, out_row_map, out, scale, rows_cap, lse, causal, s);
  return launch_attention_tc3(qkv, cu_seqlens, B, max_len_cap, heads, out_row_map, out, scale, rows_cap, lse, causal, s);
}
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
  // VEX_ATTN_IMPL=mma selects the HMMA baseline kernel; default is the tcgen05/TMEM path
  const char* impl = std::getenv("VEX_ATTN_IMPL");
  if (impl && std::strcmp(impl, "mma") == 0 && lse != nullptr) return VEX_E_UNSUPPORTED;  // baseline: forward only
  if (impl && std::strcmp(impl, "mma") == 0)
    return vex::launch_attention_mma(qkv, cu_seqlens, B, max_len_cap, heads, out_row_map, out, scale, s);
  const int64_t rows_cap = static_cast<int64_t>(B) * max_len_cap;
  if (rows_cap > 0x7fffffff) return VEX_E_UNSUPPORTED;
  return vex::launch_attention_tcgen05(qkv, cu_seqlens, B, max_len_cap, heads, out_row_map, out, scale,
                                       static_cast<int>(rows_cap), lse, /*causal=*/1, s);
}

extern "C" int vex_attention_blockdiag(const void* qkv, const int32_t* cu_seqlens, int B, int max_len_cap, int heads,
                                       const int32_t* out_row_map, void* out, float scale, vexStream stream) {
  if (!qkv || !cu_seqlens || !out || B <= 0 || max_len_cap <= 0 || heads <= 0) return VEX_E_INVALID;
  if (B > 65535 || heads > 65535) return VEX_E_UNSUPPORTED;
  const int64_t rows_cap = static_cast<int64_t>(B) * max_len_cap;
  if (rows_cap > 0x7fffffff) return VEX_E_UNSUPPORTED;
  return vex::launch_attention_tcgen05(qkv, cu_seqlens, B, max_len_cap, heads, out_row_map, out, scale,
                                       static_cast<int>(rows_cap), nullptr, /*causal=*/0,
                                       static_cast<cudaStream_t>(stream));
}
