// K10 -- integer / reduction kernels around the fused lm_head + cross-entropy:
// CogVLMForCausalLM.forward computes logits = lm_head(h).float() for EVERY position ([B, L, V] fp32, 1.5 GB per 8
// samples) and _sample_weighted_ce (modeling_cogvlm.py:610-627) then keeps only the rows with labels != -100.
// Here the rows are selected first (vex_label_rows, bit-exact ascending order like boolean-mask indexing), the
// vocabulary GEMM runs over those rows only with the softmax statistics computed in its epilogue (VEX_EPI_CE), and
// vex_ce_reduce combines the per-tile partials into the loss.
#include <algorithm>

#include "common.cuh"

namespace vex {

constexpr int K10_THREADS = 1024;

// single CTA: ordered compaction of the positions with labels != ignore_index
__global__ void __launch_bounds__(K10_THREADS)
    k10_label_rows(const int64_t* __restrict__ labels, const void* __restrict__ weight, int weight_is_fp32, int n,
                   int64_t ignore_index, int32_t* __restrict__ row_idx, int32_t* __restrict__ label_sel,
                   float* __restrict__ w_sel, int32_t* __restrict__ count) {
  __shared__ int warp_tot[K10_THREADS / 32];
  __shared__ int base_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) base_s = 0;
  __syncthreads();
  for (int i0 = 0; i0 < n; i0 += K10_THREADS) {
    const int i = i0 + threadIdx.x;
    const int64_t lab = i < n ? labels[i] : ignore_index;
    const bool keep = lab != ignore_index;
    const uint32_t bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_tot[warp] = __popc(bal);
    __syncthreads();
    int before = 0;
    for (int w = 0; w < warp; ++w) before += warp_tot[w];
    const int pos = base_s + before + __popc(bal & ((1u << lane) - 1));
    if (keep) {
      row_idx[pos] = i;
      label_sel[pos] = static_cast<int32_t>(lab);
      float wv = 1.0f;
      if (weight) wv = weight_is_fp32 ? static_cast<const float*>(weight)[i]
                                      : __bfloat162float(static_cast<const __nv_bfloat16*>(weight)[i]);
      w_sel[pos] = wv;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int w = 0; w < K10_THREADS / 32; ++w) tot += warp_tot[w];
      base_s += tot;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) count[0] = base_s;
}

// one warp per selected row: combine the per-tile (max, sum-exp) partials, accumulate the weighted loss
__global__ void __launch_bounds__(256)
    k10_ce_reduce(const float* __restrict__ pmax, const float* __restrict__ psum, const float* __restrict__ zlabel,
                  const float* __restrict__ w_sel, const int32_t* __restrict__ count, int rows_cap, int n_tiles,
                  float* __restrict__ lse, float* __restrict__ loss) {
  const int n_rows = min(count[0], rows_cap);
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
  float acc = 0.f;
  for (int r = warp; r < n_rows; r += n_warps) {
    const float* pm = pmax + static_cast<int64_t>(r) * n_tiles;
    const float* ps = psum + static_cast<int64_t>(r) * n_tiles;
    float m = -INFINITY;
    for (int t = lane; t < n_tiles; t += 32) m = fmaxf(m, pm[t]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.f;
    for (int t = lane; t < n_tiles; t += 32) s += ps[t] * __expf(pm[t] - m);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float l = m + logf(s);
    if (lane == 0) {
      lse[r] = l;
      acc += (l - zlabel[r]) * w_sel[r];
    }
  }
  if (lane == 0 && acc != 0.f) atomicAdd(loss, acc / static_cast<float>(n_rows));
  // an all-ignored batch: the reference divides 0 by 0
  if (blockIdx.x == 0 && threadIdx.x == 0 && n_rows == 0) loss[0] = __int_as_float(0x7fc00000);
}

}  // namespace vex

extern "C" int vex_label_rows(const int64_t* labels, const void* weight, int weight_is_fp32, int n,
                              int64_t ignore_index, int32_t* row_idx, int32_t* label_sel, float* w_sel,
                              int32_t* count, vexStream stream) {
  if (!labels || !row_idx || !label_sel || !w_sel || !count || n <= 0) return VEX_E_INVALID;
  if (n > (1 << 24)) return VEX_E_UNSUPPORTED;
  vex::k10_label_rows<<<1, vex::K10_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      labels, weight, weight_is_fp32, n, ignore_index, row_idx, label_sel, w_sel, count);
  VEX_LAUNCH_CHECK();
  return VEX_OK;
}

extern "C" int vex_ce_reduce(const float* pmax, const float* psum, const float* zlabel, const float* w_sel,
                             const int32_t* count, int rows_cap, int n_tiles, float* lse, float* loss,
                             vexStream stream) {
  if (!pmax || !psum || !zlabel || !w_sel || !count || !lse || !loss || rows_cap <= 0 || n_tiles <= 0)
    return VEX_E_INVALID;
  const int grid = std::min(vex::ceil_div(rows_cap, 8), 148 * 4);
  vex::k10_ce_reduce<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(pmax, psum, zlabel, w_sel, count, rows_cap,
                                                                          n_tiles, lse, loss);
  VEX_LAUNCH_CHECK();
  return VEX_OK;
}
