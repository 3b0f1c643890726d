// K11 -- row-wise (HBM-bound) kernels of the vision encoder that feeds the decoder:
//   k11_layernorm      nn.LayerNorm on the BRANCH OUTPUT + residual add of TransformerLayer.forward
//                      (visual.py:128-135: h = x + LN(attn(x)); out = h + LN(mlp(h))) and LayerNorm + GELU of the GLU
//                      projector (visual.py:173-174)
//   k11_patchify       im2col of the strided patch convolution (PatchEmbedding.forward visual.py:65 ->
//                      Downsample.forward mmmm/models/resample.py:56-63): the GEMM A operand [patches, C*pd*ph*pw]
//   k11_maxpool_tokens class-token drop + 3-D max-pool over the patch grid (EVA2CLIPModel.forward visual.py:197-202)
//   k11_scatter_rows   boi / eoi rows (visual.py:204-206) and generic row scatter
// 16-byte coalesced streaming accesses; fp32 statistics; one bf16 rounding per eager op of the reference.
#include <algorithm>

#include "common.cuh"

namespace vex {

constexpr int K11_WARPS = 8;

// ---------------------------------------------------------------------------------------------
// LayerNorm: t = bf16((x - mean) * rstd * w + b) [-> bf16(gelu(t))] [-> bf16(residual + t)]
// one warp per row, the row stays in registers (NCHUNK x 16 B per lane), two-pass variance
// ---------------------------------------------------------------------------------------------
template <int NCHUNK>  // H = NCHUNK * 256
__global__ void __launch_bounds__(K11_WARPS * 32, NCHUNK <= 8 ? 2 : 1)  // <= 128 registers: 16 warps per SM
    k11_layernorm(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ weight,
                  const __nv_bfloat16* __restrict__ bias, float eps, const __nv_bfloat16* residual, int act,
                  const int32_t* __restrict__ n_rows_ptr, __nv_bfloat16* y, int rows_cap) {
  constexpr int H = NCHUNK * 256;
  __shared__ __align__(16) float w_s[H];
  __shared__ __align__(16) float b_s[H];
#pragma unroll  // exactly H / 256 trips: unrolled so that all the (independent) loads are in flight at once
  for (int i = threadIdx.x; i < H; i += K11_WARPS * 32) {
    w_s[i] = __bfloat162float(weight[i]);
    b_s[i] = bias ? __bfloat162float(bias[i]) : 0.f;
  }
  __syncthreads();
  const int lane = lane_id();
  const int warp = threadIdx.x >> 5;
  const int n_rows = min(*n_rows_ptr, rows_cap);
  for (int r = blockIdx.x * K11_WARPS + warp; r < n_rows; r += gridDim.x * K11_WARPS) {
    const uint4* xp = reinterpret_cast<const uint4*>(x + static_cast<int64_t>(r) * H);
    uint4 v[NCHUNK], rv[NCHUNK];
#pragma unroll
    for (int i = 0; i < NCHUNK; ++i) v[i] = ld_stream(xp + i * 32 + lane);
    // the residual row is requested together with x: one memory round trip per row instead of two
    const uint4* rp = residual ? reinterpret_cast<const uint4*>(residual + static_cast<int64_t>(r) * H) : nullptr;
    if (rp) {
#pragma unroll
      for (int i = 0; i < NCHUNK; ++i) rv[i] = ld_stream(rp + i * 32 + lane);
    }
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NCHUNK; ++i) {
      const uint32_t u[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) sum += bf16_lo(u[j]) + bf16_hi(u[j]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * (1.0f / H);
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NCHUNK; ++i) {
      const uint32_t u[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float a = bf16_lo(u[j]) - mean, b = bf16_hi(u[j]) - mean;
        sq = fmaf(a, a, sq);
        sq = fmaf(b, b, sq);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq * (1.0f / H) + eps);
    uint4* yp = reinterpret_cast<uint4*>(y + static_cast<int64_t>(r) * H);
#pragma unroll
    for (int i = 0; i < NCHUNK; ++i) {
      const int col = (i * 32 + lane) * 8;
      const float4 w0 = *reinterpret_cast<const float4*>(&w_s[col]);
      const float4 w1 = *reinterpret_cast<const float4*>(&w_s[col + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&b_s[col]);
      const float4 b1 = *reinterpret_cast<const float4*>(&b_s[col + 4]);
      const float ws[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
      const float bs[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      const uint32_t u[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
      float t[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        t[2 * j] = fmaf((bf16_lo(u[j]) - mean) * rstd, ws[2 * j], bs[2 * j]);
        t[2 * j + 1] = fmaf((bf16_hi(u[j]) - mean) * rstd, ws[2 * j + 1], bs[2 * j + 1]);
      }
      if (act == VEX_ACT_GELU) {
#pragma unroll
        for (int j = 0; j < 8; ++j) t[j] = gelu_erf(bf16r(t[j]));
      }
      if (rp) {
        const uint4 r4 = rv[i];
        const uint32_t ru[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          t[2 * j] = bf16_lo(ru[j]) + bf16r(t[2 * j]);
          t[2 * j + 1] = bf16_hi(ru[j]) + bf16r(t[2 * j + 1]);
        }
      }
      uint4 o;
      o.x = pack_bf16(t[0], t[1]);
      o.y = pack_bf16(t[2], t[3]);
      o.z = pack_bf16(t[4], t[5]);
      o.w = pack_bf16(t[6], t[7]);
      st_stream(yp + i * 32 + lane, o);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// im2col of a kernel == stride 3-D convolution: patch (gz, gy, gx) of image [C, D, H, W] becomes row
// (gz * gh + gy) * gw + gx with columns ((c * pd + kz) * ph + ky) * pw + kx  (== weight.reshape(C_out, -1) order).
// One thread per contiguous pw-element segment; 16-byte copies when pw, W and the bases allow.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    k11_patchify(const __nv_bfloat16* __restrict__ img, int C, int D, int H, int W, int pd, int ph, int pw, int gd,
                 int gh, int gw, __nv_bfloat16* __restrict__ out, int64_t ldo, int vec_ok) {
  const int64_t segs_per_row = static_cast<int64_t>(C) * pd * ph;
  const int64_t n_rows = static_cast<int64_t>(gd) * gh * gw;
  const int64_t total = n_rows * segs_per_row;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    // consecutive threads walk gx fastest so that a warp reads consecutive pw-segments of one image line
    const int gx = static_cast<int>(i % gw);
    int64_t t = i / gw;
    const int ky = static_cast<int>(t % ph);
    t /= ph;
    const int gy = static_cast<int>(t % gh);
    t /= gh;
    const int kz = static_cast<int>(t % pd);
    t /= pd;
    const int gz = static_cast<int>(t % gd);
    const int c = static_cast<int>(t / gd);
    const __nv_bfloat16* src =
        img + ((static_cast<int64_t>(c) * D + (gz * pd + kz)) * H + (gy * ph + ky)) * W + static_cast<int64_t>(gx) * pw;
    __nv_bfloat16* dst = out + ((static_cast<int64_t>(gz) * gh + gy) * gw + gx) * ldo +
                         ((static_cast<int64_t>(c) * pd + kz) * ph + ky) * pw;
    if (vec_ok) {
      for (int k = 0; k < pw; k += 8)
        *reinterpret_cast<uint4*>(dst + k) = __ldg(reinterpret_cast<const uint4*>(src + k));
    } else {
      for (int k = 0; k < pw; ++k) dst[k] = src[k];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// out[o] = max over the (pz, py, px) window of token rows x[(z * gh + y) * gw + x_] (floor semantics of
// F.max_pool3d: incomplete windows are dropped).  pool == (1, 1, 1) is a row copy (class-token drop).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t bf16x2_max(uint32_t a, uint32_t b) {
  __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}

__global__ void __launch_bounds__(256)
    k11_maxpool_tokens(const uint4* __restrict__ x, int64_t ldx_vec, int gh, int gw, int pz, int py, int px, int od,
                       int oh, int ow, uint4* __restrict__ out, int64_t ldo_vec, int vec_per_row) {
  const int64_t total = static_cast<int64_t>(od) * oh * ow * vec_per_row;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % vec_per_row);
    int64_t o = i / vec_per_row;
    const int ox = static_cast<int>(o % ow);
    const int oy = static_cast<int>((o / ow) % oh);
    const int oz = static_cast<int>(o / (static_cast<int64_t>(ow) * oh));
    uint4 m;
    bool first = true;
    for (int dz = 0; dz < pz; ++dz)
      for (int dy = 0; dy < py; ++dy)
        for (int dx = 0; dx < px; ++dx) {
          const int64_t row = (static_cast<int64_t>(oz * pz + dz) * gh + (oy * py + dy)) * gw + (ox * px + dx);
          const uint4 a = ld_stream(x + row * ldx_vec + v);
          if (first) {
            m = a;
            first = false;
          } else {
            m.x = bf16x2_max(m.x, a.x);
            m.y = bf16x2_max(m.y, a.y);
            m.z = bf16x2_max(m.z, a.z);
            m.w = bf16x2_max(m.w, a.w);
          }
        }
    st_stream(out + o * ldo_vec + v, m);
  }
}

// out[row_dst[r]] = x[row_src ? row_src[r] : r] for r < n; rows with row_dst[r] < 0 are skipped
__global__ void __launch_bounds__(256)
    k11_scatter_rows(const uint4* __restrict__ x, const int32_t* __restrict__ row_src,
                     const int32_t* __restrict__ row_dst, int n, uint4* __restrict__ out, int vec_per_row) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  for (int r = warp; r < n; r += n_warps) {
    const int dst = row_dst[r];
    if (dst < 0) continue;
    const int src = row_src ? row_src[r] : r;
    const uint4* xp = x + static_cast<int64_t>(src) * vec_per_row;
    uint4* op = out + static_cast<int64_t>(dst) * vec_per_row;
    for (int i = lane; i < vec_per_row; i += 32) op[i] = __ldg(xp + i);
  }
}

}  // namespace vex

extern "C" int vex_layernorm(const void* x, const void* weight, const void* bias, float eps, const void* residual,
                             int act, const int32_t* n_rows, void* y, int rows_cap, int H, vexStream stream) {
  if (!x || !weight || !n_rows || !y || rows_cap <= 0) return VEX_E_INVALID;
  if (H % 256 != 0 || H <= 0 || H > 4096) return VEX_E_UNSUPPORTED;
  if (act != VEX_ACT_NONE && act != VEX_ACT_GELU) return VEX_E_INVALID;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int grid = std::min(vex::ceil_div(rows_cap, vex::K11_WARPS), 148 * 4);
  auto xp = static_cast<const __nv_bfloat16*>(x);
  auto wp = static_cast<const __nv_bfloat16*>(weight);
  auto bp = static_cast<const __nv_bfloat16*>(bias);
  auto rp = static_cast<const __nv_bfloat16*>(residual);
  auto yp = static_cast<__nv_bfloat16*>(y);
#define VEX_K11_CASE(NC)                                                                                         \
  case NC:                                                                                                       \
    vex::k11_layernorm<NC><<<grid, vex::K11_WARPS * 32, 0, s>>>(xp, wp, bp, eps, rp, act, n_rows, yp, rows_cap); \
    break;
  switch (H / 256) {
    VEX_K11_CASE(1) VEX_K11_CASE(2) VEX_K11_CASE(3) VEX_K11_CASE(4) VEX_K11_CASE(5) VEX_K11_CASE(6) VEX_K11_CASE(7)
    VEX_K11_CASE(8) VEX_K11_CASE(16)
    default:
      return VEX_E_UNSUPPORTED;
  }
#undef VEX_K11_CASE
  VEX_LAUNCH_CHECK();
  return VEX_OK;
}

extern "C" int vex_patchify(const void* image, int C, int D, int H, int W, int pd, int ph, int pw, void* out,
                            int64_t ldo, vexStream stream) {
  if (!image || !out || C <= 0 || D <= 0 || H <= 0 || W <= 0 || pd <= 0 || ph <= 0 || pw <= 0) return VEX_E_INVALID;
  const int gd = D / pd, gh = H / ph, gw = W / pw;  // conv3d with stride == kernel: incomplete patches are dropped
  if (gd <= 0 || gh <= 0 || gw <= 0) return VEX_E_INVALID;
  if (ldo < static_cast<int64_t>(C) * pd * ph * pw) return VEX_E_INVALID;
  const int vec_ok = pw % 8 == 0 && W % 8 == 0 && ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(image) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  const int64_t total = static_cast<int64_t>(gd) * gh * gw * C * pd * ph;
  const int grid = static_cast<int>(std::min<int64_t>((total + 255) / 256, 148 * 8));
  vex::k11_patchify<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(image), C, D, H, W, pd, ph, pw, gd, gh, gw, static_cast<__nv_bfloat16*>(out),
      ldo, vec_ok);
  VEX_LAUNCH_CHECK();
  return VEX_OK;
}

extern "C" int vex_maxpool_tokens(const void* x, int64_t ldx, int gd, int gh, int gw, int pz, int py, int px,
                                  void* out, int64_t ldo, int C, vexStream stream) {
  if (!x || !out || gd <= 0 || gh <= 0 || gw <= 0 || pz <= 0 || py <= 0 || px <= 0 || C <= 0) return VEX_E_INVALID;
  if (C % 8 != 0 || ldx % 8 != 0 || ldo % 8 != 0) return VEX_E_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return VEX_E_INVALID;
  const int od = gd / pz, oh = gh / py, ow = gw / px;
  if (od <= 0 || oh <= 0 || ow <= 0) return VEX_E_INVALID;
  const int64_t total = static_cast<int64_t>(od) * oh * ow * (C / 8);
  const int grid = static_cast<int>(std::min<int64_t>((total + 255) / 256, 148 * 8));
  vex::k11_maxpool_tokens<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(x), ldx / 8, gh, gw, pz, py, px, od, oh, ow, static_cast<uint4*>(out), ldo / 8, C / 8);
  VEX_LAUNCH_CHECK();
  return VEX_OK;
}

extern "C" int vex_scatter_rows(const void* x, const int32_t* row_src, const int32_t* row_dst, int n, void* out, int H,
                                vexStream stream) {
  if (!x || !row_dst || !out || n < 0 || H <= 0) return VEX_E_INVALID;
  if (H % 8 != 0) return VEX_E_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return VEX_E_INVALID;
  if (n == 0) return VEX_OK;
  const int grid = std::min(vex::ceil_div(n, 8), 148 * 4);
  vex::k11_scatter_rows<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(x), row_src, row_dst, n, static_cast<uint4*>(out), H / 8);
  VEX_LAUNCH_CHECK();
  return VEX_OK;
}
