// K7 -- row-wise (HBM-bound) kernels of the training-step backward: row gather, SwiGLU backward and RMSNorm
// backward.  They are the adjoints of K5 / K2 (MLP.forward modeling_cogvlm.py:55, RMSNorm.forward :36-41), i.e.
// what torch.autograd runs for those lines in the reference's LoRA training step (mmmm.py:299-306), with fp32
// internals and one bf16 rounding per produced tensor.  16-byte coalesced streaming accesses, live row counts
// read from the device (no host sync).
#include <algorithm>

#include "common.cuh"

namespace vex {

// ---------------------------------------------------------------------------------------------
// out[r] = x[row_src[r]], r < *n_rows  (d_out[B, L] -> expert-sorted rows: the A operand of the dgrad GEMMs)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k7_gather_rows(const uint4* __restrict__ x, const int32_t* __restrict__ row_src,
                                                      const int32_t* __restrict__ n_rows_ptr, uint4* __restrict__ out,
                                                      int rows_cap, int vec_per_row) {
  const int n_rows = min(*n_rows_ptr, rows_cap);
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int n_warps = (gridDim.x * blockDim.x) >> 5;
  for (int r = warp; r < n_rows; r += n_warps) {
    const int src = row_src ? row_src[r] : r;
    const uint4* xp = x + static_cast<int64_t>(src) * vec_per_row;
    uint4* op = out + static_cast<int64_t>(r) * vec_per_row;
    for (int i = lane; i < vec_per_row; i += 32) st_stream(op + i, ld_stream(xp + i));
  }
}

// ---------------------------------------------------------------------------------------------
// LoRA input dropout (PEFT lora.Linear: lora_A(dropout(x)), conf/lora.yaml lora_dropout = 0.05):
//   out[r, c] = keep(r, c) ? bf16(x[r, c] * 1/(1-p)) : 0,  r < *n_rows, mask from dropout_hash(element index, seed)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    k7_dropout_rows(const uint4* __restrict__ x, uint4* __restrict__ out, const int32_t* __restrict__ n_rows_ptr,
                    int rows_cap, int vec_per_row, uint32_t thresh16, float scale, uint32_t seed_lo, uint32_t seed_hi) {
  const int64_t n = static_cast<int64_t>(min(*n_rows_ptr, rows_cap)) * vec_per_row;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint4 v = ld_stream(x + i);
    const uint32_t vv[4] = {v.x, v.y, v.z, v.w};
    uint4 o;
    uint32_t* op = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) {  // element pair 4 * i + j
      const uint32_t hsh = dropout_hash(static_cast<uint64_t>(i) * 4 + j, seed_lo, seed_hi);
      const float a = (hsh & 0xffffu) >= thresh16 ? bf16_lo(vv[j]) * scale : 0.f;
      const float b = (hsh >> 16) >= thresh16 ? bf16_hi(vv[j]) * scale : 0.f;
      op[j] = pack_bf16(a, b);
    }
    st_stream(out + i, o);
  }
}

// ---------------------------------------------------------------------------------------------
// SwiGLU backward.  Forward (eager bf16): s = bf16(silu(g)), act = bf16(s * u).
//   dup = bf16(dact * s);  ds = bf16(dact * u);  dgate = bf16(ds * sigma(g) * (1 + g * (1 - sigma(g))))
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    k7_silu_mul_backward(const uint4* __restrict__ dact, const uint4* __restrict__ gate, const uint4* __restrict__ up,
                         uint4* __restrict__ dgate, uint4* __restrict__ dup, const int32_t* __restrict__ n_rows_ptr,
                         int rows_cap, int vec_per_row) {
  const int64_t n = static_cast<int64_t>(min(*n_rows_ptr, rows_cap)) * vec_per_row;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint4 d4 = ld_stream(dact + i), g4 = ld_stream(gate + i), u4 = ld_stream(up + i);
    const uint32_t dd[4] = {d4.x, d4.y, d4.z, d4.w}, gg[4] = {g4.x, g4.y, g4.z, g4.w},
                   uu[4] = {u4.x, u4.y, u4.z, u4.w};
    uint4 og, ou;
    uint32_t* ogp = reinterpret_cast<uint32_t*>(&og);
    uint32_t* oup = reinterpret_cast<uint32_t*>(&ou);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float dgv[2], duv[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float d = h ? bf16_hi(dd[j]) : bf16_lo(dd[j]);
        const float g = h ? bf16_hi(gg[j]) : bf16_lo(gg[j]);
        const float u = h ? bf16_hi(uu[j]) : bf16_lo(uu[j]);
        const float sig = 1.0f / (1.0f + __expf(-g));
        duv[h] = d * bf16r(g * sig);
        dgv[h] = bf16r(d * u) * (sig * (1.0f + g * (1.0f - sig)));
      }
      ogp[j] = pack_bf16(dgv[0], dgv[1]);
      oup[j] = pack_bf16(duv[0], duv[1]);
    }
    st_stream(dgate + i, og);
    st_stream(dup + i, ou);
  }
}

// ---------------------------------------------------------------------------------------------
// RMSNorm backward.  Forward: y = bf16(w * (x * inv)), inv = rsqrt(mean(x^2) + eps), all fp32 inside.
//   dx = inv * (dy*w) - x * inv^3 * sum(dy*w*x) / H          dw += sum_rows dy * x * inv   (fp32 atomics)
//   out[dx_map[r]] = bf16(dx + add[add_map[r]])              (the residual branch's gradient, fused)
// Same row slicing as K2: WPR warps per row so that a lane keeps <= 8 x 16 B of x and of dy in registers.
// ---------------------------------------------------------------------------------------------
constexpr int K7_WARPS = 8;

template <int NCHUNK>  // H = NCHUNK * 256
__global__ void __launch_bounds__(K7_WARPS * 32, NCHUNK >= 8 ? 1 : 2)
    k7_rmsnorm_backward(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                        const int32_t* __restrict__ x_map, const void* __restrict__ weight, int weight_is_fp32, float eps,
                        const __nv_bfloat16* __restrict__ add, const int32_t* __restrict__ add_map,
                        __nv_bfloat16* __restrict__ dx, const int32_t* __restrict__ dx_map, float* __restrict__ dweight,
                        const int32_t* __restrict__ n_rows_ptr, int rows_cap) {
  constexpr int H = NCHUNK * 256;
  constexpr int WPR = NCHUNK > 8 ? 2 : 1;
  constexpr int CPL = NCHUNK / WPR;
  constexpr int ROWS = K7_WARPS / WPR;
  static_assert(NCHUNK % WPR == 0, "row slices must be whole chunks");
  __shared__ __align__(16) float w_s[H];
  __shared__ __align__(16) float dw_s[H];
  __shared__ float part[K7_WARPS][2];
#pragma unroll  // exactly H / 256 trips: unrolled so that all the (independent) loads are in flight at once
  for (int i = threadIdx.x; i < H; i += K7_WARPS * 32) {
    w_s[i] = weight_is_fp32 ? static_cast<const float*>(weight)[i]
                            : __bfloat162float(static_cast<const __nv_bfloat16*>(weight)[i]);
    dw_s[i] = 0.f;
  }
  __syncthreads();
  const int lane = lane_id();
  const int warp = threadIdx.x >> 5;
  const int slot = warp / WPR, half = warp % WPR;
  const int n_rows = min(*n_rows_ptr, rows_cap);
  float dwacc[CPL * 8];
#pragma unroll
  for (int i = 0; i < CPL * 8; ++i) dwacc[i] = 0.f;

  for (int r0 = blockIdx.x * ROWS; r0 < n_rows; r0 += gridDim.x * ROWS) {
    const int r = r0 + slot;
    const bool live = r < n_rows;
    uint4 xv[CPL], dv[CPL];
    float ss = 0.f, sd = 0.f;
    if (live) {
      const int src = x_map ? x_map[r] : r;
      const uint4* xp = reinterpret_cast<const uint4*>(x + static_cast<int64_t>(src) * H) + half * CPL * 32;
      const uint4* dp = reinterpret_cast<const uint4*>(dy + static_cast<int64_t>(r) * H) + half * CPL * 32;
#pragma unroll
      for (int i = 0; i < CPL; ++i) {
        xv[i] = ld_stream(xp + i * 32 + lane);
        dv[i] = ld_stream(dp + i * 32 + lane);
      }
#pragma unroll
      for (int i = 0; i < CPL; ++i) {
        const int col = ((half * CPL + i) * 32 + lane) * 8;
        const uint32_t xu[4] = {xv[i].x, xv[i].y, xv[i].z, xv[i].w}, du[4] = {dv[i].x, dv[i].y, dv[i].z, dv[i].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float xa = bf16_lo(xu[j]), xb = bf16_hi(xu[j]);
          ss = fmaf(xa, xa, ss);
          ss = fmaf(xb, xb, ss);
          sd = fmaf(bf16_lo(du[j]) * w_s[col + 2 * j], xa, sd);
          sd = fmaf(bf16_hi(du[j]) * w_s[col + 2 * j + 1], xb, sd);
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ss += __shfl_xor_sync(0xffffffffu, ss, o);
      sd += __shfl_xor_sync(0xffffffffu, sd, o);
    }
    if constexpr (WPR == 2) {
      if (lane == 0) {
        part[warp][0] = ss;
        part[warp][1] = sd;
      }
      asm volatile("bar.sync %0, 64;" ::"r"(slot + 1) : "memory");  // pair-scoped: other rows keep streaming
      ss = part[slot * 2][0] + part[slot * 2 + 1][0];
      sd = part[slot * 2][1] + part[slot * 2 + 1][1];
      asm volatile("bar.sync %0, 64;" ::"r"(slot + 1) : "memory");
    }
    if (live) {
      const float inv = rsqrtf(ss * (1.0f / H) + eps);
      const float c = inv * inv * inv * sd * (1.0f / H);
      const int dst = dx_map ? dx_map[r] : r;
      uint4* op = reinterpret_cast<uint4*>(dx + static_cast<int64_t>(dst) * H) + half * CPL * 32;
      const uint4* ap = nullptr;
      if (add) {
        const int asrc = add_map ? add_map[r] : r;
        ap = reinterpret_cast<const uint4*>(add + static_cast<int64_t>(asrc) * H) + half * CPL * 32;
      }
#pragma unroll
      for (int i = 0; i < CPL; ++i) {
        const int col = ((half * CPL + i) * 32 + lane) * 8;
        const uint32_t xu[4] = {xv[i].x, xv[i].y, xv[i].z, xv[i].w}, du[4] = {dv[i].x, dv[i].y, dv[i].z, dv[i].w};
        uint4 a4 = make_uint4(0, 0, 0, 0);
        if (ap) a4 = ld_stream(ap + i * 32 + lane);
        const uint32_t au[4] = {a4.x, a4.y, a4.z, a4.w};
        uint4 o;
        uint32_t* ou = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float xa = bf16_lo(xu[j]), xb = bf16_hi(xu[j]);
          const float da = bf16_lo(du[j]), db = bf16_hi(du[j]);
          dwacc[i * 8 + 2 * j] = fmaf(da, xa * inv, dwacc[i * 8 + 2 * j]);
          dwacc[i * 8 + 2 * j + 1] = fmaf(db, xb * inv, dwacc[i * 8 + 2 * j + 1]);
          const float ga = inv * (da * w_s[col + 2 * j]) - xa * c;
          const float gb = inv * (db * w_s[col + 2 * j + 1]) - xb * c;
          ou[j] = pack_bf16(ga + bf16_lo(au[j]), gb + bf16_hi(au[j]));
        }
        st_stream(op + i * 32 + lane, o);
      }
    }
  }
  // weight gradient: registers -> shared (fp32 atomics across the CTA's row slots) -> one global atomic per column
  if (dweight) {
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
      const int col = ((half * CPL + i) * 32 + lane) * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&dw_s[col + j], dwacc[i * 8 + j]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < H; i += K7_WARPS * 32) {
      const float v = dw_s[i];
      if (v != 0.f) atomicAdd(dweight + i, v);
    }
  }
}

}  // namespace vex

extern "C" int vex_gather_rows(const void* x, const int32_t* row_src, const int32_t* n_rows, void* out, int rows_cap,
                               int H, vexStream stream) {
  if (!x || !n_rows || !out || rows_cap <= 0 || H <= 0) return VEX_E_INVALID;
  if (H % 8 != 0) return VEX_E_UNSUPPORTED;
  static const int resident = vex::resident_ctas(vex::k7_gather_rows, 256);
  const int grid = std::min(vex::ceil_div(rows_cap, 8), resident);
  vex::k7_gather_rows<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(x), row_src, n_rows, static_cast<uint4*>(out), rows_cap, H / 8);
  VEX_LAUNCH_CHECK();
  return VEX_OK;
}

extern "C" int vex_dropout_rows(const void* x, void* out, const int32_t* n_rows, int rows_cap, int K, float p,
                                uint64_t seed, vexStream stream) {
  if (!x || !out || !n_rows || rows_cap <= 0 || K <= 0 || !(p >= 0.f && p < 1.f)) return VEX_E_INVALID;
  if (K % 8 != 0) return VEX_E_UNSUPPORTED;
  const int64_t total = static_cast<int64_t>(rows_cap) * (K / 8);
  const int grid = static_cast<int>(std::min<int64_t>((total + 255) / 256, 148 * 16));
  const uint32_t thresh16 = static_cast<uint32_t>(p * 65536.0f + 0.5f);
  vex::k7_dropout_rows<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(x), static_cast<uint4*>(out), n_rows, rows_cap, K / 8, thresh16, 1.0f / (1.0f - p),
      static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
  VEX_LAUNCH_CHECK();
  return VEX_OK;
}

extern "C" int vex_silu_mul_backward(const void* dact, const void* gate, const void* up, void* dgate, void* dup,
                                     const int32_t* n_rows, int rows_cap, int I, vexStream stream) {
  if (!dact || !gate || !up || !dgate || !dup || !n_rows || rows_cap <= 0 || I <= 0) return VEX_E_INVALID;
  if (I % 8 != 0) return VEX_E_UNSUPPORTED;
  const int64_t total = static_cast<int64_t>(rows_cap) * (I / 8);
  // two waves of grid-stride CTAs on purpose: for these one-vector-per-thread streaming kernels a resident-only grid
  // measured 7 % slower (K5: 197 -> 211 us) -- the second wave back-fills SMs as the first one drains
  const int grid = static_cast<int>(std::min<int64_t>((total + 255) / 256, 148 * 16));
  vex::k7_silu_mul_backward<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(dact), static_cast<const uint4*>(gate), static_cast<const uint4*>(up),
      static_cast<uint4*>(dgate), static_cast<uint4*>(dup), n_rows, rows_cap, I / 8);
  VEX_LAUNCH_CHECK();
  return VEX_OK;
}

extern "C" int vex_rmsnorm_backward(const void* dy, const void* x, const int32_t* x_map, const void* weight,
                                    int weight_is_fp32, float eps, const void* add, const int32_t* add_map, void* dx,
                                    const int32_t* dx_map, float* dweight, const int32_t* n_rows, int rows_cap, int H,
                                    vexStream stream) {
  if (!dy || !x || !weight || !dx || !n_rows || rows_cap <= 0) return VEX_E_INVALID;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  auto dyp = static_cast<const __nv_bfloat16*>(dy);
  auto xp = static_cast<const __nv_bfloat16*>(x);
  auto ap = static_cast<const __nv_bfloat16*>(add);
  auto op = static_cast<__nv_bfloat16*>(dx);
#define VEX_K7_CASE(NC)                                                                                             \
  case NC: {                                                                                                        \
    static const int resident = vex::resident_ctas(vex::k7_rmsnorm_backward<NC>, vex::K7_WARPS * 32);               \
    const int grid = std::min(vex::ceil_div(rows_cap, vex::K7_WARPS / 2), resident);                                \
    vex::k7_rmsnorm_backward<NC><<<grid, vex::K7_WARPS * 32, 0, s>>>(dyp, xp, x_map, weight, weight_is_fp32, eps, ap, \
                                                                      add_map, op, dx_map, dweight, n_rows, rows_cap); \
    break;                                                                                                          \
  }
  switch (H % 256 == 0 ? H / 256 : 0) {
    VEX_K7_CASE(1) VEX_K7_CASE(2) VEX_K7_CASE(4) VEX_K7_CASE(8) VEX_K7_CASE(16)
    default:
      return VEX_E_UNSUPPORTED;
  }
#undef VEX_K7_CASE
  VEX_LAUNCH_CHECK();
  return VEX_OK;
}
