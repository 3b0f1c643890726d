// K4d -- decode-step attention (q_len == 1 against the KV cache), memory-bound: every cached key and value is
// read once (2 * L * 128 * 2 bytes per (sample, head)).
//
// Restates the generation branch of attention_fn (modeling_cogvlm.py:129-141) with its eager-bf16 rounding
// points: the query is scaled in bf16 (`query_layer *= d ** -0.5`), scores are a bf16 einsum output, masked
// positions become -inf, softmax runs in fp32 and is cast back to bf16 before the weighted sum over the values.
// Cache layout is the reference's: k, v [B, heads, L, 128] (what prefill returns and torch.cat extends, :258-262).
//
// One CTA per (head, sample), 128 threads.  Scores: one key per thread (16 x 16-byte loads of its row, fp32
// dot product against the query held in shared memory).  Values: warp w takes keys w, w+4, ...; a lane owns
// 4 of the 128 output dims (8-byte loads, a 256-byte row per warp instruction, coalesced).
#include "common.cuh"

namespace vex {

constexpr int DEC_THREADS = 128;

__device__ __forceinline__ float block_reduce_max(float v, float* red) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  if (lane_id() == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  v = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  __syncthreads();
  return v;
}
__device__ __forceinline__ float block_reduce_sum(float v, float* red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane_id() == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  v = (red[0] + red[1]) + (red[2] + red[3]);
  __syncthreads();
  return v;
}

// `kv_len` (device, may be null): number of cache positions already filled BEFORE this step; the step attends to
// *kv_len + 1 positions (the current token's K / V were appended by the QKV epilogue).  Null = attend to all L.
// Cache rows of one (sample, head) are `cap` positions apart (cap >= L: pre-allocated headroom); the mask row
// stride is ld_mask.
__global__ void __launch_bounds__(DEC_THREADS)
    k4_attention_decode(const __nv_bfloat16* __restrict__ q, int64_t ldq, const __nv_bfloat16* __restrict__ k,
                        const __nv_bfloat16* __restrict__ v, const uint8_t* __restrict__ mask, int64_t ld_mask,
                        __nv_bfloat16* __restrict__ out, int heads, int L_host, int cap,
                        const int32_t* __restrict__ kv_len, float scale) {
  extern __shared__ float sc[];  // L scores, then probabilities
  __shared__ float qs[128];
  __shared__ float red[4];
  __shared__ float osum[4][128];
  const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const int L = kv_len ? min(max(kv_len[0] + 1, 1), L_host) : L_host;
  qs[tid] = bf16r(__bfloat162float(q[static_cast<int64_t>(b) * ldq + h * 128 + tid]) * scale);
  __syncthreads();
  const int64_t base = (static_cast<int64_t>(b) * heads + h) * cap;
  const uint8_t* mrow = mask + static_cast<int64_t>(b) * ld_mask;

  float mx = -INFINITY;
  for (int l = tid; l < L; l += DEC_THREADS) {
    float s = -INFINITY;
    if (mrow[l]) {
      const uint4* kr = reinterpret_cast<const uint4*>(k + (base + l) * 128);
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const uint4 u = ld_stream(kr + i);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc = fmaf(qs[i * 8 + 2 * j], bf16_lo(w[j]), acc);
          acc = fmaf(qs[i * 8 + 2 * j + 1], bf16_hi(w[j]), acc);
        }
      }
      s = bf16r(acc);
    }
    sc[l] = s;
    mx = fmaxf(mx, s);
  }
  mx = block_reduce_max(mx, red);
  float sum = 0.f;
  for (int l = tid; l < L; l += DEC_THREADS) {
    const float e = __expf(sc[l] - mx);
    sc[l] = e;
    sum += e;
  }
  sum = block_reduce_sum(sum, red);  // also orders the sc[] writes before the reads below
  const float inv = 1.0f / sum;

  const int warp = tid >> 5, lane = tid & 31;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 4
  for (int l = warp; l < L; l += 4) {
    const float p = bf16r(sc[l] * inv);
    if (p != 0.f) {  // masked keys have p == 0 (their values are zeroed in the reference, :135)
      const uint2 u = *reinterpret_cast<const uint2*>(v + (base + l) * 128 + lane * 4);
      a0 = fmaf(p, bf16_lo(u.x), a0);
      a1 = fmaf(p, bf16_hi(u.x), a1);
      a2 = fmaf(p, bf16_lo(u.y), a2);
      a3 = fmaf(p, bf16_hi(u.y), a3);
    }
  }
  osum[warp][lane * 4 + 0] = a0;
  osum[warp][lane * 4 + 1] = a1;
  osum[warp][lane * 4 + 2] = a2;
  osum[warp][lane * 4 + 3] = a3;
  __syncthreads();
  const float o = (osum[0][tid] + osum[1][tid]) + (osum[2][tid] + osum[3][tid]);
  out[static_cast<int64_t>(b) * heads * 128 + h * 128 + tid] = __float2bfloat16_rn(o);
}

__global__ void k4_advance_counter(int32_t* p, int by) { if (threadIdx.x == 0) p[0] += by; }

// scores live in dynamic shared memory: opt in beyond the 48 KB default once, bound by the 227 KB per-CTA limit
constexpr int DEC_MAX_L = (200 * 1024) / 4;

static int launch_decode(const void* q, int64_t ldq, const void* k, const void* v, const uint8_t* mask,
                         int64_t ld_mask, void* out, int B, int heads, int L, int cap, const int32_t* kv_len,
                         float scale, cudaStream_t s) {
  if (!q || !k || !v || !mask || !out || B <= 0 || heads <= 0 || L <= 0 || cap < L) return VEX_E_INVALID;
  if (B > 65535 || L > DEC_MAX_L) return VEX_E_UNSUPPORTED;
  static bool configured = false;
  if (!configured) {
    VEX_CUDA_TRY(cudaFuncSetAttribute(k4_attention_decode, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      DEC_MAX_L * 4));
    configured = true;
  }
  dim3 grid(heads, B);
  k4_attention_decode<<<grid, DEC_THREADS, L * sizeof(float), s>>>(
      static_cast<const __nv_bfloat16*>(q), ldq, static_cast<const __nv_bfloat16*>(k),
      static_cast<const __nv_bfloat16*>(v), mask, ld_mask, static_cast<__nv_bfloat16*>(out), heads, L, cap, kv_len,
      scale);
  VEX_LAUNCH_CHECK();
  return VEX_OK;
}

}  // namespace vex

extern "C" int vex_attention_decode(const void* q, int64_t ldq, const void* k, const void* v, const uint8_t* mask,
                                    void* out, int B, int heads, int L, float scale, vexStream stream) {
  return vex::launch_decode(q, ldq, k, v, mask, L, out, B, heads, L, L, nullptr, scale,
                            static_cast<cudaStream_t>(stream));
}

extern "C" int vex_attention_decode_cache(const void* q, int64_t ldq, const void* k_cache, const void* v_cache,
                                          const uint8_t* mask, int64_t ld_mask, int mask_len, void* out, int B,
                                          int heads, int kv_capacity, const int32_t* kv_len, float scale,
                                          vexStream stream) {
  if (!kv_len || mask_len <= 0 || ld_mask < mask_len || kv_capacity <= 0) return VEX_E_INVALID;
  const int l_max = mask_len < kv_capacity ? mask_len : kv_capacity;  // host bound on the positions attended to
  return vex::launch_decode(q, ldq, k_cache, v_cache, mask, ld_mask, out, B, heads, l_max, kv_capacity, kv_len, scale,
                            static_cast<cudaStream_t>(stream));
}

extern "C" int vex_advance_counter(int32_t* counter, int by, vexStream stream) {
  if (!counter) return VEX_E_INVALID;
  vex::k4_advance_counter<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(counter, by);
  VEX_LAUNCH_CHECK();
  return VEX_OK;
}
