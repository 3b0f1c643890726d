// K4 (baseline variant) -- causal block-diagonal varlen flash attention with mma.sync (HMMA).
//
// Restates the prefill branch of attention_fn (modeling_cogvlm.py:106-128): xformers'
// memory_efficient_attention under BlockDiagonalCausalMask -- per sample, token i attends to tokens j <= i of
// the same sample in token-rank order, scale = d^-0.5, fp32 online softmax, P rounded to bf16 before P.V.
// This register-accumulator kernel is the correctness baseline and the fallback shape handler; the tcgen05/TMEM
// kernels are the fast path (this file is built into libvex_baselines.so only, see abi_baselines.cu).
//
// CTA = 64 query rows x 1 head, 4 warps (16 rows each); K/V blocks of 64 keys double-buffered with cp.async;
// 16-byte-chunk XOR swizzle in shared memory so ldmatrix is conflict-free.  q/k/v are read in place from the
// token-order QKV buffer [T, 3, heads, 128]; output rows are scattered through out_row_map.
#include "../common.cuh"

namespace vex {

constexpr int AT_BQ = 64, AT_BK = 64, AT_D = 128, AT_THREADS = 128;
constexpr int AT_TILE_BYTES = 64 * AT_D * 2;  // 16 KB

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool pred) {
  const int sz = pred ? 16 : 0;  // src-size 0 -> zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// byte offset of (row, 16-byte chunk) in a [64][128] bf16 tile with XOR swizzle
__device__ __forceinline__ uint32_t sw_off(int row, int chunk) { return row * 256 + ((chunk ^ (row & 7)) << 4); }

// 64 rows x 128 columns of one of q/k/v for head h, rows [row0, row0 + 64) of the sample clipped at row_end
__device__ __forceinline__ void load_tile_async(uint32_t smem_base, const __nv_bfloat16* src, int64_t ld, int row0,
                                                int row_end) {
#pragma unroll
  for (int i = 0; i < (64 * 16) / AT_THREADS; ++i) {
    const int idx = i * AT_THREADS + threadIdx.x;
    const int r = idx >> 4, c = idx & 15;
    const bool ok = row0 + r < row_end;
    const __nv_bfloat16* g = src + static_cast<int64_t>(ok ? row0 + r : row0) * ld + c * 8;
    cp_async16(smem_base + sw_off(r, c), g, ok);
  }
}

__global__ void __launch_bounds__(AT_THREADS)
    k4_attention_mma(const __nv_bfloat16* __restrict__ qkv, const int32_t* __restrict__ cu_seqlens, int heads,
                     const int32_t* __restrict__ out_row_map, __nv_bfloat16* __restrict__ out, float scale_log2) {
  extern __shared__ __align__(128) uint8_t at_smem[];
  const int b = blockIdx.z, h = blockIdx.y;
  const int seq0 = cu_seqlens[b], len = cu_seqlens[b + 1] - seq0;
  const int qb = gridDim.x - 1 - blockIdx.x;  // heaviest (last) query blocks first
  const int q0 = qb * AT_BQ;
  if (q0 >= len) return;
  const int H = heads * AT_D;
  const int64_t ld = 3 * static_cast<int64_t>(H);
  const __nv_bfloat16* qp = qkv + static_cast<int64_t>(seq0) * ld + h * AT_D;
  const __nv_bfloat16* kp = qp + H;
  const __nv_bfloat16* vp = qp + 2 * H;

  const uint32_t sQ = smem_u32(at_smem);
  const uint32_t sK0 = sQ + AT_TILE_BYTES, sV0 = sK0 + 2 * AT_TILE_BYTES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;

  const int n_kv = qb + 1;  // causal: key blocks 0..qb
  load_tile_async(sQ, qp, ld, q0, len);
  load_tile_async(sK0, kp, ld, 0, len);
  load_tile_async(sV0, vp, ld, 0, len);
  cp_async_commit();

  uint32_t qf[8][4];  // Q fragments: 8 k-steps of 16 over d = 128
  float o[16][4];     // O accumulators: 16 n-tiles of 8 over d = 128
#pragma unroll
  for (int i = 0; i < 16; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};

  for (int j = 0; j < n_kv; ++j) {
    const int buf = j & 1;
    if (j + 1 < n_kv) {  // prefetch the next K/V block into the other buffer
      load_tile_async(sK0 + (buf ^ 1) * AT_TILE_BYTES, kp, ld, (j + 1) * AT_BK, len);
      load_tile_async(sV0 + (buf ^ 1) * AT_TILE_BYTES, vp, ld, (j + 1) * AT_BK, len);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (j == 0) {
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const int row = warp * 16 + (lane & 15), chunk = ks * 2 + (lane >> 4);
        ldsm_x4(sQ + sw_off(row, chunk), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
      }
    }
    const uint32_t sK = sK0 + buf * AT_TILE_BYTES, sV = sV0 + buf * AT_TILE_BYTES;

    // ---- S = Q K^T : 16 x 64 per warp ----
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {  // pairs of 8-key n-tiles
        const int row = np * 16 + ((lane >> 4) << 3) + (lane & 7);
        const int chunk = ks * 2 + ((lane >> 3) & 1);
        uint32_t b0, b1, b2, b3;
        ldsm_x4(sK + sw_off(row, chunk), b0, b1, b2, b3);
        mma_bf16(s[2 * np], qf[ks], b0, b1);
        mma_bf16(s[2 * np + 1], qf[ks], b2, b3);
      }
    }

    // ---- causal mask (diagonal block only; earlier blocks are fully visible) + online softmax ----
    const int qrow0 = q0 + warp * 16 + g;  // rows qrow0 and qrow0 + 8
    if (j == n_kv - 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int key = j * AT_BK + i * 8 + 2 * t;
        if (key > qrow0) s[i][0] = -INFINITY;
        if (key + 1 > qrow0) s[i][1] = -INFINITY;
        if (key > qrow0 + 8) s[i][2] = -INFINITY;
        if (key + 1 > qrow0 + 8) s[i][3] = -INFINITY;
      }
    }
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      mx[0] = fmaxf(mx[0], fmaxf(s[i][0], s[i][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[i][2], s[i][3]));
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    }
    float alpha[2], msc[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const float m_new = fmaxf(m_run[r], mx[r]);  // key 0 is always visible, so m_new is finite
      alpha[r] = exp2f((m_run[r] - m_new) * scale_log2);
      m_run[r] = m_new;
      msc[r] = m_new * scale_log2;
    }
    float rs[2] = {0.f, 0.f};
    uint32_t pf[4][4];  // P as A fragments: 4 k-steps of 16 keys
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float p0 = exp2f(fmaf(s[i][0], scale_log2, -msc[0]));
      const float p1 = exp2f(fmaf(s[i][1], scale_log2, -msc[0]));
      const float p2 = exp2f(fmaf(s[i][2], scale_log2, -msc[1]));
      const float p3 = exp2f(fmaf(s[i][3], scale_log2, -msc[1]));
      rs[0] += p0 + p1;
      rs[1] += p2 + p3;
      // accumulator layout of two adjacent n-tiles == A-fragment layout of one 16-wide k-step
      pf[i >> 1][(i & 1) * 2 + 0] = pack_bf16(p0, p1);
      pf[i >> 1][(i & 1) * 2 + 1] = pack_bf16(p2, p3);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * alpha[r] + rs[r];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      o[i][0] *= alpha[0];
      o[i][1] *= alpha[0];
      o[i][2] *= alpha[1];
      o[i][3] *= alpha[1];
    }

    // ---- O += P V ----
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {      // 16 keys per step
#pragma unroll
      for (int np = 0; np < 8; ++np) {    // pairs of 8-wide d tiles
        const int row = ks * 16 + (((lane >> 3) & 1) << 3) + (lane & 7);
        const int chunk = np * 2 + (lane >> 4);
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(sV + sw_off(row, chunk), b0, b1, b2, b3);
        mma_bf16(o[2 * np], pf[ks], b0, b1);
        mma_bf16(o[2 * np + 1], pf[ks], b2, b3);
      }
    }
    __syncthreads();  // everyone is done with this K/V buffer before it is refilled
  }

  // ---- finalise: O / l, stage through shared memory (Q tile is dead), coalesced scattered store ----
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const float inv0 = 1.0f / l_run[0], inv1 = 1.0f / l_run[1];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int r0 = warp * 16 + g;
    const uint32_t lo = pack_bf16(o[i][0] * inv0, o[i][1] * inv0);
    const uint32_t hi = pack_bf16(o[i][2] * inv1, o[i][3] * inv1);
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(sQ + sw_off(r0, i) + t * 4), "r"(lo) : "memory");
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(sQ + sw_off(r0 + 8, i) + t * 4), "r"(hi) : "memory");
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < (64 * 16) / AT_THREADS; ++i) {
    const int idx = i * AT_THREADS + threadIdx.x;
    const int r = idx >> 4, c = idx & 15;
    if (q0 + r < len) {
      uint4 v;
      asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                   : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                   : "r"(sQ + sw_off(r, c)));
      const int tok = seq0 + q0 + r;
      const int64_t dst = out_row_map ? out_row_map[tok] : tok;
      *reinterpret_cast<uint4*>(out + dst * H + h * AT_D + c * 8) = v;
    }
  }
}

int launch_attention_mma(const void* qkv, const int32_t* cu_seqlens, int B, int max_len_cap, int heads,
                         const int32_t* out_row_map, void* out, float scale, cudaStream_t s) {
  static bool configured = false;
  constexpr int smem = 5 * AT_TILE_BYTES;
  if (!configured) {
    VEX_CUDA_TRY(cudaFuncSetAttribute(k4_attention_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  dim3 grid(ceil_div(max_len_cap, AT_BQ), heads, B);
  const float scale_log2 = scale * 1.4426950408889634f;
  k4_attention_mma<<<grid, AT_THREADS, smem, s>>>(static_cast<const __nv_bfloat16*>(qkv), cu_seqlens, heads,
                                                  out_row_map, static_cast<__nv_bfloat16*>(out), scale_log2);
  VEX_LAUNCH_CHECK();
  return VEX_OK;
}

}  // namespace vex
