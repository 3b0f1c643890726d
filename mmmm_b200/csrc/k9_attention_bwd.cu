// K9 -- backward of the causal block-diagonal (varlen) attention on tcgen05 / TMEM, fused with the adjoint of the
// rotary embedding and the scatter back to expert-sorted rows.
//
// Adjoint of attention_fn's prefill branch (modeling_cogvlm.py:106-128) and of apply_rotary_pos_emb_index_bhs
// (:188-193) -- what torch.autograd runs for those lines in the reference's LoRA training step
// (mmmm/models/mmmm.py:299-306).  With P = softmax(scale * Q K^T + causal), O = P V, delta = rowsum(dO * O):
//     dV = P^T dO        dP = dO V^T        dS = P * (dP - delta)        dQ = scale * dS K        dK = scale * dS^T Q
// P is recomputed from the log-sum-exp the forward kernel stores (log2 domain).  Two kernels, no atomics:
//   k9_attn_bwd_dkdv : CTA = 128 keys of one (sample, head); loops over 64-query steps at or below the diagonal.
//       S^T = K Q^T and dP^T = V dO^T (M = 128 keys, N = 64 queries, K-major operands) land in double-buffered
//       TMEM; one softmax thread per KEY row turns them into P^T and dS^T (bf16, shared memory, K-major);
//       dV += P^T dO and dK += dS^T Q accumulate in TMEM with dO / Q read as MN-major B operands from the very
//       tiles TMA delivered.  512 TMEM columns: S^T x2, dP^T x2 (64 each), dV, dK (128 each).
//   k9_attn_bwd_dq   : CTA = 128 queries; loops over 64-key steps; S = Q K^T, dP = dO V^T (N = 64), one softmax
//       thread per QUERY row (its lse / delta are scalars), dQ += dS K with K as an MN-major B operand.
// Epilogues: dQ and dK are multiplied by `scale`, pushed through the transpose of the rotary rotation with the
// same bf16 tables the forward used, and all three gradients are written to the row token_to_sorted[t] of
// dqkv [rows_cap, 3 * heads * 128] -- the A operand of the QKV dgrad GEMM.
#include <cuda.h>

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace vex {

int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);

// rows [T, T + 128) of a token-order buffer can be touched by the last tile: keep them finite (0 * NaN = NaN)
__global__ void k9_zero_tail_rows(__nv_bfloat16* buf, const int32_t* __restrict__ cu_seqlens, int B, int rows_cap,
                                  int row_elems) {
  const int T = cu_seqlens[B];
  const int n_rows = min(rows_cap - T, 128);
  const int64_t n_vec = static_cast<int64_t>(max(n_rows, 0)) * (row_elems / 8);
  uint4* p = reinterpret_cast<uint4*>(buf + static_cast<int64_t>(T) * row_elems);
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n_vec;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    p[i] = make_uint4(0, 0, 0, 0);
}

constexpr int AB_THREADS = 384;  // TMA, MMA, TMEM-alloc, idle + 8 softmax warps (2 threads per row)
constexpr int AB_T128 = 128 * 128 * 2;  // 32 KB tile of 128 rows x 128 d: two 16 KB atoms (64 d each)
constexpr int AB_T64 = 64 * 128 * 2;    // 16 KB tile of 64 rows x 128 d: two 8 KB chunks
constexpr int AB_P = 128 * 64 * 2;      // 16 KB: 128 rows x 64 bf16 (one 128-byte swizzle row each)

struct BwdParams {
  const int32_t* cu_seqlens;
  const float* lse;     // [heads, rows_cap] log2-domain log-sum-exp (forward)
  const float* delta;   // [heads, rows_cap] rowsum(dO * O)
  const int32_t* token_to_sorted;
  const int32_t* token_to_flat;
  const int64_t* position_ids;
  const __nv_bfloat16* rope_cos;
  const __nv_bfloat16* rope_sin;
  __nv_bfloat16* dqkv;
  int heads, rows_cap, rope_len;
  float scale, scale_log2;
};

// ---------------------------------------------------------------------------------------------
// delta[h, t] = sum_d dO[t, h, d] * O[token_to_sorted[t], h, d]   (one warp per token, 16 lanes per head)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    k9_attn_delta(const __nv_bfloat16* __restrict__ d_out_tok, const __nv_bfloat16* __restrict__ out_sorted,
                  const int32_t* __restrict__ token_to_sorted, const int32_t* __restrict__ cu_seqlens, int B,
                  float* __restrict__ delta, int heads, int rows_cap) {
  const int T = min(cu_seqlens[B], rows_cap);
  const int H = heads * 128;
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
  for (int t = warp; t < T; t += n_warps) {
    const int src = token_to_sorted ? token_to_sorted[t] : t;
    const uint4* dp = reinterpret_cast<const uint4*>(d_out_tok + static_cast<int64_t>(t) * H);
    const uint4* op = reinterpret_cast<const uint4*>(out_sorted + static_cast<int64_t>(src) * H);
    for (int h0 = 0; h0 < heads; h0 += 2) {
      const int head = h0 + (lane >> 4);
      float acc = 0.f;
      if (head < heads) {
        const uint4 a = ld_stream(dp + head * 16 + (lane & 15)), b = ld_stream(op + head * 16 + (lane & 15));
        const uint32_t au[4] = {a.x, a.y, a.z, a.w}, bu[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc = fmaf(bf16_lo(au[j]), bf16_lo(bu[j]), acc);
          acc = fmaf(bf16_hi(au[j]), bf16_hi(bu[j]), acc);
        }
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if ((lane & 15) == 0 && head < heads) delta[static_cast<int64_t>(head) * rows_cap + t] = acc;
    }
  }
}

// 32 fp32 values -> 16 packed bf16 pairs
__device__ __forceinline__ void pack32(const float (&v)[32], uint32_t (&o)[16]) {
#pragma unroll
  for (int i = 0; i < 16; ++i) o[i] = pack_bf16(v[2 * i], v[2 * i + 1]);
}

// one 128-byte row (64 bf16 = 8 x 16 B) of a K-major SWIZZLE_128B tile; `half` selects columns [32*half, +32)
__device__ __forceinline__ void store_row_half(uint32_t tile, int row, int half, const uint32_t (&o)[16]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int chunk = half * 4 + i;
    const uint32_t addr = tile + row * 128 + ((chunk ^ (row & 7)) << 4);
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(o[4 * i]), "r"(o[4 * i + 1]),
                 "r"(o[4 * i + 2]), "r"(o[4 * i + 3])
                 : "memory");
  }
}

__device__ __forceinline__ uint4 ld_shared_v4_k9(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
  return r;
}

__device__ __forceinline__ void named_bar_sync(int id, int n) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}

// gradient row (128 fp32 in TMEM at taddr) -> optional rotary transpose * mul -> bf16 -> dst[0..128)
//   forward: y[j] = x[j] c[j] - x[j+64] s[j],  y[j+64] = x[j+64] c[j+64] + x[j] s[j+64]          (j < 64)
//   adjoint: dx[j] = dy[j] c[j] + dy[j+64] s[j+64],  dx[j+64] = dy[j+64] c[j+64] - dy[j] s[j]
__device__ __forceinline__ void ld_bf16x32(const __nv_bfloat16* p, float (&o)[32]) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint4 u = __ldg(q + i);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      o[8 * i + 2 * j] = bf16_lo(w[j]);
      o[8 * i + 2 * j + 1] = bf16_hi(w[j]);
    }
  }
}

__device__ __forceinline__ void st_packed32(__nv_bfloat16* dst, const uint32_t (&pk)[16]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
    *reinterpret_cast<uint4*>(dst + i * 8) = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
}

__device__ __forceinline__ void store_grad_row(uint32_t taddr, bool rope, const __nv_bfloat16* cos_row,
                                               const __nv_bfloat16* sin_row, float mul, __nv_bfloat16* dst,
                                               bool valid, int q_begin, int q_end) {
#pragma unroll 1
  for (int q = q_begin; q < q_end; ++q) {
    uint32_t lo[32], hi[32];
    tmem_ld_32x32b_x32(taddr + q * 32, lo);       // columns [32q, 32q + 32)
    tmem_ld_32x32b_x32(taddr + 64 + q * 32, hi);  // their rotary partners, + 64
    tmem_ld_wait();
    if (valid) {
      float o[32];
      uint32_t pk[16];
      if (rope) {
        float c[32], sn[32];
        ld_bf16x32(cos_row + q * 32, c);
        ld_bf16x32(sin_row + 64 + q * 32, sn);
#pragma unroll
        for (int j = 0; j < 32; ++j) o[j] = mul * (__uint_as_float(lo[j]) * c[j] + __uint_as_float(hi[j]) * sn[j]);
        pack32(o, pk);
        st_packed32(dst + q * 32, pk);
        ld_bf16x32(cos_row + 64 + q * 32, c);
        ld_bf16x32(sin_row + q * 32, sn);
#pragma unroll
        for (int j = 0; j < 32; ++j) o[j] = mul * (__uint_as_float(hi[j]) * c[j] - __uint_as_float(lo[j]) * sn[j]);
        pack32(o, pk);
        st_packed32(dst + 64 + q * 32, pk);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) o[j] = mul * __uint_as_float(lo[j]);
        pack32(o, pk);
        st_packed32(dst + q * 32, pk);
#pragma unroll
        for (int j = 0; j < 32; ++j) o[j] = mul * __uint_as_float(hi[j]);
        pack32(o, pk);
        st_packed32(dst + 64 + q * 32, pk);
      }
    }
    __syncwarp();
  }
}

#ifdef VEX_ATTN_TRACE
// timing experiment build (tools/attn_trace.py k9): cycle stamps of CTA (0, 0, 0) of the dK/dV kernel
__device__ long long* g_k9_trace = nullptr;  // grid kernel: [64 steps][8 stamps], row 63 = CTA-level stamps; dq_p: [256][16]
extern "C" int vex_debug_k9_trace(long long* buf) {
  return cudaMemcpyToSymbol(g_k9_trace, &buf, sizeof(buf)) == cudaSuccess ? 0 : -1;
}
#define K9_TRACE(ptr, j, k)                                   \
  do {                                                        \
    if ((ptr) && (j) < 64) (ptr)[(j) * 8 + (k)] = clock64();  \
  } while (0)
#else
#define K9_TRACE(ptr, j, k) \
  do {                      \
  } while (0)
#endif

// =============================================================================================
// dK, dV
// =============================================================================================
struct BarsA {
  uint64_t kv_full, q_full[2], q_empty[2], s_full[2], p_full[2], p_empty[2], acc_full;
  uint32_t tmem_base;
};
constexpr int ABA_SMEM = 2 * AB_T128 + 4 * AB_T64 + 4 * AB_P + 2 * 128 * 4 + 256 + 1024;

__global__ void __launch_bounds__(AB_THREADS, 1)
    k9_attn_bwd_dkdv(const __grid_constant__ CUtensorMap tm_qkv128, const __grid_constant__ CUtensorMap tm_qkv64,
                     const __grid_constant__ CUtensorMap tm_do64, const BwdParams p) {
  const int b = blockIdx.z, h = blockIdx.y, jb = blockIdx.x;
#ifdef VEX_ATTN_TRACE
  long long* trc = (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 128) ? g_k9_trace : nullptr;
  K9_TRACE(trc, 63, 0);
#endif
  const int seq0 = p.cu_seqlens[b], len = p.cu_seqlens[b + 1] - seq0;
  const int kv0 = jb * 128;
  if (kv0 >= len) return;
  const int i0 = 2 * jb;                        // first 64-query step touching this key block
  const int n_steps = (len + 63) / 64 - i0;     // >= 1
  const int H = p.heads * 128;

  extern __shared__ uint8_t ab_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ab_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = smem + AB_T128;
  uint8_t* sQ = smem + 2 * AB_T128;              // 2 stages of AB_T64
  uint8_t* sdO = sQ + 2 * AB_T64;                // 2 stages
  uint8_t* sP = sdO + 2 * AB_T64;                // 2 stages of AB_P
  uint8_t* sdS = sP + 2 * AB_P;                  // 2 stages
  float* sStat = reinterpret_cast<float*>(sdS + 2 * AB_P);  // [2][128]: lse (64) | delta (64) of the step's queries
  BarsA* bars = reinterpret_cast<BarsA*>(sStat + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bars->kv_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->q_full[i], 1);
      mbar_init(&bars->q_empty[i], 1);
      mbar_init(&bars->s_full[i], 1);
      mbar_init(&bars->p_full[i], 8);
      mbar_init(&bars->p_empty[i], 1);
    }
    mbar_init(&bars->acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(&bars->tmem_base, 512);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_qkv128);
    tma_prefetch_desc(&tm_qkv64);
    tma_prefetch_desc(&tm_do64);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
  const uint32_t tS = tmem, tdP = tmem + 128, tdV = tmem + 256, tdK = tmem + 384;

  if (warp == 0 && lane == 0) {
    // =============================== TMA producer ===============================
    const int colq = h * 128, colk = H + h * 128, colv = 2 * H + h * 128;
    mbar_arrive_expect_tx(&bars->kv_full, 2 * AB_T128);
    tma_load_2d(sK, &tm_qkv128, &bars->kv_full, colk, seq0 + kv0);
    tma_load_2d(sK + AB_T128 / 2, &tm_qkv128, &bars->kv_full, colk + 64, seq0 + kv0);
    tma_load_2d(sV, &tm_qkv128, &bars->kv_full, colv, seq0 + kv0);
    tma_load_2d(sV + AB_T128 / 2, &tm_qkv128, &bars->kv_full, colv + 64, seq0 + kv0);
    for (int s = 0; s < n_steps; ++s) {
      const int st = s & 1;
      const uint32_t ph = (s >> 1) & 1;
      const int row = seq0 + (i0 + s) * 64;
      mbar_wait(&bars->q_empty[st], ph ^ 1);
      mbar_arrive_expect_tx(&bars->q_full[st], 2 * AB_T64);
      tma_load_2d(sQ + st * AB_T64, &tm_qkv64, &bars->q_full[st], colq, row);
      tma_load_2d(sQ + st * AB_T64 + AB_T64 / 2, &tm_qkv64, &bars->q_full[st], colq + 64, row);
      tma_load_2d(sdO + st * AB_T64, &tm_do64, &bars->q_full[st], colq, row);
      tma_load_2d(sdO + st * AB_T64 + AB_T64 / 2, &tm_do64, &bars->q_full[st], colq + 64, row);
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    // The whole warp runs the loop (warp-uniform control flow) and one elected lane issues: in a `lane == 0` branch
    // ptxas wraps every tcgen05.mma in an elect / waterfall loop with four R2UR moves and rebuilds both descriptors
    // (17 instructions per 32-cycle N = 64 MMA: the issuer, not the tensor pipe, set the pace).  Descriptors are one
    // base per operand tile plus constant k-step offsets in the 14-bit address field.
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 64, 0, 0);
    constexpr uint32_t idesc_acc = umma_idesc_bf16(128, 128, 0, 1);  // B MN-major
    const uint64_t dK = umma_desc_kmajor_sw128(smem_u32(sK)), dV = umma_desc_kmajor_sw128(smem_u32(sV));
    const uint64_t dQk0 = umma_desc_kmajor_sw128(smem_u32(sQ));  // stage st: + st * AB_T64 / 16
    const uint64_t ddOk0 = umma_desc_kmajor_sw128(smem_u32(sdO));  // stage st: + st * AB_T64 / 16
    const uint64_t dQm0 = umma_desc_mnmajor_sw128(smem_u32(sQ), AB_T64 / 2, 1024);  // stage st: + st * AB_T64 / 16
    const uint64_t ddOm0 = umma_desc_mnmajor_sw128(smem_u32(sdO), AB_T64 / 2, 1024);  // stage st: + st * AB_T64 / 16
    const uint64_t dPk0 = umma_desc_kmajor_sw128(smem_u32(sP));  // stage st: + st * AB_P / 16
    const uint64_t ddSk0 = umma_desc_kmajor_sw128(smem_u32(sdS));  // stage st: + st * AB_P / 16
    auto issue_s = [&](int s) {
      const int st = s & 1;
      mbar_wait(&bars->q_full[st], (s >> 1) & 1);
      tc_fence_after();
      if (elect_one_sync()) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)  // d = 128 in K = 16 steps
          umma_ss(tS + st * 64, dK + static_cast<uint64_t>(((kk >> 2) * (AB_T128 / 2) + (kk & 3) * 32) >> 4),
                  (dQk0 + static_cast<uint64_t>(st * (AB_T64 >> 4))) + static_cast<uint64_t>(((kk >> 2) * (AB_T64 / 2) + (kk & 3) * 32) >> 4), idesc_s, kk > 0);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_ss(tdP + st * 64, dV + static_cast<uint64_t>(((kk >> 2) * (AB_T128 / 2) + (kk & 3) * 32) >> 4),
                  (ddOk0 + static_cast<uint64_t>(st * (AB_T64 >> 4))) + static_cast<uint64_t>(((kk >> 2) * (AB_T64 / 2) + (kk & 3) * 32) >> 4), idesc_s, kk > 0);
        umma_commit(&bars->s_full[st]);
      }
      __syncwarp();
    };
    mbar_wait(&bars->kv_full, 0);
    issue_s(0);
    for (int s = 0; s < n_steps; ++s) {
      if (s + 1 < n_steps) issue_s(s + 1);
      const int st = s & 1;
      mbar_wait(&bars->p_full[st], (s >> 1) & 1);
      tc_fence_after();
      if (elect_one_sync()) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)  // 64 queries in K = 16 steps
          umma_ss(tdV, (dPk0 + static_cast<uint64_t>(st * (AB_P >> 4))) + static_cast<uint64_t>((kk * 32) >> 4), (ddOm0 + static_cast<uint64_t>(st * (AB_T64 >> 4))) + static_cast<uint64_t>((kk * 2048) >> 4),
                  idesc_acc, (s > 0) || (kk > 0));
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_ss(tdK, (ddSk0 + static_cast<uint64_t>(st * (AB_P >> 4))) + static_cast<uint64_t>((kk * 32) >> 4), (dQm0 + static_cast<uint64_t>(st * (AB_T64 >> 4))) + static_cast<uint64_t>((kk * 2048) >> 4),
                  idesc_acc, (s > 0) || (kk > 0));
        umma_commit(&bars->q_empty[st]);
        umma_commit(&bars->p_empty[st]);
      }
      __syncwarp();
    }
    if (elect_one_sync()) umma_commit(&bars->acc_full);
    __syncwarp();
  } else if (warp >= 4) {
    // =============================== softmax threads: two per KEY row (32 queries each) =====================
    const int sw = warp - 4;           // 0..7
    const int ew = sw & 3;             // TMEM lane quarter
    const int half = sw >> 2;          // which 32 of the step's 64 query columns
    const int c = ew * 32 + lane;      // key row inside the block == TMEM lane
    const int sid = sw * 32 + lane;    // 0..255
    const uint32_t lane_sel = static_cast<uint32_t>(ew * 32) << 16;
    const int kv_idx = kv0 + c;
    const float* lse_h = p.lse + static_cast<int64_t>(h) * p.rows_cap;
    const float* delta_h = p.delta + static_cast<int64_t>(h) * p.rows_cap;
    // lse (threads 0..63) / delta (threads 64..127) of the 64 queries of a step, prefetched one step ahead
    auto load_stat = [&](int s) {
      float v = 0.f;
      if (sid < 128 && s < n_steps) {
        const int qi = (i0 + s) * 64 + (sid & 63);
        if (qi < len) v = (sid < 64 ? lse_h : delta_h)[seq0 + qi];
      }
      return v;
    };
    float stat_next = load_stat(0);
    K9_TRACE(trc, 63, 1);
    for (int s = 0; s < n_steps; ++s) {
      const int st = s & 1;
      const uint32_t ph = (s >> 1) & 1;
      const int q_base = (i0 + s) * 64;
      K9_TRACE(trc, s, 0);
      if (sid < 128) sStat[st * 128 + sid] = stat_next;
      stat_next = load_stat(s + 1);
      named_bar_sync(1, 256);
      K9_TRACE(trc, s, 1);
      mbar_wait(&bars->s_full[st], ph);
      tc_fence_after();
      K9_TRACE(trc, s, 2);
      uint32_t sraw[32], draw[32];
      tmem_ld_32x32b_x32(tS + lane_sel + st * 64 + half * 32, sraw);
      tmem_ld_32x32b_x32(tdP + lane_sel + st * 64 + half * 32, draw);
      tmem_ld_wait();
      const float4* l4 = reinterpret_cast<const float4*>(sStat + st * 128 + half * 32);
      const float4* d4 = reinterpret_cast<const float4*>(sStat + st * 128 + 64 + half * 32);
      uint32_t pk[16], dk[16];
      // interior steps (every query at or after every key of the block, all inside the sample) need no masks
      const bool interior = (q_base >= kv0 + 127) && (q_base + 64 <= len);
      if (interior) {
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 l = l4[j4], d = d4[j4];
          const float lv[4] = {l.x, l.y, l.z, l.w}, dv[4] = {d.x, d.y, d.z, d.w};
          float pr[4], ds[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            pr[k] = ex2_approx(fmaf(__uint_as_float(sraw[4 * j4 + k]), p.scale_log2, -lv[k]));
            ds[k] = pr[k] * (__uint_as_float(draw[4 * j4 + k]) - dv[k]);
          }
          pk[2 * j4] = pack_bf16(pr[0], pr[1]);
          pk[2 * j4 + 1] = pack_bf16(pr[2], pr[3]);
          dk[2 * j4] = pack_bf16(ds[0], ds[1]);
          dk[2 * j4 + 1] = pack_bf16(ds[2], ds[3]);
        }
      } else {
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 l = l4[j4], d = d4[j4];
          const float lv[4] = {l.x, l.y, l.z, l.w}, dv[4] = {d.x, d.y, d.z, d.w};
          float pr[4], ds[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int q_idx = q_base + half * 32 + 4 * j4 + k;
            const bool keep = (q_idx >= kv_idx) && (q_idx < len);
            pr[k] = keep ? ex2_approx(fmaf(__uint_as_float(sraw[4 * j4 + k]), p.scale_log2, -lv[k])) : 0.f;
            ds[k] = keep ? pr[k] * (__uint_as_float(draw[4 * j4 + k]) - dv[k]) : 0.f;
          }
          pk[2 * j4] = pack_bf16(pr[0], pr[1]);
          pk[2 * j4 + 1] = pack_bf16(pr[2], pr[3]);
          dk[2 * j4] = pack_bf16(ds[0], ds[1]);
          dk[2 * j4 + 1] = pack_bf16(ds[2], ds[3]);
        }
      }
      K9_TRACE(trc, s, 3);
      mbar_wait(&bars->p_empty[st], ph ^ 1);  // the MMAs of step s - 2 have finished reading this P / dS buffer
      K9_TRACE(trc, s, 4);
      store_row_half(smem_u32(sP + st * AB_P), c, half, pk);
      store_row_half(smem_u32(sdS + st * AB_P), c, half, dk);
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->p_full[st]);
      K9_TRACE(trc, s, 5);
    }
    // ---- epilogue: dV (plain), dK (scale, rotary transpose) -> dqkv[token_to_sorted[tok]] ----
    K9_TRACE(trc, 63, 2);
    mbar_wait(&bars->acc_full, 0);
    tc_fence_after();
    K9_TRACE(trc, 63, 3);
    const bool valid = kv_idx < len;
    const int tok = seq0 + kv_idx;
    int dst = 0, pos = 0;
    if (valid) {
      dst = p.token_to_sorted ? p.token_to_sorted[tok] : tok;
      const int64_t pz = p.position_ids[p.token_to_flat ? p.token_to_flat[tok] : tok];
      pos = static_cast<int>(min(max(pz, int64_t(0)), int64_t(p.rope_len - 1)));
    }
    __nv_bfloat16* row = p.dqkv + static_cast<int64_t>(dst) * (3 * H) + h * 128;
    if (half == 0)
      store_grad_row(tdV + lane_sel, false, nullptr, nullptr, 1.0f, row + 2 * H, valid, 0, 2);
    else
      store_grad_row(tdK + lane_sel, true, p.rope_cos + static_cast<int64_t>(pos) * 128,
                     p.rope_sin + static_cast<int64_t>(pos) * 128, p.scale, row + H, valid, 0, 2);
    K9_TRACE(trc, 63, 4);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// =============================================================================================
// dQ
// =============================================================================================
struct BarsB {
  uint64_t q_full, kv_full[2], kv_empty[2], s_full[2], p_full[2], p_empty[2], acc_full;
  uint32_t tmem_base;
};
constexpr int ABB_SMEM = 2 * AB_T128 + 4 * AB_T64 + 2 * AB_P + 256 + 1024;

__global__ void __launch_bounds__(AB_THREADS, 1)
    k9_attn_bwd_dq(const __grid_constant__ CUtensorMap tm_qkv128, const __grid_constant__ CUtensorMap tm_qkv64,
                   const __grid_constant__ CUtensorMap tm_do128, const BwdParams p) {
  const int b = blockIdx.z, h = blockIdx.y;
  const int seq0 = p.cu_seqlens[b], len = p.cu_seqlens[b + 1] - seq0;
  const int qb = gridDim.x - 1 - blockIdx.x;  // heaviest query blocks first
  const int q0 = qb * 128;
  if (q0 >= len) return;
  const int n_steps = min((len + 63) / 64, 2 * qb + 2);  // 64-key steps at or below the diagonal
  const int H = p.heads * 128;

  extern __shared__ uint8_t ab_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ab_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sdO = smem + AB_T128;
  uint8_t* sK = smem + 2 * AB_T128;  // 2 stages of AB_T64
  uint8_t* sV = sK + 2 * AB_T64;     // 2 stages
  uint8_t* sdS = sV + 2 * AB_T64;    // 2 stages of AB_P
  BarsB* bars = reinterpret_cast<BarsB*>(sdS + 2 * AB_P);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bars->q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->kv_full[i], 1);
      mbar_init(&bars->kv_empty[i], 1);
      mbar_init(&bars->s_full[i], 1);
      mbar_init(&bars->p_full[i], 8);
      mbar_init(&bars->p_empty[i], 1);
    }
    mbar_init(&bars->acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(&bars->tmem_base, 512);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_qkv128);
    tma_prefetch_desc(&tm_qkv64);
    tma_prefetch_desc(&tm_do128);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
  const uint32_t tS = tmem, tdP = tmem + 128, tdQ = tmem + 256;

  if (warp == 0 && lane == 0) {
    // =============================== TMA producer ===============================
    const int colq = h * 128, colk = H + h * 128, colv = 2 * H + h * 128;
    mbar_arrive_expect_tx(&bars->q_full, 2 * AB_T128);
    tma_load_2d(sQ, &tm_qkv128, &bars->q_full, colq, seq0 + q0);
    tma_load_2d(sQ + AB_T128 / 2, &tm_qkv128, &bars->q_full, colq + 64, seq0 + q0);
    tma_load_2d(sdO, &tm_do128, &bars->q_full, colq, seq0 + q0);
    tma_load_2d(sdO + AB_T128 / 2, &tm_do128, &bars->q_full, colq + 64, seq0 + q0);
    for (int s = 0; s < n_steps; ++s) {
      const int st = s & 1;
      const uint32_t ph = (s >> 1) & 1;
      const int row = seq0 + s * 64;
      mbar_wait(&bars->kv_empty[st], ph ^ 1);
      mbar_arrive_expect_tx(&bars->kv_full[st], 2 * AB_T64);
      tma_load_2d(sK + st * AB_T64, &tm_qkv64, &bars->kv_full[st], colk, row);
      tma_load_2d(sK + st * AB_T64 + AB_T64 / 2, &tm_qkv64, &bars->kv_full[st], colk + 64, row);
      tma_load_2d(sV + st * AB_T64, &tm_qkv64, &bars->kv_full[st], colv, row);
      tma_load_2d(sV + st * AB_T64 + AB_T64 / 2, &tm_qkv64, &bars->kv_full[st], colv + 64, row);
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (whole warp, elected lane issues; see k9_attn_bwd_dkdv) ===============
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 64, 0, 0);
    constexpr uint32_t idesc_acc = umma_idesc_bf16(128, 128, 0, 1);
    const uint64_t dQ = umma_desc_kmajor_sw128(smem_u32(sQ)), ddO = umma_desc_kmajor_sw128(smem_u32(sdO));
    const uint64_t dKk0 = umma_desc_kmajor_sw128(smem_u32(sK));  // stage st: + st * AB_T64 / 16
    const uint64_t dVk0 = umma_desc_kmajor_sw128(smem_u32(sV));  // stage st: + st * AB_T64 / 16
    const uint64_t dKm0 = umma_desc_mnmajor_sw128(smem_u32(sK), AB_T64 / 2, 1024);  // stage st: + st * AB_T64 / 16
    const uint64_t ddSk0 = umma_desc_kmajor_sw128(smem_u32(sdS));  // stage st: + st * AB_P / 16
    auto issue_s = [&](int s) {
      const int st = s & 1;
      mbar_wait(&bars->kv_full[st], (s >> 1) & 1);
      tc_fence_after();
      if (elect_one_sync()) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_ss(tS + st * 64, dQ + static_cast<uint64_t>(((kk >> 2) * (AB_T128 / 2) + (kk & 3) * 32) >> 4),
                  (dKk0 + static_cast<uint64_t>(st * (AB_T64 >> 4))) + static_cast<uint64_t>(((kk >> 2) * (AB_T64 / 2) + (kk & 3) * 32) >> 4), idesc_s, kk > 0);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_ss(tdP + st * 64, ddO + static_cast<uint64_t>(((kk >> 2) * (AB_T128 / 2) + (kk & 3) * 32) >> 4),
                  (dVk0 + static_cast<uint64_t>(st * (AB_T64 >> 4))) + static_cast<uint64_t>(((kk >> 2) * (AB_T64 / 2) + (kk & 3) * 32) >> 4), idesc_s, kk > 0);
        umma_commit(&bars->s_full[st]);
      }
      __syncwarp();
    };
    mbar_wait(&bars->q_full, 0);
    issue_s(0);
    for (int s = 0; s < n_steps; ++s) {
      if (s + 1 < n_steps) issue_s(s + 1);
      const int st = s & 1;
      mbar_wait(&bars->p_full[st], (s >> 1) & 1);
      tc_fence_after();
      if (elect_one_sync()) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)  // 64 keys in K = 16 steps
          umma_ss(tdQ, (ddSk0 + static_cast<uint64_t>(st * (AB_P >> 4))) + static_cast<uint64_t>((kk * 32) >> 4), (dKm0 + static_cast<uint64_t>(st * (AB_T64 >> 4))) + static_cast<uint64_t>((kk * 2048) >> 4),
                  idesc_acc, (s > 0) || (kk > 0));
        umma_commit(&bars->kv_empty[st]);
        umma_commit(&bars->p_empty[st]);
      }
      __syncwarp();
    }
    if (elect_one_sync()) umma_commit(&bars->acc_full);
    __syncwarp();
  } else if (warp >= 4) {
    // =============================== softmax threads: two per QUERY row (32 keys each) ======================
    const int sw = warp - 4;
    const int ew = sw & 3;
    const int half = sw >> 2;
    const int c = ew * 32 + lane;
    const uint32_t lane_sel = static_cast<uint32_t>(ew * 32) << 16;
    const int q_idx = q0 + c;
    const bool valid = q_idx < len;
    const int tok = seq0 + q_idx;
    float l2 = 0.f, dl = 0.f;
    if (valid) {
      l2 = p.lse[static_cast<int64_t>(h) * p.rows_cap + tok];
      dl = p.delta[static_cast<int64_t>(h) * p.rows_cap + tok];
    }
    for (int s = 0; s < n_steps; ++s) {
      const int st = s & 1;
      const uint32_t ph = (s >> 1) & 1;
      mbar_wait(&bars->s_full[st], ph);
      tc_fence_after();
      uint32_t sraw[32], draw[32];
      tmem_ld_32x32b_x32(tS + lane_sel + st * 64 + half * 32, sraw);
      tmem_ld_32x32b_x32(tdP + lane_sel + st * 64 + half * 32, draw);
      tmem_ld_wait();
      uint32_t dk[16];
      // interior steps: every key of the step is at or before every query of the block, block inside the sample
      const bool interior = (s * 64 + 63 <= q0) && (q0 + 128 <= len);
      if (interior) {
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float p0 = ex2_approx(fmaf(__uint_as_float(sraw[j]), p.scale_log2, -l2));
          const float p1 = ex2_approx(fmaf(__uint_as_float(sraw[j + 1]), p.scale_log2, -l2));
          dk[j >> 1] = pack_bf16(p0 * (__uint_as_float(draw[j]) - dl), p1 * (__uint_as_float(draw[j + 1]) - dl));
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const int kv_idx = s * 64 + half * 32 + j;
          const bool k0 = valid && (kv_idx <= q_idx), k1 = valid && (kv_idx + 1 <= q_idx);
          const float p0 = k0 ? ex2_approx(fmaf(__uint_as_float(sraw[j]), p.scale_log2, -l2)) : 0.f;
          const float p1 = k1 ? ex2_approx(fmaf(__uint_as_float(sraw[j + 1]), p.scale_log2, -l2)) : 0.f;
          dk[j >> 1] = pack_bf16(k0 ? p0 * (__uint_as_float(draw[j]) - dl) : 0.f,
                                 k1 ? p1 * (__uint_as_float(draw[j + 1]) - dl) : 0.f);
        }
      }
      mbar_wait(&bars->p_empty[st], ph ^ 1);
      store_row_half(smem_u32(sdS + st * AB_P), c, half, dk);
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->p_full[st]);
    }
    // ---- epilogue: dQ * scale through the rotary transpose -> dqkv[token_to_sorted[tok]] ----
    mbar_wait(&bars->acc_full, 0);
    tc_fence_after();
    int dst = 0, pos = 0;
    if (valid) {
      dst = p.token_to_sorted ? p.token_to_sorted[tok] : tok;
      const int64_t pz = p.position_ids[p.token_to_flat ? p.token_to_flat[tok] : tok];
      pos = static_cast<int>(min(max(pz, int64_t(0)), int64_t(p.rope_len - 1)));
    }
    store_grad_row(tdQ + lane_sel, true, p.rope_cos + static_cast<int64_t>(pos) * 128,
                   p.rope_sin + static_cast<int64_t>(pos) * 128, p.scale,
                   p.dqkv + static_cast<int64_t>(dst) * (3 * H) + h * 128, valid, half, half + 1);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}


// Early-release epilogue of the persistent kernels: the gradient row is first pulled out of TMEM into packed bf16
// registers (what eager autograd holds at this point: the bf16 gradient w.r.t. the rotated q / k), `release()` hands the
// accumulator columns back to the MMA warp, and only then the rotary transpose and the global stores run -- so the next
// item's first dV / dK (dQ) MMAs do not wait for ~7 000 cycles of table loads and stores.
//   NQ = 2: the whole 128-column row (dK / dV threads);  NQ = 1: columns [32 q0, 32 q0 + 32) and their + 64 partners (dQ)
template <int NQ, class Release>
__device__ __forceinline__ void store_grad_row_early(uint32_t taddr, bool rope, const __nv_bfloat16* cos_row,
                                                     const __nv_bfloat16* sin_row, float mul, __nv_bfloat16* dst,
                                                     bool valid, int q0, Release release) {
  uint32_t lo[NQ][16], hi[NQ][16];  // packed bf16 pairs: columns 32 q + 2 i, + 1 and their rotary partners (+ 64)
#pragma unroll
  for (int qq = 0; qq < NQ; ++qq) {
    const int q = q0 + qq;
    uint32_t a[32], b[32];
    tmem_ld_32x32b_x32(taddr + q * 32, a);
    tmem_ld_32x32b_x32(taddr + 64 + q * 32, b);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      lo[qq][i] = pack_bf16(__uint_as_float(a[2 * i]) * mul, __uint_as_float(a[2 * i + 1]) * mul);
      hi[qq][i] = pack_bf16(__uint_as_float(b[2 * i]) * mul, __uint_as_float(b[2 * i + 1]) * mul);
    }
  }
  release();
  if (!valid) return;
#pragma unroll
  for (int qq = 0; qq < NQ; ++qq) {
    const int q = q0 + qq;
    if (rope) {
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {  // 8 columns (one 16-byte store) of each half at a time
        const int col = q * 32 + ch * 8;
        const uint4 c_lo = __ldg(reinterpret_cast<const uint4*>(cos_row + col));
        const uint4 s_hi = __ldg(reinterpret_cast<const uint4*>(sin_row + 64 + col));
        const uint4 c_hi = __ldg(reinterpret_cast<const uint4*>(cos_row + 64 + col));
        const uint4 s_lo = __ldg(reinterpret_cast<const uint4*>(sin_row + col));
        const uint32_t cl[4] = {c_lo.x, c_lo.y, c_lo.z, c_lo.w}, sh[4] = {s_hi.x, s_hi.y, s_hi.z, s_hi.w};
        const uint32_t chh[4] = {c_hi.x, c_hi.y, c_hi.z, c_hi.w}, sl[4] = {s_lo.x, s_lo.y, s_lo.z, s_lo.w};
        uint32_t o_lo[4], o_hi[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t x = lo[qq][ch * 4 + i], y = hi[qq][ch * 4 + i];
          // dx[j] = dy[j] c[j] + dy[j+64] s[j+64];   dx[j+64] = dy[j+64] c[j+64] - dy[j] s[j]
          o_lo[i] = pack_bf16(bf16_lo(x) * bf16_lo(cl[i]) + bf16_lo(y) * bf16_lo(sh[i]),
                              bf16_hi(x) * bf16_hi(cl[i]) + bf16_hi(y) * bf16_hi(sh[i]));
          o_hi[i] = pack_bf16(bf16_lo(y) * bf16_lo(chh[i]) - bf16_lo(x) * bf16_lo(sl[i]),
                              bf16_hi(y) * bf16_hi(chh[i]) - bf16_hi(x) * bf16_hi(sl[i]));
        }
        *reinterpret_cast<uint4*>(dst + col) = make_uint4(o_lo[0], o_lo[1], o_lo[2], o_lo[3]);
        *reinterpret_cast<uint4*>(dst + 64 + col) = make_uint4(o_hi[0], o_hi[1], o_hi[2], o_hi[3]);
      }
    } else {
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        const int col = q * 32 + ch * 8;
        *reinterpret_cast<uint4*>(dst + col) =
            make_uint4(lo[qq][ch * 4], lo[qq][ch * 4 + 1], lo[qq][ch * 4 + 2], lo[qq][ch * 4 + 3]);
        *reinterpret_cast<uint4*>(dst + 64 + col) =
            make_uint4(hi[qq][ch * 4], hi[qq][ch * 4 + 1], hi[qq][ch * 4 + 2], hi[qq][ch * 4 + 3]);
      }
    }
  }
}

// One 8-column slice (and its + 64 partner slice) of a drained gradient row: rotary transpose with the row's table
// entries and two 16-byte stores.  `lo` / `hi` hold the row's packed bf16 pairs of columns [32 q, 32 q + 32) and
// [64 + 32 q, ...); CH selects the slice statically (no dynamic register indexing).
template <int CH>
__device__ __forceinline__ void flush_rope_chunk(const uint32_t (&lo)[16], const uint32_t (&hi)[16],
                                                 const __nv_bfloat16* cos_row, const __nv_bfloat16* sin_row,
                                                 __nv_bfloat16* dst, int q) {
  const int col = q * 32 + CH * 8;
  const uint4 c_lo = __ldg(reinterpret_cast<const uint4*>(cos_row + col));
  const uint4 s_hi = __ldg(reinterpret_cast<const uint4*>(sin_row + 64 + col));
  const uint4 c_hi = __ldg(reinterpret_cast<const uint4*>(cos_row + 64 + col));
  const uint4 s_lo = __ldg(reinterpret_cast<const uint4*>(sin_row + col));
  const uint32_t cl[4] = {c_lo.x, c_lo.y, c_lo.z, c_lo.w}, sh[4] = {s_hi.x, s_hi.y, s_hi.z, s_hi.w};
  const uint32_t chh[4] = {c_hi.x, c_hi.y, c_hi.z, c_hi.w}, sl[4] = {s_lo.x, s_lo.y, s_lo.z, s_lo.w};
  uint32_t o_lo[4], o_hi[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t x = lo[CH * 4 + i], y = hi[CH * 4 + i];
    o_lo[i] = pack_bf16(bf16_lo(x) * bf16_lo(cl[i]) + bf16_lo(y) * bf16_lo(sh[i]),
                        bf16_hi(x) * bf16_hi(cl[i]) + bf16_hi(y) * bf16_hi(sh[i]));
    o_hi[i] = pack_bf16(bf16_lo(y) * bf16_lo(chh[i]) - bf16_lo(x) * bf16_lo(sl[i]),
                        bf16_hi(y) * bf16_hi(chh[i]) - bf16_hi(x) * bf16_hi(sl[i]));
  }
  *reinterpret_cast<uint4*>(dst + col) = make_uint4(o_lo[0], o_lo[1], o_lo[2], o_lo[3]);
  *reinterpret_cast<uint4*>(dst + 64 + col) = make_uint4(o_hi[0], o_hi[1], o_hi[2], o_hi[3]);
}

// =============================================================================================
// Persistent variants (round 2).  Same math, same buffers, same roles as the two kernels above; what changes is the
// life cycle: ONE CTA per SM pulls (sample, head, block) items from a per-launch atomic counter and keeps its
// barriers, TMEM allocation and pipelines alive across items, so that
//   * the TMA producer loads the next item's resident operands (K / V, resp. Q / dO) as soon as the current item's last
//     S / dP MMAs have been issued, and its first streamed tiles as ring slots free up;
//   * the MMA warp issues the next item's first S / dP while the softmax warps run the current item's epilogue;
//   * barrier init, TMEM allocation and descriptor prefetch are paid once per CTA instead of once per block.
// The cycle trace of the grid version (profiles/r1_attn_bwd_s8.md) put ~19 000 of ~44 000 cycles of an average dK/dV
// CTA into exactly these boundaries.  Items are handed out group-major (the key blocks of one (sample, head) are in
// flight together: Q / dO are shared through L2), heaviest block of a group first.
// Every mbarrier phase comes from a running counter kept identically by the roles that share the barrier:
//   g  : 64-row steps so far (q / s / p rings, 2 deep)         it : valid items so far (resident operands, accumulators)
// =============================================================================================
constexpr int K9P_SCHED = 2;
constexpr int K9P_MAXQ = 5;
// Ring depth of the streamed 64-row tiles.  A slot is only free again when the step's LAST MMAs (dV / dK resp. dQ, which
// read the tile as an MN-major B operand) have completed, so with 2 slots the load of step g + 2 could not start before
// step g was completely done and every step paid a full TMA latency (~2 000 cycles against 1 024 cycles of tensor work:
// ncu showed the tensor pipe 29 % active whatever the softmax arrangement).  dK/dV: 3 slots (shared memory is then full:
// 226 KB); dQ: 4.
// P^T / dS^T (dK/dV) and dS (dQ) live in TMEM, written by the softmax thread of the row with tcgen05.st over the first
// 32 columns of the S / dP buffer it has just read (A operand of tcgen05.mma from TMEM): the 2 x 32 KB (2 x 16 KB) of
// shared memory they used go to the ring -- 5 slots for both kernels (3 / 4 with P / dS in shared memory).
constexpr int K9P_NQ_A = 5;
constexpr int K9P_NQ_B = 5;
constexpr int ABA_SMEM_P = ABA_SMEM - 4 * AB_P + (K9P_NQ_A - 2) * 2 * AB_T64;
constexpr int ABB_SMEM_P = ABB_SMEM - 2 * AB_P + (K9P_NQ_B - 2) * 2 * AB_T64;
static_assert(ABA_SMEM_P <= 232448 && ABB_SMEM_P <= 232448, "shared memory per CTA");
constexpr int K9P_READERS = 9;  // MMA warp + 8 softmax warps (lane 0 arrives)
constexpr int K9P_NCOUNTERS = 2048;
__device__ unsigned int g_k9_counters[K9P_NCOUNTERS];

struct K9Item {  // valid: 1 = work, 0 = block past the sample's end, -1 = work exhausted
  int valid, b, h, blk, seq0, len;
};

struct BarsP {
  uint64_t res_full, res_empty;                       // resident operands (K,V / Q,dO) of the item
  uint64_t q_full[K9P_MAXQ], q_empty[K9P_MAXQ];       // streamed 64-row tiles (ring of NQ stages)
  uint64_t s_full[3], p_full[3];                      // S / dP buffers in TMEM: 2 (dK/dV) or 3 (dQ)
  uint64_t acc_full, acc_empty;
  uint64_t a_ready;                                   // dQ: the item's Q / dO rows have been copied into TMEM
  uint64_t sched_full[K9P_SCHED], sched_empty[K9P_SCHED];
  K9Item item[K9P_SCHED];
  uint32_t tmem_base;
};
static_assert(sizeof(BarsP) <= 256, "barrier block");  // the layouts reserve 256 bytes behind the last tile

// `reverse`: block index counted from the end of the group (dQ: the LAST query block is the heaviest, take it first)
__device__ __forceinline__ void k9p_decode(int it, int nblk, int heads, const int32_t* __restrict__ cu_seqlens,
                                           bool reverse, K9Item& w) {
  const int g = it / nblk;
  w.blk = it - g * nblk;
  if (reverse) w.blk = nblk - 1 - w.blk;
  w.b = g / heads;
  w.h = g - w.b * heads;
  w.seq0 = __ldg(cu_seqlens + w.b);
  w.len = __ldg(cu_seqlens + w.b + 1) - w.seq0;
  w.valid = (w.blk * 128 < w.len) ? 1 : 0;
}

__device__ __forceinline__ bool k9p_next_item(BarsP* bars, int& n_fetch, int lane, K9Item& w) {
  const int slot = n_fetch % K9P_SCHED;
  mbar_wait(&bars->sched_full[slot], (n_fetch / K9P_SCHED) & 1);
  w = bars->item[slot];
  __syncwarp();
  if (lane == 0) mbar_arrive(&bars->sched_empty[slot]);
  ++n_fetch;
  return w.valid >= 0;
}

// scheduler half of the producer thread: fetch, decode, publish.  Returns the item (valid == -1 when exhausted).
__device__ __forceinline__ K9Item k9p_publish(BarsP* bars, int& n_fetch, unsigned int* counter, int n_items, int nblk,
                                              int heads, const int32_t* __restrict__ cu_seqlens, bool reverse) {
  const int slot = n_fetch % K9P_SCHED;
  mbar_wait(&bars->sched_empty[slot], ((n_fetch / K9P_SCHED) & 1) ^ 1);
  const unsigned int fetched = atomicAdd(counter, 1u);
  K9Item w;
  w.valid = -1;
  w.b = w.h = w.blk = w.seq0 = w.len = 0;
  if (fetched < static_cast<unsigned int>(n_items))
    k9p_decode(static_cast<int>(fetched), nblk, heads, cu_seqlens, reverse, w);
  bars->item[slot] = w;
  mbar_arrive(&bars->sched_full[slot]);  // release: the item is visible to whoever observes the phase
  ++n_fetch;
  return w;
}

__device__ __forceinline__ void k9p_init(BarsP* bars, int res_empty_count) {
  mbar_init(&bars->res_full, 1);
  mbar_init(&bars->res_empty, res_empty_count);  // dK/dV: one tcgen05.commit; dQ: the 8 softmax warps (TMEM copy done)
  mbar_init(&bars->a_ready, 8);
  for (int i = 0; i < K9P_MAXQ; ++i) {
    mbar_init(&bars->q_full[i], 1);
    mbar_init(&bars->q_empty[i], 1);
  }
  for (int i = 0; i < 3; ++i) {
    mbar_init(&bars->s_full[i], 1);
    mbar_init(&bars->p_full[i], 4);   // the four warps of the softmax group that handles the step
  }
  mbar_init(&bars->acc_full, 1);
  mbar_init(&bars->acc_empty, 8);
  for (int i = 0; i < K9P_SCHED; ++i) {
    mbar_init(&bars->sched_full[i], 1);
    mbar_init(&bars->sched_empty[i], K9P_READERS);
  }
  fence_mbar_init();
}

__global__ void __launch_bounds__(AB_THREADS, 1)
    k9_attn_bwd_dkdv_p(const __grid_constant__ CUtensorMap tm_qkv128, const __grid_constant__ CUtensorMap tm_qkv64,
                       const __grid_constant__ CUtensorMap tm_do64, const BwdParams p, int B, int nkb,
                       unsigned int* __restrict__ work_counter) {
  const int H = p.heads * 128;
  const int n_items = nkb * B * p.heads;
  extern __shared__ uint8_t ab_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ab_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = smem + AB_T128;
  constexpr int NQ = K9P_NQ_A;
  uint8_t* sQ = smem + 2 * AB_T128;              // NQ stages of AB_T64
  uint8_t* sdO = sQ + NQ * AB_T64;               // NQ stages
  float* sStat = reinterpret_cast<float*>(sdO + NQ * AB_T64);  // [2][128]: lse (64) | delta (64) of a group's step
  BarsP* bars = reinterpret_cast<BarsP*>(sStat + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) k9p_init(bars, 1);
  if (warp == 2) tmem_alloc(&bars->tmem_base, 512);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_qkv128);
    tma_prefetch_desc(&tm_qkv64);
    tma_prefetch_desc(&tm_do64);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
  const uint32_t tS = tmem, tdP = tmem + 128, tdV = tmem + 256, tdK = tmem + 384;

  if (warp == 0 && lane == 0) {
    // =============================== scheduler + TMA producer ===============================
    int n_fetch = 0, it = 0, g = 0;
    for (;;) {
      const K9Item w = k9p_publish(bars, n_fetch, work_counter, n_items, nkb, p.heads, p.cu_seqlens, false);
      if (w.valid < 0) break;
      if (w.valid == 0) continue;
      const int kv0 = w.blk * 128;
      const int i0 = 2 * w.blk;
      const int n_steps = (w.len + 63) / 64 - i0;
      const int colq = w.h * 128, colk = H + w.h * 128, colv = 2 * H + w.h * 128;
      if (it > 0) mbar_wait(&bars->res_empty, (it - 1) & 1);  // every S / dP of the previous item has read K, V
      mbar_arrive_expect_tx(&bars->res_full, 2 * AB_T128);
      tma_load_2d(sK, &tm_qkv128, &bars->res_full, colk, w.seq0 + kv0);
      tma_load_2d(sK + AB_T128 / 2, &tm_qkv128, &bars->res_full, colk + 64, w.seq0 + kv0);
      tma_load_2d(sV, &tm_qkv128, &bars->res_full, colv, w.seq0 + kv0);
      tma_load_2d(sV + AB_T128 / 2, &tm_qkv128, &bars->res_full, colv + 64, w.seq0 + kv0);
      for (int s = 0; s < n_steps; ++s, ++g) {
        const int st = g % NQ;
        const int row = w.seq0 + (i0 + s) * 64;
        mbar_wait(&bars->q_empty[st], ((g / NQ) & 1) ^ 1);
        mbar_arrive_expect_tx(&bars->q_full[st], 2 * AB_T64);
        tma_load_2d(sQ + st * AB_T64, &tm_qkv64, &bars->q_full[st], colq, row);
        tma_load_2d(sQ + st * AB_T64 + AB_T64 / 2, &tm_qkv64, &bars->q_full[st], colq + 64, row);
        tma_load_2d(sdO + st * AB_T64, &tm_do64, &bars->q_full[st], colq, row);
        tma_load_2d(sdO + st * AB_T64 + AB_T64 / 2, &tm_do64, &bars->q_full[st], colq + 64, row);
      }
      ++it;
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (whole warp, elected lane issues) ===============================
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 64, 0, 0);
    constexpr uint32_t idesc_acc = umma_idesc_bf16(128, 128, 0, 1);  // B MN-major
    const uint64_t dK = umma_desc_kmajor_sw128(smem_u32(sK)), dV = umma_desc_kmajor_sw128(smem_u32(sV));
    const uint64_t dQk0 = umma_desc_kmajor_sw128(smem_u32(sQ));
    const uint64_t ddOk0 = umma_desc_kmajor_sw128(smem_u32(sdO));
    const uint64_t dQm0 = umma_desc_mnmajor_sw128(smem_u32(sQ), AB_T64 / 2, 1024);
    const uint64_t ddOm0 = umma_desc_mnmajor_sw128(smem_u32(sdO), AB_T64 / 2, 1024);
    auto issue_s = [&](int gg) {  // S^T = K Q^T and dP^T = V dO^T of global step gg into the S / dP buffer gg & 1
      const int st = gg & 1, sq = gg % NQ;
      mbar_wait(&bars->q_full[sq], (gg / NQ) & 1);
      tc_fence_after();
      if (elect_one_sync()) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_ss(tS + st * 64, dK + static_cast<uint64_t>(((kk >> 2) * (AB_T128 / 2) + (kk & 3) * 32) >> 4),
                  (dQk0 + static_cast<uint64_t>(sq * (AB_T64 >> 4))) + static_cast<uint64_t>(((kk >> 2) * (AB_T64 / 2) + (kk & 3) * 32) >> 4), idesc_s, kk > 0);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_ss(tdP + st * 64, dV + static_cast<uint64_t>(((kk >> 2) * (AB_T128 / 2) + (kk & 3) * 32) >> 4),
                  (ddOk0 + static_cast<uint64_t>(sq * (AB_T64 >> 4))) + static_cast<uint64_t>(((kk >> 2) * (AB_T64 / 2) + (kk & 3) * 32) >> 4), idesc_s, kk > 0);
        umma_commit(&bars->s_full[st]);
      }
      __syncwarp();
    };
    int n_fetch = 0, it = 0, g = 0;
    for (;;) {
      K9Item w;
      if (!k9p_next_item(bars, n_fetch, lane, w)) break;
      if (w.valid == 0) continue;
      const int n_steps = (w.len + 63) / 64 - 2 * w.blk;
      mbar_wait(&bars->res_full, it & 1);
      issue_s(g);
      if (n_steps == 1) {  // the item's last S / dP is on its way: K and V may be overwritten once it completes
        if (elect_one_sync()) umma_commit(&bars->res_empty);
        __syncwarp();
      }
      for (int s = 0; s < n_steps; ++s, ++g) {
        if (s + 1 < n_steps) {
          issue_s(g + 1);
          if (s + 2 == n_steps) {
            if (elect_one_sync()) umma_commit(&bars->res_empty);
            __syncwarp();
          }
        }
        const int st = g & 1, sq = g % NQ;
        mbar_wait(&bars->p_full[st], (g >> 1) & 1);
        if (s == 0 && it > 0) mbar_wait(&bars->acc_empty, (it - 1) & 1);  // the previous item's dV / dK have been read out
        tc_fence_after();
        if (elect_one_sync()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)  // A = P^T from TMEM: 16 queries = 8 packed columns per K step
            umma_ts(tdV, tS + st * 64 + kk * 8, (ddOm0 + static_cast<uint64_t>(sq * (AB_T64 >> 4))) + static_cast<uint64_t>((kk * 2048) >> 4),
                    idesc_acc, (s > 0) || (kk > 0));
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)  // A = dS^T from TMEM
            umma_ts(tdK, tdP + st * 64 + kk * 8, (dQm0 + static_cast<uint64_t>(sq * (AB_T64 >> 4))) + static_cast<uint64_t>((kk * 2048) >> 4),
                    idesc_acc, (s > 0) || (kk > 0));
          umma_commit(&bars->q_empty[sq]);
        }
        __syncwarp();
      }
      if (elect_one_sync()) umma_commit(&bars->acc_full);
      __syncwarp();
      ++it;
    }
  } else if (warp >= 4) {
    // =============================== softmax threads ===============================
    // Two GROUPS of four warps (one thread per KEY row each) take ALTERNATE steps: group 0 the even global steps (S / dP
    // / P / dS buffer 0), group 1 the odd ones (buffer 1).  A thread handles all 64 queries of its step in two 32-column
    // passes.  With all eight warps on the same step (round 1 and the first persistent version) the softmax stage was a
    // serial ~1 400-cycle section between two 512-cycle MMA bursts -- TMEM-load, barrier and shared-memory latencies
    // with nothing to overlap them; with two steps in flight the latencies of one group hide behind the other's work.
    const int sw = warp - 4;
    const int ew = sw & 3;             // TMEM lane quarter
    const int grp = sw >> 2;           // softmax group == buffer index == epilogue half
    const int half = grp;              // epilogue: group 0 stores dV, group 1 dK
    const int c = ew * 32 + lane;      // key row inside the block == TMEM lane
    const int sid = ew * 32 + lane;    // thread index inside the group, 0..127
    const uint32_t lane_sel = static_cast<uint32_t>(ew * 32) << 16;
    float* myStat = sStat + grp * 128;  // [64 lse | 64 delta] of the group's current step
    int n_fetch = 0, it = 0, g = 0;
    for (;;) {
      K9Item w;
      if (!k9p_next_item(bars, n_fetch, lane, w)) break;
      if (w.valid == 0) continue;
      const int kv0 = w.blk * 128, i0 = 2 * w.blk, len = w.len, seq0 = w.seq0;
      const int n_steps = (len + 63) / 64 - i0;
      const int kv_idx = kv0 + c;
      const float* lse_h = p.lse + static_cast<int64_t>(w.h) * p.rows_cap;
      const float* delta_h = p.delta + static_cast<int64_t>(w.h) * p.rows_cap;
      auto load_stat = [&](int s) {
        float v = 0.f;
        if (s < n_steps) {
          const int qi = (i0 + s) * 64 + (sid & 63);
          if (qi < len) v = (sid < 64 ? lse_h : delta_h)[seq0 + qi];
        }
        return v;
      };
      const int s_first = ((g & 1) == grp) ? 0 : 1;  // this group's first step of the item
      float stat_next = load_stat(s_first);
      // the epilogue's row constants, loaded up front so that the epilogue does not start with a global round trip
      const bool valid = kv_idx < len;
      const int tok = seq0 + kv_idx;
      int dst = 0, pos = 0;
      if (valid) {
        dst = p.token_to_sorted ? __ldg(p.token_to_sorted + tok) : tok;
        const int64_t pz = __ldg(p.position_ids + (p.token_to_flat ? __ldg(p.token_to_flat + tok) : tok));
        pos = static_cast<int>(min(max(pz, int64_t(0)), int64_t(p.rope_len - 1)));
      }
      for (int s = s_first; s < n_steps; s += 2) {
        const int gs = g + s;            // global step: gs & 1 == grp
        const int st = grp;
        const uint32_t ph = (gs >> 1) & 1;
        const int q_base = (i0 + s) * 64;
        named_bar_sync(1 + grp, 128);    // every thread of the group has read the previous step's statistics
        myStat[sid] = stat_next;
        stat_next = load_stat(s + 2);
        named_bar_sync(1 + grp, 128);
        mbar_wait(&bars->s_full[st], ph);
        tc_fence_after();
        const bool interior = (q_base >= kv0 + 127) && (q_base + 64 <= len);
        uint32_t pk[2][16], dk[2][16];
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t sraw[32], draw[32];
          tmem_ld_32x32b_x32(tS + lane_sel + st * 64 + hf * 32, sraw);
          tmem_ld_32x32b_x32(tdP + lane_sel + st * 64 + hf * 32, draw);
          tmem_ld_wait();
          const float4* l4 = reinterpret_cast<const float4*>(myStat + hf * 32);
          const float4* d4 = reinterpret_cast<const float4*>(myStat + 64 + hf * 32);
          if (interior) {
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 l = l4[j4], d = d4[j4];
              const float lv[4] = {l.x, l.y, l.z, l.w}, dv[4] = {d.x, d.y, d.z, d.w};
              float pr[4], ds[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                pr[k] = ex2_approx(fmaf(__uint_as_float(sraw[4 * j4 + k]), p.scale_log2, -lv[k]));
                ds[k] = pr[k] * (__uint_as_float(draw[4 * j4 + k]) - dv[k]);
              }
              pk[hf][2 * j4] = pack_bf16(pr[0], pr[1]);
              pk[hf][2 * j4 + 1] = pack_bf16(pr[2], pr[3]);
              dk[hf][2 * j4] = pack_bf16(ds[0], ds[1]);
              dk[hf][2 * j4 + 1] = pack_bf16(ds[2], ds[3]);
            }
          } else {
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 l = l4[j4], d = d4[j4];
              const float lv[4] = {l.x, l.y, l.z, l.w}, dv[4] = {d.x, d.y, d.z, d.w};
              float pr[4], ds[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const int q_idx = q_base + hf * 32 + 4 * j4 + k;
                const bool keep = (q_idx >= kv_idx) && (q_idx < len);
                pr[k] = keep ? ex2_approx(fmaf(__uint_as_float(sraw[4 * j4 + k]), p.scale_log2, -lv[k])) : 0.f;
                ds[k] = keep ? pr[k] * (__uint_as_float(draw[4 * j4 + k]) - dv[k]) : 0.f;
              }
              pk[hf][2 * j4] = pack_bf16(pr[0], pr[1]);
              pk[hf][2 * j4 + 1] = pack_bf16(pr[2], pr[3]);
              dk[hf][2 * j4] = pack_bf16(ds[0], ds[1]);
              dk[hf][2 * j4 + 1] = pack_bf16(ds[2], ds[3]);
            }
          }
        }
        // P^T / dS^T of this key row -> the first 32 columns of the S / dP buffer just read (lane-private rows: no other
        // thread touches them; the next S / dP into this buffer is issued behind the MMAs that read P / dS, in pipe order)
        tmem_st_32x32b_x32(tS + lane_sel + st * 64, *reinterpret_cast<uint32_t(*)[32]>(&pk[0][0]));
        tmem_st_32x32b_x32(tdP + lane_sel + st * 64, *reinterpret_cast<uint32_t(*)[32]>(&dk[0][0]));
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->p_full[st]);
      }
      g += n_steps;
      // ---- epilogue: dV (plain), dK (scale, rotary transpose) -> dqkv[token_to_sorted[tok]] ----
      mbar_wait(&bars->acc_full, it & 1);
      tc_fence_after();
      __nv_bfloat16* row = p.dqkv + static_cast<int64_t>(dst) * (3 * H) + w.h * 128;
      auto release = [&]() {  // the accumulators are in registers: the next item's first dV / dK MMAs may overwrite them
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->acc_empty);
      };
      if (half == 0)
        store_grad_row_early<2>(tdV + lane_sel, false, nullptr, nullptr, 1.0f, row + 2 * H, valid, 0, release);
      else
        store_grad_row_early<2>(tdK + lane_sel, true, p.rope_cos + static_cast<int64_t>(pos) * 128,
                                p.rope_sin + static_cast<int64_t>(pos) * 128, p.scale, row + H, valid, 0, release);
      ++it;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

__global__ void __launch_bounds__(AB_THREADS, 1)
    k9_attn_bwd_dq_p(const __grid_constant__ CUtensorMap tm_qkv128, const __grid_constant__ CUtensorMap tm_qkv64,
                     const __grid_constant__ CUtensorMap tm_do128, const BwdParams p, int B, int nqb,
                     unsigned int* __restrict__ work_counter) {
  const int H = p.heads * 128;
  const int n_items = nqb * B * p.heads;
  extern __shared__ uint8_t ab_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ab_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sdO = smem + AB_T128;
  constexpr int NQ = K9P_NQ_B;
  uint8_t* sK = smem + 2 * AB_T128;  // NQ stages of AB_T64
  uint8_t* sV = sK + NQ * AB_T64;    // NQ stages
  BarsP* bars = reinterpret_cast<BarsP*>(sV + NQ * AB_T64);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef VEX_ATTN_TRACE
  // timing experiment build (tools/k9_trace.py): stamps of CTA 0 -- row = global step (< 256); columns 0..5 softmax warp 4
  // (group 0: even steps), 8..11 MMA warp, 12 producer
  long long* trc = (blockIdx.x == 0 && (threadIdx.x == 128 || threadIdx.x == 32 || threadIdx.x == 0)) ? g_k9_trace : nullptr;
#define K9P_TRACE(row, col)                                                 \
  do {                                                                      \
    if (trc && (row) < 256) trc[(row) * 16 + (col)] = clock64();           \
  } while (0)
#else
#define K9P_TRACE(row, col) \
  do {                      \
  } while (0)
#endif
  if (threadIdx.x == 0) k9p_init(bars, 8);
  if (warp == 2) tmem_alloc(&bars->tmem_base, 512);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_qkv128);
    tma_prefetch_desc(&tm_qkv64);
    tma_prefetch_desc(&tm_do128);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
  // TMEM: S / dP double-buffered (2 x 64 + 2 x 64), dQ accumulator (128), and the item's Q and dO rows as TMEM-RESIDENT A
  // OPERANDS (64 + 64 packed columns).  With A from shared memory an N = 64 MMA reads 6 KB per instruction and issues at
  // ~68 cycles instead of 32 (cycle trace, profiles/r2_k9_attention_bwd.md: the steady state was operand-bandwidth
  // bound); with Q / dO in TMEM it reads the 2 KB K / V slice only.  (A third S / dP buffer in these columns was
  // measured first: no change.)
  constexpr int NS = 2;
  const uint32_t tS = tmem, tdP = tmem + 128, tdQ = tmem + 256, tQ = tmem + 384, tdO = tmem + 448;

  if (warp == 0 && lane == 0) {
    // =============================== scheduler + TMA producer ===============================
    int n_fetch = 0, it = 0, g = 0;
    for (;;) {
      const K9Item w = k9p_publish(bars, n_fetch, work_counter, n_items, nqb, p.heads, p.cu_seqlens, true);
      if (w.valid < 0) break;
      if (w.valid == 0) continue;
      const int q0 = w.blk * 128;
      const int n_steps = min((w.len + 63) / 64, 2 * w.blk + 2);
      const int colq = w.h * 128, colk = H + w.h * 128, colv = 2 * H + w.h * 128;
      if (it > 0) mbar_wait(&bars->res_empty, (it - 1) & 1);  // every S / dP of the previous item has read Q, dO
      mbar_arrive_expect_tx(&bars->res_full, 2 * AB_T128);
      tma_load_2d(sQ, &tm_qkv128, &bars->res_full, colq, w.seq0 + q0);
      tma_load_2d(sQ + AB_T128 / 2, &tm_qkv128, &bars->res_full, colq + 64, w.seq0 + q0);
      tma_load_2d(sdO, &tm_do128, &bars->res_full, colq, w.seq0 + q0);
      tma_load_2d(sdO + AB_T128 / 2, &tm_do128, &bars->res_full, colq + 64, w.seq0 + q0);
      for (int s = 0; s < n_steps; ++s, ++g) {
        const int st = g % NQ;
        const int row = w.seq0 + s * 64;
        mbar_wait(&bars->q_empty[st], ((g / NQ) & 1) ^ 1);
        K9P_TRACE(g, 12);
        mbar_arrive_expect_tx(&bars->q_full[st], 2 * AB_T64);
        tma_load_2d(sK + st * AB_T64, &tm_qkv64, &bars->q_full[st], colk, row);
        tma_load_2d(sK + st * AB_T64 + AB_T64 / 2, &tm_qkv64, &bars->q_full[st], colk + 64, row);
        tma_load_2d(sV + st * AB_T64, &tm_qkv64, &bars->q_full[st], colv, row);
        tma_load_2d(sV + st * AB_T64 + AB_T64 / 2, &tm_qkv64, &bars->q_full[st], colv + 64, row);
      }
      ++it;
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 64, 0, 0);
    constexpr uint32_t idesc_acc = umma_idesc_bf16(128, 128, 0, 1);
    const uint64_t dKk0 = umma_desc_kmajor_sw128(smem_u32(sK));
    const uint64_t dVk0 = umma_desc_kmajor_sw128(smem_u32(sV));
    const uint64_t dKm0 = umma_desc_mnmajor_sw128(smem_u32(sK), AB_T64 / 2, 1024);
    auto issue_s = [&](int gg) {
      const int st = gg % NS, sq = gg % NQ;
      mbar_wait(&bars->q_full[sq], (gg / NQ) & 1);
      tc_fence_after();
      if (elect_one_sync()) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_ts(tS + st * 64, tQ + kk * 8,   // A = Q rows from TMEM: 16 d = 8 packed columns per K step
                  (dKk0 + static_cast<uint64_t>(sq * (AB_T64 >> 4))) + static_cast<uint64_t>(((kk >> 2) * (AB_T64 / 2) + (kk & 3) * 32) >> 4), idesc_s, kk > 0);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_ts(tdP + st * 64, tdO + kk * 8,
                  (dVk0 + static_cast<uint64_t>(sq * (AB_T64 >> 4))) + static_cast<uint64_t>(((kk >> 2) * (AB_T64 / 2) + (kk & 3) * 32) >> 4), idesc_s, kk > 0);
        umma_commit(&bars->s_full[st]);
      }
      __syncwarp();
    };
    int n_fetch = 0, it = 0, g = 0;
    for (;;) {
      K9Item w;
      if (!k9p_next_item(bars, n_fetch, lane, w)) break;
      if (w.valid == 0) continue;
      const int n_steps = min((w.len + 63) / 64, 2 * w.blk + 2);
      mbar_wait(&bars->a_ready, it & 1);  // Q / dO of this item are in TMEM
      tc_fence_after();
      const int g_end = g + n_steps;
      int issued = g;  // next step whose S / dP has not been issued yet; runs up to NS - 1 steps ahead of the dQ MMAs
      for (int s = 0; s < n_steps; ++s, ++g) {
        K9P_TRACE(g, 8);
        while (issued < g_end && issued < g + NS) {
          issue_s(issued);
          ++issued;
        }
        const int st = g % NS, sq = g % NQ;
        K9P_TRACE(g, 9);
        mbar_wait(&bars->p_full[st], (g / NS) & 1);
        if (s == 0 && it > 0) mbar_wait(&bars->acc_empty, (it - 1) & 1);
        K9P_TRACE(g, 10);
        tc_fence_after();
        if (elect_one_sync()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)  // A = dS from TMEM: 16 keys = 8 packed columns per K step
            umma_ts(tdQ, tdP + st * 64 + kk * 8, (dKm0 + static_cast<uint64_t>(sq * (AB_T64 >> 4))) + static_cast<uint64_t>((kk * 2048) >> 4),
                    idesc_acc, (s > 0) || (kk > 0));
          umma_commit(&bars->q_empty[sq]);
        }
        __syncwarp();
        K9P_TRACE(g, 11);
      }
      if (elect_one_sync()) umma_commit(&bars->acc_full);
      __syncwarp();
      ++it;
    }
  } else if (warp >= 4) {
    // =============================== softmax threads ===============================
    // two groups of four warps on alternate steps, one thread per QUERY row handling the step's 64 keys in two
    // 32-column passes (see k9_attn_bwd_dkdv_p)
    const int sw = warp - 4;
    const int ew = sw & 3;
    const int grp = sw >> 2;
    const int half = grp;              // epilogue: which 32 + 32 columns of dQ this thread stores
    const int c = ew * 32 + lane;
    const uint32_t lane_sel = static_cast<uint32_t>(ew * 32) << 16;
    // DEFERRED EPILOGUE: the previous item's dQ row sits in packed registers (drained from TMEM right after its last
    // step, accumulator released at once); its rotary transpose and stores are done one 8-column slice at a time in
    // front of the s_full waits of the NEXT item's steps, where these warps idle anyway (cycle trace: ~1 400 cycles per
    // step) -- the ~7 000-cycle store phase no longer delays the next item's first dS.
    uint32_t pend_lo[16], pend_hi[16];
    const __nv_bfloat16* pend_cos = nullptr;
    const __nv_bfloat16* pend_sin = nullptr;
    __nv_bfloat16* pend_dst = nullptr;
    int pend_ch = 4;  // next slice to flush; 4 = nothing pending
    auto flush_one = [&]() {
      if (pend_ch == 0) flush_rope_chunk<0>(pend_lo, pend_hi, pend_cos, pend_sin, pend_dst, half);
      else if (pend_ch == 1) flush_rope_chunk<1>(pend_lo, pend_hi, pend_cos, pend_sin, pend_dst, half);
      else if (pend_ch == 2) flush_rope_chunk<2>(pend_lo, pend_hi, pend_cos, pend_sin, pend_dst, half);
      else if (pend_ch == 3) flush_rope_chunk<3>(pend_lo, pend_hi, pend_cos, pend_sin, pend_dst, half);
      if (pend_ch < 4) ++pend_ch;
    };
    int n_fetch = 0, it = 0, g = 0;
    for (;;) {
      K9Item w;
      if (!k9p_next_item(bars, n_fetch, lane, w)) break;
      if (w.valid == 0) continue;
      const int q0 = w.blk * 128, len = w.len, seq0 = w.seq0;
      const int n_steps = min((len + 63) / 64, 2 * w.blk + 2);
      const int q_idx = q0 + c;
      const bool valid = q_idx < len;
      const int tok = seq0 + q_idx;
      float l2 = 0.f, dl = 0.f;
      int dst = 0, pos = 0;
      if (valid) {
        l2 = p.lse[static_cast<int64_t>(w.h) * p.rows_cap + tok];
        dl = p.delta[static_cast<int64_t>(w.h) * p.rows_cap + tok];
        dst = p.token_to_sorted ? __ldg(p.token_to_sorted + tok) : tok;
        const int64_t pz = __ldg(p.position_ids + (p.token_to_flat ? __ldg(p.token_to_flat + tok) : tok));
        pos = static_cast<int>(min(max(pz, int64_t(0)), int64_t(p.rope_len - 1)));
      }
      {
        // this thread's row of Q (group 0) / dO (group 1): shared memory (TMA's SWIZZLE_128B K-major layout) -> TMEM.
        // All MMAs of the previous item are complete (its epilogue waited for acc_full), so tQ / tdO are free.
        mbar_wait(&bars->res_full, it & 1);
        const uint32_t src = smem_u32(grp == 0 ? sQ : sdO) + c * 128;
        uint32_t rowv[64];
#pragma unroll
        for (int a2 = 0; a2 < 2; ++a2)
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) {
            const uint4 v4 = ld_shared_v4_k9(src + a2 * (AB_T128 / 2) + ((ch ^ (c & 7)) << 4));
            rowv[a2 * 32 + ch * 4 + 0] = v4.x;
            rowv[a2 * 32 + ch * 4 + 1] = v4.y;
            rowv[a2 * 32 + ch * 4 + 2] = v4.z;
            rowv[a2 * 32 + ch * 4 + 3] = v4.w;
          }
        const uint32_t ta = (grp == 0 ? tQ : tdO) + lane_sel;
        tmem_st_32x32b_x32(ta, *reinterpret_cast<uint32_t(*)[32]>(&rowv[0]));
        tmem_st_32x32b_x32(ta + 32, *reinterpret_cast<uint32_t(*)[32]>(&rowv[32]));
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&bars->a_ready);    // the MMA warp may issue the item's S / dP
          mbar_arrive(&bars->res_empty);  // ... and the producer may load the next item's Q / dO
        }
      }
      for (int s = ((g & 1) == grp) ? 0 : 1; s < n_steps; s += 2) {
        const int gs = g + s;
        const int st = gs % NS;
        const uint32_t ph = (gs / NS) & 1;
        K9P_TRACE(gs, 0);
        flush_one();  // one slice of the previous item's deferred stores (no warp-level sync inside: rows may diverge)
        mbar_wait(&bars->s_full[st], ph);
        tc_fence_after();
        K9P_TRACE(gs, 1);
        const bool interior = (s * 64 + 63 <= q0) && (q0 + 128 <= len);
        uint32_t dk[2][16];
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t sraw[32], draw[32];
          tmem_ld_32x32b_x32(tS + lane_sel + st * 64 + hf * 32, sraw);
          tmem_ld_32x32b_x32(tdP + lane_sel + st * 64 + hf * 32, draw);
          tmem_ld_wait();
          if (interior) {
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const float p0 = ex2_approx(fmaf(__uint_as_float(sraw[j]), p.scale_log2, -l2));
              const float p1 = ex2_approx(fmaf(__uint_as_float(sraw[j + 1]), p.scale_log2, -l2));
              dk[hf][j >> 1] = pack_bf16(p0 * (__uint_as_float(draw[j]) - dl), p1 * (__uint_as_float(draw[j + 1]) - dl));
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const int kv_idx = s * 64 + hf * 32 + j;
              const bool k0 = valid && (kv_idx <= q_idx), k1 = valid && (kv_idx + 1 <= q_idx);
              const float p0 = k0 ? ex2_approx(fmaf(__uint_as_float(sraw[j]), p.scale_log2, -l2)) : 0.f;
              const float p1 = k1 ? ex2_approx(fmaf(__uint_as_float(sraw[j + 1]), p.scale_log2, -l2)) : 0.f;
              dk[hf][j >> 1] = pack_bf16(k0 ? p0 * (__uint_as_float(draw[j]) - dl) : 0.f,
                                         k1 ? p1 * (__uint_as_float(draw[j + 1]) - dl) : 0.f);
            }
          }
          K9P_TRACE(gs, 2 + hf);
        }
        tmem_st_32x32b_x32(tdP + lane_sel + st * 64, *reinterpret_cast<uint32_t(*)[32]>(&dk[0][0]));  // dS over dP
        tmem_st_wait();
        K9P_TRACE(gs, 4);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->p_full[st]);
        K9P_TRACE(gs, 5);
      }
      K9P_TRACE(g + n_steps - 1, 6);   // item end (before the epilogue)
      g += n_steps;
      // ---- epilogue, first half: what is left of the previous item's stores, then drain this item's dQ row ----
      while (pend_ch < 4) flush_one();
      mbar_wait(&bars->acc_full, it & 1);
      tc_fence_after();
      {
        uint32_t a[32], b2[32];
        tmem_ld_32x32b_x32(tdQ + lane_sel + half * 32, a);
        tmem_ld_32x32b_x32(tdQ + lane_sel + 64 + half * 32, b2);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          pend_lo[i] = pack_bf16(__uint_as_float(a[2 * i]) * p.scale, __uint_as_float(a[2 * i + 1]) * p.scale);
          pend_hi[i] = pack_bf16(__uint_as_float(b2[2 * i]) * p.scale, __uint_as_float(b2[2 * i + 1]) * p.scale);
        }
      }
      tc_fence_before();   // the accumulator is in registers: the next item's first dQ MMAs may overwrite it
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->acc_empty);
      pend_cos = p.rope_cos + static_cast<int64_t>(pos) * 128;
      pend_sin = p.rope_sin + static_cast<int64_t>(pos) * 128;
      pend_dst = p.dqkv + static_cast<int64_t>(dst) * (3 * H) + w.h * 128;
      pend_ch = valid ? 0 : 4;   // NOTE: per-thread (rows past the sample's end store nothing)
      K9P_TRACE(g - 1, 7);             // epilogue (drain) done
      ++it;
    }
    while (pend_ch < 4) flush_one();   // the last item's stores
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// zeroes the tail rows of dO (see k9_zero_tail_rows) and the two work counters of this launch
__global__ void k9p_prepare(__nv_bfloat16* buf, const int32_t* __restrict__ cu_seqlens, int B, int rows_cap, int row_elems,
                            unsigned int* counters) {
  if (blockIdx.x == 0 && threadIdx.x < 2) counters[threadIdx.x] = 0u;
  const int T = cu_seqlens[B];
  const int n_rows = min(rows_cap - T, 128);
  const int64_t n_vec = static_cast<int64_t>(max(n_rows, 0)) * (row_elems / 8);
  uint4* p = reinterpret_cast<uint4*>(buf + static_cast<int64_t>(T) * row_elems);
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n_vec;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    p[i] = make_uint4(0, 0, 0, 0);
}

}  // namespace vex

extern "C" int vex_attention_backward(const void* qkv, const void* out_sorted, void* d_out_tok, const float* lse,
                                      float* delta_ws, const int32_t* cu_seqlens, const int32_t* token_to_sorted,
                                      const int32_t* token_to_flat, const int64_t* position_ids, const void* rope_cos,
                                      const void* rope_sin, int rope_len, int B, int max_len_cap, int heads, void* dqkv,
                                      float scale, vexStream stream) {
  using namespace vex;
  if (!qkv || !out_sorted || !d_out_tok || !lse || !delta_ws || !cu_seqlens || !position_ids || !rope_cos ||
      !rope_sin || !dqkv || B <= 0 || max_len_cap <= 0 || heads <= 0 || rope_len <= 0)
    return VEX_E_INVALID;
  if (B > 65535 || heads > 65535) return VEX_E_UNSUPPORTED;
  const int64_t rows_cap64 = static_cast<int64_t>(B) * max_len_cap;
  if (rows_cap64 > 0x7fffffff) return VEX_E_UNSUPPORTED;
  const int rows_cap = static_cast<int>(rows_cap64);
  const uint64_t H = static_cast<uint64_t>(heads) * 128;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  static bool configured = false;
  if (!configured) {
    VEX_CUDA_TRY(cudaFuncSetAttribute(k9_attn_bwd_dkdv, cudaFuncAttributeMaxDynamicSharedMemorySize, ABA_SMEM));
    VEX_CUDA_TRY(cudaFuncSetAttribute(k9_attn_bwd_dq, cudaFuncAttributeMaxDynamicSharedMemorySize, ABB_SMEM));
    configured = true;
  }
  CUtensorMap tm_qkv128, tm_qkv64, tm_do128, tm_do64;
  int rc;
  if ((rc = make_tmap_2d(&tm_qkv128, qkv, rows_cap, 3 * H, 3 * H, 128)) != VEX_OK) return rc;
  if ((rc = make_tmap_2d(&tm_qkv64, qkv, rows_cap, 3 * H, 3 * H, 64)) != VEX_OK) return rc;
  if ((rc = make_tmap_2d(&tm_do128, d_out_tok, rows_cap, H, H, 128)) != VEX_OK) return rc;
  if ((rc = make_tmap_2d(&tm_do64, d_out_tok, rows_cap, H, H, 64)) != VEX_OK) return rc;
  // VEX_K9_IMPL=grid selects the round-1 kernels (one CTA per block, A/B switch); default: the persistent kernels
  const char* impl_env = std::getenv("VEX_K9_IMPL");
  const bool persistent = !(impl_env && std::strcmp(impl_env, "grid") == 0);
  unsigned int* counters = nullptr;
  int sm_count = 0;
  if (persistent) {
    constexpr int kMaxDev = 64;
    static int sm_counts[kMaxDev] = {};
    static unsigned int* counter_base[kMaxDev] = {};
    static std::atomic<unsigned int> next_counter{0};
    int dev = 0;
    VEX_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDev) return VEX_E_UNSUPPORTED;
    if (sm_counts[dev] == 0) {
      int n = 0;
      VEX_CUDA_TRY(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
      VEX_CUDA_TRY(cudaFuncSetAttribute(k9_attn_bwd_dkdv_p, cudaFuncAttributeMaxDynamicSharedMemorySize, ABA_SMEM_P));
      VEX_CUDA_TRY(cudaFuncSetAttribute(k9_attn_bwd_dq_p, cudaFuncAttributeMaxDynamicSharedMemorySize, ABB_SMEM_P));
      VEX_CUDA_TRY(cudaGetSymbolAddress(reinterpret_cast<void**>(&counter_base[dev]), g_k9_counters));
      sm_counts[dev] = n;
    }
    sm_count = sm_counts[dev];
    // two counters per launch (dK/dV, dQ), round-robin: launches in flight at the same time own different pairs
    counters = counter_base[dev] + 2 * (next_counter.fetch_add(1) % (K9P_NCOUNTERS / 2));
    k9p_prepare<<<32, 256, 0, s>>>(static_cast<__nv_bfloat16*>(d_out_tok), cu_seqlens, B, rows_cap, static_cast<int>(H),
                                   counters);
  } else {
    // rows [T, T + 128) of dO can be touched by the last tile: keep them finite (0 * NaN = NaN in the MMAs)
    k9_zero_tail_rows<<<32, 256, 0, s>>>(static_cast<__nv_bfloat16*>(d_out_tok), cu_seqlens, B, rows_cap,
                                         static_cast<int>(H));
  }
  VEX_LAUNCH_CHECK();
  k9_attn_delta<<<std::min(ceil_div(rows_cap, 8), 148 * 8), 256, 0, s>>>(
      static_cast<const __nv_bfloat16*>(d_out_tok), static_cast<const __nv_bfloat16*>(out_sorted), token_to_sorted,
      cu_seqlens, B, delta_ws, heads, rows_cap);
  VEX_LAUNCH_CHECK();
  BwdParams p;
  p.cu_seqlens = cu_seqlens;
  p.lse = lse;
  p.delta = delta_ws;
  p.token_to_sorted = token_to_sorted;
  p.token_to_flat = token_to_flat;
  p.position_ids = position_ids;
  p.rope_cos = static_cast<const __nv_bfloat16*>(rope_cos);
  p.rope_sin = static_cast<const __nv_bfloat16*>(rope_sin);
  p.dqkv = static_cast<__nv_bfloat16*>(dqkv);
  p.heads = heads;
  p.rows_cap = rows_cap;
  p.rope_len = rope_len;
  p.scale = scale;
  p.scale_log2 = scale * 1.4426950408889634f;
  const int nblk = ceil_div(max_len_cap, 128);
  if (persistent) {
    const int64_t n_items = static_cast<int64_t>(nblk) * heads * B;
    if (n_items > 0x7fffffff - 65536) return VEX_E_UNSUPPORTED;
    const int grid_p = static_cast<int>(n_items < sm_count ? n_items : sm_count);
    k9_attn_bwd_dkdv_p<<<grid_p, AB_THREADS, ABA_SMEM_P, s>>>(tm_qkv128, tm_qkv64, tm_do64, p, B, nblk, counters);
    VEX_LAUNCH_CHECK();
    k9_attn_bwd_dq_p<<<grid_p, AB_THREADS, ABB_SMEM_P, s>>>(tm_qkv128, tm_qkv64, tm_do128, p, B, nblk, counters + 1);
    VEX_LAUNCH_CHECK();
    return VEX_OK;
  }
  dim3 grid(nblk, heads, B);
  k9_attn_bwd_dkdv<<<grid, AB_THREADS, ABA_SMEM, s>>>(tm_qkv128, tm_qkv64, tm_do64, p);
  VEX_LAUNCH_CHECK();
  k9_attn_bwd_dq<<<grid, AB_THREADS, ABB_SMEM, s>>>(tm_qkv128, tm_qkv64, tm_do128, p);
  VEX_LAUNCH_CHECK();
  return VEX_OK;
}
