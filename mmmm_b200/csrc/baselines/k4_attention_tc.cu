// K4 (fast path) -- causal block-diagonal varlen flash attention on tcgen05 / TMEM for sm_100a.
//
// Restates the prefill branch of attention_fn (modeling_cogvlm.py:106-128; xformers
// memory_efficient_attention + BlockDiagonalCausalMask): per sample, token i attends to tokens j <= i in
// token-rank order, scale d^-0.5, fp32 online softmax, P rounded to bf16 before P.V.
//
// One CTA = 128 query rows of one (sample, head); key/value blocks of 128; 256 threads, warp-specialised:
//   warp 0 / lane 0 : TMA producer -- Q once, then K_j and V_j (two 64-column SWIZZLE_128B boxes each) through
//                     2-deep rings, straight out of the token-order QKV buffer [rows_cap, 3*heads*128].
//   warp 1 / lane 0 : MMA issuer.  S_j = Q.K_j^T (8 x tcgen05.mma M128 N128 K16, both operands K-major) into
//                     one of two TMEM S buffers; O (+)= P_j.V_j with P from shared memory (K-major) and V as an
//                     MN-major B operand (V is [key][d] in memory, d contiguous).  S_{j+1} is issued before
//                     waiting for P_j so the tensor core works while the softmax warps run.
//   warp 2          : TMEM allocate / free (S0, S1, O = 3 x 128 columns).
//   warps 4..7      : softmax, one query row per thread (TMEM lane == row, so row max / row sum need no
//                     shuffles): tcgen05.ld S -> mask (diagonal block only) -> lazy-rescaled online softmax
//                     (O is only rescaled in TMEM when the running max grows by more than 2^8) -> P (bf16) into
//                     shared memory in the UMMA K-major 128B-swizzle layout -> signal the MMA warp.
//                     Epilogue: O / l -> bf16 -> per-warp staging -> coalesced rows scattered via out_row_map.
#include <cuda.h>

#include <cstring>

#include "../common.cuh"

namespace vex {

int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);

constexpr int TC_BQ = 128, TC_BK = 128, TC_D = 128;
constexpr int TC_THREADS = 384;  // TMA, MMA, TMEM-alloc, idle + 8 softmax warps
constexpr int TC_TILE = 128 * 128 * 2;    // 32 KB: one 128 x 128 bf16 tile = two 64-column atoms of 16 KB
constexpr int TC_ATOM = 128 * 64 * 2;     // 16 KB
constexpr int TC_SMEM = 6 * TC_TILE + 256 + 2048 + 1024;  // Q, K x2, V x2, P + barriers + alignment slack
constexpr float TC_RESCALE_THRESHOLD = 8.0f;        // log2 units

struct AttnBars {
  uint64_t q_full, k_full[2], k_empty[2], v_full[2], v_empty[2], s_full[2], p_full, pv_done;
  uint32_t tmem_base;
  float xchg[2 * 2 * 128];  // row-max / row-sum exchange between the two threads of a query row
};

__device__ __forceinline__ void named_bar_sync_256() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__global__ void zero_tail_rows(__nv_bfloat16* buf, const int32_t* __restrict__ cu_seqlens, int B, int rows_cap,
                               int row_elems) {
  // rows [T, T + 128) can be touched by the last key block of the last sample; keep them finite (0 * NaN = NaN)
  const int T = cu_seqlens[B];
  const int n_rows = min(rows_cap - T, TC_BK);
  const int64_t n_vec = static_cast<int64_t>(max(n_rows, 0)) * (row_elems / 8);
  uint4* p = reinterpret_cast<uint4*>(buf + static_cast<int64_t>(T) * row_elems);
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n_vec;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    p[i] = make_uint4(0, 0, 0, 0);
}

__global__ void __launch_bounds__(TC_THREADS, 1)
    k4_attention_tc(const __grid_constant__ CUtensorMap tm_qkv, const int32_t* __restrict__ cu_seqlens, int heads,
                    const int32_t* __restrict__ out_row_map, __nv_bfloat16* __restrict__ out, float scale_log2,
                    float* __restrict__ lse, int rows_cap, int causal) {
  const int b = blockIdx.z, h = blockIdx.y;
  const int seq0 = cu_seqlens[b], len = cu_seqlens[b + 1] - seq0;
  const int qb = gridDim.x - 1 - blockIdx.x;  // heaviest query blocks first
  const int q0 = qb * TC_BQ;
  if (q0 >= len) return;
  // causal: key blocks up to the diagonal; non-causal (vision encoder, visual.py:96-98): every key block of the sample
  const int n_kv = causal ? qb + 1 : (len + TC_BK - 1) / TC_BK;
  const int H = heads * TC_D;

  extern __shared__ uint8_t tc_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + TC_TILE;      // 2 stages
  uint8_t* sV = smem + 3 * TC_TILE;  // 2 stages
  uint8_t* sP = smem + 5 * TC_TILE;
  AttnBars* bars = reinterpret_cast<AttnBars*>(smem + 6 * TC_TILE);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bars->q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->k_full[i], 1);
      mbar_init(&bars->k_empty[i], 1);
      mbar_init(&bars->v_full[i], 1);
      mbar_init(&bars->v_empty[i], 1);
      mbar_init(&bars->s_full[i], 1);
    }
    mbar_init(&bars->p_full, 8);
    mbar_init(&bars->pv_done, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(&bars->tmem_base, 512);
  if (warp == 0 && lane == 0) tma_prefetch_desc(&tm_qkv);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
  const uint32_t tS0 = tmem, tO = tmem + 256;  // S buffers at columns 0 and 128, O at 256

  if (warp == 0 && lane == 0) {
    // =============================== TMA producer ===============================
    const int colq = h * TC_D, colk = H + h * TC_D, colv = 2 * H + h * TC_D;
    mbar_arrive_expect_tx(&bars->q_full, TC_TILE);
    tma_load_2d(sQ, &tm_qkv, &bars->q_full, colq, seq0 + q0);
    tma_load_2d(sQ + TC_ATOM, &tm_qkv, &bars->q_full, colq + 64, seq0 + q0);
    for (int j = 0; j < n_kv; ++j) {
      const int s = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      const int row = seq0 + j * TC_BK;
      mbar_wait(&bars->k_empty[s], ph ^ 1);
      mbar_arrive_expect_tx(&bars->k_full[s], TC_TILE);
      tma_load_2d(sK + s * TC_TILE, &tm_qkv, &bars->k_full[s], colk, row);
      tma_load_2d(sK + s * TC_TILE + TC_ATOM, &tm_qkv, &bars->k_full[s], colk + 64, row);
      mbar_wait(&bars->v_empty[s], ph ^ 1);
      mbar_arrive_expect_tx(&bars->v_full[s], TC_TILE);
      tma_load_2d(sV + s * TC_TILE, &tm_qkv, &bars->v_full[s], colv, row);
      tma_load_2d(sV + s * TC_TILE + TC_ATOM, &tm_qkv, &bars->v_full[s], colv + 64, row);
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (whole warp, elected lane issues; see k4_attention_tc2.cu) ==========
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 128, 0, 1);  // B (= V) is MN-major
    const uint64_t dQ = umma_desc_kmajor_sw128(smem_u32(sQ)), dP = umma_desc_kmajor_sw128(smem_u32(sP));
    const uint64_t dK0 = umma_desc_kmajor_sw128(smem_u32(sK));                  // stage s adds s * TILE / 16
    const uint64_t dV0 = umma_desc_mnmajor_sw128(smem_u32(sV), TC_ATOM, 1024);  // 16 keys = 2 groups of 8 rows
    auto issue_s = [&](int j) {
      const int s = j & 1;
      mbar_wait(&bars->k_full[s], (j >> 1) & 1);
      tc_fence_after();
      const uint64_t dk = dK0 + static_cast<uint64_t>(s * (TC_TILE >> 4));
      if (elect_one_sync()) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {  // d = 128 in K=16 steps; atom = kk / 4, 32 B per step inside the atom
          const uint64_t off = static_cast<uint64_t>(((kk >> 2) * TC_ATOM + (kk & 3) * 32) >> 4);
          umma_ss(tS0 + s * 128, dQ + off, dk + off, idesc_s, kk > 0);
        }
        umma_commit(&bars->k_empty[s]);
        umma_commit(&bars->s_full[s]);
      }
      __syncwarp();
    };
    mbar_wait(&bars->q_full, 0);
    issue_s(0);
    for (int j = 0; j < n_kv; ++j) {
      if (j + 1 < n_kv) issue_s(j + 1);
      const int s = j & 1;
      mbar_wait(&bars->p_full, j & 1);
      mbar_wait(&bars->v_full[s], (j >> 1) & 1);
      tc_fence_after();
      const uint64_t dv = dV0 + static_cast<uint64_t>(s * (TC_TILE >> 4));
      if (elect_one_sync()) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)  // 128 keys in K=16 steps: P K-major (atom = kk / 4), V MN-major (2048 B per step)
          umma_ss(tO, dP + static_cast<uint64_t>(((kk >> 2) * TC_ATOM + (kk & 3) * 32) >> 4),
                  dv + static_cast<uint64_t>(kk * (2048 >> 4)), idesc_pv, (j > 0) || (kk > 0));
        umma_commit(&bars->v_empty[s]);
        umma_commit(&bars->pv_done);
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    // =============================== softmax + epilogue ===============================
    // 8 warps: TWO threads per query row (64 keys each), so every SM sub-partition has two warps to interleave
    // (one warp per scheduler could not hide the TMEM-load / MUFU latencies: the softmax ran ~3x slower than its
    // exp2 throughput bound).  The row maximum is exchanged through shared memory; the row sums stay per-thread
    // partial sums (same running max, same rescale decisions) and are added in the epilogue.
    const int sw = warp - 4;
    const int ew = sw & 3;       // TMEM lane quarter
    const int hf = sw >> 2;      // key half inside a block / output-column half in the epilogue
    const int r = ew * 32 + lane;  // query row inside the block == TMEM lane
    const uint32_t lane_sel = static_cast<uint32_t>(ew * 32) << 16;
    float m_run = -INFINITY, l_run = 0.f;
    const uint32_t aP = smem_u32(sP);
    const uint32_t p_row = aP + hf * TC_ATOM + r * 128;  // this thread's 128-byte row of key atom hf
    float* xchg = bars->xchg;                             // [2 (block parity)][2 (half)][128 rows]

    for (int j = 0; j < n_kv; ++j) {
      const int sb = j & 1;
      mbar_wait(&bars->s_full[sb], (j >> 1) & 1);
      tc_fence_after();
      uint32_t sraw[2][32];
#pragma unroll
      for (int c = 0; c < 2; ++c) tmem_ld_32x32b_x32(tS0 + lane_sel + sb * 128 + hf * 64 + c * 32, sraw[c]);
      tmem_ld_wait();
      float mx0 = -INFINITY, mx1 = -INFINITY;  // two chains
      if (j == n_kv - 1) {
        // last block.  causal: it is the diagonal block, key (j*128 + k) visible iff k <= r.  non-causal: keys past the
        // end of the sample (the next sample's tokens / the zeroed tail) are masked, k <= len - 1 - j*128
        const int kmax = causal ? r : len - 1 - j * TC_BK;
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            if (hf * 64 + c * 32 + i > kmax) sraw[c][i] = 0xff800000u;  // -inf
            if (hf * 64 + c * 32 + i + 1 > kmax) sraw[c][i + 1] = 0xff800000u;
            mx0 = fmaxf(mx0, __uint_as_float(sraw[c][i]));
            mx1 = fmaxf(mx1, __uint_as_float(sraw[c][i + 1]));
          }
      } else {
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            mx0 = fmaxf(mx0, __uint_as_float(sraw[c][i]));
            mx1 = fmaxf(mx1, __uint_as_float(sraw[c][i + 1]));
          }
      }
      xchg[(sb * 2 + hf) * 128 + r] = fmaxf(mx0, mx1);
      named_bar_sync_256();
      const float mx = fmaxf(xchg[(sb * 2) * 128 + r], xchg[(sb * 2 + 1) * 128 + r]);
      // lazy rescale: keep the stale max unless the new one is more than 2^8 larger
      float alpha = 1.0f;
      bool rescale = false;
      if (mx > m_run && (mx - m_run) * scale_log2 > TC_RESCALE_THRESHOLD) {  // also true for m_run == -inf
        alpha = ex2_approx((m_run - mx) * scale_log2);
        m_run = mx;
        rescale = j > 0;
      }
      const float msc = m_run * scale_log2;
      // exp2(s * scale - m): packed FFMA2 for the affine part, one MUFU.EX2 per element (ex2.approx.ftz -- exp2f()
      // costs FSETP + 2 FMUL around the MUFU for denormal results that vanish in the bf16 rounding of P anyway),
      // packed FADD2 row sums on two chains.
      const uint64_t sc2 = f2_pack(scale_log2, scale_log2), nm2 = f2_pack(-msc, -msc);
      uint64_t rs2[2] = {f2_pack(0.f, 0.f), f2_pack(0.f, 0.f)};
      // P (bf16 pairs) is packed in place: pair i of chunk c lands in sraw[c][i / 2], which is already consumed
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float x0, x1;
          f2_unpack(f2_fma(f2_pack(__uint_as_float(sraw[c][i]), __uint_as_float(sraw[c][i + 1])), sc2, nm2), x0, x1);
          const float p0 = ex2_approx(x0), p1 = ex2_approx(x1);
          rs2[(i >> 1) & 1] = f2_add(rs2[(i >> 1) & 1], f2_pack(p0, p1));
          sraw[c][i >> 1] = pack_bf16(p0, p1);
        }
      float rs0, rs1;
      f2_unpack(f2_add(rs2[0], rs2[1]), rs0, rs1);
      l_run = l_run * alpha + (rs0 + rs1);

      if (j > 0) {  // P buffer and O are free once P_{j-1}.V_{j-1} has completed
        mbar_wait(&bars->pv_done, (j - 1) & 1);
        tc_fence_after();
      }
      // P -> shared memory, UMMA K-major SWIZZLE_128B: (row, 16-byte chunk c16) at row*128 + ((c16 ^ row%8) * 16)
#pragma unroll
      for (int c16 = 0; c16 < 8; ++c16) {
        const uint32_t addr = p_row + ((c16 ^ (r & 7)) << 4);
        // chunk c16 = this half's keys [8*c16, 8*c16 + 8) = packed pairs 4*(c16 % 4) .. +3 of S chunk c16 / 4
        const uint32_t* pp = &sraw[c16 >> 2][(c16 & 3) * 4];
        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(pp[0]), "r"(pp[1]), "r"(pp[2]),
                     "r"(pp[3])
                     : "memory");
      }
      if (__any_sync(0xffffffffu, rescale)) {  // this thread's half of the O columns
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          uint32_t o[32];
          tmem_ld_32x32b_x32(tO + lane_sel + hf * 64 + c * 32, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float a0, a1;
            f2_unpack(f2_mul(f2_pack(__uint_as_float(o[i]), __uint_as_float(o[i + 1])), f2_pack(alpha, alpha)), a0, a1);
            o[i] = __float_as_uint(a0);
            o[i + 1] = __float_as_uint(a1);
          }
          tmem_st_32x32b_x32(tO + lane_sel + hf * 64 + c * 32, o);
        }
        tmem_st_wait();
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->p_full);
    }

    // ---- epilogue: this thread normalises and stores output columns [64*hf, 64*hf + 64) of its row ----
    xchg[((n_kv & 1) * 2 + hf) * 128 + r] = l_run;   // parity slot not used by the last block's max exchange
    mbar_wait(&bars->pv_done, (n_kv - 1) & 1);
    tc_fence_after();
    named_bar_sync_256();
    const float l_tot = xchg[((n_kv & 1) * 2) * 128 + r] + xchg[((n_kv & 1) * 2 + 1) * 128 + r];
    const float inv_l = 1.0f / l_tot;
    const uint32_t stage = aP + sw * 4096;  // this warp's 32 rows x 128 B (P is dead now)
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t o[32];
      tmem_ld_32x32b_x32(tO + lane_sel + hf * 64 + c * 32, o);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int piece = c * 4 + i;
        const uint32_t addr = stage + lane * 128 + ((piece ^ (lane & 7)) << 4);
        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr),
                     "r"(pack_bf16(__uint_as_float(o[8 * i + 0]) * inv_l, __uint_as_float(o[8 * i + 1]) * inv_l)),
                     "r"(pack_bf16(__uint_as_float(o[8 * i + 2]) * inv_l, __uint_as_float(o[8 * i + 3]) * inv_l)),
                     "r"(pack_bf16(__uint_as_float(o[8 * i + 4]) * inv_l, __uint_as_float(o[8 * i + 5]) * inv_l)),
                     "r"(pack_bf16(__uint_as_float(o[8 * i + 6]) * inv_l, __uint_as_float(o[8 * i + 7]) * inv_l))
                     : "memory");
      }
    }
    __syncwarp();
    const int tok = seq0 + q0 + r;
    const int my_dst = (q0 + r < len) ? (out_row_map ? out_row_map[tok] : tok) : -1;
    // training: log2-domain log-sum-exp of the scaled scores, [heads, rows_cap] (read back by the backward kernels)
    if (lse != nullptr && hf == 0 && q0 + r < len)
      lse[static_cast<int64_t>(h) * rows_cap + tok] = fmaf(m_run, scale_log2, log2f(l_tot));
#pragma unroll 4
    for (int it = 0; it < 8; ++it) {  // 4 rows of 128 B per iteration, 8 lanes each
      const int rr = it * 4 + (lane >> 3), piece = lane & 7;
      const int dst = __shfl_sync(0xffffffffu, my_dst, rr);
      if (dst >= 0) {
        uint4 v;
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                     : "r"(stage + rr * 128 + ((piece ^ (rr & 7)) << 4)));
        *reinterpret_cast<uint4*>(out + static_cast<int64_t>(dst) * H + h * TC_D + hf * 64 + piece * 8) = v;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// shared with k4_attention_tc2.cu (a __global__ function cannot be launched from another translation unit without -rdc)
int launch_zero_tail_rows(void* qkv, const int32_t* cu_seqlens, int B, int rows_cap, int row_elems, cudaStream_t s) {
  zero_tail_rows<<<32, 256, 0, s>>>(static_cast<__nv_bfloat16*>(qkv), cu_seqlens, B, rows_cap, row_elems);
  VEX_LAUNCH_CHECK();
  return VEX_OK;
}

int launch_attention_tc(const void* qkv, const int32_t* cu_seqlens, int B, int max_len_cap, int heads,
                        const int32_t* out_row_map, void* out, float scale, int rows_cap, float* lse, int causal,
                        cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    VEX_CUDA_TRY(cudaFuncSetAttribute(k4_attention_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
    configured = true;
  }
  const int H = heads * TC_D;
  CUtensorMap tm;
  std::memset(&tm, 0, sizeof(tm));
  int rc = make_tmap_2d(&tm, qkv, rows_cap, 3 * static_cast<uint64_t>(H), 3 * static_cast<uint64_t>(H), 128);
  if (rc != VEX_OK) return rc;
  if ((rc = launch_zero_tail_rows(const_cast<void*>(qkv), cu_seqlens, B, rows_cap, 3 * H, s)) != VEX_OK) return rc;
  dim3 grid(ceil_div(max_len_cap, TC_BQ), heads, B);
  k4_attention_tc<<<grid, TC_THREADS, TC_SMEM, s>>>(tm, cu_seqlens, heads, out_row_map,
                                                    static_cast<__nv_bfloat16*>(out),
                                                    scale * 1.4426950408889634f, lse, rows_cap, causal);
  VEX_LAUNCH_CHECK();
  return VEX_OK;
}

}  // namespace vex
