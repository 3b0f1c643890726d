// K4 (fast path, second generation) -- two-query-tile ping-pong flash attention on tcgen05 / TMEM for sm_100a.
//
// Same contract as k4_attention_tc.cu (attention_fn prefill branch, modeling_cogvlm.py:106-128: per sample, token i
// attends to tokens j <= i in token-rank order, scale d^-0.5, fp32 online softmax, P rounded to bf16 before P.V;
// non-causal block-diagonal form for the vision encoder, visual.py:96-98).
//
// Why a second kernel: in the one-tile kernel the softmax of a 128 x 128 score block (16 384 MUFU.EX2 = 1 024 cycles
// at 16 / clk / SM, plus TMEM loads, row max, packing) is a serial stage between S = Q.K^T and O += P.V -- two threads
// per query row had to exchange the row maximum through shared memory behind a 256-thread barrier, which put both
// warps of every scheduler in lock-step; ncu: tensor pipe 32 % active, ~3 200 cycles per block
// (profiles/r1_attn_tc_s5a.md).  Here one CTA owns TWO query tiles (A, B: 2 x 128 rows of one (sample, head)) that
// share the K/V stream:
//   warp 0 / lane 0 : TMA producer -- Q_A, Q_B once, then K_j, V_j through ONE 3-slot ring (K_j, V_j, K_{j+1} in
//                     flight), straight out of the token-order QKV buffer [rows_cap, 3*heads*128].
//   warp 1, warp 3  : MMA issuers of tile A / tile B.  The whole warp runs the loop and one elected lane issues
//                     (a single-thread issuer for both tiles set the pace of the first version: ~110 cycles per
//                     64-cycle MMA, profiles/r1_attn_tc2_s8.md).  Per key block and tile: S_x(j+1) and PV_x(j), in the
//                     order the schedule (below) prescribes; ring slots are released by both issuers.
//   warp 2          : TMEM allocate / free: S_A, S_B, O_A, O_B = 4 x 128 columns.
//   warps 4..7      : softmax of tile A, ONE thread per query row (TMEM lane == row: no shuffles, no exchange);
//   warps 8..11     : softmax of tile B.  The 128 scores of a row stay in registers (setmaxnreg: 224 for the softmax
//                     warpgroups, 56 for warps 0..3); half of the exp2 pairs run on the FMA pipe (A2_EMU).
// Schedules (VEX_ATTN_P; the persistent kernel k4_attention_tc3.cu uses the default one):
//   "early" (default, ES = 1): P in shared memory in the UMMA K-major 128B-swizzle layout; S_x(j+1) is issued as soon
//                     as the softmax warps hold S_x(j) in registers (s_free), so the softmax of block j+1 starts the
//                     moment block j is done and overlaps PV_x(j); pv_done gates the next P store / O rescale.
//   "token" (ES = 2): "early" plus exp loops of the two tiles of a scheduler taking turns (named-barrier pairs).
//   "tmem"  (PT)    : bf16 pairs written back into the first 64 columns of the tile's S accumulator (tcgen05.st) and
//                     consumed as the TMEM A operand of O += P.V; S_x(j+1) is issued after PV_x(j), and tcgen05.commit
//                     tracks every earlier MMA of the issuing thread, so s_full_x(j+1) also means "PV_x(j) done".
//   "smem"          : the "tmem" order with P in shared memory.
#include <cuda.h>

#include <cstdlib>
#include <cstring>

#include "../common.cuh"

#ifdef VEX_ATTN_TRACE
// timing experiment build (tools/attn_trace.py): cycle stamps of one CTA's softmax warps and MMA issuer
__device__ long long* g_attn_trace = nullptr;  // [4 roles][64 steps][8 stamps]
extern "C" int vex_debug_attn_trace(long long* buf) {
  return cudaMemcpyToSymbol(g_attn_trace, &buf, sizeof(buf)) == cudaSuccess ? 0 : -1;
}
#define A2_TRACE(role, j, k)                                                                          \
  do {                                                                                                \
    if (g_attn_trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (j) < 64)            \
      g_attn_trace[((role) * 64 + (j)) * 8 + (k)] = clock64();                                        \
  } while (0)
#else
#define A2_TRACE(role, j, k) \
  do {                       \
  } while (0)
#endif

namespace vex {

int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);
int launch_zero_tail_rows(void* qkv, const int32_t* cu_seqlens, int B, int rows_cap, int row_elems, cudaStream_t s);

constexpr int A2_BQ = 128, A2_BK = 128, A2_D = 128;
constexpr int A2_THREADS = 384;          // TMA, MMA, TMEM-alloc, idle + 2 x 4 softmax warps
constexpr int A2_TILE = 128 * 128 * 2;   // 32 KB: one 128 x 128 bf16 tile = two 64-column atoms of 16 KB
constexpr int A2_ATOM = 128 * 64 * 2;    // 16 KB
constexpr int A2_KV_SLOTS = 3;
constexpr float A2_RESCALE_THRESHOLD = 8.0f;  // log2 units
constexpr int A2_REGS_SOFTMAX = 224, A2_REGS_OTHER = 56;
#ifndef A2_EXP_BATCH
#define A2_EXP_BATCH 32
#endif
#ifndef A2_EMU
#define A2_EMU 4  // pairs out of every 8 whose exp2 runs on the FMA pipe
#endif

// PT: P lives in TMEM (aliases S; PV is issued before the tile's next S).  ES ("early S", P in shared memory only):
// S_x(j+1) is issued as soon as the softmax warps have pulled S_x(j) into registers, so a tile's softmax of block j+1
// starts the moment block j is done and overlaps PV_x(j) -- each tile is software-pipelined on its own and the two
// tiles only share the tensor pipe and the K/V stream.
template <bool PT>
struct A2Cfg {
  static constexpr int TILES = 2 + A2_KV_SLOTS + (PT ? 0 : 2);  // Q_A, Q_B, KV ring, (P_A, P_B)
  static constexpr int SMEM = TILES * A2_TILE + 256 + 1024;     // + barriers + alignment slack
};

struct Attn2Bars {
  uint64_t q_full, kv_full[A2_KV_SLOTS], kv_empty[A2_KV_SLOTS], s_full[2], p_full[2], pv_done[2], s_free[2];
  uint32_t tmem_base;
};

template <int N>
__device__ __forceinline__ void reg_alloc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void reg_dealloc() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}

template <bool PT, int ES>
__global__ void __launch_bounds__(A2_THREADS, 1)
    k4_attention_tc2(const __grid_constant__ CUtensorMap tm_qkv, const int32_t* __restrict__ cu_seqlens, int heads,
                     const int32_t* __restrict__ out_row_map, __nv_bfloat16* __restrict__ out, float scale_log2,
                     float* __restrict__ lse, int rows_cap, int causal) {
  const int b = blockIdx.z, h = blockIdx.y;
  const int seq0 = cu_seqlens[b], len = cu_seqlens[b + 1] - seq0;
  const int qp = gridDim.x - 1 - blockIdx.x;  // heaviest query-tile pairs first
  const int q0 = qp * 2 * A2_BQ;              // first row of tile A; tile B starts at q0 + 128
  if (q0 >= len) return;
  // key blocks per tile.  causal: up to the tile's diagonal block; non-causal: every key block of the sample.
  const int n_all = (len + A2_BK - 1) / A2_BK;
  const bool b_active = q0 + A2_BQ < len;
  const int nA = causal ? 2 * qp + 1 : n_all;
  const int nB = b_active ? (causal ? 2 * qp + 2 : n_all) : 0;
  const int n_max = max(nA, nB);
  const int H = heads * A2_D;

  extern __shared__ uint8_t a2_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(a2_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                     // 2 tiles
  uint8_t* sKV = smem + 2 * A2_TILE;      // ring of A2_KV_SLOTS tiles
  uint8_t* sP = smem + (2 + A2_KV_SLOTS) * A2_TILE;  // 2 tiles (PT = false only)
  Attn2Bars* bars = reinterpret_cast<Attn2Bars*>(smem + A2Cfg<PT>::TILES * A2_TILE);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bars->q_full, 1);
    for (int i = 0; i < A2_KV_SLOTS; ++i) {
      mbar_init(&bars->kv_full[i], 1);
      mbar_init(&bars->kv_empty[i], 2);  // both issuers
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->s_full[i], 1);
      mbar_init(&bars->p_full[i], 4);  // one arrive per softmax warp of the tile
      mbar_init(&bars->pv_done[i], 1);
      mbar_init(&bars->s_free[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(&bars->tmem_base, 512);
  if (warp == 0 && lane == 0) tma_prefetch_desc(&tm_qkv);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
  // columns: S_A [0,128), S_B [128,256), O_A [256,384), O_B [384,512); P_X aliases S_X columns [0,64) (PT)

  if (warp < 4) {
    reg_dealloc<A2_REGS_OTHER>();
    if (warp == 0 && lane == 0) {
      // =============================== TMA producer ===============================
      const int colq = h * A2_D, colk = H + h * A2_D, colv = 2 * H + h * A2_D;
      mbar_arrive_expect_tx(&bars->q_full, 2 * A2_TILE);
#pragma unroll
      for (int x = 0; x < 2; ++x) {  // rows past the buffer are zero-filled by TMA; rows past `len` are never stored
        tma_load_2d(sQ + x * A2_TILE, &tm_qkv, &bars->q_full, colq, seq0 + q0 + x * A2_BQ);
        tma_load_2d(sQ + x * A2_TILE + A2_ATOM, &tm_qkv, &bars->q_full, colq + 64, seq0 + q0 + x * A2_BQ);
      }
      int slot = 0;
      uint32_t phase = 0;
      auto load_item = [&](int col, int row) {
        mbar_wait(&bars->kv_empty[slot], phase ^ 1);
        mbar_arrive_expect_tx(&bars->kv_full[slot], A2_TILE);
        tma_load_2d(sKV + slot * A2_TILE, &tm_qkv, &bars->kv_full[slot], col, row);
        tma_load_2d(sKV + slot * A2_TILE + A2_ATOM, &tm_qkv, &bars->kv_full[slot], col + 64, row);
        if (++slot == A2_KV_SLOTS) {
          slot = 0;
          phase ^= 1;
        }
      };
      if constexpr (ES) {  // consumption order: K_0, then per key block K_{j+1} (if any), V_j
        load_item(colk, seq0);
        for (int j = 0; j < n_max; ++j) {
          if (j + 1 < n_max) load_item(colk, seq0 + (j + 1) * A2_BK);
          load_item(colv, seq0 + j * A2_BK);
        }
      } else {             // item 2j = K_j, item 2j + 1 = V_j
        for (int j = 0; j < n_max; ++j) {
          load_item(colk, seq0 + j * A2_BK);
          load_item(colv, seq0 + j * A2_BK);
        }
      }
    } else if (warp == 1 || warp == 3) {
      // =============================== MMA issuers ===============================
      // warp 1 issues tile A's GEMMs, warp 3 tile B's: one issuer took ~8 serialised steps per key block (barrier poll,
      // elect, 8 MMAs, commits; >= 230 cycles each on a scheduler it shares with two busy softmax warps) and, not the
      // tensor pipe or the softmax, set the pace (tools/attn_trace.py).  Nothing orders the two tiles' MMAs against
      // each other; the K/V ring slots are released by BOTH issuers (kv_empty counts 2 arrivals; an issuer whose tile
      // does not need an item passes it through with a plain arrive).
      // The whole warp runs this branch (warp-uniform control flow, every lane polls the barriers) and ONE elected
      // lane issues tcgen05.mma / tcgen05.commit: with a single-thread branch ptxas wraps every MMA in an
      // elect/waterfall loop with four R2UR moves and rebuilds both descriptors (19 instructions, ~110 cycles per
      // 64-cycle MMA on a scheduler shared with two softmax warps -- the issuer, not the tensor pipe, set the pace).
      // Descriptors: one base per operand tile, k-steps are constant adds to the 14-bit address field.
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 128, 0, 1);  // B (= V) is MN-major
      const uint32_t aKV = smem_u32(sKV);
      const uint64_t dQ[2] = {umma_desc_kmajor_sw128(smem_u32(sQ)), umma_desc_kmajor_sw128(smem_u32(sQ + A2_TILE))};
      const uint64_t dP[2] = {umma_desc_kmajor_sw128(smem_u32(sP)), umma_desc_kmajor_sw128(smem_u32(sP + A2_TILE))};
      const uint64_t dK0 = umma_desc_kmajor_sw128(aKV);                 // slot 0; slot s adds s * TILE / 16
      const uint64_t dV0 = umma_desc_mnmajor_sw128(aKV, A2_ATOM, 1024);  // 16 keys = 2 groups of 8 rows (SBO 1024 B),
                                                                        // d chunks of 64 are 16 KB apart (LBO)
      auto wait_item = [&](int idx) {  // ring item idx (old schedule: K_j = 2j, V_j = 2j + 1)
        mbar_wait(&bars->kv_full[idx % A2_KV_SLOTS], (idx / A2_KV_SLOTS) & 1);
        tc_fence_after();
      };
      auto pass_item = [&](int idx) {  // not needed by this tile: arrive once it has landed (keeps the phases aligned)
        wait_item(idx);
        if (lane == 0) mbar_arrive(&bars->kv_empty[idx % A2_KV_SLOTS]);
        __syncwarp();
      };
      auto issue_s = [&](int x, int item) {  // S_x = Q_x . K^T, K = ring item `item`
        wait_item(item);
        const uint64_t dq = dQ[x], dk = dK0 + static_cast<uint64_t>((item % A2_KV_SLOTS) * (A2_TILE >> 4));
        if (elect_one_sync()) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {  // d = 128 in K = 16 steps; atom = kk / 4, 32 B per step inside the atom
            const uint64_t off = static_cast<uint64_t>(((kk >> 2) * A2_ATOM + (kk & 3) * 32) >> 4);
            umma_ss(tmem + x * 128, dq + off, dk + off, idesc_s, kk > 0);
          }
          umma_commit(&bars->s_full[x]);
          umma_commit(&bars->kv_empty[item % A2_KV_SLOTS]);
        }
        __syncwarp();
      };
      auto issue_pv = [&](int x, int j, int item) {  // O_x (+)= P_x(j) . V_j, V_j = ring item `item`
        mbar_wait(&bars->p_full[x], j & 1);
        wait_item(item);
        const uint64_t dv = dV0 + static_cast<uint64_t>((item % A2_KV_SLOTS) * (A2_TILE >> 4));
        const uint64_t dp = dP[x];
        const uint32_t tO = tmem + 256 + x * 128;
        if (elect_one_sync()) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {  // 128 keys in K = 16 steps of 2048 B of V
            if constexpr (PT) {
              umma_ts(tO, tmem + x * 128 + kk * 8, dv + static_cast<uint64_t>(kk * (2048 >> 4)), idesc_pv,
                      (j > 0) || (kk > 0));  // 16 keys = 8 packed columns
            } else {
              umma_ss(tO, dp + static_cast<uint64_t>(((kk >> 2) * A2_ATOM + (kk & 3) * 32) >> 4),
                      dv + static_cast<uint64_t>(kk * (2048 >> 4)), idesc_pv, (j > 0) || (kk > 0));
            }
          }
          umma_commit(&bars->pv_done[x]);
          umma_commit(&bars->kv_empty[item % A2_KV_SLOTS]);
        }
        __syncwarp();
      };
      const int x = warp >> 1;  // 0: tile A, 1: tile B
      const int nX = x ? nB : nA;
      mbar_wait(&bars->q_full, 0);
      tc_fence_after();
      if (nX > 0) issue_s(x, 0); else pass_item(0);
      if constexpr (ES) {  // ring order K_0, then per key block K_{j+1} (if any), V_j
        int cons = 1;
        for (int j = 0; j < n_max; ++j) {
          if (lane == 0) A2_TRACE(2 + x, j, 0);
          if (j + 1 < n_max) {
            const int ki = cons++;
            if (j + 1 < nX) {
              mbar_wait(&bars->s_free[x], j & 1);  // the softmax warps have S_x(j) in registers
              tc_fence_after();
              if (lane == 0) A2_TRACE(2 + x, j, 1);
              issue_s(x, ki);
            } else {
              pass_item(ki);
            }
          }
          if (lane == 0) A2_TRACE(2 + x, j, 2);
          const int vi = cons++;
          if (j < nX) issue_pv(x, j, vi); else pass_item(vi);
          if (lane == 0) A2_TRACE(2 + x, j, 3);
        }
      } else {             // ring order K_j, V_j
        for (int j = 0; j < n_max; ++j) {
          if (j < nX) issue_pv(x, j, 2 * j + 1); else pass_item(2 * j + 1);
          if (j + 1 < n_max) {
            if (j + 1 < nX) issue_s(x, 2 * j + 2); else pass_item(2 * j + 2);
          }
        }
      }
    }
  } else {
    // =============================== softmax + epilogue ===============================
    reg_alloc<A2_REGS_SOFTMAX>();
    const int x = (warp - 4) >> 2;  // tile: 0 = A, 1 = B
    const int ew = warp & 3;        // TMEM lane quarter
    const int r = ew * 32 + lane;   // query row inside the tile == TMEM lane
    const int nX = x ? nB : nA;
    const int qx0 = q0 + x * A2_BQ;
    const uint32_t lane_sel = static_cast<uint32_t>(ew * 32) << 16;
    const uint32_t tS = tmem + lane_sel + x * 128;
    const uint32_t tO = tmem + lane_sel + 256 + x * 128;
    float m_run = -INFINITY, l_run = 0.f;  // m_run: running reference in scaled log2 units (integer-valued)
    const float inv_scale_log2 = 1.0f / scale_log2;
    // ES == 2: the exp loops of the tile-A and tile-B warp of one scheduler take turns (named barrier pair per
    // scheduler, 64 threads each), so the two never split the 4-lane MUFU unit; the token starts with tile A.
    const int n_turns = (ES == 2) ? min(nA, nB) : 0;
    if (ES == 2 && x == 1 && n_turns > 0) asm volatile("bar.arrive %0, 64;" ::"r"(1 + ew) : "memory");

    for (int j = 0; j < nX; ++j) {
      if (lane == 0 && ew == 0) A2_TRACE(x, j, 0);
      mbar_wait(&bars->s_full[x], j & 1);
      tc_fence_after();
      if (lane == 0 && ew == 0) A2_TRACE(x, j, 1);
      uint32_t s[128];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld_32x32b_x32(tS + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&s[c * 32]));
      tmem_ld_wait();
      if constexpr (ES) {  // S_x is in registers: the tensor core may overwrite it with S_x(j+1)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->s_free[x]);
      }
      if (j == nX - 1) {
        // last block.  causal: the tile's diagonal block, key (j*128 + k) visible iff k <= r.  non-causal: keys past
        // the end of the sample (the next sample's tokens / the zeroed tail) are masked, k <= len - 1 - j*128
        const int kmax = causal ? r : len - 1 - j * A2_BK;
#pragma unroll
        for (int i = 0; i < 128; ++i)
          if (i > kmax) s[i] = 0xff800000u;  // -inf
      }
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 128; i += 4) {
        mx0 = fmaxf(mx0, __uint_as_float(s[i]));
        mx1 = fmaxf(mx1, __uint_as_float(s[i + 1]));
        mx2 = fmaxf(mx2, __uint_as_float(s[i + 2]));
        mx3 = fmaxf(mx3, __uint_as_float(s[i + 3]));
      }
      const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      if (lane == 0 && ew == 0) A2_TRACE(x, j, 2);
      // lazy rescale: keep the stale reference unless the new maximum is more than 2^8 above it.  The reference m_run
      // lives in scaled (log2) units and is INTEGER-valued (ceil of the scaled maximum): any reference cancels in the
      // normalisation, and an integer one makes the range reduction of the emulated exp2 below exact.
      const float mxs = mx * scale_log2;
      float alpha = 1.0f;
      bool rescale = false;
      if (mxs > m_run + A2_RESCALE_THRESHOLD) {  // also true for m_run == -inf
        const float m_new = ceilf(mxs);
        alpha = ex2_approx(m_run - m_new);
        m_run = m_new;
        rescale = j > 0;
      }
      if (ES == 2 && j < n_turns) asm volatile("bar.sync %0, 64;" ::"r"(1 + 4 * x + ew) : "memory");
      // P = exp2(s * scale - m_run), row sums on two packed FADD2 chains; P (bf16 pairs) is packed in place: pair i
      // lands in s[i / 2], which is already consumed.  One warp gets one MUFU.EX2 through per ~14 cycles (measured:
      // a 128-element row costs ~1 800 cycles whether or not the scheduler's other softmax warp is in its exp loop), so
      // A2_EMU of every 8 pairs are computed on the FMA pipe instead (FA4's trick): n = rint(x) via the 1.5 * 2^23
      // magic-number add (exact: m_run is an integer), f = x - n in [-0.5, 0.5], 2^f by a degree-3 minimax polynomial
      // (max rel. error 7.5e-5, far below the bf16 rounding of P), 2^n by adding n to the exponent field.
      const uint64_t sc2 = f2_pack(scale_log2, scale_log2), nm2 = f2_pack(-m_run, -m_run);
      const float kmag = 12582912.0f - m_run;  // exact (|m_run| << 2^22)
      const uint64_t k2 = f2_pack(kmag, kmag), neg1 = f2_pack(-1.0f, -1.0f);
      const uint64_t c0 = f2_pack(0.9999281168f, 0.9999281168f), c1 = f2_pack(0.6932610273f, 0.6932610273f),
                     c2 = f2_pack(0.2426109761f, 0.2426109761f), c3 = f2_pack(0.0551715381f, 0.0551715381f);
      const float s_lo = (m_run - 126.0f) * inv_scale_log2;  // x >= -126 keeps 2^n a normal number (masked -inf too)
      uint64_t rs2[2] = {f2_pack(0.f, 0.f), f2_pack(0.f, 0.f)};
#pragma unroll
      for (int i0 = 0; i0 < 128; i0 += A2_EXP_BATCH) {
        float xv[A2_EXP_BATCH];
#pragma unroll
        for (int i = 0; i < A2_EXP_BATCH; i += 2) {
          const int q = i >> 1;
          float a = __uint_as_float(s[i0 + i]), b = __uint_as_float(s[i0 + i + 1]);
          if (((q * A2_EMU) & 7) < A2_EMU) {  // FMA-pipe exp2
            a = fmaxf(a, s_lo);
            b = fmaxf(b, s_lo);
            const uint64_t s2 = f2_pack(a, b);
            const uint64_t t = f2_fma(s2, sc2, k2);    // magic + rint(x)
            const uint64_t g = f2_fma(t, neg1, k2);    // -m_run - rint(x), exact
            const uint64_t f = f2_fma(s2, sc2, g);     // x - rint(x)
            uint64_t pl = f2_fma(f, c3, c2);
            pl = f2_fma(pl, f, c1);
            pl = f2_fma(pl, f, c0);
            float p0, p1, t0, t1;
            f2_unpack(pl, p0, p1);
            f2_unpack(t, t0, t1);
            xv[i] = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(t0) << 23));
            xv[i + 1] = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(t1) << 23));
          } else {
            f2_unpack(f2_fma(f2_pack(a, b), sc2, nm2), xv[i], xv[i + 1]);
            xv[i] = ex2_approx(xv[i]);
            xv[i + 1] = ex2_approx(xv[i + 1]);
          }
        }
#pragma unroll
        for (int i = 0; i < A2_EXP_BATCH; i += 2) {
          rs2[(i >> 1) & 1] = f2_add(rs2[(i >> 1) & 1], f2_pack(xv[i], xv[i + 1]));
          s[(i0 + i) >> 1] = pack_bf16(xv[i], xv[i + 1]);
        }
      }
      float rs0, rs1;
      f2_unpack(f2_add(rs2[0], rs2[1]), rs0, rs1);
      l_run = l_run * alpha + (rs0 + rs1);
      // hand the token over (tile B's last turn is not handed back: nobody would take it)
      if (ES == 2 && j < n_turns && !(x == 1 && j == n_turns - 1))
        asm volatile("bar.arrive %0, 64;" ::"r"(1 + 4 * (x ^ 1) + ew) : "memory");
      if (lane == 0 && ew == 0) A2_TRACE(x, j, 3);

      // old schedule: s_full(j) was committed after PV(j-1), so P and O of this tile are free here; early-S: wait for it
      if constexpr (ES) {
        if (j > 0) {
          mbar_wait(&bars->pv_done[x], (j - 1) & 1);
          tc_fence_after();
        }
      }
      if (lane == 0 && ew == 0) A2_TRACE(x, j, 4);
      if constexpr (PT) {
        tmem_st_32x32b_x32(tS, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
        tmem_st_32x32b_x32(tS + 32, *reinterpret_cast<uint32_t(*)[32]>(&s[32]));
      } else {
        // UMMA K-major SWIZZLE_128B: key atom a = keys [64a, 64a + 64); (row, 16-byte chunk c16) at
        // row*128 + ((c16 ^ row%8) * 16); chunk c16 of atom a = packed pairs s[32a + 4*c16 .. + 3]
        const uint32_t p_row = smem_u32(sP + x * A2_TILE) + r * 128;
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int c16 = 0; c16 < 8; ++c16) {
            const uint32_t addr = p_row + a * A2_ATOM + ((c16 ^ (r & 7)) << 4);
            const uint32_t* pp = &s[32 * a + 4 * c16];
            asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(pp[0]), "r"(pp[1]), "r"(pp[2]),
                         "r"(pp[3])
                         : "memory");
          }
      }
      if (__any_sync(0xffffffffu, rescale)) {
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t o[32];
          tmem_ld_32x32b_x32(tO + c * 32, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float a0, a1;
            f2_unpack(f2_mul(f2_pack(__uint_as_float(o[i]), __uint_as_float(o[i + 1])), f2_pack(alpha, alpha)), a0, a1);
            o[i] = __float_as_uint(a0);
            o[i + 1] = __float_as_uint(a1);
          }
          tmem_st_32x32b_x32(tO + c * 32, o);
        }
      }
      if constexpr (PT) {
        tmem_st_wait();
      } else {
        tmem_st_wait();
        fence_proxy_async_smem();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->p_full[x]);
      if (lane == 0 && ew == 0) A2_TRACE(x, j, 5);
    }

    if (nX > 0) {
      // ---- epilogue: normalise the row, stage it in the tile's (dead) Q buffer, scatter coalesced rows ----
      mbar_wait(&bars->pv_done[x], (nX - 1) & 1);  // every MMA of this tile (S and PV) has completed
      tc_fence_after();
      const float inv_l = 1.0f / l_run;
      const uint32_t stage = smem_u32(sQ + x * A2_TILE) + ew * 8192;  // this warp's 32 rows x 256 B
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t o[32];
        tmem_ld_32x32b_x32(tO + c * 32, o);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int piece = c * 4 + i;  // 16 pieces of 16 B per row
          const uint32_t addr = stage + lane * 256 + ((piece ^ (lane & 7)) << 4);
          asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr),
                       "r"(pack_bf16(__uint_as_float(o[8 * i + 0]) * inv_l, __uint_as_float(o[8 * i + 1]) * inv_l)),
                       "r"(pack_bf16(__uint_as_float(o[8 * i + 2]) * inv_l, __uint_as_float(o[8 * i + 3]) * inv_l)),
                       "r"(pack_bf16(__uint_as_float(o[8 * i + 4]) * inv_l, __uint_as_float(o[8 * i + 5]) * inv_l)),
                       "r"(pack_bf16(__uint_as_float(o[8 * i + 6]) * inv_l, __uint_as_float(o[8 * i + 7]) * inv_l))
                       : "memory");
        }
      }
      __syncwarp();
      const int tok = seq0 + qx0 + r;
      const bool row_ok = qx0 + r < len;
      const int my_dst = row_ok ? (out_row_map ? out_row_map[tok] : tok) : -1;
      // training: log2-domain log-sum-exp of the scaled scores, [heads, rows_cap] (read back by the backward kernels)
      if (lse != nullptr && row_ok) lse[static_cast<int64_t>(h) * rows_cap + tok] = m_run + log2f(l_run);
#pragma unroll 4
      for (int it = 0; it < 16; ++it) {  // 2 rows of 256 B per iteration, 16 lanes each
        const int rr = it * 2 + (lane >> 4), piece = lane & 15;
        const int dst = __shfl_sync(0xffffffffu, my_dst, rr);
        if (dst >= 0) {
          uint4 v;
          asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                       : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                       : "r"(stage + rr * 256 + ((piece ^ (rr & 7)) << 4)));
          *reinterpret_cast<uint4*>(out + static_cast<int64_t>(dst) * H + h * A2_D + piece * 8) = v;
        }
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

template <bool PT, int ES>
static int launch_tc2(const CUtensorMap& tm, const int32_t* cu_seqlens, int B, int max_len_cap, int heads,
                      const int32_t* out_row_map, void* out, float scale, int rows_cap, float* lse, int causal,
                      cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    VEX_CUDA_TRY(cudaFuncSetAttribute(k4_attention_tc2<PT, ES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      A2Cfg<PT>::SMEM));
    configured = true;
  }
  dim3 grid(ceil_div(max_len_cap, 2 * A2_BQ), heads, B);
  k4_attention_tc2<PT, ES><<<grid, A2_THREADS, A2Cfg<PT>::SMEM, s>>>(tm, cu_seqlens, heads, out_row_map,
                                                                 static_cast<__nv_bfloat16*>(out),
                                                                 scale * 1.4426950408889634f, lse, rows_cap, causal);
  VEX_LAUNCH_CHECK();
  return VEX_OK;
}

int launch_attention_tc2(const void* qkv, const int32_t* cu_seqlens, int B, int max_len_cap, int heads,
                         const int32_t* out_row_map, void* out, float scale, int rows_cap, float* lse, int causal,
                         cudaStream_t s) {
  const int H = heads * A2_D;
  CUtensorMap tm;
  std::memset(&tm, 0, sizeof(tm));
  int rc = make_tmap_2d(&tm, qkv, rows_cap, 3 * static_cast<uint64_t>(H), 3 * static_cast<uint64_t>(H), 128);
  if (rc != VEX_OK) return rc;
  if ((rc = launch_zero_tail_rows(const_cast<void*>(qkv), cu_seqlens, B, rows_cap, 3 * H, s)) != VEX_OK) return rc;
  // A/B switch (tools/bench_kernels.py): where P lives and when S(j+1) is issued.  Default "early": P in shared
  // memory, S_x(j+1) issued as soon as the softmax warps hold S_x(j) (fastest on B200, profiles/r1_attn_tc2_s8.md);
  // "tmem": P aliases S in TMEM (TS-mode PV); "smem": the "tmem" schedule with P in shared memory; "token": "early"
  // plus exp loops of the two tiles taking turns.
  const char* pe = std::getenv("VEX_ATTN_P");
  if (pe && std::strcmp(pe, "smem") == 0)
    return launch_tc2<false, 0>(tm, cu_seqlens, B, max_len_cap, heads, out_row_map, out, scale, rows_cap, lse,
                                causal, s);
  if (pe && std::strcmp(pe, "tmem") == 0)
    return launch_tc2<true, 0>(tm, cu_seqlens, B, max_len_cap, heads, out_row_map, out, scale, rows_cap, lse,
                               causal, s);
  if (pe && std::strcmp(pe, "token") == 0)
    return launch_tc2<false, 2>(tm, cu_seqlens, B, max_len_cap, heads, out_row_map, out, scale, rows_cap, lse,
                                causal, s);
  return launch_tc2<false, 1>(tm, cu_seqlens, B, max_len_cap, heads, out_row_map, out, scale, rows_cap, lse,
                              causal, s);
}

}  // namespace vex
