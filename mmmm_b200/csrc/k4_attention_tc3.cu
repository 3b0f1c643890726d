// K4 (fast path, third generation) -- PERSISTENT two-query-tile flash attention on tcgen05 / TMEM for sm_100a.
//
// Same contract and the same inner loop as k4_attention_tc2.cu in its default ("early S") schedule (attention_fn prefill
// branch, modeling_cogvlm.py:106-128; non-causal block-diagonal form for the vision encoder, visual.py:96-98).
//
// Why a third kernel: a two-tile CTA needs all 512 TMEM columns and 224 KB of shared memory, so one CTA runs per SM and
// nothing overlapped a CTA's ~8 000 cycles of prologue (barrier init, TMEM allocation, first Q / K loads: ~2 000 cycles
// until the first S), pipeline fill and epilogue with another CTA's key blocks -- 26 % of an average c2 work item
// (6.5 key-block iterations of ~3 450 cycles; tools/attn_trace.py, profiles/r1_attn_tc2_s8.md).  Here one CTA per SM
// pulls work items (sample, head, query-tile pair) from a global atomic counter.  Items are ordered in waves of a few
// (sample, head) groups, heaviest pair first inside a wave (a3_decode): the CTAs of the chip work on the pairs of the same
// groups at the same time, so K / V come out of L2 (a chip-wide heaviest-first static deal read 753 MB from DRAM per c2
// launch instead of 292 MB), and the greedy hand-out balances ragged batches.  The last three waves are merged into ONE
// heaviest-first wave, so the launch ends on one-block items instead of a few CTAs finishing 12-block items while the
// rest of the chip idles (round 2: c2 181 -> 169 us, profiles/r2_k4_schedule.md).  The TMA producer thread is the
// scheduler: it publishes each fetched item through a small shared-memory ring (sched_full / sched_empty) that the
// issuer and softmax warps follow:
//   * barriers, TMEM and tensor-map prefetch are set up once per CTA;
//   * the TMA producer runs ahead across items: the next item's Q tiles are loaded as soon as the last S of the current
//     item has been issued (q_empty), its K/V blocks simply continue in the 3-slot ring;
//   * the issuer warps start the next item's S(0) while the softmax warps normalise and store the current item's O (the
//     epilogue stages in the tile's P buffer, rows of the owning warp only), and wait for o_free before PV(0);
//   * every mbarrier phase comes from a running counter kept identically by the roles that share the barrier.
// Roles: warp 0 lane 0 TMA producer; warp 1 / warp 3 MMA issuers of tile A / tile B (warp-uniform loop, elected lane
// issues); warp 2 TMEM allocation; warps 4..7 / 8..11 softmax + epilogue of tile A / B, one thread per query row.
#include <cuda.h>

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

#ifdef VEX_ATTN_TRACE
// timing experiment build (tools/a3_trace.py): cycle stamps of tile A's softmax warp 0 in CTA 0
__device__ long long* g_a3_trace = nullptr;  // [128 blocks][8 stamps]
extern "C" int vex_debug_a3_trace(long long* buf) {
  return cudaMemcpyToSymbol(g_a3_trace, &buf, sizeof(buf)) == cudaSuccess ? 0 : -1;
}
#define A3_TRACE(ptr, j, k)                                    \
  do {                                                         \
    if ((ptr) && (j) < 128) (ptr)[(j) * 8 + (k)] = clock64();  \
  } while (0)
#else
#define A3_TRACE(ptr, j, k) \
  do {                      \
  } while (0)
#endif

namespace vex {

int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);

constexpr int A3_BQ = 128, A3_BK = 128, A3_D = 128;
constexpr int A3_THREADS = 384;
constexpr int A3_TILE = 128 * 128 * 2;   // 32 KB: one 128 x 128 bf16 tile = two 64-column atoms of 16 KB
constexpr int A3_ATOM = 128 * 64 * 2;    // 16 KB
constexpr int A3_KV_SLOTS = 3;
constexpr float A3_RESCALE_THRESHOLD = 8.0f;  // log2 units
constexpr int A3_REGS_SOFTMAX = 224, A3_REGS_OTHER = 56;
constexpr int A3_EXP_BATCH = 32;
constexpr int A3_EMU = 4;  // pairs out of every 8 whose exp2 runs on the FMA pipe
constexpr int A3_SCHED = 4;                                // scheduler ring depth
constexpr int A3_TAIL_WAVES = 3;                           // waves merged into the final heaviest-first wave
constexpr int A3_SCHED_READERS = 10;                       // 2 issuer warps + 8 softmax warps (lane 0 arrives)
constexpr int A3_NCOUNTERS = 1024;                         // work counters, one per launch in flight (round-robin)
constexpr int A3_TILES = 2 + A3_KV_SLOTS + 2;              // Q_A, Q_B, K/V ring, P_A, P_B
constexpr int A3_SMEM = A3_TILES * A3_TILE + 512 + 1024;   // + barriers / scheduler ring + alignment slack

struct A3Item {  // a decoded work item as the scheduler publishes it; valid == -1: work exhausted, 0: nothing to do
  int valid, seq0, len, h, q0, nA, nB, n_max;
};

struct Attn3Bars {
  uint64_t q_full, q_empty, kv_full[A3_KV_SLOTS], kv_empty[A3_KV_SLOTS];
  uint64_t s_full[2], s_free[2], p_full[2], pv_done[2], o_free[2];
  uint64_t sched_full[A3_SCHED], sched_empty[A3_SCHED];
  A3Item sched_item[A3_SCHED];
  uint32_t tmem_base;
};
static_assert(sizeof(Attn3Bars) <= 512, "barrier block");

__device__ unsigned int g_a3_counters[A3_NCOUNTERS];

// item -> (sample, head, query-tile pair).  Groups (sample, head) are taken in waves of `wave_groups` (about one item
// per CTA and wave); inside a wave the items are ordered heaviest pair first across the wave's groups, so the groups of a
// wave are in flight together (K / V shared through L2) and the last wave drains longest-first.  False when the pair
// lies past the sample's end.
__device__ __forceinline__ bool a3_decode(int item, int heads, int nqp, int n_groups, int wave_groups, int tail_waves,
                                          const int32_t* __restrict__ cu_seqlens, int causal, A3Item& it) {
  const int per_wave = wave_groups * nqp;
  // the last A3_TAIL_WAVES waves form ONE wave: heaviest pair first over all of its groups, so the kernel ends on the
  // one-block items instead of on a few CTAs finishing 12-block items while the others idle (greedy hand-out model,
  // profiles/r2_k4_schedule.md: 0.89 -> 0.98 schedule efficiency at c2, 0.83 -> 0.96 for a two-sample c4 shard); its
  // K / V working set (3 x 148 / nqp groups, 57 MB at c2) still fits L2
  const int n_waves = (n_groups + wave_groups - 1) / wave_groups;
  const int tail_wave = max(0, n_waves - tail_waves);
  int wave = item / per_wave, r = item % per_wave, gw = wave_groups;
  if (wave >= tail_wave) {
    wave = tail_wave;
    r = item - tail_wave * per_wave;
    gw = n_groups - tail_wave * wave_groups;
  }
  const int g = wave * wave_groups + r % gw, qp = nqp - 1 - r / gw;
  const int b = g / heads;
  it.h = g % heads;
  it.seq0 = __ldg(cu_seqlens + b);
  it.len = __ldg(cu_seqlens + b + 1) - it.seq0;
  it.q0 = qp * 2 * A3_BQ;
  it.valid = it.q0 < it.len;
  const int n_all = (it.len + A3_BK - 1) / A3_BK;
  it.nA = causal ? 2 * qp + 1 : n_all;
  it.nB = (it.q0 + A3_BQ < it.len) ? (causal ? 2 * qp + 2 : n_all) : 0;
  it.n_max = max(it.nA, it.nB);
  return it.valid != 0;
}

// Readers of the scheduler ring (issuer / softmax warps): next DECODED item of this CTA (the scheduler thread did the
// cu_seqlens loads once, so no reader waits on global memory between two items); false when the work is exhausted.
__device__ __forceinline__ bool a3_next_item(Attn3Bars* bars, int& n_fetch, int lane, A3Item& w) {
  const int slot = n_fetch % A3_SCHED;
  mbar_wait(&bars->sched_full[slot], (n_fetch / A3_SCHED) & 1);
  w = bars->sched_item[slot];
  __syncwarp();
  if (lane == 0) mbar_arrive(&bars->sched_empty[slot]);
  ++n_fetch;
  return w.valid >= 0;
}

template <int N>
__device__ __forceinline__ void a3_reg_alloc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void a3_reg_dealloc() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}

__global__ void __launch_bounds__(A3_THREADS, 1)
    k4_attention_tc3(const __grid_constant__ CUtensorMap tm_qkv, const int32_t* __restrict__ cu_seqlens, int heads, int B,
                     int nqp, const int32_t* __restrict__ out_row_map, __nv_bfloat16* __restrict__ out,
                     float scale_log2, float* __restrict__ lse, int rows_cap, int causal,
                     unsigned int* __restrict__ work_counter, int wave_groups, int tail_waves) {
  const int H = heads * A3_D;
  const int n_groups = B * heads, n_items = nqp * n_groups;

  extern __shared__ uint8_t a3_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(a3_smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                // 2 tiles
  uint8_t* sKV = smem + 2 * A3_TILE;                 // ring of A3_KV_SLOTS tiles
  uint8_t* sP = smem + (2 + A3_KV_SLOTS) * A3_TILE;  // 2 tiles
  Attn3Bars* bars = reinterpret_cast<Attn3Bars*>(smem + A3_TILES * A3_TILE);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bars->q_full, 1);
    mbar_init(&bars->q_empty, 2);  // one arrival per issuer
    for (int i = 0; i < A3_KV_SLOTS; ++i) {
      mbar_init(&bars->kv_full[i], 1);
      mbar_init(&bars->kv_empty[i], 2);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars->s_full[i], 1);
      mbar_init(&bars->s_free[i], 4);  // one arrive per softmax warp of the tile
      mbar_init(&bars->p_full[i], 4);
      mbar_init(&bars->pv_done[i], 1);
      mbar_init(&bars->o_free[i], 4);
    }
    for (int i = 0; i < A3_SCHED; ++i) {
      mbar_init(&bars->sched_full[i], 1);
      mbar_init(&bars->sched_empty[i], A3_SCHED_READERS);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(&bars->tmem_base, 512);
  if (warp == 0 && lane == 0) tma_prefetch_desc(&tm_qkv);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
  // columns: S_A [0,128), S_B [128,256), O_A [256,384), O_B [384,512)

  if (warp < 4) {
    a3_reg_dealloc<A3_REGS_OTHER>();
    if (warp == 0 && lane == 0) {
      // =============================== TMA producer ===============================
      int slot = 0;
      uint32_t phase = 0;
      auto load_item = [&](int col, int row) {
        mbar_wait(&bars->kv_empty[slot], phase ^ 1);
        mbar_arrive_expect_tx(&bars->kv_full[slot], A3_TILE);
        tma_load_2d(sKV + slot * A3_TILE, &tm_qkv, &bars->kv_full[slot], col, row);
        tma_load_2d(sKV + slot * A3_TILE + A3_ATOM, &tm_qkv, &bars->kv_full[slot], col + 64, row);
        if (++slot == A3_KV_SLOTS) {
          slot = 0;
          phase ^= 1;
        }
      };
      int n_done = 0;   // valid items so far
      int n_fetch = 0;  // items fetched / published so far
      // scheduler: fetch the next item (atomic + cu_seqlens loads, ~2 500 cycles of latency), publish it decoded
      auto fetch_publish = [&]() {
        const int slot_s = n_fetch % A3_SCHED;
        mbar_wait(&bars->sched_empty[slot_s], ((n_fetch / A3_SCHED) & 1) ^ 1);
        const unsigned int fetched = atomicAdd(work_counter, 1u);
        A3Item it;
        it.valid = -1;
        if (fetched < static_cast<unsigned int>(n_items))
          a3_decode(static_cast<int>(fetched), heads, nqp, n_groups, wave_groups, tail_waves, cu_seqlens, causal, it);
        bars->sched_item[slot_s] = it;
        mbar_arrive(&bars->sched_full[slot_s]);  // release: the item is visible to whoever observes the phase
        ++n_fetch;
        return it;
      };
      A3Item w = fetch_publish();
      while (w.valid >= 0) {
        if (w.valid == 0) {
          w = fetch_publish();
          continue;
        }
        const int colq = w.h * A3_D, colk = H + w.h * A3_D, colv = 2 * H + w.h * A3_D;
        if (n_done > 0) mbar_wait(&bars->q_empty, (n_done - 1) & 1);  // every S of the previous item has read Q
        mbar_arrive_expect_tx(&bars->q_full, 2 * A3_TILE);
#pragma unroll
        for (int x = 0; x < 2; ++x) {  // rows past the buffer are zero-filled by TMA; rows past `len` are never stored
          tma_load_2d(sQ + x * A3_TILE, &tm_qkv, &bars->q_full, colq, w.seq0 + w.q0 + x * A3_BQ);
          tma_load_2d(sQ + x * A3_TILE + A3_ATOM, &tm_qkv, &bars->q_full, colq + 64, w.seq0 + w.q0 + x * A3_BQ);
        }
        // consumption order: K_0, then per key block K_{j+1} (if any), V_j
        load_item(colk, w.seq0);
        for (int j = 0; j < w.n_max; ++j) {
          if (j + 1 < w.n_max) load_item(colk, w.seq0 + (j + 1) * A3_BK);
          load_item(colv, w.seq0 + j * A3_BK);
        }
        ++n_done;
        // fetched only now (not an item ahead): an item claimed early is an item no idle CTA can take, and with ~5 items
        // per CTA (c4 shard) the early claim cost 6 % in tail imbalance
        w = fetch_publish();
      }
    } else if (warp == 1 || warp == 3) {
      // =============================== MMA issuers: warp 1 tile A, warp 3 tile B ===============================
      const int x = warp >> 1;
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 128, 0, 1);  // B (= V) is MN-major
      const uint64_t dq = umma_desc_kmajor_sw128(smem_u32(sQ + x * A3_TILE));
      const uint64_t dp = umma_desc_kmajor_sw128(smem_u32(sP + x * A3_TILE));
      const uint64_t dK0 = umma_desc_kmajor_sw128(smem_u32(sKV));                  // slot s adds s * TILE / 16
      const uint64_t dV0 = umma_desc_mnmajor_sw128(smem_u32(sKV), A3_ATOM, 1024);  // 16 keys = 2 groups of 8 rows
      const uint32_t tS = tmem + x * 128, tO = tmem + 256 + x * 128;
      int cons = 0;    // running ring item index
      int n_s = 0;     // S GEMMs issued for this tile so far (s_full / s_free phases)
      int n_pv = 0;    // PV GEMMs issued for this tile so far (p_full / pv_done phases)
      int n_act = 0;   // items in which this tile was active (o_free phase)
      int n_done = 0;  // valid items so far (q_full / q_empty phases)
      auto wait_item = [&](int idx) {
        mbar_wait(&bars->kv_full[idx % A3_KV_SLOTS], (idx / A3_KV_SLOTS) & 1);
        tc_fence_after();
      };
      auto pass_item = [&](int idx) {  // not needed by this tile: arrive once it has landed (keeps the phases aligned)
        wait_item(idx);
        if (lane == 0) mbar_arrive(&bars->kv_empty[idx % A3_KV_SLOTS]);
        __syncwarp();
      };
      auto issue_s = [&](int idx, bool last) {  // S_x = Q_x . K^T, K = ring item idx; `last`: Q is dead afterwards
        if (n_s > 0) {
          mbar_wait(&bars->s_free[x], (n_s - 1) & 1);  // the softmax warps hold the previous S_x in registers
          tc_fence_after();
        }
        wait_item(idx);
        const uint64_t dk = dK0 + static_cast<uint64_t>((idx % A3_KV_SLOTS) * (A3_TILE >> 4));
        if (elect_one_sync()) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {  // d = 128 in K = 16 steps; atom = kk / 4, 32 B per step inside the atom
            const uint64_t off = static_cast<uint64_t>(((kk >> 2) * A3_ATOM + (kk & 3) * 32) >> 4);
            umma_ss(tS, dq + off, dk + off, idesc_s, kk > 0);
          }
          umma_commit(&bars->s_full[x]);
          umma_commit(&bars->kv_empty[idx % A3_KV_SLOTS]);
          if (last) umma_commit(&bars->q_empty);
        }
        __syncwarp();
        ++n_s;
      };
      auto issue_pv = [&](int idx, bool first) {  // O_x (+)= P_x . V, V = ring item idx
        if (first && n_act > 0) {
          mbar_wait(&bars->o_free[x], (n_act - 1) & 1);  // the previous item's O has been read out
          tc_fence_after();
        }
        mbar_wait(&bars->p_full[x], n_pv & 1);
        wait_item(idx);
        const uint64_t dv = dV0 + static_cast<uint64_t>((idx % A3_KV_SLOTS) * (A3_TILE >> 4));
        if (elect_one_sync()) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)  // 128 keys in K = 16 steps of 2048 B of V
            umma_ss(tO, dp + static_cast<uint64_t>(((kk >> 2) * A3_ATOM + (kk & 3) * 32) >> 4),
                    dv + static_cast<uint64_t>(kk * (2048 >> 4)), idesc_pv, !first || (kk > 0));
          umma_commit(&bars->pv_done[x]);
          umma_commit(&bars->kv_empty[idx % A3_KV_SLOTS]);
        }
        __syncwarp();
        ++n_pv;
      };
      int n_fetch = 0;
      for (;;) {
        A3Item w;
        if (!a3_next_item(bars, n_fetch, lane, w)) break;
        if (w.valid == 0) continue;
        const int nX = x ? w.nB : w.nA;
        mbar_wait(&bars->q_full, n_done & 1);
        tc_fence_after();
        if (nX == 0) {  // tile past the end of the sample: Q is not read, ring items are passed through
          if (lane == 0) mbar_arrive(&bars->q_empty);
          __syncwarp();
        }
        const int k0 = cons++;
        if (nX > 0) issue_s(k0, nX == 1); else pass_item(k0);
        for (int j = 0; j < w.n_max; ++j) {
          if (j + 1 < w.n_max) {
            const int ki = cons++;
            if (j + 1 < nX) issue_s(ki, j + 2 == nX); else pass_item(ki);
          }
          const int vi = cons++;
          if (j < nX) issue_pv(vi, j == 0); else pass_item(vi);
        }
        if (nX > 0) ++n_act;
        ++n_done;
      }
    }
  } else {
    // =============================== softmax + epilogue ===============================
    a3_reg_alloc<A3_REGS_SOFTMAX>();
    const int x = (warp - 4) >> 2;  // tile: 0 = A, 1 = B
    const int ew = warp & 3;        // TMEM lane quarter
    const int r = ew * 32 + lane;   // query row inside the tile == TMEM lane
    const uint32_t lane_sel = static_cast<uint32_t>(ew * 32) << 16;
    const uint32_t tS = tmem + lane_sel + x * 128;
    const uint32_t tO = tmem + lane_sel + 256 + x * 128;
    const uint32_t p_tile = smem_u32(sP + x * A3_TILE);
    const float inv_scale_log2 = 1.0f / scale_log2;
    int c = 0;  // key blocks of this tile so far (s_full / s_free / p_full / pv_done phases)
#ifdef VEX_ATTN_TRACE
    long long* trc = (blockIdx.x == 0 && threadIdx.x == 128) ? g_a3_trace : nullptr;  // pointer stays in a register
#endif

    int n_fetch = 0;
    for (;;) {
      A3Item w;
      if (!a3_next_item(bars, n_fetch, lane, w)) break;
      if (w.valid == 0) continue;
      const int nX = x ? w.nB : w.nA;
      if (nX == 0) continue;
      const int qx0 = w.q0 + x * A3_BQ;
      // destination row of this thread's query row, loaded now so that the epilogue does not wait on global memory
      const int tok = w.seq0 + qx0 + r;
      const bool row_ok = qx0 + r < w.len;
      const int my_dst = row_ok ? (out_row_map ? __ldg(out_row_map + tok) : tok) : -1;
      float m_run = -INFINITY, l_run = 0.f;  // m_run: running reference in scaled log2 units (integer-valued)

      for (int j = 0; j < nX; ++j, ++c) {
        A3_TRACE(trc, c, 0);
        mbar_wait(&bars->s_full[x], c & 1);
        tc_fence_after();
        A3_TRACE(trc, c, 1);
        uint32_t s[128];
#pragma unroll
        for (int q = 0; q < 4; ++q) tmem_ld_32x32b_x32(tS + q * 32, *reinterpret_cast<uint32_t(*)[32]>(&s[q * 32]));
        tmem_ld_wait();
        // S_x is in registers: the tensor core may overwrite it with the next S_x (next key block or next item)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->s_free[x]);
        A3_TRACE(trc, c, 2);
        if (j == nX - 1) {
          // last block.  causal: the tile's diagonal block, key (j*128 + k) visible iff k <= r.  non-causal: keys past
          // the end of the sample (the next sample's tokens / the zeroed tail) are masked, k <= len - 1 - j*128
          const int kmax = causal ? r : w.len - 1 - j * A3_BK;
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
        const float mxs = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * scale_log2;
        A3_TRACE(trc, c, 3);
        // lazy rescale against an integer-valued reference (see k4_attention_tc2.cu)
        float alpha = 1.0f;
        bool rescale = false;
        if (mxs > m_run + A3_RESCALE_THRESHOLD) {  // also true for m_run == -inf
          const float m_new = ceilf(mxs);
          alpha = ex2_approx(m_run - m_new);
          m_run = m_new;
          rescale = j > 0;
        }
        // P = exp2(s * scale - m_run); A3_EMU of every 8 pairs on the FMA pipe (magic-number rint, degree-3 minimax
        // polynomial on [-0.5, 0.5], exponent add), the rest on MUFU.EX2; bf16 pairs packed in place
        const uint64_t sc2 = f2_pack(scale_log2, scale_log2), nm2 = f2_pack(-m_run, -m_run);
        const float kmag = 12582912.0f - m_run;  // exact (|m_run| << 2^22)
        const uint64_t k2 = f2_pack(kmag, kmag), neg1 = f2_pack(-1.0f, -1.0f);
        const uint64_t c0 = f2_pack(0.9999281168f, 0.9999281168f), c1 = f2_pack(0.6932610273f, 0.6932610273f),
                       c2 = f2_pack(0.2426109761f, 0.2426109761f), c3 = f2_pack(0.0551715381f, 0.0551715381f);
        const float s_lo = (m_run - 126.0f) * inv_scale_log2;  // x >= -126 keeps 2^n a normal number (masked -inf too)
        uint64_t rs2[2] = {f2_pack(0.f, 0.f), f2_pack(0.f, 0.f)};
#pragma unroll
        for (int i0 = 0; i0 < 128; i0 += A3_EXP_BATCH) {
          float xv[A3_EXP_BATCH];
#pragma unroll
          for (int i = 0; i < A3_EXP_BATCH; i += 2) {
            const int q = i >> 1;
            float a = __uint_as_float(s[i0 + i]), b = __uint_as_float(s[i0 + i + 1]);
            if (((q * A3_EMU) & 7) < A3_EMU) {  // FMA-pipe exp2
              a = fmaxf(a, s_lo);
              b = fmaxf(b, s_lo);
              const uint64_t s2 = f2_pack(a, b);
              const uint64_t t = f2_fma(s2, sc2, k2);  // magic + rint(x)
              const uint64_t g = f2_fma(t, neg1, k2);  // -m_run - rint(x), exact
              const uint64_t f = f2_fma(s2, sc2, g);   // x - rint(x)
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
          for (int i = 0; i < A3_EXP_BATCH; i += 2) {
            rs2[(i >> 1) & 1] = f2_add(rs2[(i >> 1) & 1], f2_pack(xv[i], xv[i + 1]));
            s[(i0 + i) >> 1] = pack_bf16(xv[i], xv[i + 1]);
          }
        }
        float rs0, rs1;
        f2_unpack(f2_add(rs2[0], rs2[1]), rs0, rs1);
        l_run = l_run * alpha + (rs0 + rs1);
        A3_TRACE(trc, c, 4);

        if (j > 0) {  // P_x and O_x are free once the previous PV_x has completed (first block: the epilogue waited)
          mbar_wait(&bars->pv_done[x], (c - 1) & 1);
          tc_fence_after();
        }
        A3_TRACE(trc, c, 5);
        // UMMA K-major SWIZZLE_128B: key atom a = keys [64a, 64a + 64); (row, 16-byte chunk c16) at
        // row*128 + ((c16 ^ row%8) * 16); chunk c16 of atom a = packed pairs s[32a + 4*c16 .. + 3]
        const uint32_t p_row = p_tile + r * 128;
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
          for (int c16 = 0; c16 < 8; ++c16) {
            const uint32_t addr = p_row + a * A3_ATOM + ((c16 ^ (r & 7)) << 4);
            const uint32_t* pp = &s[32 * a + 4 * c16];
            asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(pp[0]), "r"(pp[1]), "r"(pp[2]),
                         "r"(pp[3])
                         : "memory");
          }
        if (__any_sync(0xffffffffu, rescale)) {
#pragma unroll 1
          for (int q = 0; q < 4; ++q) {
            uint32_t o[32];
            tmem_ld_32x32b_x32(tO + q * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              float a0, a1;
              f2_unpack(f2_mul(f2_pack(__uint_as_float(o[i]), __uint_as_float(o[i + 1])), f2_pack(alpha, alpha)), a0,
                        a1);
              o[i] = __float_as_uint(a0);
              o[i + 1] = __float_as_uint(a1);
            }
            tmem_st_32x32b_x32(tO + q * 32, o);
          }
        }
        tmem_st_wait();
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->p_full[x]);
        A3_TRACE(trc, c, 6);
      }
      A3_TRACE(trc, c - 1, 7);  // marks the item's last block; the next block's stamp 0 closes the item boundary

      // ---- epilogue: normalise the row, stage it in this warp's own rows of the (dead) P tile, scatter coalesced ----
      mbar_wait(&bars->pv_done[x], (c - 1) & 1);  // the item's last PV_x (and every earlier MMA of the tile) is done
      tc_fence_after();
      const float inv_l = 1.0f / l_run;
      // row rr of this warp, 16-byte piece p (0..15): atom p / 8, chunk (p % 8) ^ (rr % 8) -- the P layout, so no other
      // warp's P rows are touched and the next item's P stores of faster warps cannot collide with a slow warp's reads
      const uint32_t stage = p_tile + ew * 32 * 128;
#pragma unroll 1
      for (int q = 0; q < 4; ++q) {
        uint32_t o[32];
        tmem_ld_32x32b_x32(tO + q * 32, o);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int piece = q * 4 + i;
          const uint32_t addr = stage + (piece >> 3) * A3_ATOM + lane * 128 + (((piece & 7) ^ (lane & 7)) << 4);
          asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr),
                       "r"(pack_bf16(__uint_as_float(o[8 * i + 0]) * inv_l, __uint_as_float(o[8 * i + 1]) * inv_l)),
                       "r"(pack_bf16(__uint_as_float(o[8 * i + 2]) * inv_l, __uint_as_float(o[8 * i + 3]) * inv_l)),
                       "r"(pack_bf16(__uint_as_float(o[8 * i + 4]) * inv_l, __uint_as_float(o[8 * i + 5]) * inv_l)),
                       "r"(pack_bf16(__uint_as_float(o[8 * i + 6]) * inv_l, __uint_as_float(o[8 * i + 7]) * inv_l))
                       : "memory");
        }
      }
      // O_x is in registers / shared memory: the next item's PV_x(0) may overwrite it
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->o_free[x]);
      // training: log2-domain log-sum-exp of the scaled scores, [heads, rows_cap] (read back by the backward kernels)
      if (lse != nullptr && row_ok) lse[static_cast<int64_t>(w.h) * rows_cap + tok] = m_run + log2f(l_run);
#pragma unroll 4
      for (int it = 0; it < 16; ++it) {  // 2 rows of 256 B per iteration, 16 lanes each
        const int rr = it * 2 + (lane >> 4), piece = lane & 15;
        const int dst = __shfl_sync(0xffffffffu, my_dst, rr);
        if (dst >= 0) {
          uint4 v;
          asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                       : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                       : "r"(stage + (piece >> 3) * A3_ATOM + rr * 128 + (((piece & 7) ^ (rr & 7)) << 4)));
          *reinterpret_cast<uint4*>(out + static_cast<int64_t>(dst) * H + w.h * A3_D + piece * 8) = v;
        }
      }
      __syncwarp();  // every lane has read its staged rows before the next item's P stores
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// zeroes the tail rows a last key block may touch (0 * NaN = NaN) and this launch's work counter
__global__ void a3_prepare(__nv_bfloat16* buf, const int32_t* __restrict__ cu_seqlens, int B, int rows_cap, int row_elems,
                           unsigned int* work_counter) {
  if (blockIdx.x == 0 && threadIdx.x == 0) *work_counter = 0u;
  const int T = cu_seqlens[B];
  const int n_rows = min(rows_cap - T, A3_BK);
  const int64_t n_vec = static_cast<int64_t>(max(n_rows, 0)) * (row_elems / 8);
  uint4* p = reinterpret_cast<uint4*>(buf + static_cast<int64_t>(T) * row_elems);
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n_vec;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    p[i] = make_uint4(0, 0, 0, 0);
}

int launch_attention_tc3(const void* qkv, const int32_t* cu_seqlens, int B, int max_len_cap, int heads,
                         const int32_t* out_row_map, void* out, float scale, int rows_cap, float* lse, int causal,
                         cudaStream_t s) {
  // per-device state (one process per GPU is the normal deployment, but a process may also drive several devices)
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
    VEX_CUDA_TRY(cudaFuncSetAttribute(k4_attention_tc3, cudaFuncAttributeMaxDynamicSharedMemorySize, A3_SMEM));
    VEX_CUDA_TRY(cudaGetSymbolAddress(reinterpret_cast<void**>(&counter_base[dev]), g_a3_counters));
    sm_counts[dev] = n;
  }
  const int sm_count = sm_counts[dev];
  unsigned int* counters = counter_base[dev];
  const int H = heads * A3_D;
  CUtensorMap tm;
  std::memset(&tm, 0, sizeof(tm));
  int rc = make_tmap_2d(&tm, qkv, rows_cap, 3 * static_cast<uint64_t>(H), 3 * static_cast<uint64_t>(H), 128);
  if (rc != VEX_OK) return rc;
  // schedule knobs (tuning only): waves merged into the final heaviest-first wave, wave size in units of the SM count
  static const int tail_waves = [] {
    const char* e = std::getenv("VEX_K4_TAIL_WAVES");
    return e ? std::max(1, std::atoi(e)) : A3_TAIL_WAVES;
  }();
  static const int wave_mult = [] {
    const char* e = std::getenv("VEX_K4_WAVE_MULT");
    return e ? std::max(1, std::atoi(e)) : 1;
  }();
  const int nqp = ceil_div(max_len_cap, 2 * A3_BQ);
  const int64_t n_items = static_cast<int64_t>(nqp) * B * heads;
  if (n_items > 0x7fffffff - 65536) return VEX_E_UNSUPPORTED;
  // one counter per launch, round-robin over A3_NCOUNTERS: launches that are in flight at the same time (other streams,
  // replays of other captured graphs) own different counters; the prepare kernel ahead of the attention kernel resets it
  unsigned int* counter = counters + next_counter.fetch_add(1) % A3_NCOUNTERS;
  a3_prepare<<<32, 256, 0, s>>>(static_cast<__nv_bfloat16*>(const_cast<void*>(qkv)), cu_seqlens, B, rows_cap, 3 * H,
                                counter);
  VEX_LAUNCH_CHECK();
  const int grid = static_cast<int>(n_items < sm_count ? n_items : sm_count);
  k4_attention_tc3<<<grid, A3_THREADS, A3_SMEM, s>>>(tm, cu_seqlens, heads, B, nqp, out_row_map,
                                                     static_cast<__nv_bfloat16*>(out),
                                                     scale * 1.4426950408889634f, lse, rows_cap, causal, counter,
                                                     max(1, ceil_div(sm_count * wave_mult, nqp)), tail_waves);
  VEX_LAUNCH_CHECK();
  return VEX_OK;
}

}  // namespace vex
