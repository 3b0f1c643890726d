// K3 -- grouped (two-expert) bf16 GEMM for sm_100a: TMA -> shared memory -> tcgen05.mma -> TMEM -> fused
// epilogue.  out[rows of expert e] = A . W_e^T (+ T . Blora_e^T).
//
// What it replaces in the reference: the routed nn.Linear calls of VisionExpertAttention / VisionExpertMLP
// (modeling_cogvlm.py:244-245, :278-279, :96-97 through MLP.forward :54-56), the PEFT LoRA delta on top of them
// (scripts/cli.py:82-88, conf/lora.yaml), and -- through the epilogue modes -- rotary (:188-193), SiLU-gate
// (:55) and residual add + scatter (:321/:330 + the boolean-mask assignments).
//
// Design (one CTA per SM, persistent, 256 threads, warp-specialised):
//   warp 0 / lane 0 : TMA producer.  Per 64-wide k-block one A box (128 rows) and two B half-boxes (BN/2 rows
//                     each; for SwiGLU the halves come from gate_proj and up_proj so that accumulator columns
//                     [0,128) and [128,256) hold matching gate/up outputs -- no repacked weight copy).
//   warp 1 / lane 0 : MMA issuer, 4 x tcgen05.mma (M128 x BN x K16) per k-block, accumulators double-buffered
//                     in TMEM (2 x BN columns) so the epilogue of tile i overlaps the mainloop of tile i+1.
//                     LoRA is a K-extension: after the K loop the A operand switches to T = s*X.A^T and the B
//                     operand to lora_B through separate tensor maps (lora_B changes every optimiser step, so
//                     it cannot be folded into a static weight copy).
//   warp 2          : TMEM allocate / free.
//   warps 4..7      : epilogue.  tcgen05.ld (one row per thread) -> mode math with the reference's bf16
//                     rounding points -> swizzled per-warp staging in shared memory -> coalesced 16-byte
//                     global stores through the row map (scatter), masked by the live row count.
// Ragged M without host sync: expert row counts are read from device memory; the tile list
// (expert, m-tile, n-tile) is derived from them by every role identically; m-tiles are anchored at the
// start of each expert's segment, TMA reads that run past the segment are harmless (row-independent math)
// and the stores are masked.  Tiles are rasterised in groups of 16 m-tiles (n fastest within a group) so a
// wave of 148 CTAs shares ~16 A tiles and ~9 B tiles in L2.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "common.cuh"

namespace vex {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 256;
constexpr int GROUP_M = 16;  // smallest raster group (m-tiles): 74 CTA pairs cover 8 x ~9 tiles, the minimal working set

// Raster group height.  Tiles are walked group by group (group_m m-tiles x all n-tiles, m fastest): the A rows of a group
// (group_m * 128 * K * 2 bytes) stay in L2 while the weights stream past once per group, so the weights of an expert
// are read from DRAM ceil(m_tiles / group_m) times.  With the minimal group the c2 SwiGLU launch read 1.62 GB for
// 0.72 GB of algorithmic traffic (ncu).  Take the tallest group whose A rows fit in a third of the 126 MB L2.
__host__ inline int raster_group_for(int K) {
  const long long per_tile = 128LL * K * 2;
  long long g = (40LL << 20) / per_tile;
  g = g < GROUP_M ? GROUP_M : (g > 128 ? 128 : g);
  return static_cast<int>(g & ~1LL);
}

struct alignas(64) GemmTmaps {
  CUtensorMap a;         // activations [rows_cap, K]
  CUtensorMap w[2][2];   // [expert][half] weights [N, K]
  CUtensorMap t[2];      // [half] LoRA T = s * X.A^T [rows_cap, r]
  CUtensorMap lb[2][2];  // [expert][half] lora_B [N, r]
  CUtensorMap lb64[2][2];  // same tensors with 64-row boxes (CTA-pair kernel, SwiGLU LoRA steps)
};

struct GemmDev {
  const int32_t* counts;
  const int32_t* row_map;
  __nv_bfloat16* out;
  const __nv_bfloat16* residual;
  const __nv_bfloat16* rope_cos;
  const __nv_bfloat16* rope_sin;
  const int64_t* position_ids;
  const int32_t* sorted_to_flat;
  int64_t ldo;
  float alpha;
  int rope_len, rope_cols;
  int rows_cap, N, K, mode, single_expert;
  int lora_steps;  // 64-wide k-blocks along r (0 = no LoRA)
  int lora_mask;   // bit e: expert e has an adapter
  int trans_b;     // 1: weights are [K, N] row-major (MN-major B operand): out = A . W  (backward dgrad)
  uint32_t drop_thresh16, drop_seed_lo, drop_seed_hi;  // VEX_EPI_DROPOUT_ACC: mask of dropout_hash(s_row * N + col)
  // VEX_EPI_CE / VEX_EPI_CE_BWD (fused lm_head + cross-entropy): per-row label, per-(row, n-tile) softmax partials
  const int32_t* ce_labels;
  float* ce_pmax;
  float* ce_psum;
  float* ce_zlabel;
  const float* ce_lse;
  const float* ce_w;
  const float* ce_dloss;
  int ce_tiles;
  // vision-encoder Linears (visual.py:84-85, :110-111: nn.Linear WITH bias; MLP activation ACT2FN['gelu'])
  const __nv_bfloat16* bias;  // [N] or nullptr: added to the fp32 accumulator before the bf16 rounding
  int act;                    // VEX_ACT_NONE / VEX_ACT_GELU (exact erf form, on the bf16-rounded Linear output)
  int raster_group;           // m-tiles (128 rows) per raster group, even; see raster_group_for()
  // VEX_EPI_ROPE second output: post-rotary K and V written straight into the KV cache
  // [B, heads, kv_cap, 128] at position l (+ *kv_pos) of sample b, (b, l) = divmod(sorted_to_flat[row], kv_seq)
  __nv_bfloat16* kv_k;
  __nv_bfloat16* kv_v;
  const int32_t* kv_pos;
  int kv_seq, kv_cap;
};

// Row of the KV cache ([B * heads * kv_cap] rows of 128) that sorted row s_row's head slot goes to, for the 128-column
// panel starting at column colp of the QKV output; -1 = this panel is a query head / the row is dead / out of capacity.
// `which` (warp-uniform) receives the cache base pointer (nullptr: nothing to write).
__device__ __forceinline__ int kv_cache_row(const GemmDev& p, int colp, int s_row, bool valid, __nv_bfloat16*& which) {
  which = nullptr;
  if (p.kv_k == nullptr) return -1;
  const int Hh = p.rope_cols >> 1;  // hidden size: q and k columns are the rotated ones
  if (colp < Hh || colp >= 3 * Hh) return -1;
  const bool is_v = colp >= 2 * Hh;
  which = is_v ? p.kv_v : p.kv_k;
  if (!valid) return -1;
  const int head = (colp - (is_v ? 2 * Hh : Hh)) >> 7;
  const int flat = p.sorted_to_flat[s_row];
  const int b = flat / p.kv_seq;
  const int l = flat - b * p.kv_seq + (p.kv_pos ? p.kv_pos[0] : 0);
  if (l < 0 || l >= p.kv_cap) return -1;
  return (b * (Hh >> 7) + head) * p.kv_cap + l;
}

template <int BN>
struct GemmCfg {
  static constexpr int STAGES = BN == 256 ? 4 : 8;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int HALF_B = B_BYTES / 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int PANEL = BN >= 128 ? 128 : BN;  // epilogue panel width (columns)
  static constexpr int PANEL_BYTES = PANEL * 2;       // staging row pitch
  static constexpr int EPI_WARP_BYTES = 32 * PANEL_BYTES;
  static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
  static constexpr int NUM_BARS = 2 * STAGES + 4;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 4 * EPI_WARP_BYTES + NUM_BARS * 8 + 16 + 1024;
};

// B-operand descriptor of one pipeline stage.  MN-major tiles are stored as 64-column chunks of 8 KB
// (64 k-rows x 128 B): LBO = 8192 (next chunk of 64 n), SBO = 1024 (next group of 8 k-rows).
constexpr int MN_CHUNK_BYTES = 64 * BK * 2;
template <bool TB>
__device__ __forceinline__ uint64_t b_desc(uint32_t smem_addr) {
  if constexpr (TB) return umma_desc_mnmajor_sw128(smem_addr, MN_CHUNK_BYTES, 1024);
  else return umma_desc_kmajor_sw128(smem_addr);
}
template <bool TB>
constexpr uint64_t B_KSTEP = TB ? (16 * 128) >> 4 : 2;  // descriptor start-address increment per K = 16 step

struct TileCoord {
  int e, m, n;
};

__device__ __forceinline__ TileCoord decode_tile(int tile, int mt0, int mt1, int n_tiles, int group_m) {
  TileCoord c;
  int mt_e = mt0;
  c.e = 0;
  const int t0 = mt0 * n_tiles;
  if (tile >= t0) {
    c.e = 1;
    tile -= t0;
    mt_e = mt1;
  }
  const int per_group = group_m * n_tiles;
  const int g = tile / per_group;
  const int r = tile - g * per_group;
  const int m_lo = g * group_m;
  const int gm = min(group_m, mt_e - m_lo);
  c.n = r / gm;
  c.m = m_lo + (r - c.n * gm);
  return c;
}

__device__ __forceinline__ float silu_acc(float x) { return x / (1.0f + __expf(-x)); }

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
  return r;
}

// 32 fp32 accumulator values of this thread's row -> bf16 -> swizzled staging (4 x 16 B pieces)
__device__ __forceinline__ void stage_piece32(uint32_t stage_row_addr, int first_piece, int lane,
                                              const float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int piece = first_piece + i;
    const uint32_t addr = stage_row_addr + static_cast<uint32_t>((piece ^ (lane & 7)) << 4);
    st_shared_v4(addr, pack_bf16(v[8 * i + 0], v[8 * i + 1]), pack_bf16(v[8 * i + 2], v[8 * i + 3]),
                 pack_bf16(v[8 * i + 4], v[8 * i + 5]), pack_bf16(v[8 * i + 6], v[8 * i + 7]));
  }
}

__device__ __forceinline__ void load_bf16x32(const __nv_bfloat16* p, float (&o)[32]) {
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

// ---------------------------------------------------------------------------------------------
// epilogue of one 128 x BN accumulator tile, executed by one of the four epilogue warps (32 rows each)
//   t_acc      : TMEM address of this warp's lane quarter at the accumulator's first column
//   release()  : called once, right after the last TMEM read of the tile (hands the accumulator back)
// ---------------------------------------------------------------------------------------------
template <int BN, class Release>
__device__ __forceinline__ void epilogue_tile(const GemmDev& p, int e, int m, int n, int cnt0, int cnt1,
                                              uint32_t t_acc, uint32_t stage_base, int lane, int ew,
                                              Release release) {
  using Cfg = GemmCfg<BN>;
  constexpr int PANEL = Cfg::PANEL;
  constexpr int PITCH = Cfg::PANEL_BYTES;
  constexpr int LPR = PITCH / 16;  // lanes per staged row during write-out
  constexpr int RPI = 32 / LPR;    // rows per write-out iteration
  const uint32_t stage_row = stage_base + static_cast<uint32_t>(lane * PITCH);

  // coalesced write-out of one staged panel: out[g, col0 + ...] for the 32 rows of this warp
  auto write_panel = [&](int g_row, int col0, bool add_residual, __nv_bfloat16* kv_base = nullptr, int kv_row = -1) {
#pragma unroll 4
    for (int it = 0; it < 32 / RPI; ++it) {
      const int rr = it * RPI + lane / LPR;
      const int piece = lane % LPR;
      const int g = __shfl_sync(0xffffffffu, g_row, rr);
      const int kr = kv_base ? __shfl_sync(0xffffffffu, kv_row, rr) : -1;  // kv_base is warp-uniform
      const int col = col0 + piece * 8;
      if (g >= 0 && col < p.N) {
        uint4 v = ld_shared_v4(stage_base + static_cast<uint32_t>(rr * PITCH + ((piece ^ (rr & 7)) << 4)));
        if (kr >= 0) *reinterpret_cast<uint4*>(kv_base + static_cast<int64_t>(kr) * 128 + piece * 8) = v;
        const int64_t off = static_cast<int64_t>(g) * p.ldo + col;
        if (add_residual) {
          const uint4 r4 = ld_stream(p.residual + off);
          const uint32_t a[4] = {v.x, v.y, v.z, v.w}, b[4] = {r4.x, r4.y, r4.z, r4.w};
          uint32_t* vp = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            vp[j] = pack_bf16(bf16_lo(a[j]) + bf16_lo(b[j]), bf16_hi(a[j]) + bf16_hi(b[j]));
        }
        *reinterpret_cast<uint4*>(p.out + off) = v;
      }
    }
  };

  const int r_local = m * BM + ew * 32 + lane;
  const bool valid = r_local < (e ? cnt1 : cnt0);
  const int s_row = (e ? cnt0 : 0) + r_local;
  int g_row = -1;
  int pos = 0;
  if (valid) {
    g_row = p.row_map ? p.row_map[s_row] : s_row;
    if (p.mode == VEX_EPI_ROPE) {
      const int64_t pz = p.position_ids[p.sorted_to_flat[s_row]];
      pos = static_cast<int>(min(max(pz, int64_t(0)), int64_t(p.rope_len - 1)));
    }
  }
  uint32_t raw[32];
  float v[32];

  if (p.mode == VEX_EPI_CE) {
    // Fused lm_head + cross-entropy, forward: the logits of this 128 x BN tile never leave the SM.  Per row: running
    // max m and sum of exp(z - m) over the tile's columns (z = bf16-rounded logit, what `lm_head(h).float()` holds,
    // modeling_cogvlm.py:701) and the label's logit if it falls into the tile.  vex_ce_reduce combines the tiles.
    constexpr float L2E = 1.4426950408889634f;
    const int label = valid ? p.ce_labels[s_row] : -1;
    float m = -INFINITY, ssum = 0.f, zl = 0.f;
    bool has = false;
#pragma unroll 1
    for (int q = 0; q < BN / 32; ++q) {
      const int cbase = n * BN + q * 32;
      tmem_ld_32x32b_x32(t_acc + q * 32, raw);
      tmem_ld_wait();
      float cm = -INFINITY;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        v[j] = (cbase + j < p.N) ? bf16r(__uint_as_float(raw[j])) : -INFINITY;
        cm = fmaxf(cm, v[j]);
        zl += (cbase + j == label) ? v[j] : 0.f;  // select chain (no dynamic register indexing)
      }
      if (cm > m) {  // never true while cm == -inf (columns past N), so m - cm is never inf - inf
        ssum *= exp2f((m - cm) * L2E);
        m = cm;
      }
      if (m > -INFINITY) {
        const float ml = m * L2E;
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          a0 += exp2f(fmaf(v[j], L2E, -ml));
          a1 += exp2f(fmaf(v[j + 1], L2E, -ml));
        }
        ssum += a0 + a1;
      }
      has = has || (label >= cbase && label < cbase + 32);
    }
    release();
    if (valid) {
      const int64_t slot = static_cast<int64_t>(s_row) * p.ce_tiles + 2 * n;
      p.ce_pmax[slot] = m;
      p.ce_psum[slot] = ssum;
      p.ce_pmax[slot + 1] = -INFINITY;  // the half-tile slot the CTA-pair kernel would fill: contributes exp(-inf) * 0
      p.ce_psum[slot + 1] = 0.f;
      if (has) p.ce_zlabel[s_row] = zl;
    }
    __syncwarp();
    return;
  }
  // VEX_EPI_CE_BWD: per-row constants of d(loss)/d(logit) = (softmax - onehot) * w_row * dloss / n_rows
  int ce_label = -1;
  float ce_lse2 = 0.f, ce_coef = 0.f;
  if (p.mode == VEX_EPI_CE_BWD && valid) {
    ce_label = p.ce_labels[s_row];
    ce_lse2 = p.ce_lse[s_row] * 1.4426950408889634f;
    ce_coef = p.ce_w[s_row] * p.ce_dloss[0] / static_cast<float>(max(cnt0, 1));
  }

  if (p.mode == VEX_EPI_SWIGLU) {
    if constexpr (BN == 256) {
#pragma unroll 1
      for (int q = 0; q < 4; ++q) {
        uint32_t raw_u[32];
        tmem_ld_32x32b_x32(t_acc + q * 32, raw);
        tmem_ld_32x32b_x32(t_acc + 128 + q * 32, raw_u);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float g = bf16r(__uint_as_float(raw[j]));    // gate_proj output, bf16 like the eager Linear
          const float u = bf16r(__uint_as_float(raw_u[j]));  // up_proj output
          v[j] = bf16r(silu_acc(g)) * u;                     // silu -> bf16, product -> bf16 (on store)
        }
        stage_piece32(stage_row, q * 4, lane, v);
      }
      release();
      write_panel(g_row, n * 128, false);
      __syncwarp();
    }
  } else {
    constexpr int NPANEL = BN / PANEL;
#pragma unroll 1
    for (int pn = 0; pn < NPANEL; ++pn) {
      const int col0 = n * BN + pn * PANEL;
      if (p.mode == VEX_EPI_ROPE && col0 < p.rope_cols) {
        if constexpr (PANEL == 128) {
          // one head: pairs (j, j + 64); thread owns the whole row so both halves are local
#pragma unroll 1
          for (int q = 0; q < 2; ++q) {
            uint32_t raw_hi[32];
            tmem_ld_32x32b_x32(t_acc + pn * PANEL + q * 32, raw);
            tmem_ld_32x32b_x32(t_acc + pn * PANEL + 64 + q * 32, raw_hi);
            const __nv_bfloat16* cr = p.rope_cos + static_cast<int64_t>(pos) * 128 + q * 32;
            const __nv_bfloat16* sr = p.rope_sin + static_cast<int64_t>(pos) * 128 + q * 32;
            uint4 ct[4], st[4];  // packed bf16 table entries for columns q*32 .. q*32+31
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              ct[i] = __ldg(reinterpret_cast<const uint4*>(cr) + i);
              st[i] = __ldg(reinterpret_cast<const uint4*>(sr) + i);
            }
            tmem_ld_wait();
            // x1 = q[j], x2 = q[j + 64] as the eager bf16 Linear would have produced them
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              raw[j] = __float_as_uint(bf16r(__uint_as_float(raw[j])));
              raw_hi[j] = __float_as_uint(bf16r(__uint_as_float(raw_hi[j])));
            }
            // first half: q*cos + rotate_half(q)*sin, rotate_half(q)[j] = -q[j + 64]; every product and the
            // sum are rounded to bf16 like the eager ops (the sum is rounded by the bf16 pack)
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const uint32_t cw = reinterpret_cast<const uint32_t*>(ct)[j >> 1];
              const uint32_t sw = reinterpret_cast<const uint32_t*>(st)[j >> 1];
              const float cj = (j & 1) ? bf16_hi(cw) : bf16_lo(cw);
              const float sj = (j & 1) ? bf16_hi(sw) : bf16_lo(sw);
              v[j] = bf16r(__uint_as_float(raw[j]) * cj) - bf16r(__uint_as_float(raw_hi[j]) * sj);
            }
            stage_piece32(stage_row, q * 4, lane, v);
            // second half uses the table entries of column j + 64 (equal to column j for the reference's
            // cat(freqs, freqs) table, but read them anyway)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              ct[i] = __ldg(reinterpret_cast<const uint4*>(cr + 64) + i);
              st[i] = __ldg(reinterpret_cast<const uint4*>(sr + 64) + i);
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const uint32_t cw = reinterpret_cast<const uint32_t*>(ct)[j >> 1];
              const uint32_t sw = reinterpret_cast<const uint32_t*>(st)[j >> 1];
              const float cj = (j & 1) ? bf16_hi(cw) : bf16_lo(cw);
              const float sj = (j & 1) ? bf16_hi(sw) : bf16_lo(sw);
              v[j] = bf16r(__uint_as_float(raw_hi[j]) * cj) + bf16r(__uint_as_float(raw[j]) * sj);
            }
            stage_piece32(stage_row, 8 + q * 4, lane, v);
          }
        }
      } else {
#pragma unroll 1
        for (int q = 0; q < PANEL / 32; ++q) {
          tmem_ld_32x32b_x32(t_acc + pn * PANEL + q * 32, raw);
          tmem_ld_wait();
          if (p.mode == VEX_EPI_PLAIN) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]) * p.alpha;
          } else if (p.mode == VEX_EPI_CE_BWD) {
            // dz = (exp(z - lse) - [col == label]) * coef, z = bf16-rounded logit; rounded to bf16 on store like the
            // eager `.float()` backward hands it to lm_head's backward
            const int cbase = col0 + q * 32;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float z = bf16r(__uint_as_float(raw[j]));
              const float sm = exp2f(fmaf(z, 1.4426950408889634f, -ce_lse2));
              v[j] = (sm - (cbase + j == ce_label ? 1.f : 0.f)) * ce_coef;
            }
          } else if (p.mode == VEX_EPI_DROPOUT_ACC) {
            // adjoint of the LoRA input dropout: keep(s_row, col) * alpha * acc, same hash as k7_dropout_rows
            const uint64_t pair0 = (static_cast<uint64_t>(s_row) * p.N + (col0 + q * 32)) >> 1;
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const uint32_t hsh = dropout_hash(pair0 + (j >> 1), p.drop_seed_lo, p.drop_seed_hi);
              v[j] = (hsh & 0xffffu) >= p.drop_thresh16 ? __uint_as_float(raw[j]) * p.alpha : 0.f;
              v[j + 1] = (hsh >> 16) >= p.drop_thresh16 ? __uint_as_float(raw[j + 1]) * p.alpha : 0.f;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
          }
          if (p.bias != nullptr && col0 + q * 32 < p.N) {
            // nn.Linear with bias: the bias joins the fp32 accumulator, one rounding (cuBLASLt bias epilogue);
            // every thread of the warp reads the same 32 bias values (broadcast)
            float bv[32];
            load_bf16x32(p.bias + col0 + q * 32, bv);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += bv[j];
          }
          if (p.act == VEX_ACT_GELU) {
            // eager bf16: y = bf16(Linear), then gelu(y) = y Phi(y) (erf form) rounded on store
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              v[j] = gelu_erf(bf16r(v[j]));
            }
          }
          stage_piece32(stage_row, q * 4, lane, v);
        }
      }
      if (pn == NPANEL - 1) {  // every TMEM read of this accumulator is done: hand it back
        release();
      } else {
        __syncwarp();
      }
      if (p.mode == VEX_EPI_ROPE && PANEL == 128) {
        __nv_bfloat16* kv_base;
        const int kv_row = kv_cache_row(p, col0, s_row, valid, kv_base);
        write_panel(g_row, col0, false, kv_base, kv_row);
      } else {
        write_panel(g_row, col0, p.mode == VEX_EPI_RESIDUAL || p.mode == VEX_EPI_DROPOUT_ACC);
      }
      __syncwarp();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// CTA-pair kernel epilogue: EIGHT epilogue warps, two per SM sub-partition.  Warps w and w + 4 share TMEM lane
// quarter w % 4 and split the 256 accumulator columns (half = 0 / 1 -> columns [128 half, 128 half + 128)); each warp
// stages 64 output columns at a time (32 rows x 128 B = 4 KB per warp, the same 32 KB in total as the 4-warp layout).
// Why: with one warp per sub-partition the epilogue issues at IPC ~0.3 (nothing to switch to while a TMEM load, a
// MUFU result or a bias load is outstanding); for short-K GEMMs (the vision encoder's K = 1792: 14 k cycles of
// mainloop per tile) that made the EPILOGUE the bound -- tensor pipe 62 % active, profiles/r1_vision_fc1_gelu_v1.md.
// ---------------------------------------------------------------------------------------------
struct CeRow {
  int label;
  float lse2, coef;
};

// one 32-column chunk of the PLAIN / RESIDUAL / DROPOUT_ACC / CE_BWD epilogues (+ bias, + GELU): raw fp32 accumulator
// bits -> v (fp32 values that the bf16 pack rounds)
__device__ __forceinline__ void chunk_math(const GemmDev& p, const uint32_t (&raw)[32], float (&v)[32], int cbase,
                                           int s_row, const CeRow& ce) {
  if (p.mode == VEX_EPI_PLAIN) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]) * p.alpha;
  } else if (p.mode == VEX_EPI_CE_BWD) {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float z = bf16r(__uint_as_float(raw[j]));
      const float sm = exp2f(fmaf(z, 1.4426950408889634f, -ce.lse2));
      v[j] = (sm - (cbase + j == ce.label ? 1.f : 0.f)) * ce.coef;
    }
  } else if (p.mode == VEX_EPI_DROPOUT_ACC) {
    const uint64_t pair0 = (static_cast<uint64_t>(s_row) * p.N + cbase) >> 1;
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      const uint32_t hsh = dropout_hash(pair0 + (j >> 1), p.drop_seed_lo, p.drop_seed_hi);
      v[j] = (hsh & 0xffffu) >= p.drop_thresh16 ? __uint_as_float(raw[j]) * p.alpha : 0.f;
      v[j + 1] = (hsh >> 16) >= p.drop_thresh16 ? __uint_as_float(raw[j + 1]) * p.alpha : 0.f;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
  }
  if (p.bias != nullptr && cbase < p.N) {
    float bv[32];
    load_bf16x32(p.bias + cbase, bv);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] += bv[j];
  }
  if (p.act == VEX_ACT_GELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = gelu_erf(bf16r(v[j]));
  }
}

// rotary on 32 (j, j + 64) column pairs of one head: lo / hi = bf16-rounded q[j], q[j + 64]; first = columns j,
// second = columns j + 64 (apply_rotary_pos_emb_index_bhs, modeling_cogvlm.py:188-193, eager-bf16 rounding points)
__device__ __forceinline__ void rope_math(const GemmDev& p, int pos, int q, uint32_t (&lo)[32], uint32_t (&hi)[32],
                                          float (&first)[32], float (&second)[32]) {
  const __nv_bfloat16* cr = p.rope_cos + static_cast<int64_t>(pos) * 128 + q * 32;
  const __nv_bfloat16* sr = p.rope_sin + static_cast<int64_t>(pos) * 128 + q * 32;
  uint4 ct[4], st[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    ct[i] = __ldg(reinterpret_cast<const uint4*>(cr) + i);
    st[i] = __ldg(reinterpret_cast<const uint4*>(sr) + i);
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    lo[j] = __float_as_uint(bf16r(__uint_as_float(lo[j])));
    hi[j] = __float_as_uint(bf16r(__uint_as_float(hi[j])));
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const uint32_t cw = reinterpret_cast<const uint32_t*>(ct)[j >> 1];
    const uint32_t sw = reinterpret_cast<const uint32_t*>(st)[j >> 1];
    const float cj = (j & 1) ? bf16_hi(cw) : bf16_lo(cw);
    const float sj = (j & 1) ? bf16_hi(sw) : bf16_lo(sw);
    first[j] = bf16r(__uint_as_float(lo[j]) * cj) - bf16r(__uint_as_float(hi[j]) * sj);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    ct[i] = __ldg(reinterpret_cast<const uint4*>(cr + 64) + i);
    st[i] = __ldg(reinterpret_cast<const uint4*>(sr + 64) + i);
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const uint32_t cw = reinterpret_cast<const uint32_t*>(ct)[j >> 1];
    const uint32_t sw = reinterpret_cast<const uint32_t*>(st)[j >> 1];
    const float cj = (j & 1) ? bf16_hi(cw) : bf16_lo(cw);
    const float sj = (j & 1) ? bf16_hi(sw) : bf16_lo(sw);
    second[j] = bf16r(__uint_as_float(hi[j]) * cj) + bf16r(__uint_as_float(lo[j]) * sj);
  }
}

//   t_acc      : TMEM address of this warp's lane quarter at the accumulator's first column
//   stage_base : this warp's 4 KB staging area (32 rows x 128 B)
//   half       : which 128 accumulator columns this warp owns
//   release()  : called exactly once by every warp, after its last TMEM read of the tile
template <class Release>
__device__ __forceinline__ void epilogue_tile_half(const GemmDev& p, int e, int m, int n, int cnt0, int cnt1,
                                                   uint32_t t_acc, uint32_t stage_base, int lane, int ew, int half,
                                                   Release release) {
  constexpr int BN = 256;
  constexpr int PITCH = 128;  // bytes per staged row: 64 bf16 columns
  const uint32_t stage_row = stage_base + static_cast<uint32_t>(lane * PITCH);

  // coalesced write-out of one staged 64-column piece: staged columns [0, 32) go to global columns [colA, colA + 32),
  // staged columns [32, 64) to [colB, colB + 32) (colB = colA + 32 except for the rotary halves)
  const int panel_col0 = n * BN + half * 128;  // first output column of this warp's 128-column panel
  __nv_bfloat16* kv_base = nullptr;            // VEX_EPI_ROPE: K / V panels are mirrored into the KV cache
  int kv_row = -1;
  auto write_piece = [&](int g_row, int colA, int colB, bool add_residual) {
#pragma unroll 4
    for (int it = 0; it < 8; ++it) {
      const int rr = it * 4 + (lane >> 3);
      const int piece = lane & 7;
      const int g = __shfl_sync(0xffffffffu, g_row, rr);
      const int kr = kv_base ? __shfl_sync(0xffffffffu, kv_row, rr) : -1;  // kv_base is warp-uniform
      const int col = piece < 4 ? colA + piece * 8 : colB + (piece - 4) * 8;
      if (g >= 0 && col < p.N) {
        uint4 v = ld_shared_v4(stage_base + static_cast<uint32_t>(rr * PITCH + ((piece ^ (rr & 7)) << 4)));
        if (kr >= 0) *reinterpret_cast<uint4*>(kv_base + static_cast<int64_t>(kr) * 128 + (col - panel_col0)) = v;
        const int64_t off = static_cast<int64_t>(g) * p.ldo + col;
        if (add_residual) {
          const uint4 r4 = ld_stream(p.residual + off);
          const uint32_t a[4] = {v.x, v.y, v.z, v.w}, b[4] = {r4.x, r4.y, r4.z, r4.w};
          uint32_t* vp = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            vp[j] = pack_bf16(bf16_lo(a[j]) + bf16_lo(b[j]), bf16_hi(a[j]) + bf16_hi(b[j]));
        }
        *reinterpret_cast<uint4*>(p.out + off) = v;
      }
    }
  };

  const int r_local = m * BM + ew * 32 + lane;
  const bool valid = r_local < (e ? cnt1 : cnt0);
  const int s_row = (e ? cnt0 : 0) + r_local;
  int g_row = -1;
  int pos = 0;
  if (valid) {
    g_row = p.row_map ? p.row_map[s_row] : s_row;
    if (p.mode == VEX_EPI_ROPE) {
      const int64_t pz = p.position_ids[p.sorted_to_flat[s_row]];
      pos = static_cast<int>(min(max(pz, int64_t(0)), int64_t(p.rope_len - 1)));
    }
  }
  if (p.mode == VEX_EPI_ROPE) kv_row = kv_cache_row(p, panel_col0, s_row, valid, kv_base);
  uint32_t raw[32];
  float v[32];

  if (p.mode == VEX_EPI_CE) {
    // per-row softmax statistics over this warp's 128 columns; the two halves of a tile write separate partial slots
    // (ce_tiles counts 128-column half tiles for the pair kernel)
    constexpr float L2E = 1.4426950408889634f;
    const int label = valid ? p.ce_labels[s_row] : -1;
    float mx = -INFINITY, ssum = 0.f, zl = 0.f;
    bool has = false;
#pragma unroll 1
    for (int q = 0; q < 4; ++q) {
      const int cbase = n * BN + half * 128 + q * 32;
      tmem_ld_32x32b_x32(t_acc + half * 128 + q * 32, raw);
      tmem_ld_wait();
      float cm = -INFINITY;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        v[j] = (cbase + j < p.N) ? bf16r(__uint_as_float(raw[j])) : -INFINITY;
        cm = fmaxf(cm, v[j]);
        zl += (cbase + j == label) ? v[j] : 0.f;
      }
      if (cm > mx) {
        ssum *= exp2f((mx - cm) * L2E);
        mx = cm;
      }
      if (mx > -INFINITY) {
        const float ml = mx * L2E;
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          a0 += exp2f(fmaf(v[j], L2E, -ml));
          a1 += exp2f(fmaf(v[j + 1], L2E, -ml));
        }
        ssum += a0 + a1;
      }
      has = has || (label >= cbase && label < cbase + 32);
    }
    release();
    if (valid) {
      const int64_t slot = static_cast<int64_t>(s_row) * p.ce_tiles + 2 * n + half;
      p.ce_pmax[slot] = mx;
      p.ce_psum[slot] = ssum;
      if (has) p.ce_zlabel[s_row] = zl;
    }
    __syncwarp();
    return;
  }
  CeRow ce{-1, 0.f, 0.f};
  if (p.mode == VEX_EPI_CE_BWD && valid) {
    ce.label = p.ce_labels[s_row];
    ce.lse2 = p.ce_lse[s_row] * 1.4426950408889634f;
    ce.coef = p.ce_w[s_row] * p.ce_dloss[0] / static_cast<float>(max(cnt0, 1));
  }

  if (p.mode == VEX_EPI_SWIGLU) {
    // output columns [64 half, 64 half + 64) of the tile's 128: gate accumulator columns q*32.., up columns 128 + q*32..
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      const int q = 2 * half + c;
      uint32_t raw_u[32];
      tmem_ld_32x32b_x32(t_acc + q * 32, raw);
      tmem_ld_32x32b_x32(t_acc + 128 + q * 32, raw_u);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float g = bf16r(__uint_as_float(raw[j]));
        const float u = bf16r(__uint_as_float(raw_u[j]));
        v[j] = bf16r(silu_acc(g)) * u;
      }
      stage_piece32(stage_row, c * 4, lane, v);
    }
    release();
    write_piece(g_row, n * 128 + half * 64, n * 128 + half * 64 + 32, false);
    __syncwarp();
    return;
  }

  const int col0 = n * BN + half * 128;
  const uint32_t t_panel = t_acc + half * 128;
  if (p.mode == VEX_EPI_ROPE && col0 < p.rope_cols) {
    // this warp's panel is one head: pairs (j, j + 64)
#pragma unroll 1
    for (int q = 0; q < 2; ++q) {
      uint32_t raw_hi[32];
      float v2[32];
      tmem_ld_32x32b_x32(t_panel + q * 32, raw);
      tmem_ld_32x32b_x32(t_panel + 64 + q * 32, raw_hi);
      tmem_ld_wait();
      rope_math(p, pos, q, raw, raw_hi, v, v2);
      stage_piece32(stage_row, 0, lane, v);
      stage_piece32(stage_row, 4, lane, v2);
      if (q == 1) release(); else __syncwarp();
      write_piece(g_row, col0 + q * 32, col0 + 64 + q * 32, false);
      __syncwarp();
    }
    return;
  }
#pragma unroll 1
  for (int pc = 0; pc < 2; ++pc) {
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      const int qq = pc * 2 + c;
      tmem_ld_32x32b_x32(t_panel + qq * 32, raw);
      tmem_ld_wait();
      chunk_math(p, raw, v, col0 + qq * 32, s_row, ce);
      stage_piece32(stage_row, c * 4, lane, v);
    }
    if (pc == 1) release(); else __syncwarp();
    write_piece(g_row, col0 + pc * 64, col0 + pc * 64 + 32,
                p.mode == VEX_EPI_RESIDUAL || p.mode == VEX_EPI_DROPOUT_ACC);
    __syncwarp();
  }
}

// TB = true: the B operand is MN-major -- the weight tensor is [K, N] row-major (N contiguous), loaded as
// BN/64 boxes of {64 n, 64 k} (8 KB each, 128B-swizzled k-rows); this is the dgrad form dX = dY . W, which
// reads the nn.Linear weight [out, in] exactly as stored (K = out, N = in) -- no transposed copy.
template <int BN, bool TB>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
    k3_grouped_gemm(const __grid_constant__ GemmTmaps tm, const GemmDev p) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operands need 1024-byte aligned tiles
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi_smem = smem + STAGES * Cfg::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + 4 * Cfg::EPI_WARP_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(bars + Cfg::NUM_BARS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);  // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_base_smem, Cfg::TMEM_COLS);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm.a);
    tma_prefetch_desc(&tm.w[0][0]);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;

  // ---- tile list, derived identically by every role from the device-side counts ----
  const int cnt0 = min(max(p.counts[0], 0), p.rows_cap);
  const int cnt1 = p.single_expert ? 0 : min(max(p.counts[1], 0), p.rows_cap - cnt0);
  const int mt0 = (cnt0 + BM - 1) / BM;
  const int mt1 = (cnt1 + BM - 1) / BM;
  const bool swiglu = p.mode == VEX_EPI_SWIGLU;
  const int n_span = swiglu ? BN / 2 : BN;  // output columns per tile
  const int n_tiles = (p.N + n_span - 1) / n_span;
  const int total_tiles = (mt0 + mt1) * n_tiles;
  const int num_kb = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    // =============================== TMA producer ===============================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const TileCoord c = decode_tile(tile, mt0, mt1, n_tiles, p.raster_group);
      const int row0 = (c.e ? cnt0 : 0) + c.m * BM;
      const int nrow0 = c.n * n_span;
      const int nrow1 = swiglu ? nrow0 : nrow0 + BN / 2;
      const CUtensorMap* w0 = &tm.w[c.e][0];
      const CUtensorMap* w1 = swiglu ? &tm.w[c.e][1] : w0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
        uint8_t* sb = sa + Cfg::A_BYTES;
        mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
        tma_load_2d(sa, &tm.a, &full_bar[stage], kb * BK, row0);
        if constexpr (TB) {
#pragma unroll
          for (int c = 0; c < BN / 64; ++c)
            tma_load_2d(sb + c * MN_CHUNK_BYTES, w0, &full_bar[stage], nrow0 + 64 * c, kb * BK);
        } else {
          tma_load_2d(sb, w0, &full_bar[stage], kb * BK, nrow0);
          tma_load_2d(sb + Cfg::HALF_B, w1, &full_bar[stage], kb * BK, nrow1);
        }
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      if ((p.lora_mask >> c.e) & 1) {
        const int halves = swiglu ? 2 : 1;
        for (int h = 0; h < halves; ++h) {
          for (int ls = 0; ls < p.lora_steps; ++ls) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
            uint8_t* sb = sa + Cfg::A_BYTES;
            if (swiglu) {
              mbar_arrive_expect_tx(&full_bar[stage], Cfg::A_BYTES + Cfg::HALF_B);
              tma_load_2d(sa, &tm.t[h], &full_bar[stage], ls * BK, row0);
              tma_load_2d(sb, &tm.lb[c.e][h], &full_bar[stage], ls * BK, nrow0);
            } else {
              mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
              tma_load_2d(sa, &tm.t[0], &full_bar[stage], ls * BK, row0);
              if constexpr (TB) {
#pragma unroll
                for (int cc = 0; cc < BN / 64; ++cc)
                  tma_load_2d(sb + cc * MN_CHUNK_BYTES, &tm.lb[c.e][0], &full_bar[stage], nrow0 + 64 * cc, ls * BK);
              } else {
                tma_load_2d(sb, &tm.lb[c.e][0], &full_bar[stage], ls * BK, nrow0);
                tma_load_2d(sb + Cfg::HALF_B, &tm.lb[c.e][0], &full_bar[stage], ls * BK, nrow1);
              }
            }
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // =============================== MMA issuer ===============================
    constexpr uint32_t idesc_full = umma_idesc_bf16(BM, BN, 0, TB ? 1 : 0);
    constexpr uint32_t idesc_half = umma_idesc_bf16(BM, BN / 2, 0, TB ? 1 : 0);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const TileCoord c = decode_tile(tile, mt0, mt1, n_tiles, p.raster_group);
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
      uint32_t accumulate = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
        const uint64_t da = umma_desc_kmajor_sw128(sa);
        const uint64_t db = b_desc<TB>(sa + Cfg::A_BYTES);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          // K-major: +32 bytes per K=16 step inside the 128-byte swizzle row (+2 in the addr >> 4 field);
          // MN-major: 16 k-rows of 128 B = +2048 bytes (+128)
          umma_ss(d_tmem, da + 2 * k, db + B_KSTEP<TB> * k, idesc_full, accumulate);
          accumulate = 1;
        }
        umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      if ((p.lora_mask >> c.e) & 1) {
        const int halves = swiglu ? 2 : 1;
        for (int h = 0; h < halves; ++h) {
          for (int ls = 0; ls < p.lora_steps; ++ls) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
            const uint64_t da = umma_desc_kmajor_sw128(sa);
            const uint64_t db = b_desc<TB>(sa + Cfg::A_BYTES);
            const uint32_t d_half = d_tmem + static_cast<uint32_t>(swiglu ? h * (BN / 2) : 0);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              umma_ss(d_half, da + 2 * k, db + B_KSTEP<TB> * k, swiglu ? idesc_half : idesc_full, 1u);
            umma_commit(&empty_bar[stage]);
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
      umma_commit(&tmem_full[acc]);  // accumulator of this tile complete -> epilogue
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp >= 4) {
    // =============================== epilogue ===============================
    const int ew = warp - 4;  // == warp % 4: the TMEM lane quarter this warp may read
    const uint32_t stage_base = smem_u32(epi_smem + ew * Cfg::EPI_WARP_BYTES);
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(ew * 32) << 16);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const TileCoord c = decode_tile(tile, mt0, mt1, n_tiles, p.raster_group);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      epilogue_tile<BN>(p, c.e, c.m, c.n, cnt0, cnt1, t_lane + static_cast<uint32_t>(acc * BN), stage_base, lane, ew,
                        [&]() {
                          tc_fence_before();
                          __syncwarp();
                          if (lane == 0) mbar_arrive(&tmem_empty[acc]);
                        });
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): a cluster of two CTAs on one TPC computes a 256 x 256 super-tile.
// Each CTA loads its own 128 A rows and HALF of the B tile (128 weight rows); tcgen05.mma.cta_group::2
// (M = 256), issued by the leader CTA only, reads both halves of B from the two CTAs' shared memory and writes
// rows [0,128) of D to the leader's TMEM and rows [128,256) to the peer's.  Per SM this halves the B bytes
// moved L2 -> SM and read from shared memory by the tensor core (the 1-CTA kernel runs shared memory at
// ~85 % of its bandwidth), which frees two more pipeline stages (6 x 32 KB) and lowers power.
//   full[stage]      : leader CTA only; both CTAs' TMA loads post their bytes on it (shared::cluster address)
//   empty[stage]     : one per CTA; released by the leader's tcgen05.commit multicast to both CTAs
//   tmem_full[acc]   : one per CTA; tcgen05.commit multicast
//   tmem_empty[acc]  : leader CTA only, 16 arrivals (8 epilogue warps x 2 CTAs; the peer arrives remotely)
// ---------------------------------------------------------------------------------------------
constexpr int PAIR_THREADS = 384;  // TMA, MMA, TMEM-alloc, idle + 8 epilogue warps

struct PairCfg {
  static constexpr int BN = 256;
  static constexpr int STAGES = 6;
  static constexpr int A_BYTES = BM * BK * 2;          // 16 KB
  static constexpr int B_BYTES = (BN / 2) * BK * 2;    // 16 KB: this CTA's half of the B tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EPI_WARPS = 8;
  static constexpr int EPI_WARP_BYTES = 32 * 128;      // 32 rows x 64 staged columns
  static constexpr int TMEM_COLS = 512;
  static constexpr int NUM_BARS = 2 * STAGES + 4;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_WARPS * EPI_WARP_BYTES + NUM_BARS * 8 + 16 + 1024;
};

struct PairTile {
  int e, m2, n;  // expert, index of the 256-row super-tile inside the expert segment, n-tile
};

__device__ __forceinline__ PairTile decode_pair_tile(int tile, int mp0, int mp1, int n_tiles, int group_pairs) {
  PairTile c;
  int mp_e = mp0;
  c.e = 0;
  const int t0 = mp0 * n_tiles;
  if (tile >= t0) {
    c.e = 1;
    tile -= t0;
    mp_e = mp1;
  }
  const int GROUP = group_pairs;  // super-tiles (256 rows) per raster group
  const int per_group = GROUP * n_tiles;
  const int g = tile / per_group;
  const int r = tile - g * per_group;
  const int m_lo = g * GROUP;
  const int gm = min(GROUP, mp_e - m_lo);
  c.n = r / gm;
  c.m2 = m_lo + (r - c.n * gm);
  return c;
}

template <bool TB>
__global__ void __launch_bounds__(PAIR_THREADS, 1)
    k3_grouped_gemm_pair(const __grid_constant__ GemmTmaps tm, const GemmDev p) {
  using Cfg = PairCfg;
  constexpr int BN = Cfg::BN;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi_smem = smem + STAGES * Cfg::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + Cfg::EPI_WARPS * Cfg::EPI_WARP_BYTES);
  uint64_t* full_bar = bars;                   // used in the leader CTA
  uint64_t* empty_bar = bars + STAGES;         // per CTA
  uint64_t* tmem_full = bars + 2 * STAGES;     // per CTA
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;  // used in the leader CTA
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(bars + Cfg::NUM_BARS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();  // 0 = leader
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 2 * Cfg::EPI_WARPS);  // every epilogue warp of both CTAs
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc_pair(tmem_base_smem, Cfg::TMEM_COLS);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm.a);
    tma_prefetch_desc(&tm.w[0][0]);
  }
  tc_fence_before();
  cluster_sync_all();  // both CTAs' barriers are initialised before any remote arrive / multicast commit
  __syncthreads();     // redundant with the cluster barrier; compute-sanitizer's racecheck only models the CTA barrier as
                       // ordering warp 2's tcgen05.alloc result (written to shared memory) before the reads below
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;

  const int cnt0 = min(max(p.counts[0], 0), p.rows_cap);
  const int cnt1 = p.single_expert ? 0 : min(max(p.counts[1], 0), p.rows_cap - cnt0);
  const int mp0 = (cnt0 + 2 * BM - 1) / (2 * BM);
  const int mp1 = (cnt1 + 2 * BM - 1) / (2 * BM);
  const bool swiglu = p.mode == VEX_EPI_SWIGLU;
  const int n_span = swiglu ? BN / 2 : BN;
  const int n_tiles = (p.N + n_span - 1) / n_span;
  const int total_tiles = (mp0 + mp1) * n_tiles;
  const int num_kb = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    // =============================== TMA producer (both CTAs) ===============================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
      const PairTile c = decode_pair_tile(tile, mp0, mp1, n_tiles, p.raster_group / 2);
      const int row0 = (c.e ? cnt0 : 0) + (2 * c.m2 + static_cast<int>(rank)) * BM;
      // this CTA's half of the 256 accumulator columns: leader = [0,128), peer = [128,256).
      // SwiGLU: leader half = gate_proj rows, peer half = up_proj rows of the same 128 output columns.
      const int nrow = swiglu ? c.n * n_span : c.n * BN + static_cast<int>(rank) * (BN / 2);
      const CUtensorMap* wmap = &tm.w[c.e][swiglu ? rank : 0];
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
        const uint32_t full_leader = mapa_shared(smem_u32(&full_bar[stage]), 0);
        if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
        tma_load_2d_pair(sa, &tm.a, full_leader, kb * BK, row0);
        if constexpr (TB) {  // this CTA's 128 columns of B = two {64 n, 64 k} boxes
          tma_load_2d_pair(sa + Cfg::A_BYTES, wmap, full_leader, nrow, kb * BK);
          tma_load_2d_pair(sa + Cfg::A_BYTES + MN_CHUNK_BYTES, wmap, full_leader, nrow + 64, kb * BK);
        } else {
          tma_load_2d_pair(sa + Cfg::A_BYTES, wmap, full_leader, kb * BK, nrow);
        }
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      if ((p.lora_mask >> c.e) & 1) {
        const int halves = swiglu ? 2 : 1;
        for (int h = 0; h < halves; ++h) {
          for (int ls = 0; ls < p.lora_steps; ++ls) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
            const uint32_t full_leader = mapa_shared(smem_u32(&full_bar[stage]), 0);
            if (swiglu) {
              // N = 128 MMA into accumulator half h: each CTA supplies 64 rows of lora_B (half-box)
              if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * (Cfg::A_BYTES + Cfg::B_BYTES / 2));
              tma_load_2d_pair(sa, &tm.t[h], full_leader, ls * BK, row0);
              tma_load_2d_pair(sa + Cfg::A_BYTES, &tm.lb64[c.e][h], full_leader, ls * BK,
                               c.n * n_span + static_cast<int>(rank) * 64);
            } else {
              if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
              tma_load_2d_pair(sa, &tm.t[0], full_leader, ls * BK, row0);
              if constexpr (TB) {
                tma_load_2d_pair(sa + Cfg::A_BYTES, &tm.lb[c.e][0], full_leader, nrow, ls * BK);
                tma_load_2d_pair(sa + Cfg::A_BYTES + MN_CHUNK_BYTES, &tm.lb[c.e][0], full_leader, nrow + 64, ls * BK);
              } else {
                tma_load_2d_pair(sa + Cfg::A_BYTES, &tm.lb[c.e][0], full_leader, ls * BK, nrow);
              }
            }
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1 && lane == 0 && rank == 0) {
    // =============================== MMA issuer (leader CTA) ===============================
    constexpr uint32_t idesc_full = umma_idesc_bf16(2 * BM, BN, 0, TB ? 1 : 0);
    constexpr uint32_t idesc_half = umma_idesc_bf16(2 * BM, BN / 2, 0, TB ? 1 : 0);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
      const PairTile c = decode_pair_tile(tile, mp0, mp1, n_tiles, p.raster_group / 2);
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
      uint32_t accumulate = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
        const uint64_t da = umma_desc_kmajor_sw128(sa);
        const uint64_t db = b_desc<TB>(sa + Cfg::A_BYTES);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          umma_ss_pair(d_tmem, da + 2 * k, db + B_KSTEP<TB> * k, idesc_full, accumulate);
          accumulate = 1;
        }
        umma_commit_pair(&empty_bar[stage], 3);  // frees the slot in both CTAs
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      if ((p.lora_mask >> c.e) & 1) {
        const int halves = swiglu ? 2 : 1;
        for (int h = 0; h < halves; ++h) {
          for (int ls = 0; ls < p.lora_steps; ++ls) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
            const uint64_t da = umma_desc_kmajor_sw128(sa);
            const uint64_t db = b_desc<TB>(sa + Cfg::A_BYTES);
            const uint32_t d_half = d_tmem + static_cast<uint32_t>(swiglu ? h * (BN / 2) : 0);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              umma_ss_pair(d_half, da + 2 * k, db + B_KSTEP<TB> * k, swiglu ? idesc_half : idesc_full, 1u);
            umma_commit_pair(&empty_bar[stage], 3);
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
      umma_commit_pair(&tmem_full[acc], 3);  // accumulators of both CTAs complete -> both epilogues
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp >= 4) {
    // =============================== epilogue (both CTAs, own 128 rows) ===============================
    const int ew = (warp - 4) & 3;    // TMEM lane quarter (== warp % 4)
    const int half = (warp - 4) >> 2;  // which 128 accumulator columns
    const uint32_t stage_base = smem_u32(epi_smem + (warp - 4) * Cfg::EPI_WARP_BYTES);
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(ew * 32) << 16);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
      const PairTile c = decode_pair_tile(tile, mp0, mp1, n_tiles, p.raster_group / 2);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      epilogue_tile_half(p, c.e, 2 * c.m2 + static_cast<int>(rank), c.n, cnt0, cnt1,
                        t_lane + static_cast<uint32_t>(acc * BN), stage_base, lane, ew, half, [&]() {
                          tc_fence_before();
                          __syncwarp();
                          if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&tmem_empty[acc]), 0));
                        });
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  cluster_sync_all();  // the peer's shared memory / TMEM stay alive until the leader's last MMA has been consumed
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// host side: launch (tensor-map encoding lives in tmap.cu)
// ---------------------------------------------------------------------------------------------
int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);
int num_sms();

template <int BN, bool TB>
static int launch_gemm(const GemmTmaps& tm, const GemmDev& dev, cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    VEX_CUDA_TRY(cudaFuncSetAttribute(k3_grouped_gemm<BN, TB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      GemmCfg<BN>::SMEM_BYTES));
    configured = true;
  }
  k3_grouped_gemm<BN, TB><<<num_sms(), GEMM_THREADS, GemmCfg<BN>::SMEM_BYTES, s>>>(tm, dev);
  VEX_LAUNCH_CHECK();
  return VEX_OK;
}

template <bool TB>
static int launch_gemm_pair(const GemmTmaps& tm, const GemmDev& dev, cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    VEX_CUDA_TRY(cudaFuncSetAttribute(k3_grouped_gemm_pair<TB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      PairCfg::SMEM_BYTES));
    configured = true;
  }
  cudaLaunchConfig_t cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(num_sms() & ~1);  // whole CTA pairs
  cfg.blockDim = dim3(PAIR_THREADS);
  cfg.dynamicSmemBytes = PairCfg::SMEM_BYTES;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  VEX_CUDA_TRY(cudaLaunchKernelEx(&cfg, k3_grouped_gemm_pair<TB>, tm, dev));
  return VEX_OK;
}

// VEX_GEMM_PAIR=0 selects the single-CTA kernel for every shape (A/B switch; default: CTA pairs)
static bool use_pair_kernel() {
  const char* e = std::getenv("VEX_GEMM_PAIR");  // read per call so tests can flip it
  return !(e && e[0] == '0');
}

}  // namespace vex

extern "C" int vex_grouped_gemm(const vexGemmArgs* a, vexStream stream) {
  using namespace vex;
  if (!a || !a->a || (!a->out && a->mode != VEX_EPI_CE) || !a->counts || !a->w[0][0]) return VEX_E_INVALID;
  if (a->rows_cap <= 0 || a->N <= 0 || a->K <= 0) return VEX_E_INVALID;
  if (a->N % 8 != 0 || a->K % 8 != 0 || a->ldo % 8 != 0 || a->lda % 8 != 0 || a->ldw % 8 != 0)
    return VEX_E_UNSUPPORTED;
  if (a->mode < VEX_EPI_PLAIN || a->mode > VEX_EPI_CE_BWD) return VEX_E_INVALID;
  const bool swiglu = a->mode == VEX_EPI_SWIGLU;
  if (!a->single_expert && !a->w[1][0]) return VEX_E_INVALID;
  if (swiglu && (!a->w[0][1] || (!a->single_expert && !a->w[1][1]))) return VEX_E_INVALID;
  if ((a->mode == VEX_EPI_RESIDUAL || a->mode == VEX_EPI_DROPOUT_ACC) && !a->residual) return VEX_E_INVALID;
  if (a->mode == VEX_EPI_DROPOUT_ACC && (!(a->dropout_p >= 0.f && a->dropout_p < 1.f) || a->N % 2 != 0))
    return VEX_E_INVALID;
  if (a->mode == VEX_EPI_ROPE) {
    if (!a->rope_cos || !a->rope_sin || !a->position_ids || !a->sorted_to_flat || a->rope_len <= 0)
      return VEX_E_INVALID;
    if (a->rope_cols % 128 != 0) return VEX_E_UNSUPPORTED;  // heads of 128
  }
  const bool tb = a->w_transposed != 0;
  if (tb && (swiglu || a->mode == VEX_EPI_ROPE)) return VEX_E_UNSUPPORTED;
  if (a->bias || a->act != VEX_ACT_NONE) {
    if (a->mode != VEX_EPI_PLAIN && a->mode != VEX_EPI_RESIDUAL) return VEX_E_UNSUPPORTED;
    if (a->act != VEX_ACT_NONE && (a->act != VEX_ACT_GELU || a->mode != VEX_EPI_PLAIN)) return VEX_E_UNSUPPORTED;
    if (a->N % 32 != 0 || tb || (reinterpret_cast<uintptr_t>(a->bias) & 15)) return VEX_E_UNSUPPORTED;
  }
  const bool ce = a->mode == VEX_EPI_CE || a->mode == VEX_EPI_CE_BWD;
  if (ce) {
    if (!a->single_expert || tb || a->N <= 64 || !a->ce_labels) return VEX_E_INVALID;
    if (a->mode == VEX_EPI_CE && (!a->ce_pmax || !a->ce_psum || !a->ce_zlabel)) return VEX_E_INVALID;
    if (a->mode == VEX_EPI_CE_BWD && (!a->ce_lse || !a->ce_w || !a->ce_dloss)) return VEX_E_INVALID;
  }
  const bool small_n = a->N <= 64 && a->mode == VEX_EPI_PLAIN;
  const int BN = small_n ? 64 : 256;
  const int half_rows = swiglu ? 128 : BN / 2;

  GemmTmaps tm;
  std::memset(&tm, 0, sizeof(tm));
  int rc;
  if ((rc = make_tmap_2d(&tm.a, a->a, a->rows_cap, a->K, a->lda, BM)) != VEX_OK) return rc;
  const int n_exp = a->single_expert ? 1 : 2;
  for (int e = 0; e < n_exp; ++e)
    for (int h = 0; h < (swiglu ? 2 : 1); ++h)
      if ((rc = tb ? make_tmap_2d(&tm.w[e][h], a->w[e][h], a->K, a->N, a->ldw, 64)  // [K, N]: boxes of {64 n, 64 k}
                   : make_tmap_2d(&tm.w[e][h], a->w[e][h], a->N, a->K, a->ldw, half_rows)) != VEX_OK)
        return rc;

  GemmDev dev;
  std::memset(&dev, 0, sizeof(dev));
  if (a->lora_r > 0) {
    if (a->lora_r % 8 != 0 || a->ldt % 8 != 0) return VEX_E_UNSUPPORTED;
    if (small_n) return VEX_E_UNSUPPORTED;
    dev.lora_steps = ceil_div(a->lora_r, BK);
    for (int h = 0; h < (swiglu ? 2 : 1); ++h) {
      bool any = false;
      for (int e = 0; e < n_exp; ++e) {
        if (!a->lora_b[e][h]) continue;
        if (swiglu && !a->lora_b[e][1 - h]) return VEX_E_UNSUPPORTED;  // gate and up adapters come in pairs
        any = true;
        dev.lora_mask |= 1 << e;
        // transposed form: the K-extension reads lora_A [r, N] (dX += (s dT) . A)
        if ((rc = tb ? make_tmap_2d(&tm.lb[e][h], a->lora_b[e][h], a->lora_r, a->N, a->N, 64)
                     : make_tmap_2d(&tm.lb[e][h], a->lora_b[e][h], a->N, a->lora_r, a->lora_r, half_rows)) != VEX_OK)
          return rc;
        if (swiglu &&
            (rc = make_tmap_2d(&tm.lb64[e][h], a->lora_b[e][h], a->N, a->lora_r, a->lora_r, 64)) != VEX_OK)
          return rc;
      }
      if (any) {
        if (!a->lora_t[h]) return VEX_E_INVALID;
        if ((rc = make_tmap_2d(&tm.t[h], a->lora_t[h], a->rows_cap, a->lora_r, a->ldt, BM)) != VEX_OK) return rc;
      }
    }
    if (dev.lora_mask == 0) dev.lora_steps = 0;
  }
  dev.counts = a->counts;
  dev.row_map = a->row_map;
  dev.out = static_cast<__nv_bfloat16*>(a->out);
  dev.residual = static_cast<const __nv_bfloat16*>(a->residual);
  dev.rope_cos = static_cast<const __nv_bfloat16*>(a->rope_cos);
  dev.rope_sin = static_cast<const __nv_bfloat16*>(a->rope_sin);
  dev.position_ids = a->position_ids;
  dev.sorted_to_flat = a->sorted_to_flat;
  dev.ldo = a->ldo;
  dev.alpha = a->alpha;
  dev.rope_len = a->rope_len;
  dev.rope_cols = a->rope_cols;
  dev.rows_cap = a->rows_cap;
  dev.N = a->N;
  dev.K = a->K;
  {
    const char* rg = std::getenv("VEX_GEMM_RASTER");  // A/B switch: m-tiles per raster group
    dev.raster_group = rg ? std::max(2, std::atoi(rg) & ~1) : raster_group_for(a->K);
  }
  dev.mode = a->mode;
  dev.single_expert = a->single_expert;
  dev.trans_b = tb;
  dev.ce_labels = a->ce_labels;
  dev.ce_pmax = a->ce_pmax;
  dev.ce_psum = a->ce_psum;
  dev.ce_zlabel = a->ce_zlabel;
  dev.ce_lse = a->ce_lse;
  dev.ce_w = a->ce_w;
  dev.ce_dloss = a->ce_dloss;
  // softmax partial slots per row: one per 128-column HALF tile (the CTA-pair kernel's two epilogue warps of a row work
  // independently; the single-CTA kernel fills slot 2n and neutralises slot 2n + 1)
  dev.ce_tiles = 2 * ceil_div(a->N, 256);
  dev.bias = static_cast<const __nv_bfloat16*>(a->bias);
  dev.act = a->act;
  if (a->kv_k || a->kv_v) {
    if (a->mode != VEX_EPI_ROPE || !a->kv_k || !a->kv_v || a->kv_seq_len <= 0 || a->kv_capacity <= 0) return VEX_E_INVALID;
    if (a->N != 3 * (a->rope_cols / 2)) return VEX_E_UNSUPPORTED;  // the QKV projection: [q | k | v], q and k rotated
    if ((reinterpret_cast<uintptr_t>(a->kv_k) | reinterpret_cast<uintptr_t>(a->kv_v)) & 15) return VEX_E_INVALID;
    dev.kv_k = static_cast<__nv_bfloat16*>(a->kv_k);
    dev.kv_v = static_cast<__nv_bfloat16*>(a->kv_v);
    dev.kv_pos = a->kv_pos;
    dev.kv_seq = a->kv_seq_len;
    dev.kv_cap = a->kv_capacity;
  }
  if (a->mode == VEX_EPI_DROPOUT_ACC) {
    dev.drop_thresh16 = static_cast<uint32_t>(a->dropout_p * 65536.0f + 0.5f);
    dev.drop_seed_lo = static_cast<uint32_t>(a->dropout_seed);
    dev.drop_seed_hi = static_cast<uint32_t>(a->dropout_seed >> 32);
    dev.alpha = a->alpha / (1.0f - a->dropout_p);
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (tb) {
    if (small_n) return launch_gemm<64, true>(tm, dev, s);
    return use_pair_kernel() ? launch_gemm_pair<true>(tm, dev, s) : launch_gemm<256, true>(tm, dev, s);
  }
  if (small_n) return launch_gemm<64, false>(tm, dev, s);
  return use_pair_kernel() ? launch_gemm_pair<false>(tm, dev, s) : launch_gemm<256, false>(tm, dev, s);
}
