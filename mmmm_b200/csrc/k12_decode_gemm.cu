// K12 -- skinny GEMM of the decode step (q_len == 1): out[b, n] = sum_k x[b, k] * W[n, k] for a batch of at most 32
// rows -- the five nn.Linear calls of a generation step through the LANGUAGE expert (get_expert_mask's L == 1 rule,
// modeling_cogvlm.py:67; Linears :244-245, :278-279, MLP.forward :54-56) with the same fused epilogues as K3: rotary +
// KV-cache append (:188-193, :258-262), residual add (:321, :330), SiLU-gate (:55), LoRA K-extension.
//
// This is an HBM-BOUND kernel (every weight is read once per token; 404.75 MB per layer), not a tensor-throughput one:
// a 128 x 256 tcgen05 tile over 8 live rows leaves most SMs idle for the small-N projections (dense / down: 16 tiles),
// which is what round 2 measured for the decode step on K3 (1.8 TB/s).  Design for bandwidth instead:
//   * one CTA (8 warps) per group of 16 output features (and their partner tile: up_proj for SwiGLU, column j + 64
//     for rotary), 256 ... 768 CTAs per projection -- every SM streams;
//   * the K dimension is interleaved over the 8 warps in 64-element chunks; a thread loads 32 bytes (16 consecutive k,
//     one LDG.256) of weight row g and of row g + 8 straight from global memory into the A fragments of four
//     mma.sync.m16n8k16 steps -- the k order inside a chunk is permuted identically for the B fragment, so no shuffle /
//     shared-memory staging of the weights is needed, 4 threads cover one full 128-byte line of a row, and the next
//     chunk is in flight while the current one is multiplied (double-buffered registers);
//   * the batch is the N dimension of the MMA (n8 tiles, 1 ... 4 of them); x is read through the read-only path
//     (L1-resident: 8 x 4096 x 2 B = 64 KB, shared by every CTA on the SM);
//   * partial sums of the 8 warps are reduced through shared memory, then 16 x B outputs take the epilogue.
// mma.sync is used on purpose: the math is 0.1 % of the tensor peak, the operand path (global -> registers) is what
// matters, and tcgen05 would force the weights through shared memory.
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace vex {

int num_sms();

constexpr int DG_THREADS = 256;
constexpr int DG_WARPS = 8;

struct DecGemm {
  const __nv_bfloat16* x;     // [B, K]
  const __nv_bfloat16* w[2];  // [N, K]; w[1] = up_proj (SWIGLU) or nullptr
  const __nv_bfloat16* lt[2];  // LoRA T = s * x . A^T  [B, r] per half, or nullptr
  const __nv_bfloat16* lb[2];  // lora_B [N, r] per half
  __nv_bfloat16* out;
  const __nv_bfloat16* residual;
  const __nv_bfloat16* rope_cos;
  const __nv_bfloat16* rope_sin;
  const int64_t* position_ids;
  __nv_bfloat16* kv_k;
  __nv_bfloat16* kv_v;
  const int32_t* kv_pos;
  int64_t ldx, ldw, ldt, ldo;
  int B, N, K, mode, lora_r, rope_len, rope_cols, kv_cap;
  float alpha;
};

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }

// One 64-element k chunk of this thread's fragments: 32 bytes (16 consecutive k) of weight rows g and g + 8 per tile.
template <int TILES>
struct WFrag {
  U32x8 lo[TILES], hi[TILES];
};

template <int TILES>
__device__ __forceinline__ void load_w(WFrag<TILES>& f, const __nv_bfloat16* const* wrow_lo,
                                       const __nv_bfloat16* const* wrow_hi, int kb) {
#pragma unroll
  for (int tl = 0; tl < TILES; ++tl) {
    f.lo[tl] = ld_stream_256(wrow_lo[tl] + kb);
    f.hi[tl] = ld_stream_256(wrow_hi[tl] + kb);
  }
}

// c[tile][nb] += W_tile[16, chunk] . x[nb*8 .., chunk]^T.  The 16 k of a thread are consumed as four m16n8k16 steps
// (step j: register pair 2j, 2j + 1); the B fragment takes the same 16 k of x row g, so the k permutation cancels.
template <int NB, int TILES>
__device__ __forceinline__ void mma_chunk(float (&c)[2][NB][4], const WFrag<TILES>& f, const __nv_bfloat16* const* xrow,
                                          const bool* xlive, int kb) {
#pragma unroll
  for (int nb = 0; nb < NB; ++nb) {
    U32x8 xb;
    if (xlive[nb]) {
      xb = ld_coherent_256(xrow[nb] + kb);  // L1-resident; written by the previous kernel of the PDL chain
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) xb.v[j] = 0u;
    }
#pragma unroll
    for (int tl = 0; tl < TILES; ++tl)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        mma_bf16_16816(c[tl][nb], f.lo[tl].v[2 * j], f.hi[tl].v[2 * j], f.lo[tl].v[2 * j + 1], f.hi[tl].v[2 * j + 1],
                       xb.v[2 * j], xb.v[2 * j + 1]);
  }
}

// accumulates over the 64-element chunks this warp owns (chunk i = warp, warp + 8, ...), software-pipelined through a
// ring of NBUF register fragments: NBUF - 1 chunks are in flight while one is multiplied, so a warp always has loads
// outstanding.  One-tile launches (dense / down: only 256 CTAs, < 2 per SM) run 4 deep, two-tile launches 2 deep
// (a fragment is 16 registers per tile; batches above 16 rows keep 2 for their extra accumulators).
template <int NB, int TILES, int WARPS = DG_WARPS>
__device__ __forceinline__ void accumulate(float (&c)[2][NB][4], const __nv_bfloat16* const* wrow_lo,
                                           const __nv_bfloat16* const* wrow_hi, const __nv_bfloat16* const* xrow,
                                           const bool* xlive, int K, int warp, int t, bool wait_after_prefetch = false) {
  constexpr int NBUF = (TILES == 1 && NB <= 2) ? 4 : 2;
  const int nchunks = K >> 6;
  WFrag<TILES> f[NBUF];
#pragma unroll
  for (int b = 0; b < NBUF - 1; ++b) {
    const int ip = warp + b * WARPS;
    if (ip < nchunks) load_w<TILES>(f[b], wrow_lo, wrow_hi, (ip << 6) + 16 * t);
  }
  // the first weight fragments are in flight; x (and everything the epilogue touches) belongs to the previous kernel
  if (wait_after_prefetch) pdl_wait();
  for (int i = warp; i < nchunks; i += WARPS * NBUF) {
#pragma unroll
    for (int b = 0; b < NBUF; ++b) {
      const int ic = i + b * WARPS;                  // chunk multiplied now: lives in f[b]
      if (ic >= nchunks) break;
      const int ip = ic + (NBUF - 1) * WARPS;        // chunk prefetched into the slot consumed one step ago
      if (ip < nchunks) load_w<TILES>(f[(b + NBUF - 1) % NBUF], wrow_lo, wrow_hi, (ip << 6) + 16 * t);
      mma_chunk<NB, TILES>(c, f[b], xrow, xlive, (ic << 6) + 16 * t);
    }
  }
}

// WARPS = 8 (256 threads, 2 CTAs per SM) or 4 (128 threads, 4 CTAs per SM: the form for launches whose CTA count lies
// between 296 and 592 -- the QKV projection's 384 -- which then run as ONE resident wave instead of 1.3)
template <int NB, int TILES, int WARPS, int CTAS_PER_SM = (WARPS == 4 ? 4 : 2)>
__global__ void __launch_bounds__(WARPS * 32, CTAS_PER_SM) k12_decode_gemm(const DecGemm p) {
  constexpr int DG_WARPS = WARPS, DG_THREADS = WARPS * 32;  // shadow the 8-warp defaults below
  __shared__ float red[DG_WARPS][TILES][NB][32][4];   // per-warp partial fragments
  __shared__ float fin[TILES][NB * 8][16];            // reduced [tile][batch row][feature]
  pdl_trigger();  // the next kernel of the step may be scheduled: its weight prefetch overlaps this kernel's stream
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int task = blockIdx.x;
  // first output feature of the two tiles
  int n_a, n_b;
  if (p.mode == VEX_EPI_ROPE) {   // (j, j + 64) inside one head of 128
    n_a = (task >> 2) * 128 + (task & 3) * 16;
    n_b = n_a + 64;
  } else {
    n_a = n_b = task * 16;
  }
  const __nv_bfloat16* wa = p.w[0];
  const __nv_bfloat16* wb = (p.mode == VEX_EPI_SWIGLU) ? p.w[1] : p.w[0];
  float c[2][NB][4];
#pragma unroll
  for (int tl = 0; tl < 2; ++tl)
#pragma unroll
    for (int nb = 0; nb < NB; ++nb)
#pragma unroll
      for (int r = 0; r < 4; ++r) c[tl][nb][r] = 0.f;

  const __nv_bfloat16* xrow[NB];
  bool xlive[NB];
#pragma unroll
  for (int nb = 0; nb < NB; ++nb) {
    xlive[nb] = nb * 8 + g < p.B;
    xrow[nb] = p.x + static_cast<int64_t>(min(nb * 8 + g, p.B - 1)) * p.ldx;
  }
  {
    const __nv_bfloat16* lo[2] = {wa + static_cast<int64_t>(n_a + g) * p.ldw, wb + static_cast<int64_t>(n_b + g) * p.ldw};
    const __nv_bfloat16* hi[2] = {wa + static_cast<int64_t>(n_a + g + 8) * p.ldw,
                                  wb + static_cast<int64_t>(n_b + g + 8) * p.ldw};
    accumulate<NB, TILES, WARPS>(c, lo, hi, xrow, xlive, p.K, warp, t, /*wait_after_prefetch=*/true);
  }
  if (p.lora_r > 0) {  // K-extension: += T . lora_B^T  (r is a multiple of 64 here; r = 64 -> one chunk, warp 0)
    const __nv_bfloat16* trow[NB];
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) trow[nb] = p.lt[0] + static_cast<int64_t>(min(nb * 8 + g, p.B - 1)) * p.ldt;
    const __nv_bfloat16* lo[2] = {p.lb[0] + static_cast<int64_t>(n_a + g) * p.lora_r,
                                  (p.mode == VEX_EPI_SWIGLU ? p.lb[1] : p.lb[0]) + static_cast<int64_t>(n_b + g) * p.lora_r};
    const __nv_bfloat16* hi[2] = {lo[0] + 8 * static_cast<int64_t>(p.lora_r), lo[1] + 8 * static_cast<int64_t>(p.lora_r)};
    if (p.mode == VEX_EPI_SWIGLU) {
      // gate and up have their own T: accumulate the two tiles separately
      float cg[2][NB][4], cu[2][NB][4];
#pragma unroll
      for (int nb = 0; nb < NB; ++nb)
#pragma unroll
        for (int r = 0; r < 4; ++r) cg[0][nb][r] = cg[1][nb][r] = cu[0][nb][r] = cu[1][nb][r] = 0.f;
      const __nv_bfloat16* urow[NB];
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) urow[nb] = p.lt[1] + static_cast<int64_t>(min(nb * 8 + g, p.B - 1)) * p.ldt;
      const __nv_bfloat16* lo_g[2] = {lo[0], lo[0]};
      const __nv_bfloat16* hi_g[2] = {hi[0], hi[0]};
      const __nv_bfloat16* lo_u[2] = {lo[1], lo[1]};
      const __nv_bfloat16* hi_u[2] = {hi[1], hi[1]};
      accumulate<NB, 1, WARPS>(cg, lo_g, hi_g, trow, xlive, p.lora_r, warp, t);
      accumulate<NB, 1, WARPS>(cu, lo_u, hi_u, urow, xlive, p.lora_r, warp, t);
#pragma unroll
      for (int nb = 0; nb < NB; ++nb)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          c[0][nb][r] += cg[0][nb][r];
          if (TILES == 2) c[1][nb][r] += cu[0][nb][r];
        }
    } else {
      accumulate<NB, TILES, WARPS>(c, lo, hi, trow, xlive, p.lora_r, warp, t);
    }
  }

  // ---- reduce the 8 warps' fragments ----
#pragma unroll
  for (int tl = 0; tl < TILES; ++tl)
#pragma unroll
    for (int nb = 0; nb < NB; ++nb)
      *reinterpret_cast<float4*>(red[warp][tl][nb][lane]) = make_float4(c[tl][nb][0], c[tl][nb][1], c[tl][nb][2], c[tl][nb][3]);
  __syncthreads();
  for (int idx = threadIdx.x; idx < TILES * NB * 128; idx += DG_THREADS) {
    const int r = idx & 3, L = (idx >> 2) & 31, nb = (idx >> 7) % NB, tl = idx / (NB * 128);
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < DG_WARPS; ++w) s += red[w][tl][nb][L][r];
    // fragment element -> (feature row, batch column): c0,c1 row g cols 2t,2t+1; c2,c3 row g+8
    const int feat = (L >> 2) + ((r & 2) ? 8 : 0), brow = nb * 8 + 2 * (L & 3) + (r & 1);
    fin[tl][brow][feat] = s;
  }
  __syncthreads();

  // ---- epilogue: 16 features x B rows (x 2 tiles) ----
  for (int idx = threadIdx.x; idx < p.B * 16; idx += DG_THREADS) {
    const int b = idx >> 4, f = idx & 15;
    const float va = fin[0][b][f];
    const float vb = TILES == 2 ? fin[TILES - 1][b][f] : 0.f;
    __nv_bfloat16* orow = p.out + static_cast<int64_t>(b) * p.ldo;
    if (p.mode == VEX_EPI_PLAIN) {
      orow[n_a + f] = __float2bfloat16_rn(va * p.alpha);
    } else if (p.mode == VEX_EPI_RESIDUAL) {
      const float r = __bfloat162float(p.residual[static_cast<int64_t>(b) * p.ldo + n_a + f]);
      orow[n_a + f] = __float2bfloat16_rn(bf16r(va) + r);   // Linear output -> bf16, then the eager bf16 add
    } else if (p.mode == VEX_EPI_SWIGLU) {
      const float gt = bf16r(va), up = bf16r(vb);
      orow[n_a + f] = __float2bfloat16_rn(bf16r(silu_f(gt)) * up);
    } else {  // VEX_EPI_ROPE
      float o1, o2;
      if (n_a < p.rope_cols) {
        const int64_t pz = p.position_ids[b];
        const int pos = static_cast<int>(min(max(pz, int64_t(0)), int64_t(p.rope_len - 1)));
        const int j = (n_a & 127) + f;  // column inside the head, < 64
        const __nv_bfloat16* cr = p.rope_cos + static_cast<int64_t>(pos) * 128;
        const __nv_bfloat16* sr = p.rope_sin + static_cast<int64_t>(pos) * 128;
        const float x1 = bf16r(va), x2 = bf16r(vb);
        o1 = bf16r(x1 * __bfloat162float(cr[j])) - bf16r(x2 * __bfloat162float(sr[j]));
        o2 = bf16r(x2 * __bfloat162float(cr[j + 64])) + bf16r(x1 * __bfloat162float(sr[j + 64]));
      } else {
        o1 = va;
        o2 = vb;
      }
      const __nv_bfloat16 h1 = __float2bfloat16_rn(o1), h2 = __float2bfloat16_rn(o2);
      orow[n_a + f] = h1;
      orow[n_b + f] = h2;
      const int Hh = p.rope_cols >> 1;
      if (p.kv_k != nullptr && n_a >= Hh) {  // K (post-rotary) and V heads are appended to the cache at *kv_pos
        const bool is_v = n_a >= 2 * Hh;
        const int head = (n_a - (is_v ? 2 * Hh : Hh)) >> 7;
        const int l = p.kv_pos ? p.kv_pos[0] : 0;
        if (l >= 0 && l < p.kv_cap) {
          __nv_bfloat16* dst = (is_v ? p.kv_v : p.kv_k) +
                               ((static_cast<int64_t>(b) * (Hh >> 7) + head) * p.kv_cap + l) * 128;
          dst[(n_a & 127) + f] = h1;
          dst[(n_b & 127) + f] = h2;
        }
      }
    }
  }
}

template <int NB, int TILES>
static int launch_dg(const DecGemm& p, int tasks, cudaStream_t s) {
  // one resident wave if 4-warp CTAs (4 per SM) make it possible and 8-warp CTAs (2 per SM) do not
  static const bool narrow_ok = [] {  // VEX_K12_NARROW=0: always 8 warps (A/B timing)
    const char* e = std::getenv("VEX_K12_NARROW");
    return !(e && e[0] == '0');
  }();
  const int sms = num_sms();
  if (narrow_ok && NB <= 2 && tasks > 2 * sms && tasks <= 4 * sms) {
    VEX_CUDA_TRY(launch_pdl(k12_decode_gemm<NB, TILES, 4>, dim3(tasks), dim3(128), 0, s, p));
    return VEX_OK;
  }
  if constexpr (NB == 1) {  // 5 CTAs of 4 warps per SM (<= 102 registers): gate / up's 688 CTAs in one wave
    static const bool five_ok = [] {
      const char* e = std::getenv("VEX_K12_FIVE");
      return !(e && e[0] == '0');
    }();
    if (narrow_ok && five_ok && tasks > 4 * sms && tasks <= 5 * sms) {
      VEX_CUDA_TRY(launch_pdl(k12_decode_gemm<NB, TILES, 4, 5>, dim3(tasks), dim3(128), 0, s, p));
      return VEX_OK;
    }
  }
  VEX_CUDA_TRY(launch_pdl(k12_decode_gemm<NB, TILES, 8>, dim3(tasks), dim3(DG_THREADS), 0, s, p));
  return VEX_OK;
}

}  // namespace vex

extern "C" int vex_decode_gemm(const vexGemmArgs* a, vexStream stream) {
  using namespace vex;
  if (!a || !a->a || !a->out || !a->w[0][0]) return VEX_E_INVALID;
  if (a->rows_cap <= 0 || a->rows_cap > 32 || a->N <= 0 || a->K <= 0) return VEX_E_UNSUPPORTED;
  if (!a->single_expert || a->w_transposed || a->bias || a->act != VEX_ACT_NONE || a->row_map) return VEX_E_UNSUPPORTED;
  if (a->K % 64 != 0 || a->lda % 16 != 0 || a->ldw % 16 != 0 || a->N % 16 != 0) return VEX_E_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(a->a) | reinterpret_cast<uintptr_t>(a->w[0][0]) | reinterpret_cast<uintptr_t>(a->w[0][1])) & 31)
    return VEX_E_UNSUPPORTED;  // 256-bit loads
  const int mode = a->mode;
  if (mode != VEX_EPI_PLAIN && mode != VEX_EPI_RESIDUAL && mode != VEX_EPI_SWIGLU && mode != VEX_EPI_ROPE)
    return VEX_E_UNSUPPORTED;
  DecGemm p;
  std::memset(&p, 0, sizeof(p));
  p.x = static_cast<const __nv_bfloat16*>(a->a);
  p.w[0] = static_cast<const __nv_bfloat16*>(a->w[0][0]);
  p.w[1] = static_cast<const __nv_bfloat16*>(a->w[0][1]);
  p.out = static_cast<__nv_bfloat16*>(a->out);
  p.ldx = a->lda;
  p.ldw = a->ldw;
  p.ldo = a->ldo;
  p.B = a->rows_cap;
  p.N = a->N;
  p.K = a->K;
  p.mode = mode;
  p.alpha = a->alpha;
  if (mode == VEX_EPI_SWIGLU && !p.w[1]) return VEX_E_INVALID;
  if (mode == VEX_EPI_RESIDUAL) {
    if (!a->residual) return VEX_E_INVALID;
    p.residual = static_cast<const __nv_bfloat16*>(a->residual);
  }
  if (a->lora_r > 0 && a->lora_b[0][0]) {
    if (a->lora_r % 64 != 0 || !a->lora_t[0] || a->ldt % 16 != 0) return VEX_E_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(a->lora_t[0]) | reinterpret_cast<uintptr_t>(a->lora_t[1]) |
         reinterpret_cast<uintptr_t>(a->lora_b[0][0]) | reinterpret_cast<uintptr_t>(a->lora_b[0][1])) & 31)
      return VEX_E_UNSUPPORTED;
    if (mode == VEX_EPI_SWIGLU && (!a->lora_b[0][1] || !a->lora_t[1])) return VEX_E_UNSUPPORTED;
    p.lora_r = a->lora_r;
    p.ldt = a->ldt;
    p.lt[0] = static_cast<const __nv_bfloat16*>(a->lora_t[0]);
    p.lt[1] = static_cast<const __nv_bfloat16*>(a->lora_t[1]);
    p.lb[0] = static_cast<const __nv_bfloat16*>(a->lora_b[0][0]);
    p.lb[1] = static_cast<const __nv_bfloat16*>(a->lora_b[0][1]);
  }
  int tasks = a->N / 16;
  if (mode == VEX_EPI_ROPE) {
    if (!a->rope_cos || !a->rope_sin || !a->position_ids || a->rope_len <= 0) return VEX_E_INVALID;
    if (a->rope_cols % 128 != 0 || a->N % 128 != 0) return VEX_E_UNSUPPORTED;
    p.rope_cos = static_cast<const __nv_bfloat16*>(a->rope_cos);
    p.rope_sin = static_cast<const __nv_bfloat16*>(a->rope_sin);
    p.position_ids = a->position_ids;   // indexed by batch row (sorted_to_flat is the identity in a decode step)
    p.rope_len = a->rope_len;
    p.rope_cols = a->rope_cols;
    if (a->kv_k || a->kv_v) {
      if (!a->kv_k || !a->kv_v || a->kv_seq_len != 1 || a->kv_capacity <= 0) return VEX_E_INVALID;
      if (a->N != 3 * (a->rope_cols / 2)) return VEX_E_UNSUPPORTED;
      p.kv_k = static_cast<__nv_bfloat16*>(a->kv_k);
      p.kv_v = static_cast<__nv_bfloat16*>(a->kv_v);
      p.kv_pos = a->kv_pos;
      p.kv_cap = a->kv_capacity;
    }
    tasks = a->N / 32;  // two tiles (j, j + 64) per task
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int nb = (a->rows_cap + 7) / 8;
  const bool two = mode == VEX_EPI_SWIGLU || mode == VEX_EPI_ROPE;
#define VEX_DG_CASE(NBV)                                                                       \
  case NBV:                                                                                    \
    return two ? launch_dg<NBV, 2>(p, tasks, s) : launch_dg<NBV, 1>(p, tasks, s);
  switch (nb) {
    VEX_DG_CASE(1) VEX_DG_CASE(2) VEX_DG_CASE(3) VEX_DG_CASE(4)
    default:
      return VEX_E_UNSUPPORTED;
  }
#undef VEX_DG_CASE
}
