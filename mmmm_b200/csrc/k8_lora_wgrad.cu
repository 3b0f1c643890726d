// K8 -- LoRA weight gradients of the routed Linears on tcgen05: per expert e (row segment of the sorted order)
//
//     OUT_e[f, j] += sum_{t in segment e}  X[t, f] * Y[t, j]          f < F (128-wide tiles), j < r (<= 64)
//
// which is both adapter gradients of  y = x W^T + s (x A^T) B^T  (PEFT lora.Linear on top of
// modeling_cogvlm.py:244-245, :278-279, :96-97; wiring scripts/cli.py:82-88):
//     dB[out, r] = dy^T . T          X = dy  [rows, out],  Y = T  = s x A^T [rows, r]   (kept by the forward)
//     dA[r, in]  = dT^T . x          X = x   [rows, in ],  Y = dT = s dy B  [rows, r]   (from the dgrad GEMM),
//                                    stored transposed (transpose_out = 1)
// The reduction runs over TOKENS, so both operands are MN-major (the contiguous dimension is f resp. j, not the
// reduction index): X tiles arrive as two {64 f, 64 t} TMA boxes (A operand, M = 128), Y as one {64 j, 64 t} box
// (B operand, N = 64); D[128, 64] fp32 lives in 64 TMEM columns.  HBM-bound (X is read once: 2 F bytes per
// token); the grid is (F / 128) x token-chunks of 1024 so every SM streams, partial sums are combined with fp32
// atomics into the caller-zeroed gradient buffers.  Rows past the segment end belong to the other expert or are
// uninitialised, so the last k-block of a chunk zeroes them in shared memory before the MMA reads it.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace vex {

int num_sms();
int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);

constexpr int WG_CHUNK = 1024;  // tokens per CTA
constexpr int WG_BK = 64;       // tokens per k-block
constexpr int WG_STAGES = 4;
constexpr int WG_A_BYTES = 128 * WG_BK * 2;  // 16 KB: two 8 KB chunks of 64 features
constexpr int WG_B_BYTES = 64 * WG_BK * 2;   // 8 KB
constexpr int WG_STAGE_BYTES = WG_A_BYTES + WG_B_BYTES;
constexpr int WG_THREADS = 192;
constexpr int WG_SMEM = WG_STAGES * WG_STAGE_BYTES + 256 + 1024;

struct WgTmaps {
  CUtensorMap x, y;
};

__global__ void __launch_bounds__(WG_THREADS, 1)
    k8_lora_wgrad(const __grid_constant__ WgTmaps tm, const int32_t* __restrict__ counts, float* __restrict__ out0,
                  float* __restrict__ out1, int64_t ldo, int transpose_out, int F, int r, int rows_cap) {
  // ---- which (expert, token chunk) is this CTA's? ----
  const int cnt0 = min(max(counts[0], 0), rows_cap);
  const int cnt1 = min(max(counts[1], 0), rows_cap - cnt0);
  // gridDim.y token chunks are dealt to the two expert segments in proportion to their rows (an expert without an adapter
  // gets none), each segment is cut into equal chunks of whole k-blocks.  The host sizes gridDim.y so that the whole
  // grid is resident at once (2 CTAs per SM): with fixed 1 024-token chunks the F = 4096 launches ran 416 CTAs in 1.4
  // waves, and the partial wave of an HBM-bound launch streams at a fraction of the bandwidth (the decode kernels'
  // lesson, profiles/r2_decode_experiments.md).
  const int G = gridDim.y;
  const int w0 = out0 ? cnt0 : 0, w1 = out1 ? cnt1 : 0;
  int n0;
  if (w0 == 0) n0 = 0;
  else if (w1 == 0) n0 = G;
  else n0 = min(G - 1, max(1, static_cast<int>((static_cast<int64_t>(w0) * G + (w0 + w1) / 2) / (w0 + w1))));
  const int n1 = G - n0;
  int c = blockIdx.y, e = 0, lo, hi;
  if (c < n0) {
    const int ch = ((cnt0 + n0 - 1) / n0 + WG_BK - 1) / WG_BK * WG_BK;
    if (c * ch >= cnt0) return;
    lo = c * ch;
    hi = min(cnt0, lo + ch);
  } else {
    c -= n0;
    if (n1 == 0 || w1 == 0) return;
    const int ch = ((cnt1 + n1 - 1) / n1 + WG_BK - 1) / WG_BK * WG_BK;
    if (c * ch >= cnt1) return;
    e = 1;
    lo = cnt0 + c * ch;
    hi = min(cnt0 + cnt1, lo + ch);
  }
  float* out = e ? out1 : out0;
  if (out == nullptr) return;  // this expert has no adapter
  const int n_kb = (hi - lo + WG_BK - 1) / WG_BK;
  const int f0 = blockIdx.x * 128;

  extern __shared__ uint8_t wg_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(wg_smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WG_STAGES * WG_STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + WG_STAGES;
  uint64_t* acc_full = bars + 2 * WG_STAGES;
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(bars + 2 * WG_STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < WG_STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_base_smem, 64);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm.x);
    tma_prefetch_desc(&tm.y);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_base_smem;

  if (warp == 0) {
    if (lane == 0) {
      // =============================== TMA producer ===============================
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < n_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * WG_STAGE_BYTES;
        const int t0 = lo + kb * WG_BK;
        mbar_arrive_expect_tx(&full_bar[stage], WG_STAGE_BYTES);
        tma_load_2d(sa, &tm.x, &full_bar[stage], f0, t0);
        tma_load_2d(sa + 8192, &tm.x, &full_bar[stage], f0 + 64, t0);
        tma_load_2d(sa + WG_A_BYTES, &tm.y, &full_bar[stage], 0, t0);
        if (++stage == WG_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (whole warp: boundary fix-up) ===============================
    constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 1, 1);  // A and B MN-major
    const uint64_t dA0 = umma_desc_mnmajor_sw128(smem_u32(smem), 8192, 1024);  // stage 0; k-steps / stages are adds
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < n_kb; ++kb) {
      mbar_wait(&full_bar[stage], phase);
      uint8_t* sa = smem + stage * WG_STAGE_BYTES;
      const int valid = hi - (lo + kb * WG_BK);  // token rows of this k-block inside the segment
      if (valid < WG_BK) {
        // token row t of a box is the 128-byte line t of each 8 KB chunk (the swizzle permutes inside the line)
        for (int i = lane; i < (WG_BK - valid) * 8 * 3; i += 32) {
          const int chunk = i / ((WG_BK - valid) * 8), rem = i % ((WG_BK - valid) * 8);
          const int row = valid + rem / 8, piece = rem % 8;
          *reinterpret_cast<uint4*>(sa + chunk * 8192 + row * 128 + piece * 16) = make_uint4(0, 0, 0, 0);
        }
        fence_proxy_async_smem();
      }
      __syncwarp();
      tc_fence_after();
      // elected lane of the converged warp: operands stay in uniform registers (a `lane == 0` branch costs an elect /
      // waterfall loop, four R2UR and two descriptor rebuilds per 32-cycle MMA)
      const uint64_t da = dA0 + static_cast<uint64_t>(stage * (WG_STAGE_BYTES >> 4)), db = da + (WG_A_BYTES >> 4);
      if (elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < WG_BK / 16; ++k)
          umma_ss(tmem, da + static_cast<uint64_t>((k * 2048) >> 4), db + static_cast<uint64_t>((k * 2048) >> 4), idesc,
                  (kb > 0) || (k > 0));
        umma_commit(&empty_bar[stage]);
        if (kb == n_kb - 1) umma_commit(acc_full);
      }
      __syncwarp();
      if (++stage == WG_STAGES) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else {
    // =============================== epilogue: TMEM -> fp32 atomics ===============================
    // A thread holds one feature row f (TMEM lane) x 32 rank columns.  out[r, F] (transpose_out): lanes = consecutive f
    // -> every atomic instruction covers one contiguous 128-byte line.  out[F, r]: the same mapping would touch 32 rows
    // 256 bytes apart per instruction (32 sectors: the dB launches ran 1.5 x slower than the dA launches of the same
    // size, ncu round 2), so the 32 x 32 block is transposed through shared memory (the pipeline stages are dead once
    // acc_full has fired) and written with lanes along the rank.
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    const int f = f0 + q * 32 + lane;
    float* tile = reinterpret_cast<float*>(smem) + (warp - 2) * (32 * 33);  // [32 f][33]: conflict-free both ways
    mbar_wait(acc_full, 0);
    tc_fence_after();
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem + (static_cast<uint32_t>(q * 32) << 16) + half * 32, v);
      tmem_ld_wait();
      if (transpose_out) {
        if (f < F) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = half * 32 + j;
            if (col < r) atomicAdd(out + static_cast<int64_t>(col) * ldo + f, __uint_as_float(v[j]));
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) tile[lane * 33 + j] = __uint_as_float(v[j]);
        __syncwarp();
        const int col = half * 32 + lane;
#pragma unroll 4
        for (int ff = 0; ff < 32; ++ff) {
          const int frow = f0 + q * 32 + ff;
          if (frow < F && col < r) atomicAdd(out + static_cast<int64_t>(frow) * ldo + col, tile[ff * 33 + lane]);
        }
        __syncwarp();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 64);
  }
}

}  // namespace vex

extern "C" int vex_lora_wgrad(const void* x, int64_t ldx, const void* y, int64_t ldy, int r, float* out_vision,
                              float* out_language, int64_t ldo, int transpose_out, const int32_t* counts, int rows_cap,
                              int F, vexStream stream) {
  using namespace vex;
  if (!x || !y || !counts || (!out_vision && !out_language) || rows_cap <= 0 || F <= 0) return VEX_E_INVALID;
  if (r <= 0 || r > 64 || r % 8 != 0 || F % 8 != 0 || ldx % 8 != 0 || ldy % 8 != 0) return VEX_E_UNSUPPORTED;
  static bool configured = false;
  if (!configured) {
    VEX_CUDA_TRY(cudaFuncSetAttribute(k8_lora_wgrad, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM));
    configured = true;
  }
  WgTmaps tm;
  std::memset(&tm, 0, sizeof(tm));
  int rc;
  if ((rc = make_tmap_2d(&tm.x, x, rows_cap, F, ldx, WG_BK)) != VEX_OK) return rc;
  if ((rc = make_tmap_2d(&tm.y, y, rows_cap, r, ldy, WG_BK)) != VEX_OK) return rc;
  // token chunks: as many as keep the whole grid resident (2 CTAs per SM), at least 2 (one per expert segment), no
  // smaller than one k-block; VEX_K8_CHUNKS overrides (tuning)
  static const int forced_chunks = [] {
    const char* e = std::getenv("VEX_K8_CHUNKS");
    return e ? std::atoi(e) : 0;
  }();
  const int f_tiles = ceil_div(F, 128);
  int chunks = std::max(2, (2 * num_sms()) / f_tiles);
  chunks = std::min(chunks, ceil_div(rows_cap, WG_BK) + 1);
  if (forced_chunks > 0) chunks = std::max(2, forced_chunks);
  dim3 grid(f_tiles, chunks);
  k8_lora_wgrad<<<grid, WG_THREADS, WG_SMEM, static_cast<cudaStream_t>(stream)>>>(
      tm, counts, out_vision, out_language, ldo, transpose_out, F, r, rows_cap);
  VEX_LAUNCH_CHECK();
  return VEX_OK;
}
