// K1 -- token-type partition and compaction (integer work, bit-exact against the reference mask split).
//
// Reference behaviour restated (not copied): get_expert_mask, modeling_cogvlm.py:58-70, and the
// ascending flat order produced by ATen boolean-mask indexing (x[mask] == index by nonzero(mask)).
// Two launches, one CTA per sample, no host round trip: counts stay on the device.
//   k1_count : per-sample (vision, language, valid) counts                 -> scratch[4*b + {0,1,2}]
//   k1_emit  : exclusive prefix over samples + in-sample ballot scan       -> index lists, cu_seqlens, counts
#include "common.cuh"

namespace vex {

constexpr int K1_THREADS = 256;
constexpr int K1_WARPS = K1_THREADS / 32;

struct Flags {
  bool vis, lang, valid;
};

__device__ __forceinline__ Flags classify(const int64_t* __restrict__ tt, const uint8_t* __restrict__ pm, int L,
                                          int l) {
  Flags f;
  // a token is vision iff it AND its right neighbour carry VISION_TOKEN_TYPE (1); the last column never is
  const bool vis_raw = (l < L - 1) && (tt[l] == 1) && (tt[l + 1] == 1);
  f.valid = pm[l] != 0;
  f.vis = vis_raw && f.valid;
  f.lang = !vis_raw && f.valid;
  return f;
}

__global__ void __launch_bounds__(K1_THREADS) k1_count(const int64_t* __restrict__ token_type_ids,
                                                       const uint8_t* __restrict__ padding_mask, int L,
                                                       int32_t* __restrict__ scratch) {
  const int b = blockIdx.x;
  const int64_t* tt = token_type_ids + static_cast<int64_t>(b) * L;
  const uint8_t* pm = padding_mask + static_cast<int64_t>(b) * L;
  int nv = 0, nl = 0;
  for (int l = threadIdx.x; l < L; l += K1_THREADS) {
    Flags f = classify(tt, pm, L, l);
    nv += f.vis;
    nl += f.lang;
  }
  __shared__ int s_v[K1_WARPS], s_l[K1_WARPS];
  for (int o = 16; o > 0; o >>= 1) {
    nv += __shfl_xor_sync(0xffffffffu, nv, o);
    nl += __shfl_xor_sync(0xffffffffu, nl, o);
  }
  if (lane_id() == 0) {
    s_v[threadIdx.x >> 5] = nv;
    s_l[threadIdx.x >> 5] = nl;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int tv = 0, tl = 0;
    for (int w = 0; w < K1_WARPS; ++w) {
      tv += s_v[w];
      tl += s_l[w];
    }
    scratch[4 * b + 0] = tv;
    scratch[4 * b + 1] = tl;
    scratch[4 * b + 2] = tv + tl;
  }
}

__global__ void __launch_bounds__(K1_THREADS)
    k1_emit(const int64_t* __restrict__ token_type_ids, const uint8_t* __restrict__ padding_mask, int B, int L,
            const int32_t* __restrict__ scratch, int32_t* __restrict__ sorted_to_flat,
            int32_t* __restrict__ flat_to_sorted, int32_t* __restrict__ sorted_to_token,
            int32_t* __restrict__ token_to_sorted, int32_t* __restrict__ token_to_flat,
            int32_t* __restrict__ cu_seqlens, int32_t* __restrict__ counts) {
  const int b = blockIdx.x;
  __shared__ int s_base[4];  // vision base, language base, token base of this sample; [3] unused
  __shared__ int s_tot[4];   // Tv, Tl, T, max len
  __shared__ int s_wv[K1_WARPS], s_wl[K1_WARPS];

  // ---- exclusive prefix over samples (B is small: one warp strides over it) ----
  if (threadIdx.x < 32) {
    int pv = 0, pl = 0, tv = 0, tl = 0, mx = 0;
    for (int i = threadIdx.x; i < B; i += 32) {
      const int v = scratch[4 * i + 0], l = scratch[4 * i + 1];
      if (i < b) {
        pv += v;
        pl += l;
      }
      tv += v;
      tl += l;
      mx = max(mx, v + l);
    }
    for (int o = 16; o > 0; o >>= 1) {
      pv += __shfl_xor_sync(0xffffffffu, pv, o);
      pl += __shfl_xor_sync(0xffffffffu, pl, o);
      tv += __shfl_xor_sync(0xffffffffu, tv, o);
      tl += __shfl_xor_sync(0xffffffffu, tl, o);
      mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (threadIdx.x == 0) {
      s_base[0] = pv;
      s_base[1] = pl;
      s_base[2] = pv + pl;
      s_tot[0] = tv;
      s_tot[1] = tl;
      s_tot[2] = tv + tl;
      s_tot[3] = mx;
    }
  }
  __syncthreads();
  const int Tv = s_tot[0], T = s_tot[2];
  int run_v = s_base[0];        // next vision slot (sorted row)
  int run_l = Tv + s_base[1];   // next language slot (sorted row)
  int run_t = s_base[2];        // next token rank
  const int tok_begin = run_t;
  if (threadIdx.x == 0) {
    const int len_b = scratch[4 * b + 2];
    cu_seqlens[b + 1] = tok_begin + len_b;
    if (b == 0) {
      cu_seqlens[0] = 0;
      counts[VEX_COUNT_VISION] = s_tot[0];
      counts[VEX_COUNT_LANGUAGE] = s_tot[1];
      counts[VEX_COUNT_VALID] = s_tot[2];
      counts[VEX_COUNT_MAXLEN] = s_tot[3];
    }
  }

  const int64_t* tt = token_type_ids + static_cast<int64_t>(b) * L;
  const uint8_t* pm = padding_mask + static_cast<int64_t>(b) * L;
  const int warp = threadIdx.x >> 5, lane = lane_id();
  const uint32_t lt_mask = (1u << lane) - 1u;

  for (int l0 = 0; l0 < L; l0 += K1_THREADS) {
    const int l = l0 + threadIdx.x;
    Flags f = {false, false, false};
    if (l < L) f = classify(tt, pm, L, l);
    const uint32_t bv = __ballot_sync(0xffffffffu, f.vis);
    const uint32_t bl = __ballot_sync(0xffffffffu, f.lang);
    if (lane == 0) {
      s_wv[warp] = __popc(bv);
      s_wl[warp] = __popc(bl);
    }
    __syncthreads();
    int off_v = 0, off_l = 0, sum_v = 0, sum_l = 0;
#pragma unroll
    for (int w = 0; w < K1_WARPS; ++w) {
      const int cv = s_wv[w], cl = s_wl[w];
      if (w < warp) {
        off_v += cv;
        off_l += cl;
      }
      sum_v += cv;
      sum_l += cl;
    }
    if (l < L) {
      const int flat = b * L + l;
      if (f.valid) {
        const int rank_v = off_v + __popc(bv & lt_mask);
        const int rank_l = off_l + __popc(bl & lt_mask);
        const int s = f.vis ? (run_v + rank_v) : (run_l + rank_l);
        const int t = run_t + rank_v + rank_l;  // valid == vis | lang, so the token rank is the sum
        sorted_to_flat[s] = flat;
        sorted_to_token[s] = t;
        token_to_sorted[t] = s;
        token_to_flat[t] = flat;
        flat_to_sorted[flat] = s;
      } else {
        flat_to_sorted[flat] = -1;
      }
    }
    run_v += sum_v;
    run_l += sum_l;
    run_t += sum_v + sum_l;
    __syncthreads();
  }

  // ---- entries past the counts: -1.  Sample b owns the slots of its own padded positions ----
  const int len_b = run_t - tok_begin;
  const int pad_b = L - len_b;
  const int pad_begin = T + (b * L - tok_begin);  // T + number of padded positions in earlier samples
  for (int j = threadIdx.x; j < pad_b; j += K1_THREADS) {
    const int i = pad_begin + j;
    sorted_to_flat[i] = -1;
    sorted_to_token[i] = -1;
    token_to_sorted[i] = -1;
    token_to_flat[i] = -1;
  }
}

}  // namespace vex

extern "C" int vex_partition(const int64_t* token_type_ids, const uint8_t* padding_mask, int B, int L,
                             int32_t* sorted_to_flat, int32_t* flat_to_sorted, int32_t* sorted_to_token,
                             int32_t* token_to_sorted, int32_t* token_to_flat, int32_t* cu_seqlens,
                             int32_t* counts, int32_t* scratch, vexStream stream) {
  if (!token_type_ids || !padding_mask || !sorted_to_flat || !flat_to_sorted || !sorted_to_token ||
      !token_to_sorted || !token_to_flat || !cu_seqlens || !counts || !scratch)
    return VEX_E_INVALID;
  if (B <= 0 || L <= 0) return VEX_E_INVALID;
  if (L == 1) return VEX_E_UNSUPPORTED;  // decode rule (modeling_cogvlm.py:67) is out of scope
  if (static_cast<int64_t>(B) * L > (1ll << 30)) return VEX_E_UNSUPPORTED;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  vex::k1_count<<<B, vex::K1_THREADS, 0, s>>>(token_type_ids, padding_mask, L, scratch);
  VEX_LAUNCH_CHECK();
  vex::k1_emit<<<B, vex::K1_THREADS, 0, s>>>(token_type_ids, padding_mask, B, L, scratch, sorted_to_flat,
                                             flat_to_sorted, sorted_to_token, token_to_sorted, token_to_flat,
                                             cu_seqlens, counts);
  VEX_LAUNCH_CHECK();
  return VEX_OK;
}
