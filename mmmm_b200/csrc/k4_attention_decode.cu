// K4d -- decode-step attention (q_len == 1 against the KV cache), memory-bound: every cached key and value is
// read once (2 * L * 128 * 2 bytes per (sample, head)).
//
// Restates the generation branch of attention_fn (modeling_cogvlm.py:129-141) with its eager-bf16 rounding
// points: the query is scaled in bf16 (`query_layer *= d ** -0.5`), scores are a bf16 einsum output, masked
// positions become -inf, softmax runs in fp32 and is cast back to bf16 before the weighted sum over the values.
// Cache layout is the reference's: k, v [B, heads, capacity, 128] (what prefill returns and torch.cat extends, :258-262).
//
// Split-KV over a THREAD-BLOCK CLUSTER: the positions of one (sample, head) are split over the 1 / 2 / 4 / 8 CTAs of a
// cluster (grid z), so that a small batch still fills the machine (8 samples x 32 heads are only 256 CTAs; round 2
// measured 1.3 TB/s with one CTA per (sample, head)).  The softmax stays EXACT -- global max and sum are exchanged
// through distributed shared memory before any probability is rounded to bf16 -- and the partial outputs are reduced
// by the cluster's rank-0 CTA through DSMEM reads: no workspace in HBM, no second kernel.
// Per CTA (256 threads): both passes read the cache with LDG.256, 8 lanes per 256-byte row (two full 128-byte lines),
// 4 rows per warp instruction; a lane owns 16 of the 128 dims (scores: shuffle-reduced inside the 8-lane group; values:
// the 4 groups of a warp are folded with shuffles, the 8 warps through shared memory).
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace vex {

int num_sms();

constexpr int DEC_THREADS = 256;
constexpr int DEC_WARPS = DEC_THREADS / 32;

__device__ __forceinline__ float ld_dsmem_f32(const float* local, uint32_t rank) {
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(mapa_shared(smem_u32(local), rank)));
  return v;
}

__device__ __forceinline__ float block_reduce_max(float v, float* red) {
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  if (lane_id() == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int i = 1; i < DEC_WARPS; ++i) r = fmaxf(r, red[i]);
  __syncthreads();
  return r;
}
__device__ __forceinline__ float block_reduce_sum(float v, float* red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane_id() == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < DEC_WARPS; ++i) r += red[i];
  __syncthreads();
  return r;
}

// `kv_len` (device, may be null): number of cache positions already filled BEFORE this step; the step attends to
// *kv_len + 1 positions (the current token's K / V were appended by the QKV epilogue).  Null = attend to all L.
// Cache rows of one (sample, head) are `cap` positions apart (cap >= L: pre-allocated headroom); the mask row
// stride is ld_mask.  `chunk_cap` = ceil(L_host / nsplit): the size of the dynamic score buffer.
__global__ void __launch_bounds__(DEC_THREADS)
    k4_attention_decode(const __nv_bfloat16* __restrict__ q, int64_t ldq, const __nv_bfloat16* __restrict__ k,
                        const __nv_bfloat16* __restrict__ v, const uint8_t* __restrict__ mask, int64_t ld_mask,
                        __nv_bfloat16* __restrict__ out, int heads, int L_host, int cap,
                        const int32_t* __restrict__ kv_len, float scale, int chunk_cap) {
  extern __shared__ float sc[];  // this CTA's scores, then exp(score - max)
  __shared__ float qs[128];
  __shared__ float red[DEC_WARPS];
  __shared__ float xch[2];                  // local max, local sum: read by the peers through DSMEM
  __shared__ float opart[128];              // this CTA's partial output: read by rank 0 through DSMEM
  __shared__ float osum[DEC_WARPS][128];
  // programmatic dependent launch: q and the newest cache row come from the QKV kernel right before this one
  pdl_trigger();
  pdl_wait();
  const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const uint32_t rank = blockIdx.z, nsplit = gridDim.z;  // cluster = the z extent of the grid
  const int L = kv_len ? min(max(kv_len[0] + 1, 1), L_host) : L_host;
  const int chunk = min((L + static_cast<int>(nsplit) - 1) / static_cast<int>(nsplit), chunk_cap);
  const int l0 = min(static_cast<int>(rank) * chunk, L), l1 = min(l0 + chunk, L);
  if (tid < 128) qs[tid] = bf16r(__bfloat162float(q[static_cast<int64_t>(b) * ldq + h * 128 + tid]) * scale);
  __syncthreads();
  const int64_t base = (static_cast<int64_t>(b) * heads + h) * cap;
  const uint8_t* mrow = mask + static_cast<int64_t>(b) * ld_mask;

  // scores: 8 lanes per key row (LDG.256 each: a 256-byte row = two full 128-byte lines per 8 lanes, 4 rows per warp
  // instruction), 16 dims per lane, 3-step shuffle reduction inside the lane group
  const int warp = tid >> 5, lane = tid & 31;
  const int grp = lane >> 3, sub = lane & 7;
  float qreg[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) qreg[j] = qs[sub * 16 + j];
  float mx = -INFINITY;
  constexpr int SU = 4;  // key rows in flight per lane group
  for (int lb = l0 + warp * 4; lb < l1; lb += DEC_WARPS * 4 * SU) {
    U32x8 u[SU];
    bool in[SU], live[SU];
#pragma unroll
    for (int i = 0; i < SU; ++i) {
      const int l = lb + i * DEC_WARPS * 4 + grp;
      in[i] = l < l1;
      live[i] = in[i] && mrow[l];
      if (live[i]) u[i] = ld_coherent_stream_256(k + (base + l) * 128 + sub * 16);
    }
#pragma unroll
    for (int i = 0; i < SU; ++i) {
      const int l = lb + i * DEC_WARPS * 4 + grp;
      float acc = 0.f;
      if (live[i]) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc = fmaf(qreg[2 * j], bf16_lo(u[i].v[j]), acc);
          acc = fmaf(qreg[2 * j + 1], bf16_hi(u[i].v[j]), acc);
        }
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      acc += __shfl_xor_sync(0xffffffffu, acc, 4);
      const float sv = live[i] ? bf16r(acc) : -INFINITY;
      if (in[i] && sub == 0) sc[l - l0] = sv;
      if (in[i]) mx = fmaxf(mx, sv);
    }
  }
  mx = block_reduce_max(mx, red);
  if (nsplit > 1) {  // exact softmax across the cluster: exchange the local maxima
    if (tid == 0) xch[0] = mx;
    cluster_sync_all();
    for (uint32_t r = 0; r < nsplit; ++r) mx = fmaxf(mx, ld_dsmem_f32(&xch[0], r));
  }
  float sum = 0.f;
  for (int l = l0 + tid; l < l1; l += DEC_THREADS) {
    const float e = __expf(sc[l - l0] - mx);
    sc[l - l0] = e;
    sum += e;
  }
  sum = block_reduce_sum(sum, red);  // also orders the sc[] writes before the reads below
  if (nsplit > 1) {
    if (tid == 0) xch[1] = sum;
    cluster_sync_all();
    sum = 0.f;
    for (uint32_t r = 0; r < nsplit; ++r) sum += ld_dsmem_f32(&xch[1], r);
  }
  const float inv = 1.0f / sum;

  // values: same access shape -- lane group `grp` takes keys lb + grp, a lane owns 16 of the 128 output dims
  float a[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) a[j] = 0.f;
  for (int lb = l0 + warp * 4; lb < l1; lb += DEC_WARPS * 4 * SU) {
    U32x8 u[SU];
    float pr[SU];
#pragma unroll
    for (int i = 0; i < SU; ++i) {
      const int l = lb + i * DEC_WARPS * 4 + grp;
      pr[i] = l < l1 ? bf16r(sc[l - l0] * inv) : 0.f;
      // masked keys have p == 0 (their values are zeroed in the reference, :135): their rows are not read
      if (pr[i] != 0.f) u[i] = ld_coherent_stream_256(v + (base + l) * 128 + sub * 16);
    }
#pragma unroll
    for (int i = 0; i < SU; ++i) {
      if (pr[i] != 0.f) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          a[2 * j] = fmaf(pr[i], bf16_lo(u[i].v[j]), a[2 * j]);
          a[2 * j + 1] = fmaf(pr[i], bf16_hi(u[i].v[j]), a[2 * j + 1]);
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) {  // fold the 4 lane groups of the warp
    a[j] += __shfl_xor_sync(0xffffffffu, a[j], 8);
    a[j] += __shfl_xor_sync(0xffffffffu, a[j], 16);
  }
  if (grp == 0) {
#pragma unroll
    for (int j = 0; j < 16; ++j) osum[warp][sub * 16 + j] = a[j];
  }
  __syncthreads();
  float o = 0.f;
  if (tid < 128) {
#pragma unroll
    for (int w = 0; w < DEC_WARPS; ++w) o += osum[w][tid];
  }
  if (nsplit > 1) {
    if (tid < 128) opart[tid] = o;
    cluster_sync_all();
    if (rank == 0 && tid < 128) {
      o = 0.f;
      for (uint32_t r = 0; r < nsplit; ++r) o += ld_dsmem_f32(&opart[tid], r);
    }
  }
  if (rank == 0 && tid < 128) out[static_cast<int64_t>(b) * heads * 128 + h * 128 + tid] = __float2bfloat16_rn(o);
  if (nsplit > 1) cluster_sync_all();  // the peers' shared memory stays alive until rank 0 has read it
}

__global__ void k4_advance_counter(int32_t* p, int by) { if (threadIdx.x == 0) p[0] += by; }

// scores live in dynamic shared memory: opt in beyond the 48 KB default once, bound by the 227 KB per-CTA limit
constexpr int DEC_MAX_L = (200 * 1024) / 4;

static int launch_decode(const void* q, int64_t ldq, const void* k, const void* v, const uint8_t* mask,
                         int64_t ld_mask, void* out, int B, int heads, int L, int cap, const int32_t* kv_len,
                         float scale, cudaStream_t s) {
  if (!q || !k || !v || !mask || !out || B <= 0 || heads <= 0 || L <= 0 || cap < L) return VEX_E_INVALID;
  if (B > 65535) return VEX_E_UNSUPPORTED;
  // KV splits (= cluster size, a power of two <= 8): as many as keep the WHOLE grid resident at once (3 CTAs per SM at
  // 78 registers), chunks of at least 64 positions.  A grid that needs a second wave loses more than the split gains:
  // the last, partial wave streams with too few CTAs to keep DRAM busy (batch 8 x 32 heads, ~1 500 positions: 1 024
  // CTAs in 2.3 waves 4.26 ms per decode step, 256 CTAs in one wave 4.07 ms; profiles/r2_decode_experiments.md).
  // Very long caches are split anyway so that the score buffer stays small enough for 3 CTAs per SM.
  int nsplit = 1;
  const int64_t resident = 3 * static_cast<int64_t>(num_sms());
  while (nsplit < 8 && static_cast<int64_t>(B) * heads * nsplit * 2 <= resident && L / (2 * nsplit) >= 64) nsplit *= 2;
  while (nsplit < 8 && L / nsplit > 16384) nsplit *= 2;
  {
    static const int forced = [] {  // VEX_K4D_NSPLIT: tuning override (1, 2, 4 or 8)
      const char* e = std::getenv("VEX_K4D_NSPLIT");
      return e ? std::atoi(e) : 0;
    }();
    if (forced == 1 || forced == 2 || forced == 4 || forced == 8) nsplit = forced;
  }
  const int chunk_cap = ceil_div(L, nsplit);
  if (chunk_cap > DEC_MAX_L) return VEX_E_UNSUPPORTED;
  static bool configured = false;
  if (!configured) {
    VEX_CUDA_TRY(cudaFuncSetAttribute(k4_attention_decode, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      DEC_MAX_L * 4));
    configured = true;
  }
  cudaLaunchConfig_t cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(heads, B, nsplit);
  cfg.blockDim = dim3(DEC_THREADS);
  cfg.dynamicSmemBytes = static_cast<size_t>(chunk_cap) * sizeof(float);
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = nsplit;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  VEX_CUDA_TRY(cudaLaunchKernelEx(&cfg, k4_attention_decode, static_cast<const __nv_bfloat16*>(q), ldq,
                                  static_cast<const __nv_bfloat16*>(k), static_cast<const __nv_bfloat16*>(v), mask,
                                  ld_mask, static_cast<__nv_bfloat16*>(out), heads, L, cap, kv_len, scale, chunk_cap));
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
