/*
 * vex.h -- C ABI of libvex.so: the B200 (sm_100a) kernels behind the visual-expert decoder layer.
 *
 * The reference (function2-llx/MMMM) has no native code and no FFI on this path: the layer is eager
 * PyTorch in mmmm/models/cogvlm/modeling_cogvlm.py:30-340 and its "operator interface" is the
 * torch.nn.Module surface of CogVLMDecoderLayer.forward (:295-340).  This header therefore defines the
 * boundary a maintainer binds instead of the ATen calls; every entry point names the reference lines
 * it replaces.  The Python mirror of the reference interface (mmmm_b200/modeling_cogvlm.py) binds these
 * symbols with ctypes and registers them as torch.library custom ops.
 *
 * Conventions
 *   - plain C, raw DEVICE pointers and sizes, no torch types; all activations are bf16 (uint16_t bits)
 *   - every call is asynchronous on `stream` (a cudaStream_t), never synchronises, never allocates or
 *     frees caller memory; outputs are pre-allocated by the caller
 *   - token counts live ON THE DEVICE (`counts`, written by vex_partition); grids are sized from the
 *     host-known upper bound rows_cap = B*L, so the path has no host sync and is CUDA-graph capturable
 *   - return 0 on success, a negative VEX_E_* code otherwise (vex_error_string); CUDA launch errors are
 *     reported as VEX_E_CUDA with the cudaError_t available from vex_last_cuda_error()
 *
 * Row orders:
 *   flat   : b*L + l, the reference's [B, L] layout
 *   token  : rank among rows with padding_mask == True  (== order of hidden_states[padding_mask])
 *   sorted : vision-expert rows first (ascending flat order == order of x[vision_token_mask]),
 *            then language-expert rows (order of x[language_token_mask]); no padding between them
 */
#ifndef VEX_H_
#define VEX_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VEX_ABI_VERSION 7

#define VEX_OK 0
#define VEX_E_INVALID (-1)     /* bad argument (null pointer, size, alignment) */
#define VEX_E_UNSUPPORTED (-2) /* shape outside what the kernels implement */
#define VEX_E_CUDA (-3)        /* CUDA runtime / driver error, see vex_last_cuda_error */
#define VEX_E_NO_DEVICE (-4)   /* not an sm_100 device */

typedef void* vexStream; /* cudaStream_t */

int vex_abi_version(void);
const char* vex_error_string(int code);
int vex_last_cuda_error(void);
/* 0 if the current device can run the kernels (compute capability 10.x), VEX_E_NO_DEVICE otherwise */
int vex_device_check(void);

/* indices into the device `counts` array written by vex_partition */
#define VEX_COUNT_VISION 0
#define VEX_COUNT_LANGUAGE 1
#define VEX_COUNT_VALID 2
#define VEX_COUNT_MAXLEN 3
#define VEX_NUM_COUNTS 4

/* K1 -- token-type partition + compaction.
 * Replaces get_expert_mask (modeling_cogvlm.py:58-70) and every boolean-mask index / nonzero derived
 * from it (:96-97, :244-245, :278-279, :307, :326) and _to_tensor_list (:100-104).
 *   vision[b,l]   = tt[b,l]==1 && tt[b,l+1]==1 (l < L-1; last column false), && padding_mask
 *   language[b,l] = !vision_raw && padding_mask
 * Requires L > 1: the L == 1 rule of :67 (every token of a decode step goes to the language expert) needs no
 * partition -- the decode step calls vex_grouped_gemm with single_expert = 1 on the language weights.  All index
 * outputs have B*L entries; entries past the respective count are -1.
 *   sorted_to_flat [s] : flat position of sorted row s       ([0,Tv) == nonzero(vision), [Tv,Tv+Tl) == nonzero(language))
 *   flat_to_sorted [f] : inverse, -1 for padded positions
 *   sorted_to_token[s] : token rank of sorted row s
 *   token_to_sorted[t] : inverse
 *   token_to_flat  [t] : flat position of token t            (== nonzero(padding_mask))
 *   cu_seqlens   [B+1] : prefix sums of valid tokens per sample
 *   counts         [4] : Tv, Tl, T, max valid length of a sample
 *   scratch      [4*B] : int32 workspace
 */
int vex_partition(const int64_t* token_type_ids, const uint8_t* padding_mask, int B, int L,
                  int32_t* sorted_to_flat, int32_t* flat_to_sorted, int32_t* sorted_to_token,
                  int32_t* token_to_sorted, int32_t* token_to_flat, int32_t* cu_seqlens, int32_t* counts,
                  int32_t* scratch, vexStream stream);

/* K2 -- fused RMSNorm with gather (and optional scatter).  Replaces RMSNorm.forward (:36-41) +
 * hidden_states[padding_mask] (:307, :326) and, with row_dst, _mask_set (:390-393, the caller's final norm :570-573):
 * y[dst[r]] = bf16( w * ( f32(x[src[r]]) * rsqrt(mean(f32(x)^2) + eps) ) ) for r < *n_rows.
 *   x [*, H] bf16 rows addressed through row_src (NULL = identity); y [*, H] bf16 rows through row_dst (NULL = identity)
 *   weight: bf16 (weight_is_fp32 == 0) or fp32; n_rows: device int32 (e.g. counts + VEX_COUNT_VALID)
 *   H in {256, 512, ..., 1536, 2048, 4096} (whole row kept in registers by one warp).
 */
int vex_rmsnorm_gather(const void* x, const void* weight, int weight_is_fp32, float eps,
                       const int32_t* row_src, const int32_t* row_dst, const int32_t* n_rows, void* y, int rows_cap,
                       int H, vexStream stream);

/* K5 -- standalone SiLU gate: out = bf16(bf16(silu(gate)) * up), rows < *n_rows.  Replaces
 * act_fn(gate_proj(x)) * up_proj(x) (:55) when the SwiGLU epilogue of vex_grouped_gemm is not used. */
int vex_silu_mul(const void* gate, const void* up, void* out, const int32_t* n_rows, int rows_cap, int I,
                 vexStream stream);

/* K6 -- standalone residual add + scatter to the reference layout:
 * out[dst[r]] = bf16(f32(residual[dst[r]]) + f32(y[r])), r < *n_rows.  Replaces out[mask] = ... and
 * residual + hidden_states (:278-279/:96-97 and :321/:330) when the GEMM epilogue does not fuse it. */
int vex_residual_scatter(const void* y, const void* residual, const int32_t* row_dst, const int32_t* n_rows,
                         void* out, int rows_cap, int H, vexStream stream);

/* Rows with padding_mask == False: out[f] = x[f] (residual + 0).  The reference leaves these rows
 * uninitialised (torch.empty, :277); copying keeps the output deterministic and NaN-free. */
int vex_copy_padded_rows(const void* x, const int32_t* flat_to_sorted, void* out, int n_flat, int H,
                         vexStream stream);

/* K3 -- grouped (two-expert) bf16 GEMM on tcgen05/TMEM fed by TMA, fp32 accumulate:
 *     out = A_sorted . W_e^T  (+ T_e . Blora_e^T)       e = expert of the row (sorted order)
 * Replaces the routed nn.Linear calls (:244-245, :278-279, :96-97 via MLP.forward :54-56) and the PEFT
 * lora.Linear delta (scripts/cli.py:82-88).  `mode` selects the fused epilogue. */
#define VEX_EPI_PLAIN 0    /* out[map(r)] = bf16(acc) */
#define VEX_EPI_ROPE 1     /* QKV: rotary on columns < 2*hidden (heads of 128), scatter to token order
                              (apply_rotary_pos_emb_index_bhs :188-193 fused) */
#define VEX_EPI_SWIGLU 2   /* gate/up pair: out = silu(A.Wg^T) * (A.Wu^T)  (MLP.forward :55 fused) */
#define VEX_EPI_RESIDUAL 3 /* out[map(r)] = bf16(acc) + residual[map(r)]  (:321, :330 fused) */
#define VEX_EPI_DROPOUT_ACC 4 /* out[map(r)] = residual[map(r)] + bf16(keep(r, col) * alpha / (1 - p) * acc): adjoint of
                                 the LoRA input dropout (PEFT lora.Linear: lora_A(dropout(x))), mask as vex_dropout_rows */

#define VEX_EPI_CE 5       /* fused lm_head + cross-entropy, forward (CogVLMForCausalLM.forward :701-706 +
                              _sample_weighted_ce :610-627): no logits are written; per row r and 128-column half tile t
                              ce_pmax[r, t] = max_j z, ce_psum[r, t] = sum_j exp(z - max), ce_zlabel[r] = z[label_r],
                              z = bf16-rounded logit.  single_expert, N = vocabulary, slots per row = 2 * ceil(N / 256)
                              (an empty slot holds max = -inf, sum = 0) */
#define VEX_EPI_CE_BWD 6   /* backward: out[r, j] = bf16((exp(z - ce_lse[r]) - [j == label_r]) * ce_w[r] * ce_dloss[0]
                              / counts[0]) -- d(loss)/d(logits), the A operand of the lm_head dgrad GEMM */

#define VEX_ACT_NONE 0
#define VEX_ACT_GELU 1     /* exact (erf) GELU: ACT2FN['gelu'] of the vision MLP (visual.py:108, :115) */

typedef struct vexGemmArgs {
  const void* a;            /* [rows_cap, K] bf16, sorted row order, row stride lda elements */
  int64_t lda;
  const void* w[2][2];      /* [expert][half]: weight [N, K] bf16 row-major (nn.Linear layout).
                               half 1 is only used by VEX_EPI_SWIGLU (w[e][0] = gate_proj, w[e][1] = up_proj) */
  int64_t ldw;
  const void* lora_t[2];    /* per half: T = scaling * (A_sorted . lora_A^T) [rows_cap, r] bf16, or NULL */
  int64_t ldt;
  const void* lora_b[2][2]; /* [expert][half]: lora_B [N, r] bf16 row-major, or NULL (adapter absent) */
  int32_t lora_r;           /* 0 = no LoRA; multiple of 8, <= 64 */
  void* out;                /* bf16, row stride ldo */
  int64_t ldo;
  const int32_t* counts;    /* device: rows of expert 0 (vision), rows of expert 1 (language) */
  const int32_t* row_map;   /* device [rows_cap]: output row of sorted row r, NULL = identity */
  const void* residual;     /* VEX_EPI_RESIDUAL: bf16, same layout as out */
  const void* rope_cos;     /* VEX_EPI_ROPE: [rope_len, 128] bf16 tables (RotaryEmbedding :162-170) */
  const void* rope_sin;
  const int64_t* position_ids; /* VEX_EPI_ROPE: int64 [B*L] (flat) */
  const int32_t* sorted_to_flat;
  int32_t rope_len;
  int32_t rope_cols;        /* VEX_EPI_ROPE: columns [0, rope_cols) are rotated (2*hidden) */
  int32_t rows_cap;         /* upper bound on total rows (B*L) */
  int32_t N;                /* output features per weight (for SWIGLU: intermediate size) */
  int32_t K;
  int32_t mode;
  int32_t single_expert;    /* 1: counts[0] rows all use w[0][*] (plain dense GEMM) */
  float alpha;              /* VEX_EPI_PLAIN: out = bf16(alpha * acc) (LoRA scaling folded into T) */
  int32_t w_transposed;     /* 0: w is [N, K] (forward, out = A . W^T).  1: w is [K, N] row-major and out = A . W --
                               the backward form dX = dY . W that reads the nn.Linear weight [out, in] as stored
                               (K = out features, N = in features; what autograd does for :244-245 etc.).  lora_b[e][0]
                               is then lora_A [r, N] and lora_t = scaling * dY . lora_B.  PLAIN / RESIDUAL /
                               DROPOUT_ACC only */
  float dropout_p;          /* VEX_EPI_DROPOUT_ACC: drop probability and seed of the forward vex_dropout_rows call; */
  uint64_t dropout_seed;    /*   the mask index is sorted_row * N + col */
  const int32_t* ce_labels; /* VEX_EPI_CE / CE_BWD: label of row r (device, from vex_label_rows) */
  float* ce_pmax;           /* VEX_EPI_CE out: [rows_cap, 2 * ceil(N/256)] fp32 */
  float* ce_psum;           /* VEX_EPI_CE out: [rows_cap, 2 * ceil(N/256)] fp32 */
  float* ce_zlabel;         /* VEX_EPI_CE out: [rows_cap] fp32 */
  const float* ce_lse;      /* VEX_EPI_CE_BWD: [rows_cap] natural-log log-sum-exp (vex_ce_reduce) */
  const float* ce_w;        /* VEX_EPI_CE_BWD: [rows_cap] per-row weight (vex_label_rows) */
  const float* ce_dloss;    /* VEX_EPI_CE_BWD: device scalar, gradient of the loss */
  const void* bias;         /* VEX_EPI_PLAIN / VEX_EPI_RESIDUAL: bf16 [N] bias of an nn.Linear WITH bias (the vision
                               encoder's Linears, visual.py:84-85, :110-111, and the patch convolution), added to the fp32
                               accumulator before the single bf16 rounding; NULL = none.  N % 32 == 0, forward form only */
  int32_t act;              /* VEX_EPI_PLAIN: activation applied to the bf16-rounded output (VEX_ACT_*) */
  /* VEX_EPI_ROPE second output -- KV-cache production (modeling_cogvlm.py:252-262): when kv_k / kv_v are
     set, the post-rotary K heads and the V heads of every live row are ALSO written to the cache
     kv_x[b][head][l + *kv_pos][0..127], layout [B, heads, kv_capacity, 128] bf16 (the reference's [B, heads, L, 128]
     with pre-allocated headroom), where (b, l) = divmod(sorted_to_flat[row], kv_seq_len).  Prefill: kv_seq_len = L,
     kv_pos = NULL.  Decode step (q_len == 1): kv_seq_len = 1, sorted_to_flat = identity, kv_pos = device counter of
     the positions already cached -- the append of :258-260 without torch.cat.  Rows with l + *kv_pos >= kv_capacity
     are dropped.  Requires N == 3 * rope_cols / 2 (the QKV projection). */
  void* kv_k;
  void* kv_v;
  const int32_t* kv_pos;
  int32_t kv_seq_len;
  int32_t kv_capacity;
} vexGemmArgs;

int vex_grouped_gemm(const vexGemmArgs* args, vexStream stream);

/* K12 -- the skinny form of vex_grouped_gemm for a decode step (q_len == 1): rows_cap <= 32 live rows
 * (= the batch; `counts` is not read), single_expert = 1 (get_expert_mask's L == 1 rule :67: every token goes to the
 * language expert), forward form only.  HBM-bound by design: the weights stream from global memory straight into
 * mma.sync fragments, 16 output features per CTA, K interleaved over 8 warps.  Modes VEX_EPI_PLAIN / RESIDUAL / SWIGLU /
 * ROPE (with the KV-cache append, kv_seq_len = 1) and the LoRA K-extension (lora_r % 64 == 0); row_map, bias, act are
 * not supported (VEX_E_UNSUPPORTED -- callers fall back to vex_grouped_gemm).  K % 64 == 0, N % 16 == 0 (N % 128 == 0
 * for ROPE), operands 32-byte aligned (256-bit loads); position_ids is indexed by batch row. */
int vex_decode_gemm(const vexGemmArgs* args, vexStream stream);

/* K4 -- causal block-diagonal (varlen) flash attention over the token-order QKV buffer.
 * Replaces attention_fn's prefill branch (:106-128): per sample, token i attends to tokens j <= i of
 * the same sample (causality by token rank, not position_ids), scale = 128^-0.5, fp32 softmax.
 *   qkv [rows_cap = B*max_len_cap, 3, heads, 128] bf16, rows [0, T) live (q and k already rotated); rows
 *   [T, T+128) are scratch and are zeroed by the call.  out rows are scattered through out_row_map
 *   (token_to_sorted; NULL = identity) into [rows_cap, heads*128] bf16.
 *   Default implementation: the persistent two-tile tcgen05 kernel (k4_attention_tc3.cu); it draws work items from
 *   one of 1024 device-side counters handed out round-robin per launch and reset by the call itself (stream-ordered;
 *   captured launches keep their counter).  The kernels it superseded live in csrc/baselines/ and are built into a
 *   separate libvex_baselines.so for A/B measurements; the product library contains this implementation only. */
int vex_attention(const void* qkv, const int32_t* cu_seqlens, int B, int max_len_cap, int heads,
                  const int32_t* out_row_map, void* out, float scale, vexStream stream);

/* K4 (training): same as vex_attention, additionally writing the log-sum-exp of the scaled scores in the log2
 * domain, lse[h * rows_cap + t] (fp32 [heads, rows_cap], rows_cap = B * max_len_cap; NULL = do not write), which
 * vex_attention_backward reads.  tcgen05 implementation only. */
int vex_attention_lse(const void* qkv, const int32_t* cu_seqlens, int B, int max_len_cap, int heads,
                      const int32_t* out_row_map, void* out, float scale, float* lse, vexStream stream);

/* K4d -- decode-step attention (q_len == 1 against the KV cache).  Replaces attention_fn's generation branch
 * (:129-141): q [B, heads*128] rows of stride ldq elements (already rotated), k / v [B, heads, L, 128] (the
 * reference cache layout, current token included), mask uint8 [B, L] (attention mask over past + current),
 * out [B, heads*128].  Query scaled in bf16, bf16 scores, fp32 softmax cast to bf16, like the eager reference. */
int vex_attention_decode(const void* q, int64_t ldq, const void* k, const void* v, const uint8_t* mask, void* out,
                         int B, int heads, int L, float scale, vexStream stream);

/* K4d over a pre-allocated cache: k_cache / v_cache [B, heads, kv_capacity, 128]; *kv_len (device) =
 * positions cached BEFORE this step, the step attends to positions [0, *kv_len] (the current token's K / V were appended
 * by the VEX_EPI_ROPE epilogue, see vexGemmArgs.kv_k), clamped to min(kv_capacity, mask_len); mask [B, mask_len] uint8,
 * rows ld_mask bytes apart.  Nothing is read from the host, so a decode step replays as a CUDA graph while *kv_len
 * advances (vex_advance_counter). */
int vex_attention_decode_cache(const void* q, int64_t ldq, const void* k_cache, const void* v_cache,
                               const uint8_t* mask, int64_t ld_mask, int mask_len, void* out, int B, int heads,
                               int kv_capacity, const int32_t* kv_len, float scale, vexStream stream);

/* *counter += by, on the stream (the decode graph's own "past length += 1"). */
int vex_advance_counter(int32_t* counter, int by, vexStream stream);

/* Cache rows of padded prefill positions: k_cache[b, :, l, :] = v_cache[b, :, l, :] = 0 where flat_to_sorted[b*L + l]
 * < 0 (padding_mask == False) -- the reference's cache holds zeros there (:243); live positions are written by the
 * VEX_EPI_ROPE epilogue. */
int vex_kv_clear_padded(void* k_cache, void* v_cache, const int32_t* flat_to_sorted, int B, int L, int heads,
                        int kv_capacity, vexStream stream);

/* ---------------------------------------------------------------------------------------------------
 * Training-step variant (BASELINE config 5): backward of the layer for LoRA fine-tuning.  The reference has no
 * backward code of its own -- torch.autograd differentiates modeling_cogvlm.py:30-340 inside
 * MMMMForCausalLM.training_step (mmmm/models/mmmm.py:299-306) under non-reentrant checkpointing (:287-291);
 * these entry points are the adjoints of the forward kernels above.  The input-gradient GEMMs are
 * vex_grouped_gemm with w_transposed = 1.
 * --------------------------------------------------------------------------------------------------- */

/* out[r] = x[row_src[r]] for r < *n_rows (bf16 rows of H elements, H % 8 == 0): gathers d_out [B*L, H] into
 * expert-sorted order -- the adjoint of the out[mask] = ... scatters (:96-97, :278-279). */
int vex_gather_rows(const void* x, const int32_t* row_src, const int32_t* n_rows, void* out, int rows_cap, int H,
                    vexStream stream);

/* LoRA input dropout (PEFT lora.Linear.forward: lora_B(lora_A(dropout(x))) * scaling, conf/lora.yaml lora_dropout):
 * out[r, c] = keep(r, c) ? bf16(x[r, c] / (1 - p)) : 0 for r < *n_rows, keep from a counter-based hash of
 * (r * K + c, seed) -- reproducible, so the checkpointed recompute and VEX_EPI_DROPOUT_ACC regenerate the same mask.
 * (PyTorch's Philox stream cannot be reproduced bit-for-bit; the mask distribution is the same Bernoulli(1 - p).) */
int vex_dropout_rows(const void* x, void* out, const int32_t* n_rows, int rows_cap, int K, float p, uint64_t seed,
                     vexStream stream);

/* Adjoint of K5 / the SwiGLU epilogue (MLP.forward :55): given d(act), gate = gate_proj(x), up = up_proj(x)
 * (bf16, [rows_cap, I]) writes dgate and dup with the eager-bf16 rounding points
 * (s = bf16(silu(gate)); dup = bf16(dact * s); dgate = bf16(bf16(dact * up) * silu'(gate))). */
int vex_silu_mul_backward(const void* dact, const void* gate, const void* up, void* dgate, void* dup,
                          const int32_t* n_rows, int rows_cap, int I, vexStream stream);

/* Adjoint of K2 (RMSNorm.forward :36-41), fp32 internals:
 *   dx[dx_map[r]] = bf16( inv*(dy[r]*w) - x*inv^3*sum(dy[r]*w*x)/H  +  add[add_map[r]] ),  x = x[x_map[r]]
 *   dweight[H] (fp32, caller-zeroed or accumulating) += sum_r dy[r] * x * inv
 * `add` (may be NULL) is the gradient arriving through the residual branch (:321, :330), fused here.
 * Maps may be NULL (identity).  H in {256, 512, 1024, 2048, 4096}. */
int vex_rmsnorm_backward(const void* dy, const void* x, const int32_t* x_map, const void* weight, int weight_is_fp32,
                         float eps, const void* add, const int32_t* add_map, void* dx, const int32_t* dx_map,
                         float* dweight, const int32_t* n_rows, int rows_cap, int H, vexStream stream);

/* K10 -- row selection and reduction around the fused lm_head + cross-entropy (VEX_EPI_CE / VEX_EPI_CE_BWD).
 * vex_label_rows: rows with labels != ignore_index (-100, CE_IGNORE_INDEX mmmm/data/defs.py), ascending flat order
 * (== the order of ce[mask], :619-626): row_idx[k] = flat position, label_sel[k] = label, w_sel[k] = weight (fp32 from
 * bf16 / fp32 `weight`, 1.0 when weight == NULL), count[0] = number of selected rows.  n <= 2^24.
 * vex_ce_reduce: lse[r] = log sum_j exp(z_rj) from the tile partials, loss[0] = sum_r (lse[r] - zlabel[r]) * w_sel[r]
 * / count (== F.cross_entropy mean when weight is None, == dot(ce[mask], weight[mask]) / mask.sum() otherwise;
 * 0/0 = NaN for an all-ignored batch, like the reference).  loss must be zeroed by the caller. */
int vex_label_rows(const int64_t* labels, const void* weight, int weight_is_fp32, int n, int64_t ignore_index,
                   int32_t* row_idx, int32_t* label_sel, float* w_sel, int32_t* count, vexStream stream);
int vex_ce_reduce(const float* pmax, const float* psum, const float* zlabel, const float* w_sel, const int32_t* count,
                  int rows_cap, int n_tiles, float* lse, float* loss, vexStream stream);

/* K8 -- LoRA weight gradients on tcgen05 (both operands MN-major, reduction over tokens), per expert segment e:
 *     out_e[f, j] += sum_{t in segment e} x[t, f] * y[t, j]        f < F, j < r
 * fp32 atomics into caller-zeroed (or accumulating) buffers; out_e == NULL skips expert e.
 *   dB[out, r] = dy^T . T   : x = dy [rows, out], y = T = scaling * a . lora_A^T (forward intermediate), transpose_out = 0, ldo = r
 *   dA[r, in]  = dT^T . a   : x = a [rows, in],  y = dT = scaling * dy . lora_B,                        transpose_out = 1, ldo = in
 * (autograd of PEFT lora.Linear over the routed Linears, scripts/cli.py:82-88; x / y rows in sorted order,
 * row strides ldx / ldy elements; r % 8 == 0, r <= 64.) */
int vex_lora_wgrad(const void* x, int64_t ldx, const void* y, int64_t ldy, int r, float* out_vision,
                   float* out_language, int64_t ldo, int transpose_out, const int32_t* counts, int rows_cap, int F,
                   vexStream stream);

/* K9 -- backward of K4 fused with the adjoint of the rotary embedding (attention_fn :106-128, rotary :188-193
 * under autograd) and the scatter to expert-sorted rows.
 *   qkv        [rows_cap, 3*heads*128] bf16, token order, q / k rotated (the forward's buffer, tail rows zeroed)
 *   out_sorted [rows_cap, heads*128]   bf16, forward output in sorted order (row token_to_sorted[t] belongs to token t)
 *   d_out_tok  [rows_cap, heads*128]   bf16, gradient of the forward output in TOKEN order; rows [T, T+128) are
 *                                      zeroed by the call
 *   lse        [heads, rows_cap] fp32 from vex_attention_lse; delta_ws [heads, rows_cap] fp32 workspace
 *   position_ids int64 [B*L] indexed through token_to_flat; rope tables [rope_len, 128] bf16 (the forward's)
 *   dqkv       [rows_cap, 3*heads*128] bf16 OUT: gradient w.r.t. the PRE-rotary q | k | v, row token_to_sorted[t]
 *              (token_to_sorted / token_to_flat NULL = identity) -- the A operand of the QKV dgrad GEMM. */
int vex_attention_backward(const void* qkv, const void* out_sorted, void* d_out_tok, const float* lse,
                           float* delta_ws, const int32_t* cu_seqlens, const int32_t* token_to_sorted,
                           const int32_t* token_to_flat, const int64_t* position_ids, const void* rope_cos,
                           const void* rope_sin, int rope_len, int B, int max_len_cap, int heads, void* dqkv,
                           float scale, vexStream stream);

/* ---------------------------------------------------------------------------------------------------
 * Vision encoder in front of the decoder: EVA2CLIPModel.forward, mmmm/models/cogvlm/
 * visual.py:24-208.  Its Linears run on vex_grouped_gemm (single_expert, `bias` / `act`), its attention on
 * vex_attention_blockdiag; the entry points below are the row-wise pieces.
 * --------------------------------------------------------------------------------------------------- */

/* K4 (non-causal): xformers memory_efficient_attention under a BlockDiagonalMask (visual.py:96-98) -- every token
 * attends to all tokens of its own image.  Same buffers and layout as vex_attention (head slots of 128; a head_dim
 * below 128 is zero-padded by the caller, `scale` = head_dim^-0.5 of the real head_dim). */
int vex_attention_blockdiag(const void* qkv, const int32_t* cu_seqlens, int B, int max_len_cap, int heads,
                            const int32_t* out_row_map, void* out, float scale, vexStream stream);

/* K11 -- nn.LayerNorm over rows r < *n_rows of x [rows_cap, H] (bf16; weight / bias bf16 [H], bias may be NULL):
 *   t = bf16((x - mean) * rsqrt(var + eps) * weight + bias)      fp32 statistics, biased variance
 *   act == VEX_ACT_GELU: t = bf16(gelu(t))                       GLU projector, visual.py:173-174
 *   residual != NULL:    y = bf16(residual + t)  else  y = t     TransformerLayer.forward visual.py:128-135
 * y may alias residual (in-place residual stream).  H % 256 == 0, H <= 2048 or H == 4096. */
int vex_layernorm(const void* x, const void* weight, const void* bias, float eps, const void* residual, int act,
                  const int32_t* n_rows, void* y, int rows_cap, int H, vexStream stream);

/* K11 -- im2col of the patch convolution (PatchEmbedding.forward visual.py:65 -> Downsample.forward
 * mmmm/models/resample.py:56-63, conv3d with stride == kernel): image [C, D, H, W] bf16 ->
 * out[(gz*gh + gy)*gw + gx][((c*pd + kz)*ph + ky)*pw + kx], row stride ldo elements (columns past C*pd*ph*pw are left
 * untouched).  The convolution is then vex_grouped_gemm against weight.reshape(C_out, -1) with `bias`. */
int vex_patchify(const void* image, int C, int D, int H, int W, int pd, int ph, int pw, void* out, int64_t ldo,
                 vexStream stream);

/* K11 -- F.max_pool3d over the patch grid in token-major layout (EVA2CLIPModel.forward visual.py:197-202):
 * x rows (z*gh + y)*gw + x_ of C bf16 (row stride ldx) -> out rows (oz*oh + oy)*ow + ox, windows (pz, py, px), floor
 * semantics.  (1, 1, 1) copies the rows (the class-token drop of visual.py:198 when x points behind it). */
int vex_maxpool_tokens(const void* x, int64_t ldx, int gd, int gh, int gw, int pz, int py, int px, void* out,
                       int64_t ldo, int C, vexStream stream);

/* K11 -- out[row_dst[r]] = x[row_src ? row_src[r] : r] for r < n (rows of H bf16, H % 8 == 0; row_dst[r] < 0 skips):
 * the boi / eoi rows around each image's features (visual.py:204-206) and the feature scatter into the text
 * embeddings (CogVLMModel.forward modeling_cogvlm.py:450-453). */
int vex_scatter_rows(const void* x, const int32_t* row_src, const int32_t* row_dst, int n, void* out, int H,
                     vexStream stream);

#ifdef __cplusplus
}
#endif
#endif /* VEX_H_ */
