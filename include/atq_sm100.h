/*
 * atq_sm100.h -- C ABI of libatq_sm100.so, the B200 (sm_100a) implementation of the ATQ
 * ternary hot path.  This is the drop-in boundary: plain pointers and sizes, no torch types.
 *
 * The reference (ak736/ATQ-Multimodal) has no FFI layer; its hot path is Python over ATen
 * (SURVEY.md 8b).  Each entry point below names the reference lines whose work it replaces
 * (paths relative to the reference root).  The Python package atq-multimodal_b200/atq binds
 * these with ctypes (atq/_native.py) and keeps the reference's module signatures.
 *
 * Conventions
 *  - Every pointer is a DEVICE pointer on `device` unless the name ends in _host.
 *  - Every compute call is asynchronous on `stream` (a cudaStream_t / CUstream handle;
 *    0 = legacy default stream), never synchronises, never allocates or frees, and never
 *    retains a pointer after the work it enqueued completes.
 *  - Scratch memory is provided by the caller: `ws` must hold at least
 *    atq_workspace_bytes_<op>(...) bytes, 256-byte aligned, and stay alive until the stream
 *    work completes.  Calls on one stream may share one workspace.
 *  - Return value: ATQ_OK or a negative atq_status; atq_last_error_string() (thread-local)
 *    describes the last failure.
 *  - Scalars that the reference keeps as 0-dim tensors (threshold, alpha) stay on the
 *    device: no entry point reads them back to the host.
 */
#ifndef ATQ_SM100_H
#define ATQ_SM100_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  ATQ_OK = 0,
  ATQ_EINVAL = -1,     /* bad shape / alignment / null pointer            */
  ATQ_EARCH = -2,      /* device is not compute capability 10.x           */
  ATQ_ECUDA = -3,      /* a CUDA runtime/driver call failed               */
  ATQ_EWORKSPACE = -4  /* workspace too small                             */
} atq_status;

typedef void* atq_stream_t; /* cudaStream_t */

/* ---- library ---------------------------------------------------------------------- */
int atq_abi_version(void);                 /* bumps on any signature change */
const char* atq_last_error_string(void);
int atq_device_check(int device);          /* ATQ_OK iff `device` is sm_100-class */
int atq_num_sms(int device);
uint64_t atq_kernel_launch_count(void);   /* kernels this library has launched in this process */

/* ---- A1: per-layer adaptive threshold  (atq/quantizers.py:21-38) -------------------- */
/* K1: grid-level |W| reduction.  stats_out (device, 16 bytes): double sum|W|; float max|W|; u32 pad. */
size_t atq_workspace_bytes_abs_stats(int64_t n);
int atq_abs_stats(int device, const float* w, int64_t n, void* stats_out,
                  void* ws, size_t ws_bytes, atq_stream_t stream);

/* K2: exact k-th smallest |x| (0-indexed, ascending), i.e. torch.sort(|x|).values[k]
 * (atq/quantizers.py:25-32) and torch.kthvalue(|x|, k+1) (atq/routing.py:48).
 * Bit-exact: the result is an element of |x|.  0 <= k < n. */
size_t atq_workspace_bytes_select_kth_abs(int64_t n);
int atq_select_kth_abs(int device, const float* x, int64_t n, int64_t k, float* thr_out,
                       void* ws, size_t ws_bytes, atq_stream_t stream);

/* Measurement switch: 1 (default) = the three digit passes of the exact select run in ONE cooperative launch with
 * layer-level barriers (re-reading the layer from L2); 0 = one launch per pass. */
void atq_set_fused_select(int enabled);

/* The whole threshold stage with the reference's three branches, keyed by
 * k = int(sparsity_target * n) computed by the caller in double precision:
 *   0 < k < n : k-th order statistic           (:31-32)
 *   k >= n    : max|W| + 1.0                   (:33-35)
 *   k <= 0    : threshold_factor * mean|W|     (:36-38)   */
size_t atq_workspace_bytes_adaptive_threshold(int64_t n);
int atq_adaptive_threshold(int device, const float* w, int64_t n, int64_t k, float threshold_factor,
                           float* thr_out, void* ws, size_t ws_bytes, atq_stream_t stream);

/* Batched form: `count` independent layers in one launch sequence (one threshold each).
 * w_ptrs_host/n_host/k_host/thr_ptrs_host are HOST arrays of length count. */
size_t atq_workspace_bytes_adaptive_threshold_batched(int count, const int64_t* n_host);
int atq_adaptive_threshold_batched(int device, int count, const float* const* w_ptrs_host,
                                   const int64_t* n_host, const int64_t* k_host, float threshold_factor,
                                   float* const* thr_ptrs_host, void* ws, size_t ws_bytes,
                                   atq_stream_t stream);

/* ---- A2/A3: ternarize (+ optional optimal-alpha statistics) (atq/quantizers.py:41-59) -- */
/* stats (nullable, device, 16 bytes, accumulated into -- caller zeroes): u64 nnz; double sum(W*T). */
int atq_ternarize_f32(int device, const float* w, int64_t n, const float* thr, float* t_out,
                      void* stats, atq_stream_t stream);
/* K4: fused ternarize -> 2-bit pack in the reference's codec layout (atq/bit_packing.py:45-69). */
int atq_ternarize_pack2(int device, const float* w, int64_t n, const float* thr, uint8_t* packed,
                        void* stats, atq_stream_t stream);
/* alpha* = sum(W*T)/nnz, or mean|W| when nnz == 0 (:49-55), resolved on the device.
 * tern_stats as above; abs_stats as written by atq_abs_stats. */
int atq_optimal_alpha(int device, const void* tern_stats, const void* abs_stats, int64_t n,
                      float* alpha_out, atq_stream_t stream);

/* ---- E1/E2: 2-bit codec (atq/bit_packing.py:22-119) --------------------------------- */
/* code = value+1, element i at bits 2(i%4).. of byte i/4, tail bits zero.
 * invalid_flag (device int32, caller zeroes): set to 1 if any input is not in {-1,0,+1}
 * (reference raises ValueError at :39) / if any 2-bit code is 3 (reference KeyError at :116). */
int atq_pack2_from_f32(int device, const float* t, int64_t n, uint8_t* packed, int32_t* invalid_flag,
                       atq_stream_t stream);
int atq_unpack2_to_f32(int device, const uint8_t* packed, int64_t n, float* out, int32_t* invalid_flag,
                       atq_stream_t stream);
int atq_unpack2_to_bf16(int device, const uint8_t* packed, int64_t n, uint16_t* out, atq_stream_t stream);
int atq_unpack2_to_i8(int device, const uint8_t* packed, int64_t n, int8_t* out, atq_stream_t stream);

/* Whole-model forms of the three codec kernels (BASELINE config 5: ~1 B weights in tens of layers): one launch for
 * all layers, 64 weights per 128-bit packed access, 16-byte aligned layers; bytes identical to the per-layer calls.
 * invalid_flag (nullable) is set to 1 if any value is not in {-1, 0, +1} (pack) / any code is 3 (unpack). */
int atq_ternarize_pack2_batched(int device, int count, const float* const* w_ptrs, const int64_t* ns,
                                const float* const* thr_ptrs, uint8_t* const* packed_ptrs, atq_stream_t stream);
int atq_pack2_from_f32_batched(int device, int count, const float* const* t_ptrs, const int64_t* ns,
                               uint8_t* const* packed_ptrs, int32_t* invalid_flag, atq_stream_t stream);
int atq_unpack2_to_f32_batched(int device, int count, uint8_t* const* packed_ptrs, const int64_t* ns,
                               float* const* out_ptrs, int32_t* invalid_flag, atq_stream_t stream);

/* ---- D2: selective gradient routing backward (atq/routing.py:53-56) ------------------ */
/* grad_in = grad_out * (|x| > *thr) */
int atq_route_mask_mul(int device, const float* x, const float* grad_out, const float* thr, int64_t n,
                       float* grad_in, atq_stream_t stream);

/* ---- GEMM operand builders ----------------------------------------------------------- */
/* Operand element formats (both are 16-bit pairs consumed by tcgen05.mma kind::f16, same MMA count):
 *   scale_slot == NULL : bf16 pair, hi = bf16(x), lo = bf16(x - hi)                      (~16 significant bits)
 *   scale_slot != NULL : scaled fp16 pair, hi = fp16(s x), lo = fp16(s x - hi), s = scale_slot[1] a power of two
 *                        chosen per tensor so that max|s x| is in [2^14, 2^15)            (~22 significant bits);
 *                        the GEMM multiplies its accumulator by scale_slot[2] = 1/s (atq_bf16_operand.inv_scale).
 * A scale slot is 4 caller-owned floats {scratch, s, 1/s, scratch}; slot[0] and slot[3] must be zero before the
 * first atq_absmax_scale on it and are zero again when that call's kernel has finished (graph replays re-arm it).
 * bound = max(max|x|, |*extra|) * bound_mul lets a producer reserve head-room for a derived tensor
 * (e.g. |dropout(gelu(y))| <= max|y| / (1-p)); extra is a nullable device scalar. */
int atq_absmax_scale(int device, const float* x, int64_t rows, int64_t cols, int64_t ld, float bound_mul,
                     const float* extra, float* slot, atq_stream_t stream);
/* the same reduction for `count` contiguous tensors in one launch (per-layer weight scales of a whole model) */
int atq_absmax_scale_batched(int device, int count, const float* const* x_ptrs, const int64_t* ns,
                             const float* const* extra_ptrs /* nullable */, float* const* slot_ptrs, float bound_mul,
                             atq_stream_t stream);
/* atq_absmax_scale + the scaled-fp16 split of a small contiguous tensor (n % 8 == 0, n <= max elems, 16-byte aligned)
 * in ONE launch: a cluster of 8 CTAs holds the tensor in registers between the max|x| exchange (distributed shared
 * memory) and the split.  Writes slot[1..2] = {s, 1/s}; slot[0], slot[3] untouched. */
int64_t atq_split_scaled_fused_max_elems(void);
int atq_split_scaled_fused(int device, const float* x, int64_t n, uint16_t* hi, uint16_t* lo, float bound_mul,
                           const float* extra, float* slot, atq_stream_t stream);
/* fp32 [rows, cols] (row pitch ld_in elements) -> hi (+ lo, nullable) [rows, pitch]; pitch % 8 == 0,
 * pitch >= cols; padding columns are left untouched. */
int atq_split_bf16(int device, const float* x, int64_t rows, int64_t cols, int64_t ld_in,
                   uint16_t* hi, uint16_t* lo, int64_t pitch, const float* scale_slot, atq_stream_t stream);
/* split of a contiguous [rows, cols] tensor (cols % 8 == 0) fused with its column sums
 * (colsum_out[c] = sum_r x[r,c], deterministic two-stage): one pass over dY yields the GEMM
 * operand and the bias gradient. */
size_t atq_workspace_bytes_split_colsum(int64_t rows, int64_t cols);
int atq_split_bf16_colsum(int device, const float* x, int64_t rows, int64_t cols, uint16_t* hi, uint16_t* lo,
                          float* colsum_out, void* ws, size_t ws_bytes, const float* scale_slot, atq_stream_t stream);
/* same, transposed output: hi_t/lo_t are [cols, pitch_t], pitch_t % 8 == 0, pitch_t >= rows.
 * colsum is reserved (must be NULL; use atq_colsum_f32). */
int atq_split_bf16_t(int device, const float* x, int64_t rows, int64_t cols, int64_t ld_in,
                     uint16_t* hi_t, uint16_t* lo_t, int64_t pitch_t, float* colsum, const float* scale_slot,
                     atq_stream_t stream);

/* Quantize a layer into everything its GEMMs consume, in one pass over W [M,K]:
 *  packed   : 2-bit codec bytes of T (public format, flat row-major), nullable; K % 4 == 0
 *  packed_t : 2-bit codec bytes of T^T ([K, M/4]; dX operand of atq_tgemm_packed), nullable; M % 4 == 0
 *  tb       : T as bf16 [M, pitch]  (forward B operand), nullable
 *  tb_t     : T^T as bf16 [K, pitch_t] (dX B operand), nullable
 * replaces atq/quantizers.py:41-43 + the `w_ternary * alpha` materialisation of atq/layers.py:43. */
int atq_build_ternary_operands(int device, const float* w, int64_t M, int64_t K, const float* thr,
                               uint8_t* packed, uint8_t* packed_t, uint16_t* tb, int64_t pitch,
                               uint16_t* tb_t, int64_t pitch_t, void* stats, int fp16 /* tb, tb_t as fp16 */,
                               atq_stream_t stream);
/* Residual-precision-boost mixed weight (atq/precision_boost.py:72):
 *  Wm = T*alpha*(1-mask) + W*mask, emitted as bf16 hi/lo pairs [M,pitch] and transposed
 *  [K,pitch_t]; lo pointers nullable (fast mode).  packed as above (nullable). */
int atq_build_mixed_operands(int device, const float* w, const float* mask, int64_t M, int64_t K,
                             const float* thr, const float* alpha, uint8_t* packed,
                             uint16_t* hi, uint16_t* lo, int64_t pitch,
                             uint16_t* hi_t, uint16_t* lo_t, int64_t pitch_t, const float* scale_slot,
                             atq_stream_t stream);

/* ---- ternary GEMMs (tcgen05 / TMEM / TMA) -------------------------------------------- */
/* Common operand description: a bf16 matrix given as hi (+ optional lo) so that the fp32 value
 * is hi + lo; pitch % 8 == 0 elements, base 16-byte aligned.
 *   mn_major == 0: memory is [rows, kdim] row-major (contraction index contiguous, "K-major")
 *   mn_major == 1: memory is [kdim, rows] row-major (the transposed view of a row-major tensor is
 *                  consumed in place through MN-major UMMA descriptors: dX reads the [out, in]
 *                  weight copy, dW reads dY [tokens, out] and X [tokens, in] - no transposes)
 * Supported combinations (A, B): (0,0), (0,1), (1,1).  */
typedef struct {
  const uint16_t* hi;
  const uint16_t* lo; /* nullable */
  int64_t pitch;
  int32_t mn_major;
  int32_t format;          /* 0 = bf16, 1 = fp16; A and B of one GEMM must agree */
  const float* inv_scale;  /* nullable device scalar: the accumulator is multiplied by it (1/s of a scaled operand) */
} atq_bf16_operand;

/* K7 forward:  Y[N,M] = scale * (X[N,K] . B[M,K]^T) + bias      (atq/layers.py:43,
 * atq/precision_boost.py:74, atq/bit_packing.py:165-176).
 *  - TernaryLinear: B = T (tb, exact in bf16), scale = alpha (device scalar).
 *  - RPB: B = Wm hi/lo, scale = NULL.
 * K8 dX:  dX[N,K] = scale * (dY[N,M] . Bt[K,M]^T); with dot_ref = X it also returns
 *  dot_out = sum(acc .* X) = d(alpha) of TernaryLinear (autograd's sum(G.*T)).
 * Both are this one entry point: D[rows,cols] = scale*(A . B^T) (+bias[cols]).           */
/* GEMMs with a lo operand part run on CTA pairs (tcgen05.mma.cta_group::2, 256 x 128 tiles) when rows >= 256,
 * cols >= 128 and the problem has at least one 128 x 128 tile per SM (smaller GEMMs are latency bound: single CTAs);
 * 256 x 256 single-buffered pair tiles when the contraction has >= 32 k-blocks.  atq_set_cta_pairs(0) forces the
 * single-CTA kernels, 1 = pairs (default), 4|1 = pairs without the 256-wide tiles (A/B measurements), 8|... = pairs also
 * for problems below one wave (tests of ragged shapes).  Returns the previous setting. */
int atq_set_cta_pairs(int enabled);
size_t atq_workspace_bytes_tgemm(int64_t rows, int64_t cols);
int atq_tgemm(int device, int64_t rows, int64_t cols, int64_t kdim,
              const atq_bf16_operand* a, const atq_bf16_operand* b,
              const float* scale, const float* bias,
              float* out, int64_t out_pitch,
              const float* dot_ref, int64_t dot_ref_pitch, float* dot_out,
              void* ws, size_t ws_bytes, atq_stream_t stream);
/* atq_tgemm (scale, bias epilogue) that also produces the scale slot of ITS OUTPUT as a scaled-fp16 operand:
 * the epilogue folds max|out| into out_scale_slot[0] (zero on entry), a one-thread kernel then writes
 * slot[1..2] = {s, 1/s} for bound = max|out| * bound_mul and re-arms slot[0].  Saves the reduction pass over the
 * output when the consumer is an operand split (the FFN activation between linear1 and linear2). */
int atq_tgemm_absmax(int device, int64_t rows, int64_t cols, int64_t kdim,
                     const atq_bf16_operand* a, const atq_bf16_operand* b, const float* scale, const float* bias,
                     float* out, int64_t out_pitch, float* out_scale_slot, float bound_mul, atq_stream_t stream);
/* Same contraction with B given as the 2-bit codec bytes of T ([cols, kdim/4] row-major, i.e. the
 * public packed format of a [cols, kdim] ternary matrix; kdim % 64 == 0, 16-byte aligned).
 * Converter warps expand each 16-byte codec row segment to a 128-byte bf16 row of the
 * SWIZZLE_128B B tile in shared memory; HBM/L2 only ever see 2 bits per weight.
 * This is the native form of atq/bit_packing.py:149-176 (fast_ternary_matmul). */
int atq_tgemm_packed(int device, int64_t rows, int64_t cols, int64_t kdim,
                     const atq_bf16_operand* a, const uint8_t* b_packed,
                     const float* scale, const float* bias, float* out, int64_t out_pitch,
                     const float* dot_ref, int64_t dot_ref_pitch, float* dot_out,
                     void* ws, size_t ws_bytes, atq_stream_t stream);
/* named wrappers (same arguments), kept for readability at the call sites */
int atq_tgemm_fwd(int device, int64_t n_tokens, int64_t out_features, int64_t in_features,
                  const atq_bf16_operand* x, const atq_bf16_operand* w,
                  const float* alpha, const float* bias, float* y, int64_t y_pitch,
                  void* ws, size_t ws_bytes, atq_stream_t stream);
int atq_tgemm_dx(int device, int64_t n_tokens, int64_t in_features, int64_t out_features,
                 const atq_bf16_operand* dy, const atq_bf16_operand* w_t,
                 const float* alpha, float* dx, int64_t dx_pitch,
                 const float* x_ref, int64_t x_pitch, float* dalpha_out,
                 void* ws, size_t ws_bytes, atq_stream_t stream);
/* K9 masked dW:  G[M,K] = dY^T[M,N] . X^T[K,N]^T;  dW = G .* mask (mask NULL = STE opt-in: dW = G);
 * dalpha_out (nullable) = sum(G .* T .* (1-mask)), T read from the 2-bit codec bytes
 * (autograd of atq/precision_boost.py:72). */
/* ws >= atq_workspace_bytes_tgemm_dw(...) enables split-K over the token dimension when the gradient
 * has few output tiles (deterministic: per-range slabs + fixed-order finalize kernel). */
size_t atq_workspace_bytes_tgemm_dw(int64_t out_features, int64_t in_features, int64_t n_tokens);
int atq_tgemm_dw_masked(int device, int64_t out_features, int64_t in_features, int64_t n_tokens,
                        const atq_bf16_operand* dy_t, const atq_bf16_operand* x_t,
                        const float* mask, const uint8_t* packed /* codec bytes of T [M*K/4], row-major */,
                        float* dw, int64_t dw_pitch, float* dalpha_out,
                        void* ws, size_t ws_bytes, atq_stream_t stream);

/* fp32 helpers used by the layer wrappers: out[i] (+)= ... tiny device-side reductions */
/* out[c] = sum_r x[r,c]  (bias gradient); deterministic two-stage reduction */
size_t atq_workspace_bytes_colsum(int64_t rows, int64_t cols);
int atq_colsum_f32(int device, const float* x, int64_t rows, int64_t cols, int64_t ld, float* out,
                   void* ws, size_t ws_bytes, atq_stream_t stream);

/* ---- fused attention core -------------------------------------------------------
 * out = dropout(softmax(scale * q k^T + key_padding)) v per (batch, head); replaces the explicit
 * matmul / masked_fill / softmax / dropout / matmul of models/text_encoder.py:117-163 and its autograd
 * backward.  q, k, v, out, dout, dq, dk, dv: fp32 [B*L, pitch] with head h in columns [64h, 64h+64)
 * (head_dim 64, 1 <= L <= 256, 16-byte aligned, pitch % 4 == 0) - the layout the q/k/v projections
 * write, no head transposes.  key_padding: [B, L] bytes, non-zero = masked key (nullable).
 * lse: [B*H, L] log-sum-exp of the scaled masked scores (forward output, backward input).
 * seed: device scalar (nullable = 0) for the counter-based dropout hash; the backward call must pass
 * the same seed / dropout_p.  terms: 3 = bf16 hi/lo operand split (parity), 1 = bf16 (fast). */
int atq_attention_fwd(int device, int B, int H, int L, int head_dim /* multiple of 8, <= 64 */, const float* q, int64_t q_pitch, const float* k, int64_t k_pitch,
                      const float* v, int64_t v_pitch, const uint8_t* key_padding, float scale, float dropout_p,
                      const unsigned long long* seed, int terms, float* out, int64_t out_pitch, float* lse,
                      atq_stream_t stream);
int atq_attention_bwd(int device, int B, int H, int L, int head_dim, const float* q, int64_t q_pitch, const float* k, int64_t k_pitch,
                      const float* v, int64_t v_pitch, const uint8_t* key_padding, float scale, float dropout_p,
                      const unsigned long long* seed, int terms, const float* out, int64_t out_pitch, const float* dout,
                      int64_t dout_pitch, const float* lse, float* dq, int64_t dq_pitch, float* dk, int64_t dk_pitch,
                      float* dv, int64_t dv_pitch, atq_stream_t stream);

/* ---- FFN activation fused with the operand split (GELU into the FFN epilogue) -----
 * forward : d = dropout(gelu(y)) (exact erf GELU, models/text_encoder.py:246) written ONLY as the bf16
 *           (hi, lo) A operand of the second FFN GEMM ([rows, cols], pitch = cols, cols % 8 == 0);
 * backward: dy = g .* keep/(1-p) .* gelu'(y) as the (hi, lo) operand of the first layer's dX / dW GEMMs plus
 *           its column sums (bias gradient); ws >= atq_workspace_bytes_split_colsum(rows, cols).
 * The dropout mask is regenerated from the counter hash of (seed, flat index): pass the same seed / p. */
int atq_gelu_dropout_split(int device, const float* y, int64_t rows, int64_t cols, float dropout_p,
                           const unsigned long long* seed, uint16_t* hi, uint16_t* lo, const float* scale_slot,
                           atq_stream_t stream);
int atq_gelu_dropout_bwd_split_colsum(int device, const float* g, const float* y, int64_t rows, int64_t cols, float dropout_p,
                                      const unsigned long long* seed, uint16_t* hi, uint16_t* lo, float* colsum_out,
                                      void* ws, size_t ws_bytes, const float* scale_slot, atq_stream_t stream);

/* LayerNorm over the last dimension of a contiguous fp32 [rows, cols] tensor (4 <= cols <= 1024, cols % 4 == 0), the op
 * in front of every ternary GEMM of the transformer block (models/text_encoder.py:77,232,244).  forward saves mean / rstd
 * per row and, when out_scale_slot != NULL (zero slot[0] on entry), leaves {s, 1/s} for y as a scaled-fp16 operand in
 * slot[1..2] (the operand split that follows needs no reduction pass).  backward: dx, dgamma, dbeta (deterministic
 * two-stage column sums); ws >= atq_workspace_bytes_layernorm_bwd(cols). */
int atq_layernorm_fwd(int device, const float* x, const float* gamma, const float* beta, int64_t rows, int64_t cols, float eps,
                      float* y, float* mean_out, float* rstd_out, float* out_scale_slot /* nullable */, atq_stream_t stream);
size_t atq_workspace_bytes_layernorm_bwd(int64_t cols);
int atq_layernorm_bwd(int device, const float* dy, const float* x, const float* gamma, const float* mean, const float* rstd,
                      int64_t rows, int64_t cols, float* dx, float* dgamma, float* dbeta, void* ws, size_t ws_bytes,
                      atq_stream_t stream);

/* Gated residual of the ternary transformer block (models/text_encoder.py:238-249):
 *   out = src + dropout(h) * g   with g = sigmoid(gate) a device scalar; n % 4 == 0, contiguous tensors.
 * backward: dh = dout * g * keep/(1-p), dgate = sum(dout .* dropout(h)) (deterministic two-stage sum); d(src) = dout. */
size_t atq_workspace_bytes_gated_residual(int64_t n);
int atq_gated_residual_fwd(int device, const float* src, const float* h, const float* gate, int64_t n, float dropout_p,
                           const unsigned long long* seed, float* out, atq_stream_t stream);
int atq_gated_residual_bwd(int device, const float* dout, const float* h, const float* gate, int64_t n, float dropout_p,
                           const unsigned long long* seed, float* dh, float* dgate, void* ws, size_t ws_bytes,
                           atq_stream_t stream);

/* AdamW over a table of tensors in one launch (the optimizer step of train_multimodal.py:361-366; same update
 * as torch.optim.AdamW, no amsgrad).  table: device array of {float* p; const float* g; float* m; float* v;
 * int64 n} (16-byte aligned pointers, the four tensors of an entry share one memory layout); work unit c =
 * elements [1024*chunk_off[c], +1024) of tensor chunk_tensor[c].  step: device scalar holding the number of
 * completed steps (bias correction uses step+1; the call increments it - CUDA-graph replays advance it). */
int atq_adamw_multi(int device, const void* table, const int* chunk_tensor, const int* chunk_off, int n_chunks, float lr,
                    float beta1, float beta2, float eps, float weight_decay, float* step, atq_stream_t stream);

/* ---- fused hard-negative-mining InfoNCE (utils/enhanced_contrastive.py:64-158) -----------------
 * s: [b, b] fp32 similarity / temperature (row pitch ld).  Hard negatives = entries at least as large as the k-th
 * largest off-diagonal entry of their row or of their column (the two torch.topk calls of :98-110 + the mask loop of
 * :118-120); W = s * (pw on the diagonal, hard_mul on hard negatives, 1 elsewhere).
 *   atq_rowkth_largest      thr[i] = k-th largest of { s[i, j] : j != i }, exact (1 <= k <= b - 1; b <= 49 152)
 *   atq_infonce_row_stats   per row of m (m = s with (thr_a, thr_b) = (row, col) thresholds, or m = s^T with them
 *                           swapped): lse_w = logsumexp(W_i), lse_s = logsumexp(m_i), exp_s = sum softmax(m_i) m_i,
 *                           wdiag (nullable) = W_ii
 *   atq_infonce_finalize    loss = (CE_rows + CE_cols)/2 + lambda (H_rows + H_cols)/2 from the vectors above
 *   atq_infonce_grad        ds = grad_out * out_scale * dL/ds for every entry; dpw (nullable) = dL/d pw            */
int atq_rowkth_largest(int device, const float* s, int64_t b, int64_t ld, int64_t k, float* thr_out, atq_stream_t stream);
int atq_infonce_row_stats(int device, const float* m, int64_t b, int64_t ld, const float* thr_a, const float* thr_b,
                          const float* pw /* nullable */, float hard_mul, float* lse_w, float* lse_s, float* exp_s,
                          float* wdiag /* nullable */, atq_stream_t stream);
int atq_infonce_finalize(int device, int64_t b, const float* lse_w_r, const float* lse_w_c, const float* lse_s_r,
                         const float* lse_s_c, const float* exp_s_r, const float* exp_s_c, const float* wdiag,
                         float lambda_reg, float* loss, atq_stream_t stream);
int atq_infonce_grad(int device, const float* s, int64_t b, int64_t ld, const float* thr_r, const float* thr_c,
                     const float* pw /* nullable */, float hard_mul, const float* lse_w_r, const float* lse_w_c,
                     const float* lse_s_r, const float* lse_s_c, const float* exp_s_r, const float* exp_s_c,
                     float lambda_reg, float out_scale, const float* grad_out, float* ds, int64_t ld_out,
                     float* dpw /* nullable */, atq_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* ATQ_SM100_H */
