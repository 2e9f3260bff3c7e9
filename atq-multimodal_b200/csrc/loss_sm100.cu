// Hard-negative-mining InfoNCE on the B x B similarity matrix.
//
// Replaces the body of HardNegativeMiningInfoNCE.forward (utils/enhanced_contrastive.py:64-158) after the similarity
// GEMM: two topk(k = B/2) over the B x B matrix, a Python loop of B indexed writes building two B x B masks (:118-120),
// ~25 full-matrix elementwise ops, two cross entropies and two entropy terms -- and autograd's backward through all
// of it.  Here the matrix S = (img . txt^T) / tau comes from the ternary path's own tcgen05 GEMM; everything else is
//
//   rowkth_kernel     per row: exact k-th largest entry off the diagonal (MSD radix select in shared memory) ->
//                     hardness thresholds; run on S (image -> text) and on S^T (text -> image)
//   row_stats_kernel  per row: logsumexp of the weighted row W, logsumexp and expectation of the plain row
//                     (cross entropy + entropy regulariser), run on S and S^T
//   finalize_kernel   the scalar loss from the 2 x 3 per-row vectors (fp64 sums, fixed order)
//   grad_kernel       dL/dS for every entry from the per-row / per-column statistics, and d/d(pos_weights)
//
// with  W_ij = S_ij * pw_i            (i == j; pw = curriculum weights, may carry a gradient)
//              S_ij * (1 + hard_w)    (i != j and (S_ij >= rowthr_i or S_ij >= colthr_j): a hard negative)
//              S_ij                   (other negatives)
//       loss = (CE_rows(W) + CE_cols(W)) / 2 + lambda * (H_rows(S) + H_cols(S)) / 2.
// The matrix is read 6 times forward and once backward; at B = 4096 it is 64 MiB and stays in B200's L2.
// "S_ij >= k-th largest" marks every entry tied with the k-th as hard; torch.topk keeps exactly k of a tie group
// (which ones is implementation-defined) -- the two only differ when fp32 similarities tie exactly at rank k.
#include "common.cuh"

namespace atq {

constexpr int kLossThreads = 256;

__device__ __forceinline__ uint32_t ordered_key(float f) {  // order-preserving float -> uint32
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_to_float(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

// thr[i] = k-th largest of { S[i, j] : j != i } (1 <= k <= B - 1).  One CTA per row; the row's keys live in shared
// memory; four 8-bit MSD passes, each a 256-bin histogram of the keys that still match the prefix.
__global__ void __launch_bounds__(kLossThreads) rowkth_kernel(const float* __restrict__ S, int B, int64_t ld, int k, float* __restrict__ thr) {
  extern __shared__ uint32_t keys[];
  __shared__ unsigned int hist[256];
  __shared__ uint32_t s_prefix;
  __shared__ int s_k;
  const int i = blockIdx.x;
  const float* row = S + (int64_t)i * ld;
  for (int j = threadIdx.x; j < B; j += kLossThreads) keys[j] = (j == i) ? 0u : ordered_key(__ldg(row + j));  // diagonal sorts last
  if (threadIdx.x == 0) { s_prefix = 0u; s_k = k; }
  __syncthreads();
  uint32_t mask = 0u;
  for (int shift = 24; shift >= 0; shift -= 8) {
    hist[threadIdx.x] = 0u;
    __syncthreads();
    const uint32_t prefix = s_prefix;
    for (int j = threadIdx.x; j < B; j += kLossThreads) {
      const uint32_t key = keys[j];
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 0xFFu], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {  // walk the bins from the top until k entries are covered
      int need = s_k;
      int b = 255;
      for (; b > 0; --b) {
        const int c = (int)hist[b];
        if (c >= need) break;
        need -= c;
      }
      s_k = need;
      s_prefix = prefix | ((uint32_t)b << shift);
    }
    mask |= 0xFFu << shift;
    __syncthreads();
  }
  if (threadIdx.x == 0) thr[i] = key_to_float(s_prefix);
}

__device__ __forceinline__ float block_max(float x, float* s) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, o));
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = x;
  __syncthreads();
  float r = s[0];
#pragma unroll
  for (int i = 1; i < kLossThreads / 32; ++i) r = fmaxf(r, s[i]);
  __syncthreads();
  return r;
}
__device__ __forceinline__ float block_sum(float x, float* s) {
  x = warp_sum(x);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = x;
  __syncthreads();
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < kLossThreads / 32; ++i) r += s[i];
  __syncthreads();
  return r;
}

__device__ __forceinline__ float weight_of(int i, int j, float s, float thr_i, float thr_j, float pw_i, float hard_mul) {
  if (i == j) return pw_i;
  return (s >= thr_i || s >= thr_j) ? hard_mul : 1.f;
}

// One CTA per row i of M (M = S with (thr_a, thr_b) = (row, col) thresholds, or M = S^T with them swapped):
//   lse_w[i] = logsumexp_j W_ij,  lse_s[i] = logsumexp_j M_ij,  exp_s[i] = sum_j softmax(M_i)_j M_ij,  wdiag[i] = W_ii
__global__ void __launch_bounds__(kLossThreads)
    row_stats_kernel(const float* __restrict__ M, int B, int64_t ld, const float* __restrict__ thr_a, const float* __restrict__ thr_b,
                     const float* __restrict__ pw, float hard_mul, float* __restrict__ lse_w, float* __restrict__ lse_s,
                     float* __restrict__ exp_s, float* __restrict__ wdiag) {
  extern __shared__ float rowbuf[];
  __shared__ float red[kLossThreads / 32];
  const int i = blockIdx.x;
  const float* row = M + (int64_t)i * ld;
  const float thr_i = __ldg(thr_a + i);
  const float pw_i = pw != nullptr ? __ldg(pw + i) : 1.f;
  float mw = -INFINITY, ms = -INFINITY;
  for (int j = threadIdx.x; j < B; j += kLossThreads) {
    const float s = __ldg(row + j);
    rowbuf[j] = s;
    mw = fmaxf(mw, s * weight_of(i, j, s, thr_i, __ldg(thr_b + j), pw_i, hard_mul));
    ms = fmaxf(ms, s);
  }
  mw = block_max(mw, red);
  ms = block_max(ms, red);
  float sw = 0.f, ss = 0.f, es = 0.f;
  for (int j = threadIdx.x; j < B; j += kLossThreads) {
    const float s = rowbuf[j];
    const float w = s * weight_of(i, j, s, thr_i, __ldg(thr_b + j), pw_i, hard_mul);
    sw += expf(w - mw);
    const float e = expf(s - ms);
    ss += e;
    es += e * s;
  }
  sw = block_sum(sw, red);
  ss = block_sum(ss, red);
  es = block_sum(es, red);
  if (threadIdx.x == 0) {
    lse_w[i] = mw + logf(sw);
    lse_s[i] = ms + logf(ss);
    exp_s[i] = es / ss;
    if (wdiag != nullptr) wdiag[i] = rowbuf[i] * pw_i;
  }
}

// loss = mean_i(lse_w_r - wdiag)/2 + mean_j(lse_w_c - wdiag)/2 + lambda/2 * (mean_i(lse_s_r - exp_s_r) + mean_j(lse_s_c - exp_s_c))
__global__ void __launch_bounds__(kLossThreads)
    finalize_kernel(int B, const float* __restrict__ lse_w_r, const float* __restrict__ lse_w_c, const float* __restrict__ lse_s_r,
                    const float* __restrict__ lse_s_c, const float* __restrict__ exp_s_r, const float* __restrict__ exp_s_c,
                    const float* __restrict__ wdiag, float lambda_reg, float* __restrict__ loss) {
  __shared__ double s[kLossThreads];
  double ce = 0.0, ent = 0.0;
  for (int i = threadIdx.x; i < B; i += kLossThreads) {
    ce += ((double)lse_w_r[i] - (double)wdiag[i]) + ((double)lse_w_c[i] - (double)wdiag[i]);
    ent += ((double)lse_s_r[i] - (double)exp_s_r[i]) + ((double)lse_s_c[i] - (double)exp_s_c[i]);
  }
  s[threadIdx.x] = 0.5 * ce + 0.5 * (double)lambda_reg * ent;
  __syncthreads();
  for (int o = kLossThreads / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss = (float)(s[0] / (double)B);
}

// dS_ij (already multiplied by the upstream gradient *go and by out_scale = 1/tau, the factor between the normalised
// dot products and S) and, when dpw != NULL, dpw_i = dL/dW_ii * S_ii.  Thread = 4 consecutive columns of one row.
__global__ void __launch_bounds__(kLossThreads)
    grad_kernel(const float* __restrict__ S, int B, int64_t ld, const float* __restrict__ thr_r, const float* __restrict__ thr_c,
                const float* __restrict__ pw, float hard_mul, const float* __restrict__ lse_w_r, const float* __restrict__ lse_w_c,
                const float* __restrict__ lse_s_r, const float* __restrict__ lse_s_c, const float* __restrict__ exp_s_r,
                const float* __restrict__ exp_s_c, float lambda_reg, float out_scale, const float* __restrict__ go_p,
                float* __restrict__ dS, int64_t ld_out, float* __restrict__ dpw) {
  const int i = blockIdx.y;
  const float go = __ldg(go_p);
  const float inv2b = 0.5f / (float)B;
  const float thr_i = __ldg(thr_r + i);
  const float pw_i = pw != nullptr ? __ldg(pw + i) : 1.f;
  const float lwr = __ldg(lse_w_r + i), lsr = __ldg(lse_s_r + i);
  const float h_r = lsr - __ldg(exp_s_r + i);
  for (int j = blockIdx.x * kLossThreads + threadIdx.x; j < B; j += gridDim.x * kLossThreads) {
    const float s = __ldg(S + (int64_t)i * ld + j);
    const float c = weight_of(i, j, s, thr_i, __ldg(thr_c + j), pw_i, hard_mul);
    const float w = s * c;
    const float lsc = __ldg(lse_s_c + j);
    const float h_c = lsc - __ldg(exp_s_c + j);
    const float delta = (i == j) ? 2.f : 0.f;
    const float dce = expf(w - lwr) + expf(w - __ldg(lse_w_c + j)) - delta;   // dL/dW_ij * 2B
    const float lpr = s - lsr, lpc = s - lsc;
    const float dent = -(expf(lpr) * (lpr + h_r)) - (expf(lpc) * (lpc + h_c));  // d(H_r + H_c)/dS_ij * B
    dS[(int64_t)i * ld_out + j] = go * out_scale * inv2b * (c * dce + lambda_reg * dent);
    if (i == j && dpw != nullptr) dpw[i] = go * inv2b * dce * s;
  }
}

}  // namespace atq

using namespace atq;

extern "C" {

int atq_rowkth_largest(int device, const float* s, int64_t b, int64_t ld, int64_t k, float* thr_out, atq_stream_t stream_) {
  ATQ_CHECK_ARG(s && thr_out && b >= 2 && ld >= b && k >= 1 && k <= b - 1, "needs b >= 2, ld >= b, 1 <= k <= b - 1");
  ATQ_CHECK_ARG(b <= 49152, "row longer than the shared-memory select supports (49 152)");
  ATQ_ENSURE_DEVICE(device);
  const size_t smem = (size_t)b * sizeof(uint32_t);
  static bool attr_done[64] = {false};
  if (smem > 48 * 1024 && !attr_done[device & 63]) {
    ATQ_CUDA(cudaFuncSetAttribute(rowkth_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152 * (int)sizeof(uint32_t)));
    attr_done[device & 63] = true;
  }
  rowkth_kernel<<<(unsigned)b, kLossThreads, smem, (cudaStream_t)stream_>>>(s, (int)b, ld, (int)k, thr_out);
  ATQ_LAUNCH_CHECK();
  return ATQ_OK;
}

int atq_infonce_row_stats(int device, const float* m, int64_t b, int64_t ld, const float* thr_a, const float* thr_b, const float* pw,
                          float hard_mul, float* lse_w, float* lse_s, float* exp_s, float* wdiag, atq_stream_t stream_) {
  ATQ_CHECK_ARG(m && thr_a && thr_b && lse_w && lse_s && exp_s && b >= 1 && ld >= b, "null pointer or bad shape");
  ATQ_CHECK_ARG(b <= 49152, "row longer than the shared-memory row buffer supports (49 152)");
  ATQ_ENSURE_DEVICE(device);
  const size_t smem = (size_t)b * sizeof(float);
  static bool attr_done[64] = {false};
  if (smem > 48 * 1024 && !attr_done[device & 63]) {
    ATQ_CUDA(cudaFuncSetAttribute(row_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152 * (int)sizeof(float)));
    attr_done[device & 63] = true;
  }
  row_stats_kernel<<<(unsigned)b, kLossThreads, smem, (cudaStream_t)stream_>>>(m, (int)b, ld, thr_a, thr_b, pw, hard_mul, lse_w, lse_s,
                                                                                 exp_s, wdiag);
  ATQ_LAUNCH_CHECK();
  return ATQ_OK;
}

int atq_infonce_finalize(int device, int64_t b, const float* lse_w_r, const float* lse_w_c, const float* lse_s_r, const float* lse_s_c,
                         const float* exp_s_r, const float* exp_s_c, const float* wdiag, float lambda_reg, float* loss,
                         atq_stream_t stream_) {
  ATQ_CHECK_ARG(lse_w_r && lse_w_c && lse_s_r && lse_s_c && exp_s_r && exp_s_c && wdiag && loss && b >= 1, "null pointer or b < 1");
  ATQ_ENSURE_DEVICE(device);
  finalize_kernel<<<1, kLossThreads, 0, (cudaStream_t)stream_>>>((int)b, lse_w_r, lse_w_c, lse_s_r, lse_s_c, exp_s_r, exp_s_c, wdiag,
                                                                   lambda_reg, loss);
  ATQ_LAUNCH_CHECK();
  return ATQ_OK;
}

int atq_infonce_grad(int device, const float* s, int64_t b, int64_t ld, const float* thr_r, const float* thr_c, const float* pw,
                     float hard_mul, const float* lse_w_r, const float* lse_w_c, const float* lse_s_r, const float* lse_s_c,
                     const float* exp_s_r, const float* exp_s_c, float lambda_reg, float out_scale, const float* grad_out,
                     float* ds, int64_t ld_out, float* dpw, atq_stream_t stream_) {
  ATQ_CHECK_ARG(s && thr_r && thr_c && lse_w_r && lse_w_c && lse_s_r && lse_s_c && exp_s_r && exp_s_c && grad_out && ds,
                "null pointer");
  ATQ_CHECK_ARG(b >= 1 && ld >= b && ld_out >= b && b <= 65535, "bad shape (b <= 65535)");
  ATQ_ENSURE_DEVICE(device);
  dim3 grid((unsigned)((b + kLossThreads - 1) / kLossThreads), (unsigned)b);
  if (grid.x > 16) grid.x = 16;
  grad_kernel<<<grid, kLossThreads, 0, (cudaStream_t)stream_>>>(s, (int)b, ld, thr_r, thr_c, pw, hard_mul, lse_w_r, lse_w_c, lse_s_r, lse_s_c,
                                                                 exp_s_r, exp_s_c, lambda_reg, out_scale, grad_out, ds, ld_out, dpw);
  ATQ_LAUNCH_CHECK();
  return ATQ_OK;
}

}  // extern "C"
