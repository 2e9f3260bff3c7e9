// K2: exact k-th order statistic of |x| -- the per-layer adaptive threshold
// (atq/quantizers.py:25-32: torch.sort(|W|).values[int(s*n)]) and the routing percentile
// (atq/routing.py:46-50: torch.kthvalue).
//
// |x| as a uint32 bit pattern is order-preserving, so the k-th smallest value is found by a
// most-significant-digit radix SELECT: three grid-level histogram reductions (11 + 10 + 10
// bits).  Each CTA histograms its grid-stride slice in shared memory (warp-shuffle free,
// shared-memory atomics), flushes non-empty bins to the layer's global histogram, and the
// last CTA to finish (ticket counter) scans the histogram, fixes the next digit of the
// answer and re-arms the state for the next pass.  No host synchronisation; the result is an
// element of |x|, so it is bit-identical to the reference's sorted[k].
//
// The kernels are natively batched: blockIdx.y selects one of up to kMaxBatch independent
// layers per launch, so a whole model's thresholds cost three launches, not three per layer.
#include "common.cuh"

namespace atq {

constexpr int kSelThreads = 512;
constexpr int kMaxBatch = 64;
constexpr int kBins0 = 2048, kBins12 = 1024;

struct SelectState {
  unsigned long long hist[kBins0];
  unsigned long long k_rem;
  unsigned int prefix;
  unsigned int blocks_done;
  unsigned int pad[4];
};

struct SelectBatch {
  const float* x[kMaxBatch];
  long long n[kMaxBatch];
  long long k[kMaxBatch];
  float* thr_out[kMaxBatch];
  SelectState* states;  // [count]
};

__global__ void __launch_bounds__(kSelThreads) select_init_kernel(SelectBatch b) {
  SelectState* st = b.states + blockIdx.x;
  for (int i = threadIdx.x; i < kBins0; i += blockDim.x) st->hist[i] = 0ull;
  if (threadIdx.x == 0) {
    st->k_rem = (unsigned long long)b.k[blockIdx.x];
    st->prefix = 0u;
    st->blocks_done = 0u;
  }
}

template <int PASS>
__device__ __forceinline__ void hist_one(uint32_t* sh, float v, uint32_t prefix) {
  const uint32_t u = __float_as_uint(v) & 0x7fffffffu;
  if constexpr (PASS == 0) {
    atomicAdd(&sh[u >> 20], 1u);
  } else if constexpr (PASS == 1) {
    if ((u >> 20) == (prefix >> 20)) atomicAdd(&sh[(u >> 10) & 0x3ffu], 1u);
  } else {
    if ((u >> 10) == (prefix >> 10)) atomicAdd(&sh[u & 0x3ffu], 1u);
  }
}

template <int PASS>
__global__ void __launch_bounds__(kSelThreads) select_pass_kernel(SelectBatch b) {
  constexpr int NB = (PASS == 0) ? kBins0 : kBins12;
  constexpr int SHIFT = (PASS == 0) ? 20 : (PASS == 1 ? 10 : 0);
  constexpr int PER_T = NB / kSelThreads;  // 4 or 2
  __shared__ uint32_t sh[NB];
  __shared__ unsigned long long s_warp_tot[kSelThreads / 32];
  __shared__ int s_last;

  const int layer = blockIdx.y;
  const float* __restrict__ x = b.x[layer];
  const long long n = b.n[layer];
  SelectState* st = b.states + layer;
  const uint32_t prefix = (PASS == 0) ? 0u : st->prefix;  // written by the previous launch

  for (int i = threadIdx.x; i < NB; i += kSelThreads) sh[i] = 0u;
  __syncthreads();

  // head (unaligned) + body (float4) + tail
  const uintptr_t addr = reinterpret_cast<uintptr_t>(x);
  long long head = (long long)(((16u - (unsigned)(addr & 15u)) & 15u) >> 2);
  if (head > n) head = n;
  const long long n4 = (n - head) >> 2;
  const float* xb = x + head;
  const long long stride = (long long)gridDim.x * kSelThreads;
  const long long tid = (long long)blockIdx.x * kSelThreads + threadIdx.x;
  for (long long g0 = tid; g0 < n4; g0 += stride * 4) {
    float4 v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      long long g = g0 + j * stride;
      if (g < n4) v[j] = (PASS == 0) ? ldg_stream4(xb + 4 * g) : __ldg(reinterpret_cast<const float4*>(xb + 4 * g));
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      long long g = g0 + j * stride;
      if (g < n4) {
        hist_one<PASS>(sh, v[j].x, prefix);
        hist_one<PASS>(sh, v[j].y, prefix);
        hist_one<PASS>(sh, v[j].z, prefix);
        hist_one<PASS>(sh, v[j].w, prefix);
      }
    }
  }
  if (blockIdx.x == 0) {
    for (long long i = threadIdx.x; i < head; i += kSelThreads) hist_one<PASS>(sh, x[i], prefix);
    for (long long i = head + n4 * 4 + threadIdx.x; i < n; i += kSelThreads) hist_one<PASS>(sh, x[i], prefix);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NB; i += kSelThreads) {
    uint32_t c = sh[i];
    if (c) atomicAdd(&st->hist[i], (unsigned long long)c);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int ticket = atomicAdd(&st->blocks_done, 1u);
    s_last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();

  // last CTA of this layer: locate the bin holding rank k_rem
  unsigned long long c[PER_T], tsum = 0ull;
#pragma unroll
  for (int j = 0; j < PER_T; ++j) {
    c[j] = __ldcg(&st->hist[threadIdx.x * PER_T + j]);
    st->hist[threadIdx.x * PER_T + j] = 0ull;  // re-arm for the next pass
    tsum += c[j];
  }
  unsigned long long incl = tsum;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp_tot[wid] = incl;
  __syncthreads();
  unsigned long long base = 0ull;
  for (int i = 0; i < wid; ++i) base += s_warp_tot[i];
  unsigned long long excl = base + incl - tsum;
  const unsigned long long k = __ldcg(&st->k_rem);
  __syncthreads();
  if (excl <= k && k < excl + tsum) {
    unsigned long long cum = excl;
#pragma unroll
    for (int j = 0; j < PER_T; ++j) {
      if (k < cum + c[j]) {
        const uint32_t bin = threadIdx.x * PER_T + j;
        const uint32_t np = prefix | (bin << SHIFT);
        st->prefix = np;
        st->k_rem = k - cum;
        if constexpr (PASS == 2) *b.thr_out[layer] = __uint_as_float(np);
        break;
      }
      cum += c[j];
    }
  }
  if (threadIdx.x == 0) st->blocks_done = 0u;
}

// out-of-range branches of the reference's threshold stage (atq/quantizers.py:33-38)
struct AbsStatsView { double sum_abs; unsigned int max_bits; unsigned int pad; };
__global__ void threshold_from_stats_kernel(const AbsStatsView* as, long long n, int use_max, float factor, float* thr_out) {
  if (use_max) *thr_out = __uint_as_float(as->max_bits) + 1.0f;
  else *thr_out = factor * (float)(as->sum_abs / (double)n);
}

static inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace atq

using namespace atq;

static int select_batched_impl(int device, int count, const float* const* xs, const int64_t* ns, const int64_t* ks,
                               float* const* thr_outs, void* ws, cudaStream_t stream) {
  // all entries here satisfy 0 <= k < n
  SelectState* states = reinterpret_cast<SelectState*>(ws);
  for (int base = 0; base < count; base += kMaxBatch) {
    const int cnt = (count - base < kMaxBatch) ? (count - base) : kMaxBatch;
    SelectBatch b;
    memset(&b, 0, sizeof(b));
    int64_t max_n = 0;
    for (int i = 0; i < cnt; ++i) {
      b.x[i] = xs[base + i];
      b.n[i] = ns[base + i];
      b.k[i] = ks[base + i];
      b.thr_out[i] = thr_outs[base + i];
      if (ns[base + i] > max_n) max_n = ns[base + i];
    }
    b.states = states + base;
    // CTAs per layer: 16 elements per thread per trip, capped so that the whole launch is
    // about 8 CTAs per SM
    int64_t need = (max_n + (int64_t)kSelThreads * 16 - 1) / ((int64_t)kSelThreads * 16);
    int64_t cap = ((int64_t)sm_count(device) * 8 + cnt - 1) / cnt;
    if (cap < 1) cap = 1;
    int gx = (int)(need < cap ? need : cap);
    if (gx < 1) gx = 1;
    dim3 grid((unsigned)gx, (unsigned)cnt);
    select_init_kernel<<<cnt, kSelThreads, 0, stream>>>(b);
    select_pass_kernel<0><<<grid, kSelThreads, 0, stream>>>(b);
    select_pass_kernel<1><<<grid, kSelThreads, 0, stream>>>(b);
    select_pass_kernel<2><<<grid, kSelThreads, 0, stream>>>(b);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
      set_error("select: kernel launch failed: %s", cudaGetErrorString(e));
      return ATQ_ECUDA;
    }
    note_launch(4);
  }
  return ATQ_OK;
}

extern "C" {

size_t atq_workspace_bytes_select_kth_abs(int64_t) { return align256(sizeof(SelectState)); }

int atq_select_kth_abs(int device, const float* x, int64_t n, int64_t k, float* thr_out, void* ws, size_t ws_bytes,
                       atq_stream_t stream) {
  ATQ_CHECK_ARG(x && thr_out && n > 0, "null pointer or n <= 0");
  ATQ_CHECK_ARG(k >= 0 && k < n, "k out of range");
  ATQ_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 3u) == 0, "x must be 4-byte aligned");
  if (ws == nullptr || ws_bytes < atq_workspace_bytes_select_kth_abs(n)) {
    set_error("atq_select_kth_abs: workspace too small");
    return ATQ_EWORKSPACE;
  }
  ATQ_ENSURE_DEVICE(device);
  return select_batched_impl(device, 1, &x, &n, &k, &thr_out, ws, (cudaStream_t)stream);
}

size_t atq_workspace_bytes_adaptive_threshold(int64_t n) { return atq_workspace_bytes_select_kth_abs(n); }

int atq_adaptive_threshold(int device, const float* w, int64_t n, int64_t k, float threshold_factor, float* thr_out,
                           void* ws, size_t ws_bytes, atq_stream_t stream) {
  const float* const xs[1] = {w};
  float* const ts[1] = {thr_out};
  return atq_adaptive_threshold_batched(device, 1, xs, &n, &k, threshold_factor, ts, ws, ws_bytes, stream);
}

size_t atq_workspace_bytes_adaptive_threshold_batched(int count, const int64_t*) {
  return align256(sizeof(SelectState)) * (size_t)(count > 0 ? count : 1);
}

int atq_adaptive_threshold_batched(int device, int count, const float* const* w_ptrs, const int64_t* ns,
                                   const int64_t* ks, float threshold_factor, float* const* thr_ptrs, void* ws,
                                   size_t ws_bytes, atq_stream_t stream_) {
  ATQ_CHECK_ARG(count > 0 && w_ptrs && ns && ks && thr_ptrs, "null pointer or count <= 0");
  if (ws == nullptr || ws_bytes < atq_workspace_bytes_adaptive_threshold_batched(count, ns)) {
    set_error("atq_adaptive_threshold_batched: workspace too small");
    return ATQ_EWORKSPACE;
  }
  ATQ_ENSURE_DEVICE(device);
  cudaStream_t stream = (cudaStream_t)stream_;
  // in-range layers go through the batched select; the two edge branches use |W| statistics
  const float* xs[kMaxBatch];
  int64_t n2[kMaxBatch], k2[kMaxBatch];
  float* t2[kMaxBatch];
  int m = 0, slot = 0;
  char* wsb = reinterpret_cast<char*>(ws);
  const size_t st_sz = align256(sizeof(SelectState));
  for (int i = 0; i < count; ++i) {
    ATQ_CHECK_ARG(w_ptrs[i] && thr_ptrs[i] && ns[i] > 0, "null layer pointer or empty layer");
    ATQ_CHECK_ARG((reinterpret_cast<uintptr_t>(w_ptrs[i]) & 3u) == 0, "weights must be 4-byte aligned");
    if (ks[i] > 0 && ks[i] < ns[i]) {
      xs[m] = w_ptrs[i]; n2[m] = ns[i]; k2[m] = ks[i]; t2[m] = thr_ptrs[i];
      if (++m == kMaxBatch) {
        int r = select_batched_impl(device, m, xs, n2, k2, t2, wsb + (size_t)slot * st_sz, stream);
        if (r != ATQ_OK) return r;
        slot += m; m = 0;
      }
    } else {
      void* stats = wsb + (size_t)slot * st_sz;  // 16 bytes of this layer's slot
      ++slot;
      int r = atq_abs_stats(device, w_ptrs[i], ns[i], stats, nullptr, 0, stream_);
      if (r != ATQ_OK) return r;
      threshold_from_stats_kernel<<<1, 1, 0, stream>>>((const AbsStatsView*)stats, (long long)ns[i], ks[i] >= ns[i] ? 1 : 0,
                                                        threshold_factor, thr_ptrs[i]);
      ATQ_LAUNCH_CHECK();
    }
  }
  if (m > 0) {
    int r = select_batched_impl(device, m, xs, n2, k2, t2, wsb + (size_t)slot * st_sz, stream);
    if (r != ATQ_OK) return r;
  }
  return ATQ_OK;
}

}  // extern "C"
