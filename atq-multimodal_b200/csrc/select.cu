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

// Sampling front-end for large layers (n >= kSampleMinN): a sorted 8192-element sample brackets the
// target rank, ONE streaming pass counts the elements below the bracket and compacts the few
// percent inside it, and the radix select then runs on that L2-resident candidate list.  The
// result is verified (rank must fall inside the candidates); if the bracket missed -- adversarial
// data, massive ties -- a device-side flag routes to the full three-pass select.  Exactness is
// therefore unconditional; only the cost is data dependent.
constexpr int kSampleN = 8192;
constexpr long long kSampleMinN = 1ll << 25;  // measured crossover: 74 us (plain) vs 85 us (sampled) at 16M, 2x faster at 67M
constexpr int kChunk = 8192;      // elements per filter CTA
constexpr int kFilterThreads = 256;
constexpr int kLocalCap = 2048;   // shared-memory candidate slots per filter CTA

struct SampleState {
  unsigned int lo, hi;             // candidate bracket on |x| bit patterns, inclusive
  unsigned int n_cand;             // candidates appended
  unsigned int fallback;           // 1 => the bracket missed; the full radix select decides
  unsigned long long count_below;  // #{u < lo}
  unsigned long long cap;          // capacity of the candidate buffer
};

enum { MODE_PLAIN = 0, MODE_CAND = 1, MODE_FALLBACK = 2 };

struct SelectBatch {
  const float* x[kMaxBatch];
  long long n[kMaxBatch];
  long long k[kMaxBatch];
  float* thr_out[kMaxBatch];
  SelectState* states;  // [count]
  SampleState* ss;      // MODE_CAND / MODE_FALLBACK only (single layer)
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
__device__ __forceinline__ void hist_one(uint32_t* sh, float v, uint32_t prefix, uint32_t bias = 0u) {
  const uint32_t u = (__float_as_uint(v) & 0x7fffffffu) - bias;
  if constexpr (PASS == 0) {
    atomicAdd(&sh[u >> 20], 1u);
  } else if constexpr (PASS == 1) {
    if ((u >> 20) == (prefix >> 20)) atomicAdd(&sh[(u >> 10) & 0x3ffu], 1u);
  } else {
    if ((u >> 10) == (prefix >> 10)) atomicAdd(&sh[u & 0x3ffu], 1u);
  }
}

template <int PASS, int MODE>
__global__ void __launch_bounds__(kSelThreads) select_pass_kernel(SelectBatch b) {
  constexpr int NB = (PASS == 0) ? kBins0 : kBins12;
  constexpr int SHIFT = (PASS == 0) ? 20 : (PASS == 1 ? 10 : 0);
  constexpr int PER_T = NB / kSelThreads;  // 4 or 2
  __shared__ uint32_t sh[NB];
  __shared__ unsigned long long s_warp_tot[kSelThreads / 32];
  __shared__ int s_last;

  const int layer = blockIdx.y;
  const float* __restrict__ x = b.x[layer];
  long long n = b.n[layer];
  SelectState* st = b.states + layer;
  uint32_t bias = 0u;
  if constexpr (MODE == MODE_CAND) {
    if (b.ss->fallback) return;  // bracket already known to have missed
    const unsigned long long nc = b.ss->n_cand;
    n = (long long)(nc < b.ss->cap ? nc : b.ss->cap);
    bias = b.ss->lo;  // keys are taken relative to the bracket's lower end
  }
  if constexpr (MODE == MODE_FALLBACK) {
    if (!b.ss->fallback) return;
  }
  const uint32_t prefix = (PASS == 0) ? 0u : st->prefix;  // written by the previous launch
  bool skip_scan = false;
  if constexpr (MODE == MODE_CAND && PASS == 0) {
    // keys are |x| - lo: if the bracket spans fewer than 2^20 values every key's top digit is 0
    skip_scan = (b.ss->hi - b.ss->lo) < (1u << 20);
  }

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
  if (skip_scan) {
    if (blockIdx.x == 0 && threadIdx.x == 0) sh[0] = (uint32_t)n;  // n <= cap < 2^32
  }
  for (long long g0 = tid; g0 < (skip_scan ? 0 : n4); g0 += stride * 4) {
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
        hist_one<PASS>(sh, v[j].x, prefix, bias);
        hist_one<PASS>(sh, v[j].y, prefix, bias);
        hist_one<PASS>(sh, v[j].z, prefix, bias);
        hist_one<PASS>(sh, v[j].w, prefix, bias);
      }
    }
  }
  if (blockIdx.x == 0 && !skip_scan) {
    for (long long i = threadIdx.x; i < head; i += kSelThreads) hist_one<PASS>(sh, x[i], prefix, bias);
    for (long long i = head + n4 * 4 + threadIdx.x; i < n; i += kSelThreads) hist_one<PASS>(sh, x[i], prefix, bias);
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
  unsigned long long k = __ldcg(&st->k_rem);
  if constexpr (MODE == MODE_CAND && PASS == 0) {
    // rank inside the candidate list; outside => the bracket missed, hand over to the full select
    const long long kc = b.k[layer] - (long long)__ldcg(&b.ss->count_below);
    const bool overflow = __ldcg(&b.ss->n_cand) > b.ss->cap;
    if (kc < 0 || kc >= n || overflow) {
      if (threadIdx.x == 0) {
        b.ss->fallback = 1u;
        st->blocks_done = 0u;
      }
      return;
    }
    k = (unsigned long long)kc;
  }
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
        if constexpr (PASS == 2) *b.thr_out[layer] = __uint_as_float(np + bias);
        break;
      }
      cum += c[j];
    }
  }
  if (threadIdx.x == 0) st->blocks_done = 0u;
}

// ---- sampling front-end ----------------------------------------------------------------------
// Bracket from a strided sample WITHOUT sorting it: every CTA holds the whole 8192-element sample
// in shared memory and ranks 8 "pivot" elements (every 16th sample) by counting, one warp per
// pivot (ties broken by index, so ranks are distinct).  The 512 (rank, key) pairs go to a small
// table; the filter CTAs pick the tightest pivots around the target rank from it.
constexpr int kPivots = 512;
constexpr int kPivotsPerCta = 8;
struct PivotTable {
  unsigned int rank[kPivots];
  unsigned int key[kPivots];
  unsigned int sample[kSampleN];
};
// one strided load per thread (the sample positions are 32 KB or more apart: pure DRAM latency)
__global__ void __launch_bounds__(256) sample_gather_kernel(const float* __restrict__ x, long long n, uint32_t* __restrict__ sample) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  const long long stride = n / kSampleN;
  sample[i] = __float_as_uint(__ldg(x + (long long)i * stride + (stride >> 1))) & 0x7fffffffu;
}

__global__ void __launch_bounds__(256)
    sample_pivots_kernel(const uint32_t* __restrict__ sample, long long n, long long k, SampleState* ss, SelectState* st_full,
                         SelectState* st_cand, PivotTable* pt, unsigned long long cap) {
  __shared__ __align__(16) uint32_t keys[kSampleN];
  {
    const uint4* src = reinterpret_cast<const uint4*>(sample);
    uint4* dst = reinterpret_cast<uint4*>(keys);
#pragma unroll
    for (int j = 0; j < kSampleN / 4 / 256; ++j) dst[threadIdx.x + 256 * j] = __ldg(src + threadIdx.x + 256 * j);
  }
  if (blockIdx.x == 0) {
    for (int i = threadIdx.x; i < kBins0; i += 256) {
      st_full->hist[i] = 0ull;
      st_cand->hist[i] = 0ull;
    }
    if (threadIdx.x == 0) {
      ss->n_cand = 0u;
      ss->fallback = 0u;
      ss->count_below = 0ull;
      ss->cap = cap;
      st_full->k_rem = (unsigned long long)k;
      st_full->prefix = 0u;
      st_full->blocks_done = 0u;
      st_cand->k_rem = 0ull;
      st_cand->prefix = 0u;
      st_cand->blocks_done = 0u;
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int p = blockIdx.x * kPivotsPerCta + wid;        // pivot id
  const int e = p * (kSampleN / kPivots) + (kSampleN / kPivots) / 2;  // its index in the sample
  const uint32_t me = keys[e];
  int rank = 0;
#pragma unroll 8
  for (int i = lane; i < kSampleN; i += 32) {
    const uint32_t v = keys[i];
    rank += (v < me) || (v == me && i < e);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) rank += __shfl_xor_sync(0xffffffffu, rank, o);
  if (lane == 0) {
    pt->rank[p] = (unsigned)rank;
    pt->key[p] = me;
  }
}

// bracket ends from the pivot table: lo = key of the highest-ranked pivot with rank <= target-delta
// (0 if none), hi = key of the lowest-ranked pivot with rank >= target+delta (max if none).
// Called by a whole CTA of >= 256 threads; result broadcast through shared memory.
__device__ __forceinline__ void bracket_from_pivots(const PivotTable* pt, long long n, long long k, uint32_t* s_lohi,
                                                    uint32_t& lo, uint32_t& hi) {
  const double q = (double)k / (double)n;
  const long long r = (long long)(q * kSampleN);
  const long long delta = (long long)ceil(5.5 * sqrt((double)kSampleN * q * (1.0 - q))) + 2;
  const long long lo_t = r - delta, hi_t = r + delta + 1;
  if (threadIdx.x == 0) {
    s_lohi[0] = 0u;           // max over candidates for lo
    s_lohi[1] = 0x7fffffffu;  // min over candidates for hi
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kPivots; i += blockDim.x) {
    const long long rk = (long long)pt->rank[i];
    const uint32_t key = pt->key[i];
    if (rk <= lo_t) atomicMax(&s_lohi[0], key);  // keys are monotone in rank
    if (rk >= hi_t) atomicMin(&s_lohi[1], key);
  }
  __syncthreads();
  lo = s_lohi[0];
  hi = s_lohi[1];
}

// one streaming pass: count |x| below the bracket, compact |x| inside it (x 16-byte aligned).
// Each warp owns a private shared-memory segment and keeps its fill count in a register
// (ballot + popc), so the hot loop has no atomics at all; one global atomic per CTA publishes.
constexpr int kWarpCap = kLocalCap / (kFilterThreads / 32);  // 256 slots per warp
__global__ void __launch_bounds__(kFilterThreads, 4)
    filter_kernel(const float* __restrict__ x, long long n, long long k, SampleState* ss, const PivotTable* pt,
                  uint32_t* __restrict__ cand) {
  __shared__ uint32_t buf[kLocalCap + kFilterThreads / 32];
  __shared__ unsigned int s_wcnt[kFilterThreads / 32], s_woff[kFilterThreads / 32], s_gbase;
  __shared__ unsigned long long s_below[kFilterThreads / 32];
  __shared__ uint32_t s_lohi[2];
  uint32_t lo, hi;
  bracket_from_pivots(pt, n, k, s_lohi, lo, hi);
  if (blockIdx.x == 0 && threadIdx.x == 0) {  // the candidate passes read the bracket from here
    ss->lo = lo;
    ss->hi = hi;
  }
  const long long start = (long long)blockIdx.x * kChunk;
  const long long end = (start + kChunk < n) ? start + kChunk : n;
  const int nvec = (int)((end - start) >> 2);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t* wbuf = buf + wid * kWarpCap;
  // Two branch-free sweeps over the 32 values this thread holds in registers: count, warp-scan the
  // counts (5 shuffles per thread instead of a ballot + popc per element), then scatter with
  // unconditional stores (non-candidates go to a dump slot).
  constexpr int kPerThread = kChunk / 4 / kFilterThreads;
  const uint32_t span = hi - lo;  // lo <= u <= hi  <=>  (u - lo) <= span   (hi >= lo always)
  const bool full = (nvec == kChunk / 4);  // CTA-uniform: every slot of the chunk is real
  uint32_t u[4 * kPerThread];
#pragma unroll
  for (int j = 0; j < kPerThread; ++j) {
    const int g = (threadIdx.x & ~31) + j * kFilterThreads + lane;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (full || g < nvec) v = ldg_stream4(x + start + 4ll * g);
    u[4 * j + 0] = __float_as_uint(v.x) & 0x7fffffffu;
    u[4 * j + 1] = __float_as_uint(v.y) & 0x7fffffffu;
    u[4 * j + 2] = __float_as_uint(v.z) & 0x7fffffffu;
    u[4 * j + 3] = __float_as_uint(v.w) & 0x7fffffffu;
  }
  if (!full) {  // park the out-of-range slots above every bracket and every `lo`
#pragma unroll
    for (int j = 0; j < kPerThread; ++j) {
      const int g = (threadIdx.x & ~31) + j * kFilterThreads + lane;
      if (g >= nvec) { u[4 * j] = u[4 * j + 1] = u[4 * j + 2] = u[4 * j + 3] = 0xffffffffu; }
    }
  }
  unsigned int below = 0u, mine = 0u;
#pragma unroll
  for (int i = 0; i < 4 * kPerThread; ++i) {
    below += (u[i] < lo) ? 1u : 0u;
    mine += (u[i] - lo <= span) ? 1u : 0u;
  }
  uint32_t tail_u = 0xffffffffu;
  if (wid == 0) {  // ragged tail of the last chunk (at most 3 elements)
    const long long i = start + 4ll * nvec + lane;
    if ((lane < 4) && i < end) tail_u = __float_as_uint(__ldg(x + i)) & 0x7fffffffu;
    below += (tail_u < lo) ? 1u : 0u;
    mine += (tail_u - lo <= span) ? 1u : 0u;
  }
  unsigned incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  const unsigned wcount = __shfl_sync(0xffffffffu, incl, 31);
  if (wcount) {  // warp-uniform
    unsigned pos = incl - mine;
    const uint32_t wbase = (uint32_t)__cvta_generic_to_shared(wbuf);
    auto scatter = [&](uint32_t val) {
      // predicated shared store, no branch: p = in-bracket && slot available
      const uint32_t c = (val - lo <= span) ? 1u : 0u;
      const uint32_t ok = (pos < (unsigned)kWarpCap) ? c : 0u;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.shared.u32 [%0], %1;\n\t}"
          ::"r"(wbase + 4u * pos), "r"(val), "r"(ok)
          : "memory");
      pos += c;
    };
#pragma unroll
    for (int i = 0; i < 4 * kPerThread; ++i) scatter(u[i]);
    if (wid == 0) scatter(tail_u);
  }
  unsigned long long below64 = warp_sum((unsigned long long)below);
  if (lane == 0) {
    s_below[wid] = below64;
    s_wcnt[wid] = wcount;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0ull;
    unsigned total = 0;
    bool over = false;
    for (int i = 0; i < kFilterThreads / 32; ++i) {
      t += s_below[i];
      s_woff[i] = total;
      total += s_wcnt[i];
      over |= s_wcnt[i] > (unsigned)kWarpCap;
    }
    if (t) atomicAdd(&ss->count_below, t);
    if (over) {
      atomicExch(&ss->fallback, 1u);
      s_gbase = 0xffffffffu;
    } else {
      s_gbase = total ? atomicAdd(&ss->n_cand, total) : 0u;
    }
  }
  __syncthreads();
  const unsigned gbase = s_gbase;
  if (gbase == 0xffffffffu) return;
  const unsigned long long cap = ss->cap;
  const unsigned my = s_wcnt[wid], off = gbase + s_woff[wid];
  for (unsigned i = lane; i < my; i += 32)
    if ((unsigned long long)off + i < cap) cand[off + i] = wbuf[i];
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
    select_pass_kernel<0, MODE_PLAIN><<<grid, kSelThreads, 0, stream>>>(b);
    select_pass_kernel<1, MODE_PLAIN><<<grid, kSelThreads, 0, stream>>>(b);
    select_pass_kernel<2, MODE_PLAIN><<<grid, kSelThreads, 0, stream>>>(b);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
      set_error("select: kernel launch failed: %s", cudaGetErrorString(e));
      return ATQ_ECUDA;
    }
    note_launch(4);
  }
  return ATQ_OK;
}

static inline bool use_sampling(const float* x, int64_t n) {
  return n >= kSampleMinN && (reinterpret_cast<uintptr_t>(x) & 15u) == 0;
}
static inline unsigned long long cand_capacity(int64_t n) { return (unsigned long long)(n / 8 + 65536); }
static inline size_t big_layer_ws(int64_t n) {
  return 2 * align256(sizeof(SelectState)) + align256(sizeof(SampleState)) + align256(sizeof(PivotTable)) +
         align256((size_t)cand_capacity(n) * 4);
}

// large layer: sample -> filter -> select among candidates (-> full select only if the bracket missed)
static int select_sampled_impl(int device, const float* x, int64_t n, int64_t k, float* thr_out, void* ws,
                               cudaStream_t stream) {
  char* base = reinterpret_cast<char*>(ws);
  SelectState* st_full = reinterpret_cast<SelectState*>(base);
  SelectState* st_cand = reinterpret_cast<SelectState*>(base + align256(sizeof(SelectState)));
  SampleState* ss = reinterpret_cast<SampleState*>(base + 2 * align256(sizeof(SelectState)));
  PivotTable* pt = reinterpret_cast<PivotTable*>(base + 2 * align256(sizeof(SelectState)) + align256(sizeof(SampleState)));
  uint32_t* cand = reinterpret_cast<uint32_t*>(base + 2 * align256(sizeof(SelectState)) + align256(sizeof(SampleState)) +
                                               align256(sizeof(PivotTable)));
  const unsigned long long cap = cand_capacity(n);
  sample_gather_kernel<<<kSampleN / 256, 256, 0, stream>>>(x, (long long)n, pt->sample);
  sample_pivots_kernel<<<kPivots / kPivotsPerCta, 256, 0, stream>>>(pt->sample, (long long)n, (long long)k, ss, st_full, st_cand, pt, cap);
  const int64_t chunks = (n + kChunk - 1) / kChunk;
  filter_kernel<<<(unsigned)chunks, kFilterThreads, 0, stream>>>(x, (long long)n, (long long)k, ss, pt, cand);
  SelectBatch b;
  memset(&b, 0, sizeof(b));
  b.x[0] = reinterpret_cast<const float*>(cand);
  b.n[0] = (long long)cap;
  b.k[0] = k;
  b.thr_out[0] = thr_out;
  b.states = st_cand;
  b.ss = ss;
  int64_t need = ((int64_t)cap + (int64_t)kSelThreads * 16 - 1) / ((int64_t)kSelThreads * 16);
  int64_t capg = (int64_t)sm_count(device) * 4;
  int gx = (int)(need < capg ? need : capg);
  select_pass_kernel<0, MODE_CAND><<<dim3(gx, 1), kSelThreads, 0, stream>>>(b);
  select_pass_kernel<1, MODE_CAND><<<dim3(gx, 1), kSelThreads, 0, stream>>>(b);
  select_pass_kernel<2, MODE_CAND><<<dim3(gx, 1), kSelThreads, 0, stream>>>(b);
  SelectBatch f;
  memset(&f, 0, sizeof(f));
  f.x[0] = x;
  f.n[0] = n;
  f.k[0] = k;
  f.thr_out[0] = thr_out;
  f.states = st_full;
  f.ss = ss;
  need = (n + (int64_t)kSelThreads * 16 - 1) / ((int64_t)kSelThreads * 16);
  capg = (int64_t)sm_count(device) * 2;  // normally exits at once; keep the launch cheap
  gx = (int)(need < capg ? need : capg);
  select_pass_kernel<0, MODE_FALLBACK><<<dim3(gx, 1), kSelThreads, 0, stream>>>(f);
  select_pass_kernel<1, MODE_FALLBACK><<<dim3(gx, 1), kSelThreads, 0, stream>>>(f);
  select_pass_kernel<2, MODE_FALLBACK><<<dim3(gx, 1), kSelThreads, 0, stream>>>(f);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("sampled select: kernel launch failed: %s", cudaGetErrorString(e));
    return ATQ_ECUDA;
  }
  note_launch(9);
  return ATQ_OK;
}

extern "C" {

size_t atq_workspace_bytes_select_kth_abs(int64_t n) {
  return n >= kSampleMinN ? big_layer_ws(n) : align256(sizeof(SelectState));
}

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
  if (use_sampling(x, n)) return select_sampled_impl(device, x, n, k, thr_out, ws, (cudaStream_t)stream);
  return select_batched_impl(device, 1, &x, &n, &k, &thr_out, ws, (cudaStream_t)stream);
}

size_t atq_workspace_bytes_adaptive_threshold(int64_t n) { return atq_workspace_bytes_adaptive_threshold_batched(1, &n); }

int atq_adaptive_threshold(int device, const float* w, int64_t n, int64_t k, float threshold_factor, float* thr_out,
                           void* ws, size_t ws_bytes, atq_stream_t stream) {
  const float* const xs[1] = {w};
  float* const ts[1] = {thr_out};
  return atq_adaptive_threshold_batched(device, 1, xs, &n, &k, threshold_factor, ts, ws, ws_bytes, stream);
}

size_t atq_workspace_bytes_adaptive_threshold_batched(int count, const int64_t* ns) {
  // one SelectState slot per layer, then a private region per large (sampled) layer
  size_t total = align256(sizeof(SelectState)) * (size_t)(count > 0 ? count : 1);
  if (ns != nullptr)
    for (int i = 0; i < count; ++i)
      if (ns[i] >= kSampleMinN) total += big_layer_ws(ns[i]);
  return total;
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
  // small in-range layers go through the batched select, large ones through the sampling front-end,
  // the two edge branches use |W| statistics
  const float* xs[kMaxBatch];
  int64_t n2[kMaxBatch], k2[kMaxBatch];
  float* t2[kMaxBatch];
  int m = 0, slot = 0;
  char* wsb = reinterpret_cast<char*>(ws);
  const size_t st_sz = align256(sizeof(SelectState));
  size_t big_off = st_sz * (size_t)count;
  for (int i = 0; i < count; ++i) {
    ATQ_CHECK_ARG(w_ptrs[i] && thr_ptrs[i] && ns[i] > 0, "null layer pointer or empty layer");
    ATQ_CHECK_ARG((reinterpret_cast<uintptr_t>(w_ptrs[i]) & 3u) == 0, "weights must be 4-byte aligned");
    const bool in_range = ks[i] > 0 && ks[i] < ns[i];
    if (in_range && use_sampling(w_ptrs[i], ns[i])) {
      int r = select_sampled_impl(device, w_ptrs[i], ns[i], ks[i], thr_ptrs[i], wsb + big_off, stream);
      if (r != ATQ_OK) return r;
    } else if (in_range) {
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
    if (ns[i] >= kSampleMinN) big_off += big_layer_ws(ns[i]);
  }
  if (m > 0) {
    int r = select_batched_impl(device, m, xs, n2, k2, t2, wsb + (size_t)slot * st_sz, stream);
    if (r != ATQ_OK) return r;
  }
  return ATQ_OK;
}

}  // extern "C"
