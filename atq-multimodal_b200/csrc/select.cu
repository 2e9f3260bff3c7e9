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
// sampling pays off for one layer from ~32M elements (fixed cost: 10 small launches) and from ~4M
// elements when the call batches >= 64M elements in total (launches shared by all layers)
constexpr long long kSampleMinN = 1ll << 25;
constexpr long long kSampleMinNBatched = 1ll << 22;
constexpr long long kBatchedTotal = 1ll << 26;
constexpr int kFilterThreads = 256;
constexpr int kWarpChunk = 512;   // elements per warp step of the filter pass (4 float4 per lane)

struct SampleState {
  unsigned int lo, hi;             // candidate bracket on |x| bit patterns, inclusive
  unsigned int n_cand;             // candidate-buffer slots reserved (slabs; unused slots hold a sentinel)
  unsigned int fallback;           // 1 => the bracket missed; the full radix select decides
  unsigned int n_valid;            // real candidates among the reserved slots
  unsigned int pad0;
  unsigned long long count_below;  // #{u < lo}
  unsigned long long cap;          // capacity of the candidate buffer
};

enum { MODE_PLAIN = 0, MODE_CAND = 1, MODE_FALLBACK = 2 };

struct PivotTable;
struct SelectBatch {
  const float* x[kMaxBatch];
  long long n[kMaxBatch];
  long long k[kMaxBatch];
  float* thr_out[kMaxBatch];
  uint32_t* cand[kMaxBatch];  // sampled layers: candidate key buffers
  SelectState* states;        // [count]  plain passes / full (fallback) passes
  SelectState* cand_states;   // [count]  candidate passes
  SampleState* ss;            // [count]
  PivotTable* pt;             // [count]
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
  const float* __restrict__ x = (MODE == MODE_CAND) ? reinterpret_cast<const float*>(b.cand[layer]) : b.x[layer];
  long long n = b.n[layer];
  SelectState* st = (MODE == MODE_CAND ? b.cand_states : b.states) + layer;
  SampleState* ss = (MODE == MODE_PLAIN) ? nullptr : b.ss + layer;
  uint32_t bias = 0u;
  if constexpr (MODE == MODE_CAND) {
    if (ss->fallback) return;  // bracket already known to have missed
    const unsigned long long nc = ss->n_cand;
    n = (long long)(nc < ss->cap ? nc : ss->cap);
    bias = ss->lo;  // keys are taken relative to the bracket's lower end
  }
  if constexpr (MODE == MODE_FALLBACK) {
    if (!ss->fallback) return;
  }
  const uint32_t prefix = (PASS == 0) ? 0u : st->prefix;  // written by the previous launch
  bool skip_scan = false;
  if constexpr (MODE == MODE_CAND && PASS == 0) {
    // keys are |x| - lo: if the bracket spans fewer than 2^20 values every key's top digit is 0
    skip_scan = (ss->hi - ss->lo) < (1u << 20);
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
    const long long kc = b.k[layer] - (long long)__ldcg(&ss->count_below);
    const bool overflow = __ldcg(&ss->n_cand) > ss->cap;
    if (kc < 0 || kc >= (long long)__ldcg(&ss->n_valid) || overflow) {
      if (threadIdx.x == 0) {
        ss->fallback = 1u;
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

// ---- fused plain select: the three digit passes in ONE cooperative launch -----------------------
// The CTAs of a layer (blockIdx.y) meet at a layer-level barrier after each pass: the last CTA to flush its histogram
// fixes the next digit and releases the others (a cooperative launch guarantees that all of them are resident).  The
// second and third pass re-read the layer while it is still in L2 (layers up to ~64 MiB), instead of three launches that
// each start cold.  State (histogram, ticket, release flag) is zeroed by a memset node in front of the kernel.
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(unsigned int* p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int PASS>
__device__ __forceinline__ void fused_scan(uint32_t* sh, const float* __restrict__ x, long long n, uint32_t prefix) {
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
      const long long g = g0 + j * stride;
      if (g < n4) v[j] = __ldg(reinterpret_cast<const float4*>(xb + 4 * g));
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long g = g0 + j * stride;
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
}

// one pass: histogram, flush, layer barrier; returns the prefix with this pass's digit fixed (all threads)
template <int PASS>
__device__ __forceinline__ uint32_t fused_pass(const SelectBatch& b, int layer, SelectState* st, uint32_t prefix, uint32_t* sh,
                                               unsigned long long* s_warp_tot, int* s_flag, uint32_t* s_prefix) {
  constexpr int NB = (PASS == 0) ? kBins0 : kBins12;
  constexpr int SHIFT = (PASS == 0) ? 20 : (PASS == 1 ? 10 : 0);
  constexpr int PER_T = NB / kSelThreads;
  for (int i = threadIdx.x; i < NB; i += kSelThreads) sh[i] = 0u;
  __syncthreads();
  fused_scan<PASS>(sh, b.x[layer], b.n[layer], prefix);
  __syncthreads();
  for (int i = threadIdx.x; i < NB; i += kSelThreads) {
    const uint32_t c = sh[i];
    if (c) atomicAdd(&st->hist[i], (unsigned long long)c);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) *s_flag = atomicAdd(&st->blocks_done, 1u) == (unsigned)(PASS + 1) * gridDim.x - 1u;
  __syncthreads();
  if (*s_flag) {
    __threadfence();
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
      const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp_tot[wid] = incl;
    __syncthreads();
    unsigned long long base = 0ull;
    for (int i = 0; i < wid; ++i) base += s_warp_tot[i];
    const unsigned long long excl = base + incl - tsum;
    const unsigned long long k = PASS == 0 ? (unsigned long long)b.k[layer] : __ldcg(&st->k_rem);
    if (excl <= k && k < excl + tsum) {
      unsigned long long cum = excl;
#pragma unroll
      for (int j = 0; j < PER_T; ++j) {
        if (k < cum + c[j]) {
          const uint32_t np = prefix | ((uint32_t)(threadIdx.x * PER_T + j) << SHIFT);
          st->prefix = np;
          st->k_rem = k - cum;
          *s_prefix = np;
          if constexpr (PASS == 2) *b.thr_out[layer] = __uint_as_float(np);
          break;
        }
        cum += c[j];
      }
    }
    __threadfence();
    __syncthreads();
    if (PASS < 2 && threadIdx.x == 0) st_release_u32(&st->pad[0], (unsigned)(PASS + 1));
  } else if (PASS < 2) {
    if (threadIdx.x == 0) {
      while (ld_acquire_u32(&st->pad[0]) < (unsigned)(PASS + 1)) __nanosleep(64);
      *s_prefix = __ldcg(&st->prefix);
    }
    __syncthreads();
  }
  return PASS < 2 ? *s_prefix : 0u;
}

__global__ void __launch_bounds__(kSelThreads) select_fused_kernel(SelectBatch b) {
  __shared__ uint32_t sh[kBins0];
  __shared__ unsigned long long s_warp_tot[kSelThreads / 32];
  __shared__ int s_flag;
  __shared__ uint32_t s_prefix;
  const int layer = blockIdx.y;
  SelectState* st = b.states + layer;
  uint32_t prefix = fused_pass<0>(b, layer, st, 0u, sh, s_warp_tot, &s_flag, &s_prefix);
  __syncthreads();
  prefix = fused_pass<1>(b, layer, st, prefix, sh, s_warp_tot, &s_flag, &s_prefix);
  __syncthreads();
  fused_pass<2>(b, layer, st, prefix, sh, s_warp_tot, &s_flag, &s_prefix);
}

// ---- sampling front-end (batched: blockIdx.y = layer) -------------------------------------------
constexpr int kPivots = 512;
constexpr int kPivotsPerCta = 8;
struct PivotTable {
  unsigned int rank[kPivots];
  unsigned int key[kPivots];
  unsigned int sample[kSampleN];
};

// one strided load per thread (the sample positions are 2 KB or more apart: pure DRAM latency)
__global__ void __launch_bounds__(256) sample_gather_kernel(SelectBatch b) {
  const int layer = blockIdx.y;
  const int i = blockIdx.x * 256 + threadIdx.x;
  const long long n = b.n[layer];
  const long long stride = n / kSampleN;
  b.pt[layer].sample[i] = __float_as_uint(__ldg(b.x[layer] + (long long)i * stride + (stride >> 1))) & 0x7fffffffu;
}

// Every CTA holds the layer's whole 8192-element sample in shared memory and ranks 8 "pivot"
// elements (every 16th sample) by counting, one warp per pivot (ties broken by index, so ranks are
// distinct) -- no sort.  CTA 0 of each layer also arms the layer's state.
__global__ void __launch_bounds__(256) sample_pivots_kernel(SelectBatch b) {
  __shared__ __align__(16) uint32_t keys[kSampleN];
  const int layer = blockIdx.y;
  PivotTable* pt = b.pt + layer;
  {
    const uint4* src = reinterpret_cast<const uint4*>(pt->sample);
    uint4* dst = reinterpret_cast<uint4*>(keys);
#pragma unroll
    for (int j = 0; j < kSampleN / 4 / 256; ++j) dst[threadIdx.x + 256 * j] = __ldg(src + threadIdx.x + 256 * j);
  }
  if (blockIdx.x == 0) {
    SelectState* st_full = b.states + layer;
    SelectState* st_cand = b.cand_states + layer;
    for (int i = threadIdx.x; i < kBins0; i += 256) {
      st_full->hist[i] = 0ull;
      st_cand->hist[i] = 0ull;
    }
    if (threadIdx.x == 0) {
      SampleState* ss = b.ss + layer;
      ss->n_cand = 0u;
      ss->n_valid = 0u;
      ss->fallback = 0u;
      ss->count_below = 0ull;
      st_full->k_rem = (unsigned long long)b.k[layer];
      st_full->prefix = 0u;
      st_full->blocks_done = 0u;
      st_cand->k_rem = 0ull;
      st_cand->prefix = 0u;
      st_cand->blocks_done = 0u;
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int p = blockIdx.x * kPivotsPerCta + wid;                      // pivot id
  const int e = p * (kSampleN / kPivots) + (kSampleN / kPivots) / 2;   // its index in the sample
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
// (0 if none), hi = key of the lowest-ranked pivot with rank >= target+delta (max if none)
__global__ void __launch_bounds__(256) bracket_kernel(SelectBatch b) {
  __shared__ uint32_t s_lohi[2];
  const int layer = blockIdx.x;
  const PivotTable* pt = b.pt + layer;
  const double q = (double)b.k[layer] / (double)b.n[layer];
  const long long r = (long long)(q * kSampleN);
  const long long delta = (long long)ceil(5.5 * sqrt((double)kSampleN * q * (1.0 - q))) + 2;
  const long long lo_t = r - delta, hi_t = r + delta + 1;
  if (threadIdx.x == 0) {
    s_lohi[0] = 0u;
    s_lohi[1] = 0x7fffffffu;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kPivots; i += 256) {
    const long long rk = (long long)pt->rank[i];
    const uint32_t key = pt->key[i];
    if (rk <= lo_t) atomicMax(&s_lohi[0], key);  // keys are monotone in rank
    if (rk >= hi_t) atomicMin(&s_lohi[1], key);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    b.ss[layer].lo = s_lohi[0];
    b.ss[layer].hi = s_lohi[1];
  }
}

// ONE streaming pass per layer: count |x| below the bracket, compact |x| inside it.
// Warp-granular and barrier-free: a warp takes 1024 elements (8 float4 per lane, all loads issued
// before the first use), tallies, scans the per-lane counts with shuffles, reserves room in the
// layer's candidate buffer with ONE global atomic and scatters its candidates with predicated
// stores.  No shared memory, no __syncthreads, no ballots.
__global__ void __launch_bounds__(kFilterThreads) filter_kernel(SelectBatch b) {
  const int layer = blockIdx.y;
  const float* __restrict__ x = b.x[layer];
  const long long n = b.n[layer];
  SampleState* ss = b.ss + layer;
  uint32_t* __restrict__ cand = b.cand[layer];
  const uint32_t lo = ss->lo, hi = ss->hi;
  const uint32_t span = hi - lo;  // lo <= u <= hi  <=>  (u - lo) <= span
  const unsigned cap32 = (unsigned)(ss->cap < 0xffffffffull ? ss->cap : 0xffffffffull);
  const int lane = threadIdx.x & 31;
  const long long warps_total = (long long)gridDim.x * (kFilterThreads / 32);
  const long long warp_id = (long long)blockIdx.x * (kFilterThreads / 32) + (threadIdx.x >> 5);
  constexpr int kV = kWarpChunk / 128;  // float4 per lane per chunk
  const long long nfull = n / kWarpChunk;  // whole chunks; the ragged remainder is handled after the loop
  unsigned int below = 0u;

  // Candidate space is reserved in warp-private slabs of kSlab slots (one returning atomic per slab,
  // ~1 per 8K input elements): same-address returning atomics run at ~1/ns chip-wide, so one per
  // 512-element chunk would cap the pass at ~2 TB/s.  Unused slab tails are filled with a sentinel
  // key that sorts above every real candidate.
  constexpr unsigned kSlab = 512;
  static_assert(kSlab >= kWarpChunk, "a chunk's candidates must fit one slab");
  unsigned slab_base = 0, slab_used = kSlab, valid_total = 0;  // warp-uniform
  bool have_slab = false;
  auto close_slab = [&]() {
    if (have_slab) {
      for (unsigned i = slab_used + lane; i < kSlab; i += 32)
        if (slab_base + i < cap32) cand[slab_base + i] = 0xffffffffu;
    }
  };
  auto process = [&](const uint32_t (&u)[4 * kV]) {
    unsigned int mine = 0u;
#pragma unroll
    for (int i = 0; i < 4 * kV; ++i) {
      below += (u[i] < lo) ? 1u : 0u;
      mine += (u[i] - lo <= span) ? 1u : 0u;
    }
    unsigned incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    const unsigned wcount = __shfl_sync(0xffffffffu, incl, 31);
    if (wcount) {  // warp-uniform
      if (slab_used + wcount > kSlab) {
        close_slab();
        unsigned gbase = 0;
        if (lane == 0) gbase = atomicAdd(&ss->n_cand, kSlab);
        slab_base = __shfl_sync(0xffffffffu, gbase, 0);
        slab_used = 0;
        have_slab = true;
      }
      unsigned pos = slab_base + slab_used + (incl - mine);
#pragma unroll
      for (int i = 0; i < 4 * kV; ++i) {
        const bool c = (u[i] - lo <= span);
        if (c && pos < cap32) cand[pos] = u[i];
        pos += c ? 1u : 0u;
      }
      slab_used += wcount;
      valid_total += wcount;
    }
  };

  // software pipeline: the next chunk's loads are in flight while this chunk is tallied / scattered
  float4 nxt[kV];
  long long ch = warp_id;
  if (ch < nfull) {
#pragma unroll
    for (int j = 0; j < kV; ++j) nxt[j] = ldg_stream4(x + ch * kWarpChunk + 4ll * (j * 32 + lane));
  }
  while (ch < nfull) {
    uint32_t u[4 * kV];
#pragma unroll
    for (int j = 0; j < kV; ++j) {
      u[4 * j + 0] = __float_as_uint(nxt[j].x) & 0x7fffffffu;
      u[4 * j + 1] = __float_as_uint(nxt[j].y) & 0x7fffffffu;
      u[4 * j + 2] = __float_as_uint(nxt[j].z) & 0x7fffffffu;
      u[4 * j + 3] = __float_as_uint(nxt[j].w) & 0x7fffffffu;
    }
    const long long nx = ch + warps_total;
    if (nx < nfull) {
#pragma unroll
      for (int j = 0; j < kV; ++j) nxt[j] = ldg_stream4(x + nx * kWarpChunk + 4ll * (j * 32 + lane));
    }
    process(u);
    ch = nx;
  }
  if (warp_id == 0 && nfull * kWarpChunk < n) {  // ragged tail (< kWarpChunk elements), one warp
    uint32_t u[4 * kV];
#pragma unroll
    for (int j = 0; j < kV; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const long long i = nfull * kWarpChunk + 4ll * (j * 32 + lane) + e;
        u[4 * j + e] = (i < n) ? (__float_as_uint(__ldg(x + i)) & 0x7fffffffu) : 0xffffffffu;  // parked above any bracket
      }
    }
    process(u);
  }
  close_slab();
  below = (unsigned int)warp_sum((unsigned long long)below);
  if (lane == 0) {
    if (below) atomicAdd(&ss->count_below, (unsigned long long)below);
    if (valid_total) atomicAdd(&ss->n_valid, valid_total);
  }
}

// out-of-range branches of the reference's threshold stage (atq/quantizers.py:33-38)
struct AbsStatsView { double sum_abs; unsigned int max_bits; unsigned int pad; };
__global__ void threshold_from_stats_kernel(const AbsStatsView* as, long long n, int use_max, float factor, float* thr_out) {
  if (use_max) *thr_out = __uint_as_float(as->max_bits) + 1.0f;
  else *thr_out = factor * (float)(as->sum_abs / (double)n);
}

static inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }
static inline int64_t sample_min_n(int count, const int64_t* ns) {
  int64_t total = 0;
  for (int i = 0; i < count; ++i) total += ns[i];
  return total >= kBatchedTotal ? kSampleMinNBatched : kSampleMinN;
}
constexpr int kFilterMaxCtas = 592;  // per layer: bounds the slab-tail waste (296 * 8 warps * 512 slots)
static inline unsigned long long cand_capacity(int64_t n) {
  return (unsigned long long)(n / 8 + 65536) + (unsigned long long)kFilterMaxCtas * (kFilterThreads / 32) * 512ull;
}

}  // namespace atq

using namespace atq;

static bool g_fused_select = true;  // atq_set_fused_select (measurement switch)

static int check_launch(const char* what, int nlaunches) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: kernel launch failed: %s", what, cudaGetErrorString(e));
    return ATQ_ECUDA;
  }
  note_launch(nlaunches);
  return ATQ_OK;
}

// plain path: up to kMaxBatch layers per launch sequence; all entries satisfy 0 <= k < n
static int select_plain_batch(int device, int cnt, const float* const* xs, const int64_t* ns, const int64_t* ks,
                              float* const* thr_outs, SelectState* states, cudaStream_t stream) {
  SelectBatch b;
  memset(&b, 0, sizeof(b));
  int64_t max_n = 0;
  for (int i = 0; i < cnt; ++i) {
    b.x[i] = xs[i]; b.n[i] = ns[i]; b.k[i] = ks[i]; b.thr_out[i] = thr_outs[i];
    if (ns[i] > max_n) max_n = ns[i];
  }
  b.states = states;
  int64_t need = (max_n + (int64_t)kSelThreads * 16 - 1) / ((int64_t)kSelThreads * 16);
  // one cooperative launch (all CTAs resident, layer-level barriers between the digit passes) when the device allows it
  static int coop_cap[64] = {0};  // resident CTAs of select_fused_kernel per device; -1 = no cooperative launch
  int& cc = coop_cap[device & 63];
  if (cc == 0) {
    int coop = 0, per_sm = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
    if (coop && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, select_fused_kernel, kSelThreads, 0) == cudaSuccess && per_sm > 0)
      cc = per_sm * sm_count(device);
    else
      cc = -1;
  }
  if (cc >= cnt && g_fused_select) {
    int64_t cap = cc / cnt;
    int gx = (int)(need < cap ? need : cap);
    if (gx < 1) gx = 1;
    if (cudaMemsetAsync(states, 0, sizeof(SelectState) * (size_t)cnt, stream) != cudaSuccess) return check_launch("select", 0);
    void* args[1] = {&b};
    cudaError_t e = cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(select_fused_kernel), dim3((unsigned)gx, (unsigned)cnt),
                                                dim3(kSelThreads), args, 0, stream);
    if (e != cudaSuccess) {
      set_error("select: cooperative launch failed: %s", cudaGetErrorString(e));
      return ATQ_ECUDA;
    }
    return check_launch("select", 1);
  }
  // CTAs per layer: 16 elements per thread per trip, capped so that the launch is ~8 CTAs per SM
  int64_t cap = ((int64_t)sm_count(device) * 8 + cnt - 1) / cnt;
  if (cap < 1) cap = 1;
  int gx = (int)(need < cap ? need : cap);
  if (gx < 1) gx = 1;
  dim3 grid((unsigned)gx, (unsigned)cnt);
  select_init_kernel<<<cnt, kSelThreads, 0, stream>>>(b);
  select_pass_kernel<0, MODE_PLAIN><<<grid, kSelThreads, 0, stream>>>(b);
  select_pass_kernel<1, MODE_PLAIN><<<grid, kSelThreads, 0, stream>>>(b);
  select_pass_kernel<2, MODE_PLAIN><<<grid, kSelThreads, 0, stream>>>(b);
  return check_launch("select", 4);
}

// sampled path: sample -> pivots -> bracket -> filter -> candidate passes (-> full passes on a miss)
static int select_sampled_batch(int device, int cnt, const float* const* xs, const int64_t* ns, const int64_t* ks,
                                float* const* thr_outs, SelectState* full_states, SelectState* cand_states,
                                SampleState* ss, PivotTable* pt, uint32_t* const* cands, const unsigned long long* caps,
                                cudaStream_t stream) {
  SelectBatch b;
  memset(&b, 0, sizeof(b));
  int64_t max_n = 0;
  unsigned long long max_cap = 0;
  for (int i = 0; i < cnt; ++i) {
    b.x[i] = xs[i]; b.n[i] = ns[i]; b.k[i] = ks[i]; b.thr_out[i] = thr_outs[i]; b.cand[i] = cands[i];
    if (ns[i] > max_n) max_n = ns[i];
    if (caps[i] > max_cap) max_cap = caps[i];
  }
  b.states = full_states; b.cand_states = cand_states; b.ss = ss; b.pt = pt;
  const int sms = sm_count(device);
  // capacities are host-known: write them with the state init (tiny async copy avoided: kernel arg)
  sample_gather_kernel<<<dim3(kSampleN / 256, cnt), 256, 0, stream>>>(b);
  sample_pivots_kernel<<<dim3(kPivots / kPivotsPerCta, cnt), 256, 0, stream>>>(b);
  bracket_kernel<<<cnt, 256, 0, stream>>>(b);
  {  // filter: ~6 CTAs of 8 warps per SM over the whole batch
    int64_t need = (max_n + (int64_t)kWarpChunk * 8 * 4 - 1) / ((int64_t)kWarpChunk * 8 * 4);
    int64_t capg = ((int64_t)sms * 8 + cnt - 1) / cnt;
    if (capg < 1) capg = 1;
    if (capg > kFilterMaxCtas) capg = kFilterMaxCtas;
    int gx = (int)(need < capg ? need : capg);
    filter_kernel<<<dim3(gx, cnt), kFilterThreads, 0, stream>>>(b);
  }
  {
    SelectBatch c = b;
    for (int i = 0; i < cnt; ++i) c.n[i] = (long long)caps[i];
    int64_t need = ((int64_t)max_cap + (int64_t)kSelThreads * 16 - 1) / ((int64_t)kSelThreads * 16);
    int64_t capg = ((int64_t)sms * 4 + cnt - 1) / cnt;
    if (capg < 1) capg = 1;
    int gx = (int)(need < capg ? need : capg);
    // k stays the layer's global rank: pass 0 converts it with count_below
    select_pass_kernel<0, MODE_CAND><<<dim3(gx, cnt), kSelThreads, 0, stream>>>(c);
    select_pass_kernel<1, MODE_CAND><<<dim3(gx, cnt), kSelThreads, 0, stream>>>(c);
    select_pass_kernel<2, MODE_CAND><<<dim3(gx, cnt), kSelThreads, 0, stream>>>(c);
  }
  {  // normally every CTA exits at once; keep the launch small
    int64_t need = (max_n + (int64_t)kSelThreads * 16 - 1) / ((int64_t)kSelThreads * 16);
    int64_t capg = ((int64_t)sms * 2 + cnt - 1) / cnt;
    if (capg < 1) capg = 1;
    int gx = (int)(need < capg ? need : capg);
    select_pass_kernel<0, MODE_FALLBACK><<<dim3(gx, cnt), kSelThreads, 0, stream>>>(b);
    select_pass_kernel<1, MODE_FALLBACK><<<dim3(gx, cnt), kSelThreads, 0, stream>>>(b);
    select_pass_kernel<2, MODE_FALLBACK><<<dim3(gx, cnt), kSelThreads, 0, stream>>>(b);
  }
  return check_launch("sampled select", 10);
}

// workspace plan shared by the size query and the launcher
struct WsPlan {
  size_t plain_states;   // offset of SelectState[count] (plain layers + stats slots)
  size_t full_states, cand_states, ss, pt, cand;  // offsets of the sampled layers' arrays / first buffer
  size_t total;
};
static WsPlan plan_ws(int count, const int64_t* ns) {
  WsPlan p;
  int sampled = 0;
  size_t cand_bytes = 0;
  if (ns != nullptr) {
    const int64_t min_n = sample_min_n(count, ns);
    for (int i = 0; i < count; ++i)
      if (ns[i] >= min_n) { ++sampled; cand_bytes += align256((size_t)cand_capacity(ns[i]) * 4); }
  }
  size_t off = 0;
  p.plain_states = off; off += align256(sizeof(SelectState) * (size_t)(count > 0 ? count : 1));
  p.full_states = off;  off += align256(sizeof(SelectState) * (size_t)sampled);
  p.cand_states = off;  off += align256(sizeof(SelectState) * (size_t)sampled);
  p.ss = off;           off += align256(sizeof(SampleState) * (size_t)sampled);
  p.pt = off;           off += align256(sizeof(PivotTable) * (size_t)sampled);
  p.cand = off;         off += cand_bytes;
  p.total = off;
  return p;
}

__global__ void set_caps_kernel(SampleState* ss, const unsigned long long c0, const unsigned long long c1,
                                const unsigned long long c2, const unsigned long long c3, int base, int cnt) {
  const unsigned long long c[4] = {c0, c1, c2, c3};
  if ((int)threadIdx.x < cnt) ss[base + threadIdx.x].cap = c[threadIdx.x];
}

extern "C" {

void atq_set_fused_select(int enabled) { g_fused_select = enabled != 0; }

size_t atq_workspace_bytes_adaptive_threshold_batched(int count, const int64_t* ns) { return plan_ws(count, ns).total; }
size_t atq_workspace_bytes_select_kth_abs(int64_t n) { return plan_ws(1, &n).total; }
size_t atq_workspace_bytes_adaptive_threshold(int64_t n) { return plan_ws(1, &n).total; }

int atq_adaptive_threshold_batched(int device, int count, const float* const* w_ptrs, const int64_t* ns,
                                   const int64_t* ks, float threshold_factor, float* const* thr_ptrs, void* ws,
                                   size_t ws_bytes, atq_stream_t stream_) {
  ATQ_CHECK_ARG(count > 0 && w_ptrs && ns && ks && thr_ptrs, "null pointer or count <= 0");
  const WsPlan plan = plan_ws(count, ns);
  if (ws == nullptr || ws_bytes < plan.total) {
    set_error("atq_adaptive_threshold_batched: workspace too small");
    return ATQ_EWORKSPACE;
  }
  ATQ_ENSURE_DEVICE(device);
  cudaStream_t stream = (cudaStream_t)stream_;
  char* wsb = reinterpret_cast<char*>(ws);
  SelectState* plain_states = reinterpret_cast<SelectState*>(wsb + plan.plain_states);
  SelectState* full_states = reinterpret_cast<SelectState*>(wsb + plan.full_states);
  SelectState* cand_states = reinterpret_cast<SelectState*>(wsb + plan.cand_states);
  SampleState* ss = reinterpret_cast<SampleState*>(wsb + plan.ss);
  PivotTable* pt = reinterpret_cast<PivotTable*>(wsb + plan.pt);
  char* cand_base = wsb + plan.cand;

  // three kinds of layers: edge branches (|W| statistics), small (batched plain radix select),
  // large and 16-byte aligned (batched sampling front-end)
  const float* px[kMaxBatch]; int64_t pn[kMaxBatch], pk[kMaxBatch]; float* pthr[kMaxBatch];
  const float* sx[kMaxBatch]; int64_t sn[kMaxBatch], sk[kMaxBatch]; float* sthr[kMaxBatch];
  uint32_t* scand[kMaxBatch]; unsigned long long scap[kMaxBatch];
  int pm = 0, sm = 0, pslot = 0, sslot = 0;
  size_t cand_off = 0;
  auto flush_plain = [&]() -> int {
    if (pm == 0) return ATQ_OK;
    int r = select_plain_batch(device, pm, px, pn, pk, pthr, plain_states + pslot, stream);
    pslot += pm; pm = 0;
    return r;
  };
  auto flush_sampled = [&]() -> int {
    if (sm == 0) return ATQ_OK;
    for (int i0 = 0; i0 < sm; i0 += 4) {  // capacities -> SampleState.cap (kernel arguments, no host buffer)
      unsigned long long c[4] = {0, 0, 0, 0};
      int m = sm - i0 < 4 ? sm - i0 : 4;
      for (int j = 0; j < m; ++j) c[j] = scap[i0 + j];
      set_caps_kernel<<<1, 32, 0, stream>>>(ss, c[0], c[1], c[2], c[3], sslot + i0, m);
    }
    int r = select_sampled_batch(device, sm, sx, sn, sk, sthr, full_states + sslot, cand_states + sslot, ss + sslot,
                                 pt + sslot, scand, scap, stream);
    sslot += sm; sm = 0;
    return r;
  };
  const int64_t min_n = sample_min_n(count, ns);
  for (int i = 0; i < count; ++i) {
    ATQ_CHECK_ARG(w_ptrs[i] && thr_ptrs[i] && ns[i] > 0, "null layer pointer or empty layer");
    ATQ_CHECK_ARG((reinterpret_cast<uintptr_t>(w_ptrs[i]) & 3u) == 0, "weights must be 4-byte aligned");
    const bool in_range = ks[i] > 0 && ks[i] < ns[i];
    const bool big = ns[i] >= min_n;
    if (in_range && big && (reinterpret_cast<uintptr_t>(w_ptrs[i]) & 15u) == 0) {
      sx[sm] = w_ptrs[i]; sn[sm] = ns[i]; sk[sm] = ks[i]; sthr[sm] = thr_ptrs[i];
      scap[sm] = cand_capacity(ns[i]);
      scand[sm] = reinterpret_cast<uint32_t*>(cand_base + cand_off);
      if (++sm == kMaxBatch) { int r = flush_sampled(); if (r != ATQ_OK) return r; }
    } else if (in_range) {
      px[pm] = w_ptrs[i]; pn[pm] = ns[i]; pk[pm] = ks[i]; pthr[pm] = thr_ptrs[i];
      if (++pm == kMaxBatch) { int r = flush_plain(); if (r != ATQ_OK) return r; }
    } else {
      int r = flush_plain();  // keep slot indices in launch order
      if (r != ATQ_OK) return r;
      void* stats = plain_states + pslot;  // 16 bytes of this layer's slot
      ++pslot;
      r = atq_abs_stats(device, w_ptrs[i], ns[i], stats, nullptr, 0, stream_);
      if (r != ATQ_OK) return r;
      threshold_from_stats_kernel<<<1, 1, 0, stream>>>((const AbsStatsView*)stats, (long long)ns[i], ks[i] >= ns[i] ? 1 : 0,
                                                        threshold_factor, thr_ptrs[i]);
      ATQ_LAUNCH_CHECK();
    }
    if (big) cand_off += align256((size_t)cand_capacity(ns[i]) * 4);
  }
  int r = flush_plain();
  if (r != ATQ_OK) return r;
  return flush_sampled();
}

int atq_adaptive_threshold(int device, const float* w, int64_t n, int64_t k, float threshold_factor, float* thr_out,
                           void* ws, size_t ws_bytes, atq_stream_t stream) {
  const float* const xs[1] = {w};
  float* const ts[1] = {thr_out};
  return atq_adaptive_threshold_batched(device, 1, xs, &n, &k, threshold_factor, ts, ws, ws_bytes, stream);
}

int atq_select_kth_abs(int device, const float* x, int64_t n, int64_t k, float* thr_out, void* ws, size_t ws_bytes,
                       atq_stream_t stream) {
  ATQ_CHECK_ARG(x && thr_out && n > 0, "null pointer or n <= 0");
  ATQ_CHECK_ARG(k >= 0 && k < n, "k out of range");
  if (k == 0) {
    // rank 0 is in range for a select but is the "k <= 0" branch of the threshold stage: route it
    // through the plain path directly
    ATQ_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 3u) == 0, "x must be 4-byte aligned");
    if (ws == nullptr || ws_bytes < plan_ws(1, &n).total) {
      set_error("atq_select_kth_abs: workspace too small");
      return ATQ_EWORKSPACE;
    }
    ATQ_ENSURE_DEVICE(device);
    return select_plain_batch(device, 1, &x, &n, &k, &thr_out, reinterpret_cast<SelectState*>(ws), (cudaStream_t)stream);
  }
  const float* const xs[1] = {x};
  float* const ts[1] = {thr_out};
  return atq_adaptive_threshold_batched(device, 1, xs, &n, &k, 0.f, ts, ws, ws_bytes, stream);
}

}  // extern "C"
