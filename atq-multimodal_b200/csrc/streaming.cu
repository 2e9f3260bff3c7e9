// HBM-streaming kernels of the ATQ hot path: |W| statistics, ternarize, 2-bit pack/unpack,
// routing mask-multiply, bf16 operand builders.  All are coalesced 128-bit load/store,
// grid-stride kernels sized in multiples of the SM count; integer outputs are bit-exact
// against the reference (atq/quantizers.py:41-43, atq/bit_packing.py:45-69,104-119).
#include <cooperative_groups.h>
#include "common.cuh"

namespace atq {

constexpr int kThreads = 256;
constexpr int kUnroll = 4;

struct TernStats {  // 16 bytes
  unsigned long long nnz;
  double sum_wt;
};
struct AbsStats {  // 16 bytes
  double sum_abs;
  unsigned int max_bits;
  unsigned int pad;
};

template <bool VEC>
__device__ __forceinline__ float4 load4(const float* __restrict__ base, int64_t g) {
  if constexpr (VEC) {
    return ldg_stream4(base + 4 * g);
  } else {
    const float* p = base + 4 * g;
    return make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3));
  }
}

// ------------------------------------------------------------------------------------
// K1: sum|W| (fp64 accumulate) and max|W|
// ------------------------------------------------------------------------------------
template <bool VEC>
__global__ void __launch_bounds__(kThreads) abs_stats_kernel(const float* __restrict__ w, int64_t n,
                                                            AbsStats* __restrict__ out) {
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double acc = 0.0;
  float mx = 0.f;
  for (int64_t g0 = tid; g0 < n4; g0 += stride * kUnroll) {
    float4 v[kUnroll];
#pragma unroll
    for (int j = 0; j < kUnroll; ++j) {
      int64_t g = g0 + j * stride;
      v[j] = (g < n4) ? load4<VEC>(w, g) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < kUnroll; ++j) {
      float a0 = fabsf(v[j].x), a1 = fabsf(v[j].y), a2 = fabsf(v[j].z), a3 = fabsf(v[j].w);
      acc += (double)a0 + (double)a1 + (double)a2 + (double)a3;
      mx = fmaxf(mx, fmaxf(fmaxf(a0, a1), fmaxf(a2, a3)));
    }
  }
  if (tid == 0) {
    for (int64_t i = n4 * 4; i < n; ++i) {
      float a = fabsf(w[i]);
      acc += (double)a;
      mx = fmaxf(mx, a);
    }
  }
  __shared__ double s_sum[kThreads / 32];
  __shared__ float s_max[kThreads / 32];
  acc = warp_sum(acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) {
    s_sum[wid] = acc;
    s_max[wid] = mx;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    float m = 0.f;
    for (int i = 0; i < kThreads / 32; ++i) {
      t += s_sum[i];
      m = fmaxf(m, s_max[i]);
    }
    atomicAdd(&out->sum_abs, t);
    atomicMax(&out->max_bits, __float_as_uint(m));  // non-negative floats order as uints
  }
}

// ------------------------------------------------------------------------------------
// K3 / K4 / K5: ternarize (threshold compare) or validate (already-ternary fp32) -> fp32 T or
// 2-bit codes.  One float4 (= one packed byte) per thread per step: 512 B loads and 32 B /
// 512 B stores per warp instruction, all fully coalesced.
// ------------------------------------------------------------------------------------
enum { OUT_F32 = 0, OUT_PACK2 = 1 };
enum { SRC_THRESHOLD = 0, SRC_TERNARY = 1 };

template <int SRC>
__device__ __forceinline__ uint32_t code_of(float v, float thr, bool& bad) {
  if constexpr (SRC == SRC_THRESHOLD) {
    return tern_code(v, thr);
  } else {
    // atq/bit_packing.py:36-39,49: only -1, 0 (either sign), +1 are valid
    uint32_t c = (v == 1.0f) ? 2u : ((v == -1.0f) ? 0u : 1u);
    bad |= !(v == 1.0f || v == -1.0f || v == 0.0f);
    return c;
  }
}

template <bool VEC, int SRC, int OUT, bool STATS>
__global__ void __launch_bounds__(kThreads)
    ternarize_kernel(const float* __restrict__ w, int64_t n, const float* __restrict__ thr_p,
                     float* __restrict__ t_out, uint8_t* __restrict__ packed, TernStats* __restrict__ stats,
                     int32_t* __restrict__ invalid_flag) {
  const float thr = (SRC == SRC_THRESHOLD) ? __ldg(thr_p) : 0.f;
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long nnz = 0;
  double swt = 0.0;
  bool bad = false;

  for (int64_t g0 = tid; g0 < n4; g0 += stride * kUnroll) {
    float4 v[kUnroll];
#pragma unroll
    for (int j = 0; j < kUnroll; ++j) {
      int64_t g = g0 + j * stride;
      if (g < n4) v[j] = load4<VEC>(w, g);
    }
#pragma unroll
    for (int j = 0; j < kUnroll; ++j) {
      int64_t g = g0 + j * stride;
      if (g < n4) {
        uint32_t c0 = code_of<SRC>(v[j].x, thr, bad), c1 = code_of<SRC>(v[j].y, thr, bad);
        uint32_t c2 = code_of<SRC>(v[j].z, thr, bad), c3 = code_of<SRC>(v[j].w, thr, bad);
        if constexpr (STATS) {
          float t0 = (float)c0 - 1.f, t1 = (float)c1 - 1.f, t2 = (float)c2 - 1.f, t3 = (float)c3 - 1.f;
          nnz += (c0 != 1u) + (c1 != 1u) + (c2 != 1u) + (c3 != 1u);
          swt += (double)(v[j].x * t0) + (double)(v[j].y * t1) + (double)(v[j].z * t2) + (double)(v[j].w * t3);
        }
        if constexpr (OUT == OUT_PACK2) {
          packed[g] = (uint8_t)(c0 | (c1 << 2) | (c2 << 4) | (c3 << 6));
        } else {
          float4 o = make_float4((float)c0 - 1.f, (float)c1 - 1.f, (float)c2 - 1.f, (float)c3 - 1.f);
          if constexpr (VEC) {
            *reinterpret_cast<float4*>(t_out + 4 * g) = o;
          } else {
            t_out[4 * g] = o.x; t_out[4 * g + 1] = o.y; t_out[4 * g + 2] = o.z; t_out[4 * g + 3] = o.w;
          }
        }
      }
    }
  }
  if (tid == 0 && (n & 3)) {  // ragged tail: last byte has zero bits above the last code
    uint32_t byte = 0;
    for (int64_t i = n4 * 4, j = 0; i < n; ++i, ++j) {
      float x = w[i];
      uint32_t c = code_of<SRC>(x, thr, bad);
      if constexpr (STATS) {
        nnz += (c != 1u);
        swt += (double)(x * ((float)c - 1.f));
      }
      if constexpr (OUT == OUT_PACK2) byte |= c << (2 * j);
      else t_out[i] = (float)c - 1.f;
    }
    if constexpr (OUT == OUT_PACK2) packed[n4] = (uint8_t)byte;
  }
  if constexpr (SRC == SRC_TERNARY) {
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicExch(invalid_flag, 1);
  }
  if constexpr (STATS) {
    __shared__ unsigned long long s_n[kThreads / 32];
    __shared__ double s_s[kThreads / 32];
    nnz = warp_sum(nnz);
    swt = warp_sum(swt);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { s_n[wid] = nnz; s_s[wid] = swt; }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long tn = 0; double ts = 0.0;
      for (int i = 0; i < kThreads / 32; ++i) { tn += s_n[i]; ts += s_s[i]; }
      atomicAdd(&stats->nnz, tn);
      atomicAdd(&stats->sum_wt, ts);
    }
  }
}

__global__ void optimal_alpha_kernel(const TernStats* ts, const AbsStats* as, int64_t n, float* alpha_out) {
  // atq/quantizers.py:49-55 resolved on the device (the reference syncs the host here)
  if (ts->nnz > 0) {
    // reference divides an fp32 sum by an fp32 count
    *alpha_out = (float)ts->sum_wt / (float)ts->nnz;
  } else {
    *alpha_out = (float)(as->sum_abs / (double)n);
  }
}

// ------------------------------------------------------------------------------------
// K6: unpack 2-bit codes.  One byte -> four values per thread step.
// ------------------------------------------------------------------------------------
enum { UNPACK_F32 = 0, UNPACK_BF16 = 1, UNPACK_I8 = 2 };

template <int KIND, bool VEC>
__global__ void __launch_bounds__(kThreads)
    unpack2_kernel(const uint8_t* __restrict__ packed, int64_t n, void* __restrict__ out_v,
                   int32_t* __restrict__ invalid_flag) {
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool bad = false;
  for (int64_t g0 = tid; g0 < n4; g0 += stride * kUnroll) {
    uint32_t b[kUnroll];
#pragma unroll
    for (int j = 0; j < kUnroll; ++j) {
      int64_t g = g0 + j * stride;
      b[j] = (g < n4) ? (uint32_t)__ldg(packed + g) : 0x55u;
    }
#pragma unroll
    for (int j = 0; j < kUnroll; ++j) {
      int64_t g = g0 + j * stride;
      if (g >= n4) continue;
      uint32_t c0 = b[j] & 3u, c1 = (b[j] >> 2) & 3u, c2 = (b[j] >> 4) & 3u, c3 = (b[j] >> 6) & 3u;
      bad |= (c0 == 3u) | (c1 == 3u) | (c2 == 3u) | (c3 == 3u);
      if constexpr (KIND == UNPACK_F32) {
        float* out = reinterpret_cast<float*>(out_v);
        float4 o = make_float4((float)c0 - 1.f, (float)c1 - 1.f, (float)c2 - 1.f, (float)c3 - 1.f);
        if constexpr (VEC) *reinterpret_cast<float4*>(out + 4 * g) = o;
        else { out[4 * g] = o.x; out[4 * g + 1] = o.y; out[4 * g + 2] = o.z; out[4 * g + 3] = o.w; }
      } else if constexpr (KIND == UNPACK_BF16) {
        uint16_t* out = reinterpret_cast<uint16_t*>(out_v);
        // bf16: -1 = 0xBF80, 0 = 0, +1 = 0x3F80
        auto bf = [](uint32_t c) -> uint32_t { return c == 1u ? 0u : (c == 0u ? 0xBF80u : 0x3F80u); };
        uint2 o = make_uint2(bf(c0) | (bf(c1) << 16), bf(c2) | (bf(c3) << 16));
        if constexpr (VEC) *reinterpret_cast<uint2*>(out + 4 * g) = o;
        else { out[4*g] = (uint16_t)bf(c0); out[4*g+1] = (uint16_t)bf(c1); out[4*g+2] = (uint16_t)bf(c2); out[4*g+3] = (uint16_t)bf(c3); }
      } else {
        int8_t* out = reinterpret_cast<int8_t*>(out_v);
        uint32_t o = ((c0 - 1u) & 0xffu) | (((c1 - 1u) & 0xffu) << 8) | (((c2 - 1u) & 0xffu) << 16) | (((c3 - 1u) & 0xffu) << 24);
        if constexpr (VEC) *reinterpret_cast<uint32_t*>(out + 4 * g) = o;
        else { out[4*g] = (int8_t)(c0 - 1); out[4*g+1] = (int8_t)(c1 - 1); out[4*g+2] = (int8_t)(c2 - 1); out[4*g+3] = (int8_t)(c3 - 1); }
      }
    }
  }
  if (tid == 0 && (n & 3)) {
    uint32_t byte = packed[n4];
    for (int64_t i = n4 * 4, j = 0; i < n; ++i, ++j) {
      uint32_t c = (byte >> (2 * j)) & 3u;
      bad |= (c == 3u);
      if constexpr (KIND == UNPACK_F32) reinterpret_cast<float*>(out_v)[i] = (float)c - 1.f;
      else if constexpr (KIND == UNPACK_BF16) reinterpret_cast<uint16_t*>(out_v)[i] = c == 1u ? 0 : (c == 0u ? 0xBF80 : 0x3F80);
      else reinterpret_cast<int8_t*>(out_v)[i] = (int8_t)((int)c - 1);
    }
  }
  if (invalid_flag != nullptr && __any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicExch(invalid_flag, 1);
}

// ------------------------------------------------------------------------------------
// Codec kernels for whole models (BASELINE config 5: 1 B weights in ~60 layers): ONE launch for all layers
// (blockIdx.y = layer, so no per-layer launch gap or tail), 128-bit global accesses on both sides.
// A warp step covers 2048 consecutive weights: 16 coalesced float4 loads per lane (8 KB per warp in flight), the
// 512 code bytes are transposed through a warp-private shared-memory strip, and every lane stores (or, for unpack,
// loads) ONE 16-byte piece of the packed stream -- 64 weights per 128-bit packed access.
// Layers must be 16-byte aligned; the last (n % 2048) weights of a layer take the scalar tail.
// ------------------------------------------------------------------------------------
constexpr int kCodecBatch = 64;
enum { CODEC_TERNARIZE = 0, CODEC_PACK = 1, CODEC_UNPACK = 2 };
struct CodecBatch {
  const float* src[kCodecBatch];      // fp32 weights (TERNARIZE) / fp32 ternary values (PACK) / unused (UNPACK)
  uint8_t* packed[kCodecBatch];       // 2-bit codes (written, or read for UNPACK)
  float* dst[kCodecBatch];            // fp32 output (UNPACK)
  const float* thr[kCodecBatch];      // per-layer threshold (TERNARIZE)
  long long n[kCodecBatch];
  int32_t* invalid_flag;              // PACK / UNPACK: set to 1 when a value / code is not ternary
};

template <int MODE>
__global__ void __launch_bounds__(kThreads) codec_batched_kernel(const CodecBatch b) {
  __shared__ __align__(16) uint8_t strip[kThreads / 32][512];
  const int layer = blockIdx.y;
  const long long n = b.n[layer];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const long long warps = (long long)gridDim.x * (kThreads / 32);
  const long long wg = (long long)blockIdx.x * (kThreads / 32) + wid;
  const long long chunks = n >> 11;  // 2048 weights per warp step
  uint8_t* const my = strip[wid];
  bool bad = false;
  if constexpr (MODE != CODEC_UNPACK) {
    const float* __restrict__ src = b.src[layer];
    const float thr = (MODE == CODEC_TERNARIZE) ? __ldg(b.thr[layer]) : 0.f;
    uint8_t* __restrict__ packed = b.packed[layer];
    for (long long c = wg; c < chunks; c += warps) {
      const float4* base = reinterpret_cast<const float4*>(src) + c * 512;
      float4 v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = ldg_stream4(reinterpret_cast<const float*>(base + j * 32 + lane));
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        uint32_t c0, c1, c2, c3;
        if constexpr (MODE == CODEC_TERNARIZE) {
          c0 = tern_code(v[j].x, thr); c1 = tern_code(v[j].y, thr); c2 = tern_code(v[j].z, thr); c3 = tern_code(v[j].w, thr);
        } else {
          c0 = code_of<SRC_TERNARY>(v[j].x, 0.f, bad); c1 = code_of<SRC_TERNARY>(v[j].y, 0.f, bad);
          c2 = code_of<SRC_TERNARY>(v[j].z, 0.f, bad); c3 = code_of<SRC_TERNARY>(v[j].w, 0.f, bad);
        }
        my[j * 32 + lane] = (uint8_t)(c0 | (c1 << 2) | (c2 << 4) | (c3 << 6));
      }
      __syncwarp();
      const uint4 out = *reinterpret_cast<const uint4*>(my + 16 * lane);
      __syncwarp();
      *reinterpret_cast<uint4*>(packed + c * 512 + 16 * lane) = out;
    }
    if (wg == 0 && lane == 0) {  // tail: whole bytes, then the ragged last byte (zero bits above the last code)
      for (long long i = chunks << 11; i < n; i += 4) {
        uint32_t byte = 0;
        for (int j = 0; j < 4 && i + j < n; ++j) {
          const float x = src[i + j];
          const uint32_t code = (MODE == CODEC_TERNARIZE) ? tern_code(x, thr) : code_of<SRC_TERNARY>(x, 0.f, bad);
          byte |= code << (2 * j);
        }
        packed[i >> 2] = (uint8_t)byte;
      }
    }
  } else {
    const uint8_t* __restrict__ packed = b.packed[layer];
    float* __restrict__ dst = b.dst[layer];
    for (long long c = wg; c < chunks; c += warps) {
      *reinterpret_cast<uint4*>(my + 16 * lane) = __ldg(reinterpret_cast<const uint4*>(packed + c * 512) + lane);
      __syncwarp();
      float4* out = reinterpret_cast<float4*>(dst) + c * 512;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const uint32_t byte = my[j * 32 + lane];
        const uint32_t c0 = byte & 3u, c1 = (byte >> 2) & 3u, c2 = (byte >> 4) & 3u, c3 = byte >> 6;
        bad |= (c0 == 3u) | (c1 == 3u) | (c2 == 3u) | (c3 == 3u);
        __stcs(out + j * 32 + lane, make_float4((float)c0 - 1.f, (float)c1 - 1.f, (float)c2 - 1.f, (float)c3 - 1.f));
      }
      __syncwarp();
    }
    if (wg == 0 && lane == 0) {
      for (long long i = chunks << 11; i < n; ++i) {
        const uint32_t code = ((uint32_t)packed[i >> 2] >> (2 * (int)(i & 3))) & 3u;
        bad |= (code == 3u);
        dst[i] = (float)code - 1.f;
      }
    }
  }
  if constexpr (MODE != CODEC_TERNARIZE) {
    if (b.invalid_flag != nullptr && __any_sync(0xffffffffu, bad) && lane == 0) atomicExch(b.invalid_flag, 1);
  }
}

// ------------------------------------------------------------------------------------
// K10: grad_in = grad_out * (|x| > thr)      (atq/routing.py:53-56)
// ------------------------------------------------------------------------------------
template <bool VEC>
__global__ void __launch_bounds__(kThreads)
    route_mask_mul_kernel(const float* __restrict__ x, const float* __restrict__ go, const float* __restrict__ thr_p,
                          int64_t n, float* __restrict__ gi) {
  const float thr = __ldg(thr_p);
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t g0 = tid; g0 < n4; g0 += stride * kUnroll) {
    float4 a[kUnroll], b[kUnroll];
#pragma unroll
    for (int j = 0; j < kUnroll; ++j) {
      int64_t g = g0 + j * stride;
      if (g < n4) { a[j] = load4<VEC>(x, g); b[j] = load4<VEC>(go, g); }
    }
#pragma unroll
    for (int j = 0; j < kUnroll; ++j) {
      int64_t g = g0 + j * stride;
      if (g >= n4) continue;
      // multiply (not select) so that NaN/Inf in grad_out propagate like `grad * mask` does
      float4 o = make_float4(b[j].x * (fabsf(a[j].x) > thr ? 1.f : 0.f), b[j].y * (fabsf(a[j].y) > thr ? 1.f : 0.f),
                             b[j].z * (fabsf(a[j].z) > thr ? 1.f : 0.f), b[j].w * (fabsf(a[j].w) > thr ? 1.f : 0.f));
      if constexpr (VEC) *reinterpret_cast<float4*>(gi + 4 * g) = o;
      else { gi[4*g] = o.x; gi[4*g+1] = o.y; gi[4*g+2] = o.z; gi[4*g+3] = o.w; }
    }
  }
  if (tid == 0) {
    for (int64_t i = n4 * 4; i < n; ++i) gi[i] = go[i] * (fabsf(x[i]) > thr ? 1.f : 0.f);
  }
}

// ------------------------------------------------------------------------------------
// Per-tensor power-of-two scale for the scaled-fp16 operand format (common.cuh: OperandFmt).
// One launch: every CTA folds max|x| of its share into slot[0] (atomicMax on the bit pattern, valid for
// non-negative floats); the last CTA to finish (ticket in slot[3]) derives the scale from
// bound = max(max|x|, |*extra|) * bound_mul, stores {scale, 1/scale} in slot[1..2] and re-arms slot[0] = slot[3] = 0
// so that a CUDA-graph replay of the same launch starts clean.  NaNs are ignored by fmaxf; an infinite
// maximum yields scale 1 (the operand then carries the infinities, as the fp32 reference would).
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
    absmax_scale_kernel(const float* __restrict__ x, int64_t rows, int64_t cols, int64_t ld, float bound_mul,
                        const float* __restrict__ extra, float* __restrict__ slot, int vec) {
  float m = 0.f;
  if (vec) {  // contiguous, 16-byte aligned: flat float4 walk
    const int64_t n = rows * cols, n4 = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n4; g += stride) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x) + g);
      m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
      for (int64_t i = n4 * 4; i < n; ++i) m = fmaxf(m, fabsf(x[i]));
  } else {
    for (int64_t r = blockIdx.x; r < rows; r += gridDim.x)
      for (int64_t c = threadIdx.x; c < cols; c += blockDim.x) m = fmaxf(m, fabsf(__ldg(x + r * ld + c)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  __shared__ float s_m[kThreads / 32];
  if ((threadIdx.x & 31) == 0) s_m[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < kThreads / 32; ++i) m = fmaxf(m, s_m[i]);
    unsigned int* bits = reinterpret_cast<unsigned int*>(slot);
    atomicMax(bits, __float_as_uint(m));
    __threadfence();
    const unsigned int ticket = atomicAdd(bits + 3, 1u);
    if (ticket == gridDim.x - 1) {
      __threadfence();
      float b = __uint_as_float(atomicExch(bits, 0u));  // read the maximum and re-arm
      if (extra != nullptr) b = fmaxf(b, fabsf(__ldg(extra)));
      b *= bound_mul;
      float sc, inv;
      pow2_scale_for(b, sc, inv);
      slot[1] = sc;
      slot[2] = inv;
      bits[3] = 0u;
    }
  }
}

// many tensors, one launch (the per-layer weight scales of a model: blockIdx.y = tensor); contiguous tensors only
constexpr int kAbsmaxBatch = 64;
struct AbsmaxBatch {
  const float* x[kAbsmaxBatch];
  long long n[kAbsmaxBatch];
  const float* extra[kAbsmaxBatch];
  float* slot[kAbsmaxBatch];
};
__global__ void __launch_bounds__(kThreads) absmax_scale_batched_kernel(const AbsmaxBatch b, float bound_mul) {
  const float* __restrict__ x = b.x[blockIdx.y];
  const long long n = b.n[blockIdx.y];
  float* slot = b.slot[blockIdx.y];
  const bool vec = (reinterpret_cast<uintptr_t>(x) & 15u) == 0;
  float m = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (vec) {
    const long long n4 = n >> 2;
    for (long long g = t; g < n4; g += stride) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x) + g);
      m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
    if (t == 0)
      for (long long i = n4 * 4; i < n; ++i) m = fmaxf(m, fabsf(x[i]));
  } else {
    for (long long i = t; i < n; i += stride) m = fmaxf(m, fabsf(__ldg(x + i)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  __shared__ float s_m[kThreads / 32];
  if ((threadIdx.x & 31) == 0) s_m[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < kThreads / 32; ++i) m = fmaxf(m, s_m[i]);
    unsigned int* bits = reinterpret_cast<unsigned int*>(slot);
    atomicMax(bits, __float_as_uint(m));
    __threadfence();
    if (atomicAdd(bits + 3, 1u) == gridDim.x - 1) {
      __threadfence();
      float bound = __uint_as_float(atomicExch(bits, 0u));
      const float* extra = b.extra[blockIdx.y];
      if (extra != nullptr) bound = fmaxf(bound, fabsf(__ldg(extra)));
      bound *= bound_mul;
      float sc, inv;
      pow2_scale_for(bound, sc, inv);
      slot[1] = sc;
      slot[2] = inv;
      bits[3] = 0u;
    }
  }
}

// ------------------------------------------------------------------------------------
// Small tensors (<= kFusedSplitMax elements, the launch-bound shapes of BASELINE configs 1-2): max|x| reduction,
// scale derivation and the scaled-fp16 split in ONE launch.  A thread-block cluster of 8 CTAs (co-scheduled by
// the hardware, so the cluster barrier cannot dead-lock against other streams) reads the tensor once into
// registers, exchanges the per-CTA maxima through distributed shared memory, derives the power-of-two scale and
// writes the operand pair straight from the registers.  slot[1..2] receive {scale, 1/scale} for the GEMM epilogue.
// ------------------------------------------------------------------------------------
constexpr int kFusedCtas = 8, kFusedThreads = 512, kFusedVec = 24;  // float4 per thread
constexpr int64_t kFusedSplitMax = (int64_t)kFusedCtas * kFusedThreads * kFusedVec * 4;  // 393 216 elements

template <bool HAS_LO>
__global__ void __cluster_dims__(kFusedCtas, 1, 1) __launch_bounds__(kFusedThreads, 1)
    split_scaled_cluster_kernel(const float* __restrict__ x, int64_t n4, uint16_t* __restrict__ hi, uint16_t* __restrict__ lo,
                                float* __restrict__ slot, float bound_mul, const float* __restrict__ extra) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ float s_w[kFusedThreads / 32];
  __shared__ float s_cta_max;
  const int64_t t = (int64_t)blockIdx.x * kFusedThreads + threadIdx.x;
  constexpr int64_t kStride = (int64_t)kFusedCtas * kFusedThreads;
  float4 v[kFusedVec];
  float m = 0.f;
#pragma unroll
  for (int j = 0; j < kFusedVec; ++j) {
    const int64_t g = t + j * kStride;
    v[j] = g < n4 ? __ldg(reinterpret_cast<const float4*>(x) + g) : make_float4(0.f, 0.f, 0.f, 0.f);
    m = fmaxf(fmaxf(m, fmaxf(fabsf(v[j].x), fabsf(v[j].y))), fmaxf(fabsf(v[j].z), fabsf(v[j].w)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < kFusedThreads / 32; ++i) m = fmaxf(m, s_w[i]);
    s_cta_max = m;
  }
  cluster.sync();  // every CTA's maximum is visible cluster-wide
  float b = 0.f;
#pragma unroll
  for (int r = 0; r < kFusedCtas; ++r) b = fmaxf(b, *cluster.map_shared_rank(&s_cta_max, r));
  cluster.sync();  // nobody reads remote shared memory after this point
  if (extra != nullptr) b = fmaxf(b, fabsf(__ldg(extra)));
  b *= bound_mul;
  OperandFmt fmt;
  float inv;
  fmt.f16 = 1;
  pow2_scale_for(b, fmt.scale, inv);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    slot[1] = fmt.scale;
    slot[2] = inv;
  }
#pragma unroll
  for (int j = 0; j < kFusedVec; ++j) {
    const int64_t g = t + j * kStride;
    if (g >= n4) continue;
    uint16_t h0, h1, h2, h3, l0, l1, l2, l3;
    split2(fmt, v[j].x, h0, l0); split2(fmt, v[j].y, h1, l1); split2(fmt, v[j].z, h2, l2); split2(fmt, v[j].w, h3, l3);
    *reinterpret_cast<uint2*>(hi + 4 * g) = make_uint2((uint32_t)h0 | ((uint32_t)h1 << 16), (uint32_t)h2 | ((uint32_t)h3 << 16));
    if constexpr (HAS_LO)
      *reinterpret_cast<uint2*>(lo + 4 * g) = make_uint2((uint32_t)l0 | ((uint32_t)l1 << 16), (uint32_t)l2 | ((uint32_t)l3 << 16));
  }
}

// ------------------------------------------------------------------------------------
// fp32 -> (hi, lo) operand split (bf16 pair, or scaled fp16 pair when a scale slot is given), row-major
// ------------------------------------------------------------------------------------
template <bool HAS_LO>
__global__ void __launch_bounds__(kThreads)
    split_flat_kernel(const float* __restrict__ x, int64_t n, uint16_t* __restrict__ hi, uint16_t* __restrict__ lo,
                      const float* __restrict__ fslot) {
  const OperandFmt fmt = load_fmt(fslot);
  // contiguous case (ld_in == cols == pitch): n % 8 == 0 guaranteed by the caller
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t g0 = tid; g0 < n4; g0 += stride * kUnroll) {
    float4 v[kUnroll];
#pragma unroll
    for (int j = 0; j < kUnroll; ++j) {
      int64_t g = g0 + j * stride;
      if (g < n4) v[j] = ldg_stream4(x + 4 * g);
    }
#pragma unroll
    for (int j = 0; j < kUnroll; ++j) {
      int64_t g = g0 + j * stride;
      if (g >= n4) continue;
      uint16_t h0, h1, h2, h3, l0, l1, l2, l3;
      split2(fmt, v[j].x, h0, l0); split2(fmt, v[j].y, h1, l1); split2(fmt, v[j].z, h2, l2); split2(fmt, v[j].w, h3, l3);
      *reinterpret_cast<uint2*>(hi + 4 * g) = make_uint2((uint32_t)h0 | ((uint32_t)h1 << 16), (uint32_t)h2 | ((uint32_t)h3 << 16));
      if constexpr (HAS_LO)
        *reinterpret_cast<uint2*>(lo + 4 * g) = make_uint2((uint32_t)l0 | ((uint32_t)l1 << 16), (uint32_t)l2 | ((uint32_t)l3 << 16));
    }
  }
}

// split + column sums in one pass over dY (bias gradient rides along): thread (cg, rl) owns column
// group cg (4 columns) and walks rows rl, rl+R, ...; its four running sums go to part[rl][4cg..],
// which colsum_stage2_kernel reduces in fixed order (deterministic).  Requires cols % 4 == 0,
// ld_in == cols == pitch and 16-byte aligned pointers (the flat-path conditions).
template <bool HAS_LO>
__global__ void __launch_bounds__(kThreads)
    split_colsum_kernel(const float* __restrict__ x, int64_t rows, int64_t cols, uint16_t* __restrict__ hi,
                        uint16_t* __restrict__ lo, float* __restrict__ part, int64_t R, const float* __restrict__ fslot) {
  const OperandFmt fmt = load_fmt(fslot);
  const int64_t cg4 = cols >> 2;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= cg4 * R) return;
  const int64_t cg = t % cg4, rl = t / cg4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t r0 = rl; r0 < rows; r0 += R * kUnroll) {
    float4 v[kUnroll];
#pragma unroll
    for (int j = 0; j < kUnroll; ++j) {
      const int64_t r = r0 + j * R;
      if (r < rows) v[j] = ldg_stream4(x + r * cols + 4 * cg);
    }
#pragma unroll
    for (int j = 0; j < kUnroll; ++j) {
      const int64_t r = r0 + j * R;
      if (r >= rows) continue;
      const int64_t g = r * cg4 + cg;
      uint16_t h0, h1, h2, h3, l0, l1, l2, l3;
      split2(fmt, v[j].x, h0, l0); split2(fmt, v[j].y, h1, l1); split2(fmt, v[j].z, h2, l2); split2(fmt, v[j].w, h3, l3);
      *reinterpret_cast<uint2*>(hi + 4 * g) = make_uint2((uint32_t)h0 | ((uint32_t)h1 << 16), (uint32_t)h2 | ((uint32_t)h3 << 16));
      if constexpr (HAS_LO)
        *reinterpret_cast<uint2*>(lo + 4 * g) = make_uint2((uint32_t)l0 | ((uint32_t)l1 << 16), (uint32_t)l2 | ((uint32_t)l3 << 16));
      acc.x += v[j].x; acc.y += v[j].y; acc.z += v[j].z; acc.w += v[j].w;
    }
  }
  *reinterpret_cast<float4*>(part + rl * cols + 4 * cg) = acc;
}

// ------------------------------------------------------------------------------------
// FFN activation fused with the operand split ("GELU into the FFN epilogue").
//   forward : d  = dropout(gelu(y))                   -> bf16 (hi, lo) operand of the second FFN GEMM
//   backward: dy = g .* keep/(1-p) .* gelu'(y)        -> bf16 (hi, lo) operand of the first layer's dX / dW
//             GEMMs + column sums (its bias gradient)
// gelu is the exact erf form (F.gelu default, models/text_encoder.py:246).  The dropout mask is never
// stored: both directions regenerate it from the counter hash of (seed, flat element index).
// Same lane layout as split_colsum_kernel: a thread owns 4 consecutive columns of every R-th row.
// ------------------------------------------------------------------------------------
struct ActParams {
  uint32_t drop_thresh;  // 0 = no dropout
  float inv_keep;
  const unsigned long long* seed;
};

__device__ __forceinline__ float gelu_fwd(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad(float x) {
  return 0.5f * (1.f + erff(x * 0.70710678118654752f)) + x * 0.39894228040143268f * __expf(-0.5f * x * x);
}

template <bool BWD, bool HAS_LO>
__global__ void __launch_bounds__(kThreads)
    act_split_kernel(const float* __restrict__ x, const float* __restrict__ y, int64_t rows, int64_t cols,
                     uint16_t* __restrict__ hi, uint16_t* __restrict__ lo, float* __restrict__ part, int64_t R, const ActParams ap,
                     const float* __restrict__ fslot) {
  const OperandFmt fmt = load_fmt(fslot);
  const int64_t cg4 = cols >> 2;
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= cg4 * R) return;
  const int64_t cg = t % cg4, rl = t / cg4;
  uint32_t key = 0;
  if (ap.drop_thresh != 0u) {
    const unsigned long long seed = ap.seed != nullptr ? *ap.seed : 0ull;
    key = drop_row_key((uint32_t)seed, (uint32_t)(seed >> 32), 0x0FF1CEu);
  }
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t r0 = rl; r0 < rows; r0 += R * kUnroll) {
    float4 v[kUnroll], w[kUnroll];
#pragma unroll
    for (int j = 0; j < kUnroll; ++j) {
      const int64_t r = r0 + j * R;
      if (r < rows) {
        v[j] = ldg_stream4(x + r * cols + 4 * cg);
        if constexpr (BWD) w[j] = ldg_stream4(y + r * cols + 4 * cg);
      }
    }
#pragma unroll
    for (int j = 0; j < kUnroll; ++j) {
      const int64_t r = r0 + j * R;
      if (r >= rows) continue;
      const int64_t g = r * cg4 + cg;  // float4 group index = flat element index / 4
      float k0 = 1.f, k1 = 1.f, k2 = 1.f, k3 = 1.f;
      if (ap.drop_thresh != 0u) {
        const uint32_t h0 = drop_hash_pair(key, (uint32_t)(2 * g)), h1 = drop_hash_pair(key, (uint32_t)(2 * g + 1));
        k0 = (h0 & 0xFFFFu) >= ap.drop_thresh ? ap.inv_keep : 0.f;
        k1 = (h0 >> 16) >= ap.drop_thresh ? ap.inv_keep : 0.f;
        k2 = (h1 & 0xFFFFu) >= ap.drop_thresh ? ap.inv_keep : 0.f;
        k3 = (h1 >> 16) >= ap.drop_thresh ? ap.inv_keep : 0.f;
      }
      float4 o;
      if constexpr (BWD) {
        o = make_float4(v[j].x * k0 * gelu_grad(w[j].x), v[j].y * k1 * gelu_grad(w[j].y), v[j].z * k2 * gelu_grad(w[j].z),
                        v[j].w * k3 * gelu_grad(w[j].w));
      } else {
        o = make_float4(gelu_fwd(v[j].x) * k0, gelu_fwd(v[j].y) * k1, gelu_fwd(v[j].z) * k2, gelu_fwd(v[j].w) * k3);
      }
      uint16_t h0, h1, h2, h3, l0, l1, l2, l3;
      split2(fmt, o.x, h0, l0); split2(fmt, o.y, h1, l1); split2(fmt, o.z, h2, l2); split2(fmt, o.w, h3, l3);
      *reinterpret_cast<uint2*>(hi + 4 * g) = make_uint2((uint32_t)h0 | ((uint32_t)h1 << 16), (uint32_t)h2 | ((uint32_t)h3 << 16));
      if constexpr (HAS_LO)
        *reinterpret_cast<uint2*>(lo + 4 * g) = make_uint2((uint32_t)l0 | ((uint32_t)l1 << 16), (uint32_t)l2 | ((uint32_t)l3 << 16));
      if constexpr (BWD) { acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w; }
    }
  }
  if constexpr (BWD) *reinterpret_cast<float4*>(part + rl * cols + 4 * cg) = acc;
}

// ------------------------------------------------------------------------------------
// Gated residual of the ternary transformer block (models/text_encoder.py:238-249):
//   out = src + dropout(h) * g          g = sigmoid(gate), a device scalar
// forward: one pass (read src, h; write out) instead of dropout + broadcast-multiply + add;
// backward: dh = dout * g * keep/(1-p) and the per-CTA partials of dg = sum(dout .* dropout(h)) in one pass
// over (dout, h); d(src) is dout itself.  The mask comes from the counter hash (seed, flat index).  n % 4 == 0.
// ------------------------------------------------------------------------------------
template <bool BWD>
__global__ void __launch_bounds__(kThreads)
    gated_residual_kernel(const float* __restrict__ a, const float* __restrict__ h, const float* __restrict__ g_p, int64_t n,
                          float* __restrict__ out, float* __restrict__ part, const ActParams ap) {
  const float gate = __ldg(g_p);
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t key = 0;
  if (ap.drop_thresh != 0u) {
    const unsigned long long seed = ap.seed != nullptr ? *ap.seed : 0ull;
    key = drop_row_key((uint32_t)seed, (uint32_t)(seed >> 32), 0x6A7EDu);
  }
  float acc = 0.f;
  for (int64_t g0 = tid; g0 < n4; g0 += stride * kUnroll) {
    float4 x[kUnroll], y[kUnroll];
#pragma unroll
    for (int j = 0; j < kUnroll; ++j) {
      const int64_t g = g0 + j * stride;
      if (g < n4) { x[j] = ldg_stream4(a + 4 * g); y[j] = ldg_stream4(h + 4 * g); }
    }
#pragma unroll
    for (int j = 0; j < kUnroll; ++j) {
      const int64_t g = g0 + j * stride;
      if (g >= n4) continue;
      float k0 = 1.f, k1 = 1.f, k2 = 1.f, k3 = 1.f;
      if (ap.drop_thresh != 0u) {
        const uint32_t h0 = drop_hash_pair(key, (uint32_t)(2 * g)), h1 = drop_hash_pair(key, (uint32_t)(2 * g + 1));
        k0 = (h0 & 0xFFFFu) >= ap.drop_thresh ? ap.inv_keep : 0.f;
        k1 = (h0 >> 16) >= ap.drop_thresh ? ap.inv_keep : 0.f;
        k2 = (h1 & 0xFFFFu) >= ap.drop_thresh ? ap.inv_keep : 0.f;
        k3 = (h1 >> 16) >= ap.drop_thresh ? ap.inv_keep : 0.f;
      }
      float4 o;
      if constexpr (BWD) {  // a = dout
        o = make_float4(x[j].x * gate * k0, x[j].y * gate * k1, x[j].z * gate * k2, x[j].w * gate * k3);
        acc += (x[j].x * (y[j].x * k0) + x[j].y * (y[j].y * k1)) + (x[j].z * (y[j].z * k2) + x[j].w * (y[j].w * k3));
      } else {              // a = src
        o = make_float4(x[j].x + (y[j].x * k0) * gate, x[j].y + (y[j].y * k1) * gate, x[j].z + (y[j].z * k2) * gate,
                        x[j].w + (y[j].w * k3) * gate);
      }
      *reinterpret_cast<float4*>(out + 4 * g) = o;
    }
  }
  if constexpr (BWD) {
    __shared__ float s_w[kThreads / 32];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < kThreads / 32; ++i) t += s_w[i];
      part[blockIdx.x] = t;
    }
  }
}

// fixed-order final sum of per-CTA partials (deterministic)
__global__ void __launch_bounds__(256) sum_partials_kernel(const float* __restrict__ part, int n, float* __restrict__ out) {
  __shared__ double s[256];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) acc += (double)part[i];
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = (float)s[0];
}

// ------------------------------------------------------------------------------------
// AdamW over a table of tensors in ONE launch (train_multimodal.py:361-366 uses torch.optim.AdamW; torch's fused
// multi-tensor kernel runs at ~13 % of the copy roofline on the ~120 small tensors of config 2).
// Work unit = 1024 consecutive elements of one tensor (chunk table built by the host); same update as
// torch.optim.AdamW (decoupled weight decay, bias correction from the device step counter so that CUDA-graph
// replays advance it).  p, g, m, v of a tensor share one memory layout (elementwise, layout-agnostic).
// ------------------------------------------------------------------------------------
struct AdamwTensor {
  float* p;
  const float* g;
  float* m;
  float* v;
  long long n;
};
constexpr int kAdamChunk = 1024;

__global__ void __launch_bounds__(256)
    adamw_multi_kernel(const AdamwTensor* __restrict__ table, const int* __restrict__ chunk_tensor, const int* __restrict__ chunk_off,
                       int n_chunks, float lr, float beta1, float beta2, float eps, float weight_decay, const float* __restrict__ step_p) {
  const float t = __ldg(step_p) + 1.f;
  const float bc1 = 1.f - powf(beta1, t);
  const float bc2_sqrt = sqrtf(1.f - powf(beta2, t));
  const float step_size = lr / bc1;
  const float decay = 1.f - lr * weight_decay;
  for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    const AdamwTensor T = table[chunk_tensor[c]];
    const long long base = (long long)chunk_off[c] * kAdamChunk;
    const long long i = base + 4 * (long long)threadIdx.x;
    // 128-bit path only when all four tensors are 16-byte aligned (gradient views into a flat all-reduce buffer may not be)
    const bool vec = ((reinterpret_cast<uintptr_t>(T.p) | reinterpret_cast<uintptr_t>(T.g) | reinterpret_cast<uintptr_t>(T.m) |
                       reinterpret_cast<uintptr_t>(T.v)) & 15u) == 0;
    if (vec && i + 3 < T.n) {
      const float4 g = *reinterpret_cast<const float4*>(T.g + i);
      float4 p = *reinterpret_cast<const float4*>(T.p + i);
      float4 m = *reinterpret_cast<const float4*>(T.m + i);
      float4 v = *reinterpret_cast<const float4*>(T.v + i);
#define ATQ_ADAM1(X)                                        \
  m.X = beta1 * m.X + (1.f - beta1) * g.X;                  \
  v.X = beta2 * v.X + (1.f - beta2) * g.X * g.X;            \
  p.X = p.X * decay - step_size * (m.X / (sqrtf(v.X) / bc2_sqrt + eps));
      ATQ_ADAM1(x) ATQ_ADAM1(y) ATQ_ADAM1(z) ATQ_ADAM1(w)
#undef ATQ_ADAM1
      *reinterpret_cast<float4*>(T.p + i) = p;
      *reinterpret_cast<float4*>(T.m + i) = m;
      *reinterpret_cast<float4*>(T.v + i) = v;
    } else {
      for (long long j = i; j < T.n && j < i + 4; ++j) {
        const float g = T.g[j];
        const float m = beta1 * T.m[j] + (1.f - beta1) * g;
        const float v = beta2 * T.v[j] + (1.f - beta2) * g * g;
        T.m[j] = m;
        T.v[j] = v;
        T.p[j] = T.p[j] * decay - step_size * (m / (sqrtf(v) / bc2_sqrt + eps));
      }
    }
  }
}
__global__ void adamw_advance_kernel(float* step_p) { *step_p += 1.f; }

template <bool HAS_LO>
__global__ void __launch_bounds__(kThreads)
    split_rows_kernel(const float* __restrict__ x, int64_t rows, int64_t cols, int64_t ld_in,
                      uint16_t* __restrict__ hi, uint16_t* __restrict__ lo, int64_t pitch, const float* __restrict__ fslot) {
  const OperandFmt fmt = load_fmt(fslot);
  // general strided case: one row per CTA step, scalar coalesced accesses
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
    const float* xr = x + r * ld_in;
    for (int64_t c = threadIdx.x; c < cols; c += blockDim.x) {
      uint16_t h, l;
      split2(fmt, __ldg(xr + c), h, l);
      hi[r * pitch + c] = h;
      if constexpr (HAS_LO) lo[r * pitch + c] = l;
    }
  }
}

// transposed split: x [rows, cols] -> hi_t/lo_t [cols, pitch_t]; 64x64 tiles through shared memory
constexpr int kTile = 64;
template <bool HAS_LO>
__global__ void __launch_bounds__(256)
    split_t_kernel(const float* __restrict__ x, int64_t rows, int64_t cols, int64_t ld_in,
                   uint16_t* __restrict__ hi_t, uint16_t* __restrict__ lo_t, int64_t pitch_t, const float* __restrict__ fslot) {
  const OperandFmt fmt = load_fmt(fslot);
  __shared__ float tile[kTile][kTile + 1];
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;  // 64 x 4
  const int64_t c0 = (int64_t)blockIdx.x * kTile, r0 = (int64_t)blockIdx.y * kTile;
#pragma unroll 4
  for (int rr = ty; rr < kTile; rr += 4) {
    int64_t r = r0 + rr, c = c0 + tx;
    tile[rr][tx] = (r < rows && c < cols) ? __ldg(x + r * ld_in + c) : 0.f;
  }
  __syncthreads();
#pragma unroll 4
  for (int cc = ty; cc < kTile; cc += 4) {
    int64_t c = c0 + cc, r = r0 + tx;
    if (c < cols && r < rows) {
      uint16_t h, l;
      split2(fmt, tile[tx][cc], h, l);
      hi_t[c * pitch_t + r] = h;
      if constexpr (HAS_LO) lo_t[c * pitch_t + r] = l;
    }
  }
}

// ------------------------------------------------------------------------------------
// Quantize-and-build: one pass over W [M,K] producing the 2-bit codec bytes and the bf16
// GEMM operands in both orientations.  MIXED adds the RPB formula
// Wm = T*alpha*(1-mask) + W*mask  (atq/precision_boost.py:72).
// ------------------------------------------------------------------------------------
template <bool MIXED>
__global__ void __launch_bounds__(256)
    build_operands_kernel(const float* __restrict__ w, const float* __restrict__ mask, int64_t M, int64_t K,
                          const float* __restrict__ thr_p, const float* __restrict__ alpha_p,
                          uint8_t* __restrict__ packed, uint8_t* __restrict__ packed_t,
                          uint16_t* __restrict__ hi, uint16_t* __restrict__ lo, int64_t pitch,
                          uint16_t* __restrict__ hi_t, uint16_t* __restrict__ lo_t, int64_t pitch_t,
                          TernStats* __restrict__ stats, const float* __restrict__ fslot, int force_f16) {
  const OperandFmt fmt = load_fmt(fslot, force_f16);
  __shared__ float tile[kTile][kTile + 1];      // value that goes to the bf16 / fp16 operands
  __shared__ uint8_t codes[kTile][kTile + 4];   // 2-bit codes of T (for the codec bytes)
  const float thr = __ldg(thr_p);
  const float alpha = MIXED ? __ldg(alpha_p) : 1.f;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  const int64_t k0 = (int64_t)blockIdx.x * kTile, m0 = (int64_t)blockIdx.y * kTile;
  unsigned long long nnz = 0;
  double swt = 0.0;
  // all 16 (+16 mask) loads of this thread are issued before any of them is consumed
  float xs[kTile / 4], mks[kTile / 4];
#pragma unroll
  for (int it = 0; it < kTile / 4; ++it) {
    const int64_t m = m0 + ty + 4 * it, k = k0 + tx;
    const bool ok = (m < M && k < K);
    xs[it] = ok ? __ldg(w + m * K + k) : 0.f;
    if constexpr (MIXED) mks[it] = ok ? __ldg(mask + m * K + k) : 0.f;
  }
#pragma unroll
  for (int it = 0; it < kTile / 4; ++it) {
    const int rr = ty + 4 * it;
    const int64_t m = m0 + rr, k = k0 + tx;
    float val = 0.f;
    uint32_t c = 1u;
    if (m < M && k < K) {
      const float x = xs[it];
      c = tern_code(x, thr);
      const float t = (float)c - 1.f;
      if constexpr (MIXED) {
        const float mk = mks[it];
        // same expression order as the reference: ((T*alpha)*(1-mask)) + (W*mask), fp32
        val = (t * alpha) * (1.f - mk) + x * mk;
      } else {
        val = t;
        nnz += (c != 1u);
        swt += (double)(x * t);
      }
      if (hi != nullptr) {
        uint16_t h, l;
        split2(fmt, val, h, l);
        hi[m * pitch + k] = h;
        if (MIXED && lo != nullptr) lo[m * pitch + k] = l;
      }
    }
    tile[rr][tx] = val;
    codes[rr][tx] = (uint8_t)c;
  }
  __syncthreads();
  if (packed != nullptr) {  // K % 4 == 0 (checked by the host): 16 bytes per tile row
    for (int i = threadIdx.x; i < kTile * (kTile / 4); i += 256) {
      int rr = i >> 4, bj = i & 15;
      int64_t m = m0 + rr, k = k0 + 4 * bj;
      if (m < M && k < K) {
        uint32_t b = codes[rr][4 * bj] | (codes[rr][4 * bj + 1] << 2) | (codes[rr][4 * bj + 2] << 4) | (codes[rr][4 * bj + 3] << 6);
        // columns beyond K inside this byte cannot occur because K % 4 == 0
        packed[(m * K + k) >> 2] = (uint8_t)b;
      }
    }
  }
  if (packed_t != nullptr) {  // codec bytes of T^T [K, M/4] (M % 4 == 0 checked by the host)
    for (int i = threadIdx.x; i < kTile * (kTile / 4); i += 256) {
      int kk = i >> 4, bj = i & 15;
      int64_t k = k0 + kk, m = m0 + 4 * bj;
      if (k < K && m < M) {
        uint32_t b = codes[4 * bj][kk] | (codes[4 * bj + 1][kk] << 2) | (codes[4 * bj + 2][kk] << 4) | (codes[4 * bj + 3][kk] << 6);
        packed_t[(k * M + m) >> 2] = (uint8_t)b;
      }
    }
  }
  if (hi_t != nullptr) {
#pragma unroll 4
    for (int cc = ty; cc < kTile; cc += 4) {
      int64_t k = k0 + cc, m = m0 + tx;
      if (k < K && m < M) {
        uint16_t h, l;
        split2(fmt, tile[tx][cc], h, l);
        hi_t[k * pitch_t + m] = h;
        if (MIXED && lo_t != nullptr) lo_t[k * pitch_t + m] = l;
      }
    }
  }
  if constexpr (!MIXED) {
    if (stats != nullptr) {
      __shared__ unsigned long long s_n[8];
      __shared__ double s_s[8];
      nnz = warp_sum(nnz);
      swt = warp_sum(swt);
      const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
      if (lane == 0) { s_n[wid] = nnz; s_s[wid] = swt; }
      __syncthreads();
      if (threadIdx.x == 0) {
        unsigned long long tn = 0; double ts = 0.0;
        for (int i = 0; i < 8; ++i) { tn += s_n[i]; ts += s_s[i]; }
        atomicAdd(&stats->nnz, tn);
        atomicAdd(&stats->sum_wt, ts);
      }
    }
  }
}

// ------------------------------------------------------------------------------------
// LayerNorm over the last dimension (the op either side of every ternary GEMM of the transformer block,
// models/text_encoder.py:77,232,244; SURVEY 8f rank 2 "fuse LayerNorm into the GEMM prologue"): one warp per row, the
// row lives in registers (cols <= 1024, cols % 4 == 0), fp32 two-pass statistics.
//   forward : y = (x - mean) * rstd * gamma + beta, saves mean / rstd per row and folds max|y| into a scale slot, so
//             the operand split that follows needs no reduction pass of its own;
//   backward: dx = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * gamma, plus per-CTA partial column sums of
//             dy * xhat and dy (reduced in fixed order by colsum_stage2_kernel: deterministic).
// ------------------------------------------------------------------------------------
constexpr int kLnVec = 8;  // float4 per lane: cols <= 32 * 8 * 4 = 1024

__global__ void __launch_bounds__(kThreads)
    layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, int64_t rows,
                         int cols, float eps, float* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                         unsigned int* __restrict__ absmax_bits) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (kThreads / 32);
  const int nv = cols >> 2;  // float4 per row
  const float inv_n = 1.f / (float)cols;
  float amax = 0.f;
  for (int64_t r = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); r < rows; r += warps) {
    const float4* xr = reinterpret_cast<const float4*>(x + r * cols);
    float4 v[kLnVec];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < kLnVec; ++i) {
      const int c = lane + 32 * i;
      v[i] = c < nv ? ldg_stream4(reinterpret_cast<const float*>(xr + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
      sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float mean = warp_sum(sum) * inv_n;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < kLnVec; ++i) {
      if (lane + 32 * i < nv) {
        const float a = v[i].x - mean, b = v[i].y - mean, c2 = v[i].z - mean, d = v[i].w - mean;
        sq += (a * a + b * b) + (c2 * c2 + d * d);
      }
    }
    const float rstd = rsqrtf(warp_sum(sq) * inv_n + eps);
    float4* yr = reinterpret_cast<float4*>(y + r * cols);
#pragma unroll
    for (int i = 0; i < kLnVec; ++i) {
      const int c = lane + 32 * i;
      if (c < nv) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c);
        const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + c);
        const float4 o = make_float4((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y,
                                     (v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w);
        yr[c] = o;
        amax = fmaxf(fmaxf(amax, fmaxf(fabsf(o.x), fabsf(o.y))), fmaxf(fabsf(o.z), fabsf(o.w)));
      }
    }
    if (lane == 0) {
      mean_out[r] = mean;
      rstd_out[r] = rstd;
    }
  }
  if (absmax_bits != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    __shared__ float s_m[kThreads / 32];
    if (lane == 0) s_m[threadIdx.x >> 5] = amax;
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
      for (int i = 0; i < kThreads / 32; ++i) amax = fmaxf(amax, s_m[i]);
      atomicMax(absmax_bits, __float_as_uint(amax));
      __threadfence();
      if (atomicAdd(absmax_bits + 3, 1u) == gridDim.x - 1) {  // last CTA: finish the slot and re-arm it (graph replays)
        __threadfence();
        float sc, inv;
        pow2_scale_for(__uint_as_float(atomicExch(absmax_bits, 0u)), sc, inv);
        float* slot = reinterpret_cast<float*>(absmax_bits);
        slot[1] = sc;
        slot[2] = inv;
        absmax_bits[3] = 0u;
      }
    }
  }
}

__global__ void __launch_bounds__(kThreads)
    layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ gamma,
                         const float* __restrict__ mean_in, const float* __restrict__ rstd_in, int64_t rows, int cols,
                         float* __restrict__ dx, float* __restrict__ part_g, float* __restrict__ part_b) {
  extern __shared__ float s_red[];  // [kThreads / 32][cols]
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t warps = (int64_t)gridDim.x * (kThreads / 32);
  const int nv = cols >> 2;
  const float inv_n = 1.f / (float)cols;
  float4 ag[kLnVec], ab[kLnVec];  // this lane's running column sums of dy * xhat and dy
#pragma unroll
  for (int i = 0; i < kLnVec; ++i) ag[i] = ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t r = (int64_t)blockIdx.x * (kThreads / 32) + wid; r < rows; r += warps) {
    const float mean = __ldg(mean_in + r), rstd = __ldg(rstd_in + r);
    const float4* xr = reinterpret_cast<const float4*>(x + r * cols);
    const float4* gr = reinterpret_cast<const float4*>(dy + r * cols);
    float4 xh[kLnVec], g[kLnVec];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < kLnVec; ++i) {
      const int c = lane + 32 * i;
      if (c < nv) {
        xh[i] = ldg_stream4(reinterpret_cast<const float*>(xr + c));
        g[i] = ldg_stream4(reinterpret_cast<const float*>(gr + c));
      } else {
        xh[i] = g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int i = 0; i < kLnVec; ++i) {
      const int c = lane + 32 * i;
      if (c < nv) {
        const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + c);
        xh[i] = make_float4((xh[i].x - mean) * rstd, (xh[i].y - mean) * rstd, (xh[i].z - mean) * rstd, (xh[i].w - mean) * rstd);
        ag[i].x += g[i].x * xh[i].x; ag[i].y += g[i].y * xh[i].y; ag[i].z += g[i].z * xh[i].z; ag[i].w += g[i].w * xh[i].w;
        ab[i].x += g[i].x; ab[i].y += g[i].y; ab[i].z += g[i].z; ab[i].w += g[i].w;
        g[i] = make_float4(g[i].x * gm.x, g[i].y * gm.y, g[i].z * gm.z, g[i].w * gm.w);
        s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
        s2 += (g[i].x * xh[i].x + g[i].y * xh[i].y) + (g[i].z * xh[i].z + g[i].w * xh[i].w);
      }
    }
    s1 = warp_sum(s1) * inv_n;
    s2 = warp_sum(s2) * inv_n;
    float4* dr = reinterpret_cast<float4*>(dx + r * cols);
#pragma unroll
    for (int i = 0; i < kLnVec; ++i) {
      const int c = lane + 32 * i;
      if (c < nv)
        dr[c] = make_float4(rstd * (g[i].x - s1 - xh[i].x * s2), rstd * (g[i].y - s1 - xh[i].y * s2),
                            rstd * (g[i].z - s1 - xh[i].z * s2), rstd * (g[i].w - s1 - xh[i].w * s2));
    }
  }
  // CTA-level column sums: the 8 warps' lanes own the same columns; fixed-order combine through shared memory
  for (int pass = 0; pass < 2; ++pass) {
    float4* mine = reinterpret_cast<float4*>(s_red + (size_t)wid * cols);
#pragma unroll
    for (int i = 0; i < kLnVec; ++i) {
      const int c = lane + 32 * i;
      if (c < nv) mine[c] = pass == 0 ? ag[i] : ab[i];
    }
    __syncthreads();
    float* out = (pass == 0 ? part_g : part_b) + (size_t)blockIdx.x * cols;
    for (int c = threadIdx.x; c < cols; c += kThreads) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < kThreads / 32; ++w) t += s_red[(size_t)w * cols + c];
      out[c] = t;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------
// column sums (bias gradient): deterministic two-stage reduction
// ------------------------------------------------------------------------------------
// rows per CTA: 64 for short inputs (enough CTAs to hide latency), up to 512 for long ones
static inline int colsum_rows_per_cta(int64_t rows) { return rows <= 8192 ? 64 : (rows <= 65536 ? 256 : 512); }
__global__ void __launch_bounds__(256)
    colsum_stage1_kernel(const float* __restrict__ x, int64_t rows, int64_t cols, int64_t ld, float* __restrict__ part,
                         int rows_per_cta) {
  // CTA = 32 columns x rows_per_cta rows; 8 warps take interleaved rows
  __shared__ float s[8][33];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t c = (int64_t)blockIdx.x * 32 + lane;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_cta;
  const int64_t r1 = min(rows, r0 + rows_per_cta);
  float acc = 0.f;
  if (c < cols)
    for (int64_t r = r0 + wid; r < r1; r += 8) acc += __ldg(x + r * ld + c);
  s[wid][lane] = acc;
  __syncthreads();
  if (wid == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += s[i][lane];
    part[(int64_t)blockIdx.y * cols + c] = t;
  }
}
__global__ void __launch_bounds__(256)
    colsum_stage2_kernel(const float* __restrict__ part, int64_t nparts, int64_t cols, float* __restrict__ out) {
  // CTA = 32 columns; its 8 warps take interleaved partial rows (4 loads in flight each), fixed-order
  // shared-memory combine: deterministic and no long serial chain
  __shared__ float s[8][33];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t c = (int64_t)blockIdx.x * 32 + lane;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (c < cols) {
    int64_t p = wid;
    for (; p + 24 < nparts; p += 32) {
      a0 += part[p * cols + c];
      a1 += part[(p + 8) * cols + c];
      a2 += part[(p + 16) * cols + c];
      a3 += part[(p + 24) * cols + c];
    }
    for (; p < nparts; p += 8) a0 += part[p * cols + c];
  }
  s[wid][lane] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (wid == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += s[i][lane];
    out[c] = t;
  }
}

// small-N path: one CTA per 32 columns walks every row (single launch, deterministic)
__global__ void __launch_bounds__(256)
    colsum_small_kernel(const float* __restrict__ x, int64_t rows, int64_t cols, int64_t ld, float* __restrict__ out) {
  __shared__ float s[8][33];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t c = (int64_t)blockIdx.x * 32 + lane;
  float acc0 = 0.f, acc1 = 0.f;
  if (c < cols) {
    int64_t r = wid;
    for (; r + 8 < rows; r += 16) {
      acc0 += __ldg(x + r * ld + c);
      acc1 += __ldg(x + (r + 8) * ld + c);
    }
    if (r < rows) acc0 += __ldg(x + r * ld + c);
  }
  s[wid][lane] = acc0 + acc1;
  __syncthreads();
  if (wid == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += s[i][lane];
    out[c] = t;
  }
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace atq

using namespace atq;

template <int KIND>
static int launch_unpack(int device, const uint8_t* packed, int64_t n, void* out, int32_t* flag, cudaStream_t stream) {
  int grid = stream_grid(device, (n >> 2) + 1, kThreads * kUnroll, 8);
  if (aligned16(out)) unpack2_kernel<KIND, true><<<grid, kThreads, 0, stream>>>(packed, n, out, flag);
  else unpack2_kernel<KIND, false><<<grid, kThreads, 0, stream>>>(packed, n, out, flag);
  ATQ_LAUNCH_CHECK();
  return ATQ_OK;
}

// Whole-model codec launcher (defined below); the per-layer entry points route large, 16-byte aligned layers through
// it with count = 1: 128-bit packed stores / loads (64 weights per access) instead of one code byte per thread.
// ATQ_CODEC_PER_LAYER_KERNELS=1 keeps the simple per-layer kernels (A/B measurements).
template <int MODE>
static int codec_batched(int device, int count, const float* const* src, uint8_t* const* packed, float* const* dst,
                         const float* const* thr, const int64_t* ns, int32_t* invalid_flag, cudaStream_t stream);
constexpr int64_t kCodecRouteMinN = 1ll << 21;  // measured: equal at 2M weights, 5-15 % faster at 16M, slower below 1M
static bool codec_route(int64_t n, const void* a, const void* b) {
  static const bool per_layer = [] { const char* e = getenv("ATQ_CODEC_PER_LAYER_KERNELS"); return e != nullptr && e[0] == '1'; }();
  return !per_layer && n >= kCodecRouteMinN && aligned16(a) && aligned16(b);
}

extern "C" {

size_t atq_workspace_bytes_abs_stats(int64_t) { return 0; }

int atq_abs_stats(int device, const float* w, int64_t n, void* stats_out, void*, size_t, atq_stream_t stream_) {
  ATQ_CHECK_ARG(w != nullptr && stats_out != nullptr && n > 0, "null pointer or n <= 0");
  ATQ_ENSURE_DEVICE(device);
  cudaStream_t stream = (cudaStream_t)stream_;
  ATQ_CUDA(cudaMemsetAsync(stats_out, 0, sizeof(AbsStats), stream));
  int grid = stream_grid(device, (n >> 2) + 1, kThreads * kUnroll, 8);
  if (aligned16(w)) abs_stats_kernel<true><<<grid, kThreads, 0, stream>>>(w, n, (AbsStats*)stats_out);
  else abs_stats_kernel<false><<<grid, kThreads, 0, stream>>>(w, n, (AbsStats*)stats_out);
  ATQ_LAUNCH_CHECK();
  return ATQ_OK;
}

static int launch_ternarize(int device, const float* w, int64_t n, const float* thr, float* t_out, uint8_t* packed,
                            void* stats, cudaStream_t stream) {
  int grid = stream_grid(device, (n >> 2) + 1, kThreads * kUnroll, 8);
  const bool vec = aligned16(w) && (t_out == nullptr || aligned16(t_out));
  TernStats* st = (TernStats*)stats;
#define ATQ_TERN(V, O, S) ternarize_kernel<V, SRC_THRESHOLD, O, S><<<grid, kThreads, 0, stream>>>(w, n, thr, t_out, packed, st, nullptr)
  if (t_out != nullptr) {
    if (vec) { if (st) ATQ_TERN(true, OUT_F32, true); else ATQ_TERN(true, OUT_F32, false); }
    else     { if (st) ATQ_TERN(false, OUT_F32, true); else ATQ_TERN(false, OUT_F32, false); }
  } else {
    if (vec) { if (st) ATQ_TERN(true, OUT_PACK2, true); else ATQ_TERN(true, OUT_PACK2, false); }
    else     { if (st) ATQ_TERN(false, OUT_PACK2, true); else ATQ_TERN(false, OUT_PACK2, false); }
  }
#undef ATQ_TERN
  ATQ_LAUNCH_CHECK();
  return ATQ_OK;
}

int atq_ternarize_f32(int device, const float* w, int64_t n, const float* thr, float* t_out, void* stats,
                      atq_stream_t stream) {
  ATQ_CHECK_ARG(w && thr && t_out && n > 0, "null pointer or n <= 0");
  ATQ_ENSURE_DEVICE(device);
  return launch_ternarize(device, w, n, thr, t_out, nullptr, stats, (cudaStream_t)stream);
}

int atq_ternarize_pack2(int device, const float* w, int64_t n, const float* thr, uint8_t* packed, void* stats,
                        atq_stream_t stream) {
  ATQ_CHECK_ARG(w && thr && packed && n > 0, "null pointer or n <= 0");
  ATQ_ENSURE_DEVICE(device);
  if (stats == nullptr && codec_route(n, w, packed))
    return codec_batched<CODEC_TERNARIZE>(device, 1, &w, &packed, nullptr, &thr, &n, nullptr, (cudaStream_t)stream);
  return launch_ternarize(device, w, n, thr, nullptr, packed, stats, (cudaStream_t)stream);
}

int atq_optimal_alpha(int device, const void* tern_stats, const void* abs_stats, int64_t n, float* alpha_out,
                      atq_stream_t stream) {
  ATQ_CHECK_ARG(tern_stats && abs_stats && alpha_out && n > 0, "null pointer or n <= 0");
  ATQ_ENSURE_DEVICE(device);
  optimal_alpha_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((const TernStats*)tern_stats, (const AbsStats*)abs_stats, n, alpha_out);
  ATQ_LAUNCH_CHECK();
  return ATQ_OK;
}

int atq_pack2_from_f32(int device, const float* t, int64_t n, uint8_t* packed, int32_t* invalid_flag,
                       atq_stream_t stream_) {
  ATQ_CHECK_ARG(t && packed && invalid_flag && n > 0, "null pointer or n <= 0");
  ATQ_ENSURE_DEVICE(device);
  cudaStream_t stream = (cudaStream_t)stream_;
  if (codec_route(n, t, packed)) return codec_batched<CODEC_PACK>(device, 1, &t, &packed, nullptr, nullptr, &n, invalid_flag, stream);
  int grid = stream_grid(device, (n >> 2) + 1, kThreads * kUnroll, 8);
  if (aligned16(t))
    ternarize_kernel<true, SRC_TERNARY, OUT_PACK2, false><<<grid, kThreads, 0, stream>>>(t, n, nullptr, nullptr, packed, nullptr, invalid_flag);
  else
    ternarize_kernel<false, SRC_TERNARY, OUT_PACK2, false><<<grid, kThreads, 0, stream>>>(t, n, nullptr, nullptr, packed, nullptr, invalid_flag);
  ATQ_LAUNCH_CHECK();
  return ATQ_OK;
}

int atq_unpack2_to_f32(int device, const uint8_t* packed, int64_t n, float* out, int32_t* invalid_flag, atq_stream_t stream) {
  ATQ_CHECK_ARG(packed && out && n > 0, "null pointer or n <= 0");
  ATQ_ENSURE_DEVICE(device);
  if (codec_route(n, packed, out)) {
    uint8_t* pk = const_cast<uint8_t*>(packed);  // read-only in CODEC_UNPACK
    return codec_batched<CODEC_UNPACK>(device, 1, nullptr, &pk, &out, nullptr, &n, invalid_flag, (cudaStream_t)stream);
  }
  return launch_unpack<UNPACK_F32>(device, packed, n, out, invalid_flag, (cudaStream_t)stream);
}
int atq_unpack2_to_bf16(int device, const uint8_t* packed, int64_t n, uint16_t* out, atq_stream_t stream) {
  ATQ_CHECK_ARG(packed && out && n > 0, "null pointer or n <= 0");
  ATQ_ENSURE_DEVICE(device);
  return launch_unpack<UNPACK_BF16>(device, packed, n, out, nullptr, (cudaStream_t)stream);
}
int atq_unpack2_to_i8(int device, const uint8_t* packed, int64_t n, int8_t* out, atq_stream_t stream) {
  ATQ_CHECK_ARG(packed && out && n > 0, "null pointer or n <= 0");
  ATQ_ENSURE_DEVICE(device);
  return launch_unpack<UNPACK_I8>(device, packed, n, out, nullptr, (cudaStream_t)stream);
}

}  // extern "C" (the launcher below is a template)

template <int MODE>
static int codec_batched(int device, int count, const float* const* src, uint8_t* const* packed, float* const* dst,
                         const float* const* thr, const int64_t* ns, int32_t* invalid_flag, cudaStream_t stream) {
  for (int base = 0; base < count; base += kCodecBatch) {
    const int m = count - base < kCodecBatch ? count - base : kCodecBatch;
    CodecBatch b;
    memset(&b, 0, sizeof(b));
    int64_t nmax = 1;
    for (int i = 0; i < m; ++i) {
      const int k = base + i;
      if (ns[k] <= 0 || packed[k] == nullptr || !aligned16(packed[k]) || (MODE != CODEC_UNPACK && (src[k] == nullptr || !aligned16(src[k]))) ||
          (MODE == CODEC_UNPACK && (dst[k] == nullptr || !aligned16(dst[k]))) || (MODE == CODEC_TERNARIZE && thr[k] == nullptr)) {
        set_error("batched codec: layer %d has a null / unaligned pointer or no elements", k);
        return ATQ_EINVAL;
      }
      b.src[i] = MODE != CODEC_UNPACK ? src[k] : nullptr;
      b.packed[i] = packed[k];
      b.dst[i] = MODE == CODEC_UNPACK ? dst[k] : nullptr;
      b.thr[i] = MODE == CODEC_TERNARIZE ? thr[k] : nullptr;
      b.n[i] = ns[k];
      if (ns[k] > nmax) nmax = ns[k];
    }
    b.invalid_flag = invalid_flag;
    // CTAs per layer: enough for the largest layer, the whole grid a few waves of the machine
    int64_t per = ((nmax >> 11) + (kThreads / 32) - 1) / (kThreads / 32);
    const int64_t cap = ((int64_t)sm_count(device) * 8 + m - 1) / m;
    if (per > cap) per = cap;
    if (per < 1) per = 1;
    codec_batched_kernel<MODE><<<dim3((unsigned)per, (unsigned)m), kThreads, 0, stream>>>(b);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
      set_error("batched codec launch failed: %s", cudaGetErrorString(e));
      return ATQ_ECUDA;
    }
    note_launch();
  }
  return ATQ_OK;
}

extern "C" {

int atq_ternarize_pack2_batched(int device, int count, const float* const* w_ptrs, const int64_t* ns, const float* const* thr_ptrs,
                                uint8_t* const* packed_ptrs, atq_stream_t stream_) {
  ATQ_CHECK_ARG(count > 0 && w_ptrs && ns && thr_ptrs && packed_ptrs, "null pointer or count <= 0");
  ATQ_ENSURE_DEVICE(device);
  return codec_batched<CODEC_TERNARIZE>(device, count, w_ptrs, packed_ptrs, nullptr, thr_ptrs, ns, nullptr, (cudaStream_t)stream_);
}

int atq_pack2_from_f32_batched(int device, int count, const float* const* t_ptrs, const int64_t* ns, uint8_t* const* packed_ptrs,
                               int32_t* invalid_flag, atq_stream_t stream_) {
  ATQ_CHECK_ARG(count > 0 && t_ptrs && ns && packed_ptrs, "null pointer or count <= 0");
  ATQ_ENSURE_DEVICE(device);
  return codec_batched<CODEC_PACK>(device, count, t_ptrs, packed_ptrs, nullptr, nullptr, ns, invalid_flag, (cudaStream_t)stream_);
}

int atq_unpack2_to_f32_batched(int device, int count, uint8_t* const* packed_ptrs, const int64_t* ns, float* const* out_ptrs,
                               int32_t* invalid_flag, atq_stream_t stream_) {
  ATQ_CHECK_ARG(count > 0 && packed_ptrs && ns && out_ptrs, "null pointer or count <= 0");
  ATQ_ENSURE_DEVICE(device);
  return codec_batched<CODEC_UNPACK>(device, count, nullptr, packed_ptrs, out_ptrs, nullptr, ns, invalid_flag, (cudaStream_t)stream_);
}

int atq_route_mask_mul(int device, const float* x, const float* grad_out, const float* thr, int64_t n, float* grad_in,
                       atq_stream_t stream_) {
  ATQ_CHECK_ARG(x && grad_out && thr && grad_in && n > 0, "null pointer or n <= 0");
  ATQ_ENSURE_DEVICE(device);
  cudaStream_t stream = (cudaStream_t)stream_;
  int grid = stream_grid(device, (n >> 2) + 1, kThreads * kUnroll, 8);
  if (aligned16(x) && aligned16(grad_out) && aligned16(grad_in))
    route_mask_mul_kernel<true><<<grid, kThreads, 0, stream>>>(x, grad_out, thr, n, grad_in);
  else
    route_mask_mul_kernel<false><<<grid, kThreads, 0, stream>>>(x, grad_out, thr, n, grad_in);
  ATQ_LAUNCH_CHECK();
  return ATQ_OK;
}

int atq_absmax_scale(int device, const float* x, int64_t rows, int64_t cols, int64_t ld, float bound_mul, const float* extra,
                     float* slot, atq_stream_t stream_) {
  ATQ_CHECK_ARG(x && slot && rows > 0 && cols > 0 && ld >= cols, "null pointer, empty shape or ld < cols");
  ATQ_CHECK_ARG(bound_mul > 0.f, "bound_mul must be positive");
  ATQ_ENSURE_DEVICE(device);
  const int vec = (ld == cols && aligned16(x)) ? 1 : 0;
  const int64_t n = rows * cols;
  int grid = vec ? stream_grid(device, (n >> 2) + 1, kThreads * 4, 4)
                 : (int)(rows < (int64_t)sm_count(device) * 4 ? rows : (int64_t)sm_count(device) * 4);
  absmax_scale_kernel<<<grid, kThreads, 0, (cudaStream_t)stream_>>>(x, rows, cols, ld, bound_mul, extra, slot, vec);
  ATQ_LAUNCH_CHECK();
  return ATQ_OK;
}

int atq_absmax_scale_batched(int device, int count, const float* const* x_ptrs, const int64_t* ns, const float* const* extra_ptrs,
                             float* const* slot_ptrs, float bound_mul, atq_stream_t stream_) {
  ATQ_CHECK_ARG(count > 0 && x_ptrs && ns && slot_ptrs && bound_mul > 0.f, "null pointer, count <= 0 or bound_mul <= 0");
  ATQ_ENSURE_DEVICE(device);
  for (int base = 0; base < count; base += kAbsmaxBatch) {
    const int m = count - base < kAbsmaxBatch ? count - base : kAbsmaxBatch;
    AbsmaxBatch b;
    memset(&b, 0, sizeof(b));
    int64_t nmax = 1;
    for (int i = 0; i < m; ++i) {
      ATQ_CHECK_ARG(x_ptrs[base + i] && slot_ptrs[base + i] && ns[base + i] > 0, "null tensor / slot or empty tensor in the batch");
      b.x[i] = x_ptrs[base + i];
      b.n[i] = ns[base + i];
      b.extra[i] = extra_ptrs ? extra_ptrs[base + i] : nullptr;
      b.slot[i] = slot_ptrs[base + i];
      if (ns[base + i] > nmax) nmax = ns[base + i];
    }
    // enough CTAs per tensor for the largest one, a whole number of waves overall
    int per = stream_grid(device, (nmax >> 2) + 1, kThreads * 4, 4);
    const int cap = (sm_count(device) * 4 + m - 1) / m;
    if (per > cap) per = cap;
    if (per < 1) per = 1;
    absmax_scale_batched_kernel<<<dim3((unsigned)per, (unsigned)m), kThreads, 0, (cudaStream_t)stream_>>>(b, bound_mul);
    ATQ_LAUNCH_CHECK();
  }
  return ATQ_OK;
}

int64_t atq_split_scaled_fused_max_elems(void) { return kFusedSplitMax; }

int atq_split_scaled_fused(int device, const float* x, int64_t n, uint16_t* hi, uint16_t* lo, float bound_mul, const float* extra,
                           float* slot, atq_stream_t stream_) {
  ATQ_CHECK_ARG(x && hi && slot && n > 0 && n <= kFusedSplitMax && (n % 8) == 0, "needs 0 < n <= max elems, n % 8 == 0");
  ATQ_CHECK_ARG(aligned16(x) && aligned16(hi) && (lo == nullptr || aligned16(lo)), "needs 16-byte aligned contiguous tensors");
  ATQ_CHECK_ARG(bound_mul > 0.f, "bound_mul must be positive");
  ATQ_ENSURE_DEVICE(device);
  cudaStream_t stream = (cudaStream_t)stream_;
  if (lo) split_scaled_cluster_kernel<true><<<kFusedCtas, kFusedThreads, 0, stream>>>(x, n >> 2, hi, lo, slot, bound_mul, extra);
  else split_scaled_cluster_kernel<false><<<kFusedCtas, kFusedThreads, 0, stream>>>(x, n >> 2, hi, lo, slot, bound_mul, extra);
  ATQ_LAUNCH_CHECK();
  return ATQ_OK;
}

int atq_split_bf16(int device, const float* x, int64_t rows, int64_t cols, int64_t ld_in, uint16_t* hi, uint16_t* lo,
                   int64_t pitch, const float* scale_slot, atq_stream_t stream_) {
  ATQ_CHECK_ARG(x && hi && rows > 0 && cols > 0, "null pointer or empty shape");
  ATQ_CHECK_ARG(ld_in >= cols && pitch >= cols && (pitch % 8) == 0, "need ld_in >= cols, pitch >= cols, pitch % 8 == 0");
  ATQ_ENSURE_DEVICE(device);
  cudaStream_t stream = (cudaStream_t)stream_;
  if (ld_in == cols && pitch == cols && aligned16(x) && aligned16(hi) && (lo == nullptr || aligned16(lo))) {
    int64_t n = rows * cols;
    int grid = stream_grid(device, n >> 2, kThreads * kUnroll, 8);
    if (lo) split_flat_kernel<true><<<grid, kThreads, 0, stream>>>(x, n, hi, lo, scale_slot);
    else split_flat_kernel<false><<<grid, kThreads, 0, stream>>>(x, n, hi, lo, scale_slot);
  } else {
    int grid = (int)(rows < (int64_t)sm_count(device) * 8 ? rows : (int64_t)sm_count(device) * 8);
    if (lo) split_rows_kernel<true><<<grid, kThreads, 0, stream>>>(x, rows, cols, ld_in, hi, lo, pitch, scale_slot);
    else split_rows_kernel<false><<<grid, kThreads, 0, stream>>>(x, rows, cols, ld_in, hi, lo, pitch, scale_slot);
  }
  ATQ_LAUNCH_CHECK();
  return ATQ_OK;
}

int atq_split_bf16_t(int device, const float* x, int64_t rows, int64_t cols, int64_t ld_in, uint16_t* hi_t,
                     uint16_t* lo_t, int64_t pitch_t, float* colsum, const float* scale_slot, atq_stream_t stream_) {
  ATQ_CHECK_ARG(x && hi_t && rows > 0 && cols > 0, "null pointer or empty shape");
  ATQ_CHECK_ARG(ld_in >= cols && pitch_t >= rows && (pitch_t % 8) == 0, "need ld_in >= cols, pitch_t >= rows, pitch_t % 8 == 0");
  ATQ_CHECK_ARG(colsum == nullptr, "fused colsum not supported in this build; use atq_colsum_f32");
  ATQ_ENSURE_DEVICE(device);
  cudaStream_t stream = (cudaStream_t)stream_;
  dim3 grid((unsigned)((cols + kTile - 1) / kTile), (unsigned)((rows + kTile - 1) / kTile));
  ATQ_CHECK_ARG(grid.y <= 65535u, "rows too large for one launch");
  if (lo_t) split_t_kernel<true><<<grid, 256, 0, stream>>>(x, rows, cols, ld_in, hi_t, lo_t, pitch_t, scale_slot);
  else split_t_kernel<false><<<grid, 256, 0, stream>>>(x, rows, cols, ld_in, hi_t, lo_t, pitch_t, scale_slot);
  ATQ_LAUNCH_CHECK();
  return ATQ_OK;
}

int atq_build_ternary_operands(int device, const float* w, int64_t M, int64_t K, const float* thr, uint8_t* packed,
                               uint8_t* packed_t, uint16_t* tb, int64_t pitch, uint16_t* tb_t, int64_t pitch_t,
                               void* stats, int fp16, atq_stream_t stream_) {
  ATQ_CHECK_ARG(w && thr && M > 0 && K > 0, "null pointer or empty shape");
  ATQ_CHECK_ARG(packed == nullptr || (K % 4) == 0, "packed output needs K % 4 == 0 (use atq_ternarize_pack2)");
  ATQ_CHECK_ARG(packed_t == nullptr || (M % 4) == 0, "packed_t output needs M % 4 == 0");
  ATQ_CHECK_ARG(tb == nullptr || (pitch >= K && pitch % 8 == 0), "bad pitch");
  ATQ_CHECK_ARG(tb_t == nullptr || (pitch_t >= M && pitch_t % 8 == 0), "bad pitch_t");
  ATQ_ENSURE_DEVICE(device);
  dim3 grid((unsigned)((K + kTile - 1) / kTile), (unsigned)((M + kTile - 1) / kTile));
  ATQ_CHECK_ARG(grid.y <= 65535u, "M too large for one launch");
  build_operands_kernel<false><<<grid, 256, 0, (cudaStream_t)stream_>>>(w, nullptr, M, K, thr, nullptr, packed, packed_t, tb, nullptr,
                                                                         pitch, tb_t, nullptr, pitch_t, (TernStats*)stats, nullptr, fp16);
  ATQ_LAUNCH_CHECK();
  return ATQ_OK;
}

int atq_build_mixed_operands(int device, const float* w, const float* mask, int64_t M, int64_t K, const float* thr,
                             const float* alpha, uint8_t* packed, uint16_t* hi, uint16_t* lo, int64_t pitch,
                             uint16_t* hi_t, uint16_t* lo_t, int64_t pitch_t, const float* scale_slot, atq_stream_t stream_) {
  ATQ_CHECK_ARG(w && mask && thr && alpha && M > 0 && K > 0, "null pointer or empty shape");
  ATQ_CHECK_ARG(packed == nullptr || (K % 4) == 0, "packed output needs K % 4 == 0 (use atq_ternarize_pack2)");
  ATQ_CHECK_ARG(hi == nullptr || (pitch >= K && pitch % 8 == 0), "bad pitch");
  ATQ_CHECK_ARG(hi_t == nullptr || (pitch_t >= M && pitch_t % 8 == 0), "bad pitch_t");
  ATQ_ENSURE_DEVICE(device);
  dim3 grid((unsigned)((K + kTile - 1) / kTile), (unsigned)((M + kTile - 1) / kTile));
  ATQ_CHECK_ARG(grid.y <= 65535u, "M too large for one launch");
  build_operands_kernel<true><<<grid, 256, 0, (cudaStream_t)stream_>>>(w, mask, M, K, thr, alpha, packed, nullptr, hi, lo, pitch, hi_t,
                                                                        lo_t, pitch_t, nullptr, scale_slot, 0);
  ATQ_LAUNCH_CHECK();
  return ATQ_OK;
}

static inline int64_t split_colsum_lanes(int device, int64_t rows, int64_t cols) {
  int64_t cg4 = cols >> 2;
  int64_t R = ((int64_t)sm_count(device) * 8 * kThreads) / (cg4 > 0 ? cg4 : 1);
  if (R > (rows + 3) / 4) R = (rows + 3) / 4;  // at least ~4 rows per lane
  if (R > 512) R = 512;                        // keeps the second stage short
  if (R < 1) R = 1;
  return R;
}

size_t atq_workspace_bytes_split_colsum(int64_t rows, int64_t cols) {
  // upper bound over devices: the lane count never exceeds rows
  int64_t R = (rows + 3) / 4;
  int64_t cap = ((int64_t)256 * 8 * kThreads) / ((cols >> 2) > 0 ? (cols >> 2) : 1);
  if (R > cap) R = cap;
  if (R > 512) R = 512;
  if (R < 1) R = 1;
  return (size_t)(R * cols * sizeof(float));
}

int atq_split_bf16_colsum(int device, const float* x, int64_t rows, int64_t cols, uint16_t* hi, uint16_t* lo,
                          float* colsum_out, void* ws, size_t ws_bytes, const float* scale_slot, atq_stream_t stream_) {
  ATQ_CHECK_ARG(x && hi && colsum_out && rows > 0 && cols > 0, "null pointer or empty shape");
  ATQ_CHECK_ARG((cols % 8) == 0 && aligned16(x) && aligned16(hi) && (lo == nullptr || aligned16(lo)),
                "needs cols % 8 == 0 and 16-byte aligned contiguous tensors");
  ATQ_ENSURE_DEVICE(device);
  const int64_t R = split_colsum_lanes(device, rows, cols);
  if (ws == nullptr || ws_bytes < (size_t)(R * cols * sizeof(float))) {
    set_error("atq_split_bf16_colsum: workspace too small");
    return ATQ_EWORKSPACE;
  }
  cudaStream_t stream = (cudaStream_t)stream_;
  const int64_t threads = (cols >> 2) * R;
  const unsigned grid = (unsigned)((threads + kThreads - 1) / kThreads);
  if (lo) split_colsum_kernel<true><<<grid, kThreads, 0, stream>>>(x, rows, cols, hi, lo, (float*)ws, R, scale_slot);
  else split_colsum_kernel<false><<<grid, kThreads, 0, stream>>>(x, rows, cols, hi, lo, (float*)ws, R, scale_slot);
  ATQ_LAUNCH_CHECK();
  colsum_stage2_kernel<<<(unsigned)((cols + 31) / 32), 256, 0, stream>>>((const float*)ws, R, cols, colsum_out);
  ATQ_LAUNCH_CHECK();
  return ATQ_OK;
}

int atq_gelu_dropout_split(int device, const float* y, int64_t rows, int64_t cols, float dropout_p,
                           const unsigned long long* seed, uint16_t* hi, uint16_t* lo, const float* scale_slot, atq_stream_t stream_) {
  ATQ_CHECK_ARG(y && hi && rows > 0 && cols > 0, "null pointer or empty shape");
  ATQ_CHECK_ARG((cols % 8) == 0 && aligned16(y) && aligned16(hi) && (lo == nullptr || aligned16(lo)),
                "needs cols % 8 == 0 and 16-byte aligned contiguous tensors");
  ATQ_CHECK_ARG(dropout_p >= 0.f && dropout_p < 1.f && rows * cols < ((int64_t)1 << 33), "dropout_p in [0,1), rows*cols < 2^33");
  ATQ_ENSURE_DEVICE(device);
  ActParams ap;
  dropout_threshold(dropout_p, &ap.drop_thresh, &ap.inv_keep);
  ap.seed = seed;
  const int64_t R = split_colsum_lanes(device, rows, cols);
  const int64_t threads = (cols >> 2) * R;
  const unsigned grid = (unsigned)((threads + kThreads - 1) / kThreads);
  cudaStream_t stream = (cudaStream_t)stream_;
  if (lo) act_split_kernel<false, true><<<grid, kThreads, 0, stream>>>(y, nullptr, rows, cols, hi, lo, nullptr, R, ap, scale_slot);
  else act_split_kernel<false, false><<<grid, kThreads, 0, stream>>>(y, nullptr, rows, cols, hi, lo, nullptr, R, ap, scale_slot);
  ATQ_LAUNCH_CHECK();
  return ATQ_OK;
}

int atq_gelu_dropout_bwd_split_colsum(int device, const float* g, const float* y, int64_t rows, int64_t cols, float dropout_p,
                                      const unsigned long long* seed, uint16_t* hi, uint16_t* lo, float* colsum_out,
                                      void* ws, size_t ws_bytes, const float* scale_slot, atq_stream_t stream_) {
  ATQ_CHECK_ARG(g && y && hi && colsum_out && rows > 0 && cols > 0, "null pointer or empty shape");
  ATQ_CHECK_ARG((cols % 8) == 0 && aligned16(g) && aligned16(y) && aligned16(hi) && (lo == nullptr || aligned16(lo)),
                "needs cols % 8 == 0 and 16-byte aligned contiguous tensors");
  ATQ_CHECK_ARG(dropout_p >= 0.f && dropout_p < 1.f && rows * cols < ((int64_t)1 << 33), "dropout_p in [0,1), rows*cols < 2^33");
  ATQ_ENSURE_DEVICE(device);
  const int64_t R = split_colsum_lanes(device, rows, cols);
  if (ws == nullptr || ws_bytes < (size_t)(R * cols * sizeof(float))) {
    set_error("atq_gelu_dropout_bwd_split_colsum: workspace too small (atq_workspace_bytes_split_colsum)");
    return ATQ_EWORKSPACE;
  }
  ActParams ap;
  dropout_threshold(dropout_p, &ap.drop_thresh, &ap.inv_keep);
  ap.seed = seed;
  const int64_t threads = (cols >> 2) * R;
  const unsigned grid = (unsigned)((threads + kThreads - 1) / kThreads);
  cudaStream_t stream = (cudaStream_t)stream_;
  if (lo) act_split_kernel<true, true><<<grid, kThreads, 0, stream>>>(g, y, rows, cols, hi, lo, (float*)ws, R, ap, scale_slot);
  else act_split_kernel<true, false><<<grid, kThreads, 0, stream>>>(g, y, rows, cols, hi, lo, (float*)ws, R, ap, scale_slot);
  ATQ_LAUNCH_CHECK();
  colsum_stage2_kernel<<<(unsigned)((cols + 31) / 32), 256, 0, stream>>>((const float*)ws, R, cols, colsum_out);
  ATQ_LAUNCH_CHECK();
  return ATQ_OK;
}

size_t atq_workspace_bytes_gated_residual(int64_t n) {
  (void)n;
  return (size_t)4096 * sizeof(float);  // one partial per CTA, grid <= 4096
}

int atq_gated_residual_fwd(int device, const float* src, const float* h, const float* gate, int64_t n, float dropout_p,
                           const unsigned long long* seed, float* out, atq_stream_t stream_) {
  ATQ_CHECK_ARG(src && h && gate && out && n > 0 && (n % 4) == 0, "null pointer or n not a positive multiple of 4");
  ATQ_CHECK_ARG(aligned16(src) && aligned16(h) && aligned16(out), "16-byte aligned contiguous tensors");
  ATQ_CHECK_ARG(dropout_p >= 0.f && dropout_p < 1.f && n < ((int64_t)1 << 33), "dropout_p in [0,1), n < 2^33");
  ATQ_ENSURE_DEVICE(device);
  ActParams ap;
  dropout_threshold(dropout_p, &ap.drop_thresh, &ap.inv_keep);
  ap.seed = seed;
  int grid = stream_grid(device, n >> 2, kThreads * kUnroll, 8);
  gated_residual_kernel<false><<<grid, kThreads, 0, (cudaStream_t)stream_>>>(src, h, gate, n, out, nullptr, ap);
  ATQ_LAUNCH_CHECK();
  return ATQ_OK;
}

int atq_gated_residual_bwd(int device, const float* dout, const float* h, const float* gate, int64_t n, float dropout_p,
                           const unsigned long long* seed, float* dh, float* dgate, void* ws, size_t ws_bytes,
                           atq_stream_t stream_) {
  ATQ_CHECK_ARG(dout && h && gate && dh && dgate && n > 0 && (n % 4) == 0, "null pointer or n not a positive multiple of 4");
  ATQ_CHECK_ARG(aligned16(dout) && aligned16(h) && aligned16(dh), "16-byte aligned contiguous tensors");
  ATQ_CHECK_ARG(dropout_p >= 0.f && dropout_p < 1.f && n < ((int64_t)1 << 33), "dropout_p in [0,1), n < 2^33");
  ATQ_ENSURE_DEVICE(device);
  if (ws == nullptr || ws_bytes < atq_workspace_bytes_gated_residual(n)) {
    set_error("atq_gated_residual_bwd: workspace too small");
    return ATQ_EWORKSPACE;
  }
  ActParams ap;
  dropout_threshold(dropout_p, &ap.drop_thresh, &ap.inv_keep);
  ap.seed = seed;
  int grid = stream_grid(device, n >> 2, kThreads * kUnroll, 8);
  if (grid > 4096) grid = 4096;
  cudaStream_t stream = (cudaStream_t)stream_;
  gated_residual_kernel<true><<<grid, kThreads, 0, stream>>>(dout, h, gate, n, dh, (float*)ws, ap);
  ATQ_LAUNCH_CHECK();
  sum_partials_kernel<<<1, 256, 0, stream>>>((const float*)ws, grid, dgate);
  ATQ_LAUNCH_CHECK();
  return ATQ_OK;
}

int atq_adamw_multi(int device, const void* table, const int* chunk_tensor, const int* chunk_off, int n_chunks, float lr,
                    float beta1, float beta2, float eps, float weight_decay, float* step, atq_stream_t stream_) {
  ATQ_CHECK_ARG(table && chunk_tensor && chunk_off && step && n_chunks > 0, "null pointer or no work");
  ATQ_ENSURE_DEVICE(device);
  cudaStream_t stream = (cudaStream_t)stream_;
  int grid = sm_count(device) * 8;
  if (grid > n_chunks) grid = n_chunks;
  adamw_multi_kernel<<<grid, 256, 0, stream>>>((const AdamwTensor*)table, chunk_tensor, chunk_off, n_chunks, lr, beta1, beta2, eps,
                                               weight_decay, step);
  ATQ_LAUNCH_CHECK();
  adamw_advance_kernel<<<1, 1, 0, stream>>>(step);
  ATQ_LAUNCH_CHECK();
  return ATQ_OK;
}

static inline int layernorm_grid(int device, int64_t rows) {
  int64_t need = (rows + (kThreads / 32) - 1) / (kThreads / 32);
  const int64_t cap = (int64_t)sm_count(device) * 2;
  return (int)(need < cap ? (need < 1 ? 1 : need) : cap);
}

size_t atq_workspace_bytes_layernorm_bwd(int64_t cols) { return (size_t)2 * 2 * 256 * (size_t)cols * sizeof(float); }

int atq_layernorm_fwd(int device, const float* x, const float* gamma, const float* beta, int64_t rows, int64_t cols, float eps, float* y,
                      float* mean_out, float* rstd_out, float* out_scale_slot, atq_stream_t stream_) {
  ATQ_CHECK_ARG(x && gamma && beta && y && mean_out && rstd_out && rows > 0, "null pointer or rows <= 0");
  ATQ_CHECK_ARG(cols >= 4 && cols <= 32 * kLnVec * 4 && (cols % 4) == 0, "needs 4 <= cols <= 1024, cols % 4 == 0");
  ATQ_CHECK_ARG(aligned16(x) && aligned16(y) && aligned16(gamma) && aligned16(beta), "needs 16-byte aligned contiguous tensors");
  ATQ_ENSURE_DEVICE(device);
  cudaStream_t stream = (cudaStream_t)stream_;
  layernorm_fwd_kernel<<<layernorm_grid(device, rows), kThreads, 0, stream>>>(x, gamma, beta, rows, (int)cols, eps, y, mean_out, rstd_out,
                                                                              reinterpret_cast<unsigned int*>(out_scale_slot));
  ATQ_LAUNCH_CHECK();
  return ATQ_OK;
}

int atq_layernorm_bwd(int device, const float* dy, const float* x, const float* gamma, const float* mean, const float* rstd, int64_t rows,
                      int64_t cols, float* dx, float* dgamma, float* dbeta, void* ws, size_t ws_bytes, atq_stream_t stream_) {
  ATQ_CHECK_ARG(dy && x && gamma && mean && rstd && dx && dgamma && dbeta && rows > 0, "null pointer or rows <= 0");
  ATQ_CHECK_ARG(cols >= 4 && cols <= 32 * kLnVec * 4 && (cols % 4) == 0, "needs 4 <= cols <= 1024, cols % 4 == 0");
  ATQ_CHECK_ARG(aligned16(dy) && aligned16(x) && aligned16(dx) && aligned16(gamma), "needs 16-byte aligned contiguous tensors");
  ATQ_ENSURE_DEVICE(device);
  const int grid = layernorm_grid(device, rows);
  if (ws == nullptr || ws_bytes < (size_t)2 * grid * cols * sizeof(float)) {
    set_error("atq_layernorm_bwd: workspace too small (atq_workspace_bytes_layernorm_bwd)");
    return ATQ_EWORKSPACE;
  }
  cudaStream_t stream = (cudaStream_t)stream_;
  float* part_g = reinterpret_cast<float*>(ws);
  float* part_b = part_g + (size_t)grid * cols;
  const size_t smem = (size_t)(kThreads / 32) * cols * sizeof(float);
  layernorm_bwd_kernel<<<grid, kThreads, smem, stream>>>(dy, x, gamma, mean, rstd, rows, (int)cols, dx, part_g, part_b);
  ATQ_LAUNCH_CHECK();
  colsum_stage2_kernel<<<(unsigned)((cols + 31) / 32), 256, 0, stream>>>(part_g, grid, cols, dgamma);
  ATQ_LAUNCH_CHECK();
  colsum_stage2_kernel<<<(unsigned)((cols + 31) / 32), 256, 0, stream>>>(part_b, grid, cols, dbeta);
  ATQ_LAUNCH_CHECK();
  return ATQ_OK;
}

size_t atq_workspace_bytes_colsum(int64_t rows, int64_t cols) {
  const int rpc = colsum_rows_per_cta(rows);
  int64_t parts = (rows + rpc - 1) / rpc;
  return (size_t)(parts * cols * sizeof(float));
}

int atq_colsum_f32(int device, const float* x, int64_t rows, int64_t cols, int64_t ld, float* out, void* ws,
                   size_t ws_bytes, atq_stream_t stream_) {
  ATQ_CHECK_ARG(x && out && rows > 0 && cols > 0 && ld >= cols, "null pointer or bad shape");
  ATQ_ENSURE_DEVICE(device);
  if (ws_bytes < atq_workspace_bytes_colsum(rows, cols) || ws == nullptr) {
    set_error("atq_colsum_f32: workspace too small");
    return ATQ_EWORKSPACE;
  }
  cudaStream_t stream = (cudaStream_t)stream_;
  if (rows <= 64) {  // a single row block: one launch
    colsum_small_kernel<<<(unsigned)((cols + 31) / 32), 256, 0, stream>>>(x, rows, cols, ld, out);
    ATQ_LAUNCH_CHECK();
    return ATQ_OK;
  }
  const int rpc = colsum_rows_per_cta(rows);
  int64_t parts = (rows + rpc - 1) / rpc;
  dim3 g1((unsigned)((cols + 31) / 32), (unsigned)parts);
  ATQ_CHECK_ARG(parts <= 65535, "rows too large for one launch");
  colsum_stage1_kernel<<<g1, 256, 0, stream>>>(x, rows, cols, ld, (float*)ws, rpc);
  ATQ_LAUNCH_CHECK();
  colsum_stage2_kernel<<<(unsigned)((cols + 31) / 32), 256, 0, stream>>>((const float*)ws, parts, cols, out);
  ATQ_LAUNCH_CHECK();
  return ATQ_OK;
}

}  // extern "C"
