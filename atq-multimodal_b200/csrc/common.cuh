// Shared helpers for libatq_sm100 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/atq_sm100.h"

namespace atq {

constexpr int kNumSMsB200 = 148;

// thread-local last-error text (atq_last_error_string)
char* last_error_buf();
void set_error(const char* fmt, ...);

// Makes `device` current on the calling thread for the duration of one entry point (autograd's backward thread
// starts on device 0 from this library's statically linked runtime's point of view) and restores the caller's.
struct DeviceGuard {
  int prev = -1;
  int set(int device);
  ~DeviceGuard();
};
int sm_count(int device);
void note_launch(int n = 1);  // kernel-launch counter reported by atq_kernel_launch_count()

#define ATQ_CHECK_ARG(cond, msg)                                   \
  do {                                                             \
    if (!(cond)) {                                                 \
      atq::set_error("%s: %s", __func__, msg);                     \
      return ATQ_EINVAL;                                           \
    }                                                              \
  } while (0)

#define ATQ_CUDA(call)                                                                   \
  do {                                                                                   \
    cudaError_t e__ = (call);                                                            \
    if (e__ != cudaSuccess) {                                                            \
      atq::set_error("%s: %s failed: %s", __func__, #call, cudaGetErrorString(e__));     \
      return ATQ_ECUDA;                                                                  \
    }                                                                                    \
  } while (0)

#define ATQ_LAUNCH_CHECK()                                                               \
  do {                                                                                   \
    cudaError_t e__ = cudaGetLastError();                                                \
    if (e__ != cudaSuccess) {                                                            \
      atq::set_error("%s: kernel launch failed: %s", __func__, cudaGetErrorString(e__)); \
      return ATQ_ECUDA;                                                                  \
    }                                                                                    \
    atq::note_launch();                                                                  \
  } while (0)

#define ATQ_ENSURE_DEVICE(dev)                  \
  atq::DeviceGuard device_guard__;              \
  do {                                          \
    int r__ = device_guard__.set(dev);          \
    if (r__ != ATQ_OK) return r__;              \
  } while (0)

// grid for a grid-stride streaming kernel: enough CTAs to cover `work_items` once at
// `per_cta` items each, capped at `waves` resident CTAs per SM, rounded to a multiple of the
// SM count when capped.
inline int stream_grid(int device, int64_t work_items, int per_cta, int ctas_per_sm) {
  int64_t need = (work_items + per_cta - 1) / per_cta;
  int64_t cap = (int64_t)sm_count(device) * ctas_per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

__device__ __forceinline__ float4 ldg_stream4(const float* p) {
  // read-once streaming data: keep it out of L1
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// 2-bit code of the reference codec (atq/bit_packing.py:49): value+1 -> {0,1,2}
__device__ __forceinline__ uint32_t tern_code(float w, float thr) {
  // strict compares as atq/quantizers.py:42-43; NaN compares false both ways -> code 1 (zero)
  return (w > thr) ? 2u : ((w < -thr) ? 0u : 1u);
}

__device__ __forceinline__ uint16_t bf16_bits(float f) {
  return __bfloat16_as_ushort(__float2bfloat16_rn(f));
}
__device__ __forceinline__ float bf16_bits_to_float(uint16_t b) {
  return __uint_as_float(((uint32_t)b) << 16);
}
// hi = bf16(x); lo = bf16(x - hi) (0 when hi is not finite)
__device__ __forceinline__ void split_bf16(float x, uint16_t& hi, uint16_t& lo) {
  hi = bf16_bits(x);
  float hf = bf16_bits_to_float(hi);
  float r = x - hf;
  lo = (fabsf(hf) <= 3.3895313892515355e38f) ? bf16_bits(r) : (uint16_t)0;
}

// GEMM operand element format.  bf16: (hi, lo) = (bf16(x), bf16(x - hi)), ~16 significant bits.
// scaled fp16: (hi, lo) = (fp16(s x), fp16(s x - hi)) with ONE power-of-two scale s per tensor chosen so that
// max|s x| lies in [2^14, 2^15): ~22 significant bits for every element within 2^-18 of the tensor's maximum and an
// absolute error below 2^-40 max|x| for the rest; the GEMM epilogue multiplies by 1/s (exact).  Same MMA count as
// the bf16 pair, 2^5 times the precision: this is what lets a whole network's gradients track the fp32 reference.
// Scale slot (4 floats, caller-owned): [0] bits of max|x| (scratch), [1] s, [2] 1/s, [3] ticket (scratch).
struct OperandFmt {
  float scale;
  int f16;
};
__device__ __forceinline__ OperandFmt load_fmt(const float* slot, int force_f16 = 0) {
  OperandFmt f;
  f.f16 = (slot != nullptr) || force_f16;
  f.scale = slot != nullptr ? __ldg(slot + 1) : 1.f;
  return f;
}
__device__ __forceinline__ void split2(const OperandFmt& f, float x, uint16_t& hi, uint16_t& lo) {
  if (f.f16) {
    const float xs = x * f.scale;
    const __half h = __float2half_rn(xs);
    const float hf = __half2float(h);
    hi = __half_as_ushort(h);
    lo = (fabsf(hf) <= 65504.f) ? __half_as_ushort(__float2half_rn(xs - hf)) : (uint16_t)0;
  } else {
    split_bf16(x, hi, lo);
  }
}
// power-of-two scale that maps a bound b >= max|x| into [2^14, 2^15); (1, 1) for b = 0 / inf / nan
__device__ __forceinline__ void pow2_scale_for(float b, float& scale, float& inv) {
  const uint32_t e = (__float_as_uint(b) >> 23) & 0xFFu;
  if (e == 0u || e == 255u) {
    scale = 1.f; inv = 1.f;
    return;
  }
  int se = 268 - (int)e;  // biased exponent of the scale
  se = se < 1 ? 1 : (se > 253 ? 253 : se);
  scale = __uint_as_float((uint32_t)se << 23);
  inv = __uint_as_float((uint32_t)(254 - se) << 23);
}

// Counter-based dropout (shared by the attention kernels and the fused FFN activation): one 32-bit hash per
// PAIR of consecutive elements (2k, 2k+1) of a row; element 2k uses the low 16 bits, 2k+1 the high 16 bits,
// kept iff >= the 16-bit threshold round(p * 65536).  atq/attention.py restates it on the host for the tests.
__device__ __forceinline__ uint32_t drop_row_key(uint32_t seed_lo, uint32_t seed_hi, uint32_t row_id) {
  uint32_t x = (row_id * 0x9E3779B1u) ^ seed_lo;
  x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13;
  return x + seed_hi;
}
__device__ __forceinline__ uint32_t drop_hash_pair(uint32_t row_key, uint32_t pair) {
  uint32_t x = row_key + pair * 0xC2B2AE35u;
  x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
  return x;
}
// 16-bit threshold and 1/(1-p_effective) for a requested rate p (host)
static inline void dropout_threshold(float p, uint32_t* thresh, float* inv_keep) {
  if (p > 0.f) {
    int t = (int)((double)p * 65536.0 + 0.5);
    if (t < 1) t = 1;
    if (t > 65535) t = 65535;
    *thresh = (uint32_t)t;
    *inv_keep = (float)(1.0 / (1.0 - (double)t / 65536.0));
  } else {
    *thresh = 0u;
    *inv_keep = 1.f;
  }
}

}  // namespace atq
