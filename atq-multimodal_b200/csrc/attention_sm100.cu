// Fused attention core for the ternary transformer blocks:
//
//   O = dropout(softmax(scale * Q K^T + key_padding_mask)) V        per (batch, head), head_dim <= 64 (multiple of 8;
//   the tiles are 64 wide: the missing columns of a smaller head are staged as zeros and never stored)
//
// replaces the reference's explicit matmul / masked_fill / softmax / dropout / matmul sequence
// (models/text_encoder.py:117-163, TernaryMultiheadAttention._scaled_dot_product_attention) and its
// autograd backward.  Q, K, V are the fp32 outputs of the ternary q/k/v projections, read in place
// from their [B*L, E] layout (head h = columns 64h..64h+63): no head transposes, no [B,h,L,L] score
// tensors in HBM.
//
// One CTA (128 threads) per (batch, head).  fp32 tiles are converted on the fly to bf16 (hi, lo) pairs
// and stored with the SWIZZLE_128B pattern UMMA descriptors expect (16-byte chunk c of row r at chunk
// c ^ (r & 7)); one thread issues tcgen05.mma (S = Q K^T with terms hi*hi, lo*hi, hi*lo; P V with V hi/lo),
// accumulators live in TMEM, every thread owns one accumulator row (query) for the softmax.
// The same shared-memory image of a [rows x 64] tile serves as a K-major operand (contraction over the
// 64 head dims) and as an MN-major operand (contraction over the rows) - only the descriptor differs.
//
// Dropout uses a counter-based hash of (seed, batch*head*L + query, key) so that the backward kernel
// regenerates the mask; the seed is read from device memory (CUDA-graph safe).
#include "common.cuh"
#include "tc_sm100.cuh"
#include "../../include/atq_sm100.h"

namespace atq {

#ifdef ATQ_ATTN_PROF
__device__ long long g_attn_prof[64];
#define ATTN_PROF(i) do { if (blockIdx.x == 200 && threadIdx.x == 0) g_attn_prof[(i)] = clock64(); } while (0)
#else
#define ATTN_PROF(i) do { } while (0)
#endif

constexpr int kHd = 64;          // head dim
constexpr int kAttThreads = 256; // 8 warps: warps w and w+4 share TMEM lane quarter w & 3 and split the columns
constexpr int kTileBytes = 128 * 128;  // 128 rows x 64 bf16

struct AttnParams {
  const float *q, *k, *v;
  int64_t q_pitch, k_pitch, v_pitch;
  int hd;  // real head dim (<= 64, multiple of 8); tiles are 64 wide, columns >= hd are staged as zeros and never stored
  const uint8_t* key_pad;  // [B, L], non-zero = padded key; nullable
  float* out;
  int64_t out_pitch;
  float* lse;  // [B*H, L] log-sum-exp of the scaled, masked scores
  int B, H, L, NK;  // NK = L rounded up to 32 (forward)
  float scale;
  uint32_t drop_thresh;  // 16-bit threshold: keep iff hash16 >= thresh; 0 = no dropout
  float inv_keep;        // 1 / (1 - thresh / 65536)
  const unsigned long long* seed;  // device scalar, nullable (= 0)
  int terms;  // 3 = hi/lo split (parity), 1 = bf16 only (fast)
  // backward
  const float *o, *dout;
  int64_t o_pitch, do_pitch;
  float *dq, *dk, *dv;
  int64_t dq_pitch, dk_pitch, dv_pitch;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// two floats -> packed bf16x2 (a in the low half), one cvt instruction
__device__ __forceinline__ uint32_t pack2_bf16(float a, float b) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
  return d;
}
// tile loads: plain global loads that bypass L1 (every byte is used once per CTA; the read-only texture path is slower
// for this access pattern).  Not volatile: the compiler keeps all loads of a tile in flight together.
__device__ __forceinline__ float4 ld_tile4(const float4* p) {
  float4 v;
  asm("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float bf16lo_f(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16hi_f(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

// fp32 [rows_valid x 64] (row pitch in floats, 16-byte aligned rows) -> bf16 hi (+ lo) tiles of rows_total
// swizzled 128-byte rows; rows >= rows_valid become zeros.  8 threads per row: thread c loads the float4 at columns
// 4c and 32 + 4c, so every warp-level load covers whole 128-byte row segments (no half-used sectors), and stores the
// two 8-byte halves of the bf16 chunks they fall into.  All global loads of up to PASSES row passes (PASSES x 32 rows)
// are in flight before the first conversion, so a 128-row tile (PASSES = 4) or a 256-row K / V tile (PASSES = 8) costs
// ONE memory round trip.
__device__ __forceinline__ void convert_store_half(const float4& a, int r, int chunk, int sub, bool lo, uint32_t s_hi, uint32_t s_lo) {
  const uint32_t h0 = pack2_bf16(a.x, a.y), h1 = pack2_bf16(a.z, a.w);
  const uint32_t off = (uint32_t)r * 128u + (((uint32_t)chunk ^ ((uint32_t)r & 7u)) << 4) + (uint32_t)sub;
  st_shared_v2(s_hi + off, h0, h1);
  if (lo) {
    const uint32_t l0 = pack2_bf16(a.x - bf16lo_f(h0), a.y - bf16hi_f(h0));
    const uint32_t l1 = pack2_bf16(a.z - bf16lo_f(h1), a.w - bf16hi_f(h1));
    st_shared_v2(s_lo + off, l0, l1);
  }
}
__device__ __forceinline__ void convert_store_row(const float4& a, const float4& b, int r, int c, bool lo, uint32_t s_hi, uint32_t s_lo) {
  convert_store_half(a, r, c >> 1, (c & 1) * 8, lo, s_hi, s_lo);
  convert_store_half(b, r, 4 + (c >> 1), (c & 1) * 8, lo, s_hi, s_lo);
}

template <bool LO, int PASSES = 4>
__device__ __forceinline__ void stage_tile(const float* __restrict__ src, int64_t pitch, int rows_valid, int rows_total,
                                           uint32_t s_hi, uint32_t s_lo, int hd) {
  const int c = threadIdx.x & 7;
  const bool va = 4 * c < hd, vb = 32 + 4 * c < hd;  // head dims < 64: the missing columns are zero padding
  constexpr int kRowsPerPass = kAttThreads / 8;  // 32
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int r0 = threadIdx.x >> 3; r0 < rows_total; r0 += PASSES * kRowsPerPass) {
    float4 a[PASSES], b[PASSES];
#pragma unroll
    for (int i = 0; i < PASSES; ++i) {
      const int r = r0 + i * kRowsPerPass;
      const float4* p = reinterpret_cast<const float4*>(src + (int64_t)r * pitch) + c;
      a[i] = (r < rows_valid && va) ? ld_tile4(p) : zero;
      b[i] = (r < rows_valid && vb) ? ld_tile4(p + 8) : zero;
    }
#ifdef ATQ_ATTN_PROF
    if (PASSES == 8) {  // when did the last load land?
      float chk = 0.f;
#pragma unroll
      for (int i = 0; i < PASSES; ++i) chk += a[i].x + b[i].w;
      if (chk == 1.2345e30f) g_attn_prof[63] = 1;
      if (blockIdx.x == 200 && threadIdx.x == 0) g_attn_prof[32 + (g_attn_prof[62]++ & 7)] = clock64();
    }
#endif
#pragma unroll
    for (int i = 0; i < PASSES; ++i) {
      const int r = r0 + i * kRowsPerPass;
      if (r < rows_total) convert_store_row(a[i], b[i], r, c, LO, s_hi, s_lo);
    }
  }
}

// two 128-row tiles (e.g. Q and dO) with the loads of both in flight together
template <bool LO>
__device__ __forceinline__ void stage_two_tiles(const float* __restrict__ src0, int64_t pitch0, uint32_t s_hi0, uint32_t s_lo0,
                                                const float* __restrict__ src1, int64_t pitch1, uint32_t s_hi1, uint32_t s_lo1,
                                                int rows_valid, int hd) {
  const int c = threadIdx.x & 7, r0 = threadIdx.x >> 3;
  const bool va = 4 * c < hd, vb = 32 + 4 * c < hd;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 a0[4], b0[4], a1[4], b1[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + i * 32;
    const float4* p0 = reinterpret_cast<const float4*>(src0 + (int64_t)r * pitch0) + c;
    const float4* p1 = reinterpret_cast<const float4*>(src1 + (int64_t)r * pitch1) + c;
    a0[i] = (r < rows_valid && va) ? ld_tile4(p0) : zero;
    b0[i] = (r < rows_valid && vb) ? ld_tile4(p0 + 8) : zero;
    a1[i] = (r < rows_valid && va) ? ld_tile4(p1) : zero;
    b1[i] = (r < rows_valid && vb) ? ld_tile4(p1 + 8) : zero;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + i * 32;
    convert_store_row(a0[i], b0[i], r, c, LO, s_hi0, s_lo0);
    convert_store_row(a1[i], b1[i], r, c, LO, s_hi1, s_lo1);
  }
}

// backward: the Q and dO tiles of one query tile travel through registers in two steps, so that the global loads of the
// NEXT (key tile, query tile) iteration are in flight while the tensor cores work on the current one.  With `with_o`,
// delta[r] = sum_d dO[r, d] * O[r, d] over this head comes from the SAME dO registers and an O load with the same
// mapping (the 8 threads of a row reduce by shuffle).
struct QdoRegs {
  float4 a0[4], b0[4], a1[4], b1[4], oa[4], ob[4];
};
__device__ __forceinline__ void load_q_do(QdoRegs& R, const float* __restrict__ qg, int64_t q_pitch, const float* __restrict__ dg,
                                          int64_t do_pitch, const float* __restrict__ og, int64_t o_pitch, bool with_o, int rows_valid,
                                          int hd) {
  const int c = threadIdx.x & 7, r0 = threadIdx.x >> 3;
  const bool va = 4 * c < hd, vb = 32 + 4 * c < hd;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + i * 32;
    const bool ok = r < rows_valid;
    const float4* p0 = reinterpret_cast<const float4*>(qg + (int64_t)r * q_pitch) + c;
    const float4* p1 = reinterpret_cast<const float4*>(dg + (int64_t)r * do_pitch) + c;
    R.a0[i] = (ok && va) ? ld_tile4(p0) : zero;
    R.b0[i] = (ok && vb) ? ld_tile4(p0 + 8) : zero;
    R.a1[i] = (ok && va) ? ld_tile4(p1) : zero;
    R.b1[i] = (ok && vb) ? ld_tile4(p1 + 8) : zero;
    if (with_o) {
      const float4* p2 = reinterpret_cast<const float4*>(og + (int64_t)r * o_pitch) + c;
      R.oa[i] = (ok && va) ? ld_tile4(p2) : zero;
      R.ob[i] = (ok && vb) ? ld_tile4(p2 + 8) : zero;
    }
  }
}
template <bool LO>
__device__ __forceinline__ void store_q_do(const QdoRegs& R, uint32_t s_qh, uint32_t s_ql, uint32_t s_dh, uint32_t s_dl, float* s_delta) {
  const int c = threadIdx.x & 7, r0 = threadIdx.x >> 3;
  if (s_delta != nullptr) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float d = (R.a1[i].x * R.oa[i].x + R.a1[i].y * R.oa[i].y) + (R.a1[i].z * R.oa[i].z + R.a1[i].w * R.oa[i].w) +
                (R.b1[i].x * R.ob[i].x + R.b1[i].y * R.ob[i].y) + (R.b1[i].z * R.ob[i].z + R.b1[i].w * R.ob[i].w);
      d += __shfl_xor_sync(0xffffffffu, d, 1);
      d += __shfl_xor_sync(0xffffffffu, d, 2);
      d += __shfl_xor_sync(0xffffffffu, d, 4);
      if (c == 0) s_delta[r0 + i * 32] = d;
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + i * 32;
    convert_store_row(R.a0[i], R.b0[i], r, c, LO, s_qh, s_ql);
    convert_store_row(R.a1[i], R.b1[i], r, c, LO, s_dh, s_dl);
  }
}

// one accumulator row (this thread's TMEM lane), 32 consecutive columns
__device__ __forceinline__ void ld_row32(uint32_t taddr, float (&f)[32]) {
  uint32_t r[32];
  tmem_ld_32x32b_x32(taddr, r);
  tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(r[j]);
}

// this thread's accumulator row: write 32 consecutive TMEM columns back (fp32)
__device__ __forceinline__ void st_row32(uint32_t taddr, const float (&f)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(__float_as_uint(f[0])), "r"(__float_as_uint(f[1])), "r"(__float_as_uint(f[2])), "r"(__float_as_uint(f[3])),
      "r"(__float_as_uint(f[4])), "r"(__float_as_uint(f[5])), "r"(__float_as_uint(f[6])), "r"(__float_as_uint(f[7])),
      "r"(__float_as_uint(f[8])), "r"(__float_as_uint(f[9])), "r"(__float_as_uint(f[10])), "r"(__float_as_uint(f[11])),
      "r"(__float_as_uint(f[12])), "r"(__float_as_uint(f[13])), "r"(__float_as_uint(f[14])), "r"(__float_as_uint(f[15])),
      "r"(__float_as_uint(f[16])), "r"(__float_as_uint(f[17])), "r"(__float_as_uint(f[18])), "r"(__float_as_uint(f[19])),
      "r"(__float_as_uint(f[20])), "r"(__float_as_uint(f[21])), "r"(__float_as_uint(f[22])), "r"(__float_as_uint(f[23])),
      "r"(__float_as_uint(f[24])), "r"(__float_as_uint(f[25])), "r"(__float_as_uint(f[26])), "r"(__float_as_uint(f[27])),
      "r"(__float_as_uint(f[28])), "r"(__float_as_uint(f[29])), "r"(__float_as_uint(f[30])), "r"(__float_as_uint(f[31]))
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 32 values of row `row` (keys 32*ch .. 32*ch+31) -> bf16 in a K-major sequence of [128 x 64-key] atoms.
// KEEP_LO: v[] is replaced by the rounding residual v - bf16(v) (the lo operand of a later pass).
template <bool KEEP_LO>
__device__ __forceinline__ void store_row32_bf16(uint32_t base, int row, int ch, float (&v)[32]) {
  const uint32_t atom = base + (uint32_t)(ch >> 1) * (uint32_t)kTileBytes + (uint32_t)row * 128u;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int e = 8 * g + 2 * j;
      w[j] = pack2_bf16(v[e], v[e + 1]);
      if (KEEP_LO) {
        v[e] -= bf16lo_f(w[j]);
        v[e + 1] -= bf16hi_f(w[j]);
      }
    }
    const uint32_t chunk = (uint32_t)((ch & 1) * 4 + g);
    st_shared_v4(atom + ((chunk ^ ((uint32_t)row & 7u)) << 4), make_uint4(w[0], w[1], w[2], w[3]));
  }
}

// One [128 x 64] fp32 accumulator tile (thread = row, columns [32 half, +32) in o[], scaled by `mul`) -> global rows
// gbase + r * pitch (r < rows_valid, columns < hd) through a 32 KB swizzled shared-memory tile, so that every warp-level
// global store covers two whole row segments instead of 32 rows x 16 bytes.  Contains a block barrier.
__device__ __forceinline__ void store_tile_coalesced(uint32_t s_stage, const float (&o)[32], float mul, int row, int half,
                                                     float* __restrict__ gbase, int64_t pitch, int rows_valid, int hd) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint32_t slot = (uint32_t)(half * 8 + j) ^ ((uint32_t)row & 7u);
    st_shared_f4(s_stage + (uint32_t)row * 256u + (slot << 4), o[4 * j] * mul, o[4 * j + 1] * mul, o[4 * j + 2] * mul, o[4 * j + 3] * mul);
  }
  __syncthreads();
  const int l = threadIdx.x & 31, w = threadIdx.x >> 5, jj = l & 15;
  if (4 * jj < hd) {
#pragma unroll
    for (int pass = 0; pass < 128 / (2 * (kAttThreads / 32)); ++pass) {
      const int r = pass * 2 * (kAttThreads / 32) + w * 2 + (l >> 4);
      if (r < rows_valid) {
        const float4 v = ld_shared_f4(s_stage + (uint32_t)r * 256u + (((uint32_t)jj ^ ((uint32_t)r & 7u)) << 4));
        *reinterpret_cast<float4*>(gbase + (int64_t)r * pitch + 4 * jj) = v;
      }
    }
  }
}

// key validity bits (key < L and not padded) for up to 256 keys -> s_valid[8]
__device__ __forceinline__ void build_valid_bits(const AttnParams& p, int b, uint32_t* s_valid) {
  const int k = (int)threadIdx.x;  // kAttThreads == 256
  bool ok = k < p.L;
  if (ok && p.key_pad != nullptr) ok = p.key_pad[(int64_t)b * p.L + k] == 0;
  const uint32_t bits = __ballot_sync(0xffffffffu, ok);
  if ((threadIdx.x & 31) == 0) s_valid[k >> 5] = bits;
}

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
// shared memory: [K hi | K lo | V hi | V lo] (NK rows x 128 B each), [Q hi | Q lo] (128 rows), P (NK/64 atoms)
__global__ void __launch_bounds__(kAttThreads, 1) attention_fwd_kernel(const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int NK = p.NK;
  const bool lo = p.terms == 3;
  const uint32_t kv_bytes = (uint32_t)NK * 128u;
  const uint32_t s_kh = base, s_kl = s_kh + kv_bytes;
  const uint32_t s_vh = s_kl + (lo ? kv_bytes : 0u), s_vl = s_vh + kv_bytes;
  const uint32_t s_qh = s_vl + (lo ? kv_bytes : 0u), s_ql = s_qh + kTileBytes;
  const uint32_t s_p = s_ql + (lo ? (uint32_t)kTileBytes : 0u);
  __shared__ __align__(8) unsigned long long s_bar;
  __shared__ uint32_t s_tmem;
  __shared__ uint32_t s_valid[8];
  __shared__ float s_red[2][128];  // row max / row sum partials of the two column halves
  const uint32_t bar = smem_u32(&s_bar);
  const int warp = threadIdx.x >> 5;
  const int half = warp >> 2;               // column half this warp works on
  const int row = threadIdx.x & 127;        // accumulator row = TMEM lane
  const int bh = blockIdx.x, b = bh / p.H, h = bh % p.H;
  const int L = p.L;

  ATTN_PROF(0);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(smem_u32(&s_tmem));
  // the Q tile travels through registers: the loads of the next query tile are issued while the tensor cores and the
  // softmax work on the current one
  float4 qa[4], qb[4];
  auto issue_q_loads = [&](int q0n) {
    const int c = threadIdx.x & 7, r0 = threadIdx.x >> 3;
    const bool va = 4 * c < p.hd, vb = 32 + 4 * c < p.hd;
    const int valid = (L - q0n) < 128 ? (L - q0n) : 128;
    const float* qg = p.q + ((int64_t)b * L + q0n) * p.q_pitch + h * p.hd;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = r0 + i * 32;
      const float4* src = reinterpret_cast<const float4*>(qg + (int64_t)r * p.q_pitch) + c;
      qa[i] = (r < valid && va) ? ld_tile4(src) : make_float4(0.f, 0.f, 0.f, 0.f);
      qb[i] = (r < valid && vb) ? ld_tile4(src + 8) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  issue_q_loads(0);
  build_valid_bits(p, b, s_valid);
  const float* kg = p.k + (int64_t)b * L * p.k_pitch + h * p.hd;
  const float* vg = p.v + (int64_t)b * L * p.v_pitch + h * p.hd;
  ATTN_PROF(20);
  if (lo) {
    stage_tile<true, 8>(kg, p.k_pitch, L, NK, s_kh, s_kl, p.hd);
    ATTN_PROF(21);
    stage_tile<true, 8>(vg, p.v_pitch, L, NK, s_vh, s_vl, p.hd);
    ATTN_PROF(22);
  } else {
    stage_tile<false, 8>(kg, p.k_pitch, L, NK, s_kh, 0, p.hd);
    stage_tile<false, 8>(vg, p.v_pitch, L, NK, s_vh, 0, p.hd);
  }
#ifdef ATQ_ATTN_PROF
  if (blockIdx.x == 200 && (threadIdx.x & 31) == 0) g_attn_prof[24 + warp] = clock64();
#endif
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  ATTN_PROF(1);
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(&s_tmem);
  const uint32_t t_s = tmem, t_o = tmem + 256u;
  const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
  const unsigned long long seed = p.seed != nullptr ? *p.seed : 0ull;
  const float c2 = p.scale * 1.4426950408889634f;
  const float log2_inv_keep = p.drop_thresh != 0u ? __log2f(p.inv_keep) : 0.f;
  const float keep = p.drop_thresh != 0u ? 1.f / p.inv_keep : 1.f;
  const uint32_t thr_hi = p.drop_thresh << 16;
  const int nch = NK / 32;
  const int ch_lo = half ? (nch + 1) / 2 : 0, ch_hi = half ? nch : (nch + 1) / 2;
  uint32_t phase = 0;

  for (int q0 = 0; q0 < L; q0 += 128) {
    const int q_valid = (L - q0) < 128 ? (L - q0) : 128;
#pragma unroll
    for (int i = 0; i < 4; ++i) convert_store_row(qa[i], qb[i], (threadIdx.x >> 3) + i * 32, threadIdx.x & 7, lo, s_qh, s_ql);
    fence_proxy_async();
    tcgen05_fence_before();
    __syncthreads();
    ATTN_PROF(2 + (q0 >> 7) * 8);
    if (threadIdx.x == 0) {
      tcgen05_fence_after();
      const uint32_t idesc = make_idesc_rt(NK, false, false);
      uint32_t acc = 0;
      for (int term = 0; term < p.terms; ++term) {
        const uint32_t sa = term == 1 ? s_ql : s_qh, sb = term == 2 ? s_kl : s_kh;
        const uint64_t da = make_smem_desc_kmajor_sw128(sa), db = make_smem_desc_kmajor_sw128(sb);
#pragma unroll
        for (int j = 0; j < kHd / UMMA_K; ++j) {
          umma_bf16(t_s, da + (uint64_t)(2 * j), db + (uint64_t)(2 * j), idesc, acc);
          acc = 1;
        }
      }
      umma_commit(bar);
    }
    if (q0 + 128 < L) issue_q_loads(q0 + 128);
    mbar_wait(bar, phase);
    phase ^= 1u;
    tcgen05_fence_after();
    ATTN_PROF(3 + (q0 >> 7) * 8);

    // ---- softmax: query q0 + row; this warp covers key chunks [ch_lo, ch_hi) ----
    const uint32_t t_row = t_s + lane_off;
    float m = -INFINITY;
    for (int ch = ch_lo; ch < ch_hi; ++ch) {
      float s[32];
      ld_row32(t_row + (uint32_t)(ch * 32), s);
      const uint32_t vb = s_valid[ch];
      if (vb == 0xffffffffu) {
#pragma unroll
        for (int j = 0; j < 32; ++j) m = fmaxf(m, s[j]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if ((vb >> j) & 1u) m = fmaxf(m, s[j]);
      }
    }
    s_red[half][row] = m;
    __syncthreads();
    m = fmaxf(s_red[0][row], s_red[1][row]);
    if (m == -INFINITY) m = 0.f;  // every key masked: probabilities are all zero below
    __syncthreads();              // s_red is reused for the row sums
    ATTN_PROF(4 + (q0 >> 7) * 8);
    // dropout keeps p / (1 - rate): the factor rides in the exponent, and the row sum (taken before the mask) is scaled back
    const float mc = m * c2 - log2_inv_keep;
    const uint32_t row_key = drop_row_key((uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)(bh * L + q0 + row));
    float sum = 0.f;
    for (int ch = ch_lo; ch < ch_hi; ++ch) {
      float s[32];
      ld_row32(t_row + (uint32_t)(ch * 32), s);
      const uint32_t vb = s_valid[ch];
      if (vb == 0xffffffffu) {  // warp-uniform: every key of this chunk is valid
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          s[j] = ex2_approx(fmaf(s[j], c2, -mc));
          sum += s[j];
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float pj = ((vb >> j) & 1u) ? ex2_approx(fmaf(s[j], c2, -mc)) : 0.f;
          sum += pj;
          s[j] = pj;
        }
      }
      if (p.drop_thresh != 0u) {
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const uint32_t hsh = drop_hash_pair(row_key, (uint32_t)(ch * 16 + (j >> 1)));
          s[j] = (hsh << 16) >= thr_hi ? s[j] : 0.f;      // low 16 bits >= threshold
          s[j + 1] = hsh >= thr_hi ? s[j + 1] : 0.f;      // high 16 bits >= threshold
        }
      }
      if (lo) {
        store_row32_bf16<true>(s_p, row, ch, s);        // P_hi to shared memory, s[] <- rounding residual
        st_row32(t_row + (uint32_t)(ch * 32), s);       // residual stays in TMEM for the lo pass
      } else {
        store_row32_bf16<false>(s_p, row, ch, s);
      }
    }
    if (lo) tmem_st_wait();
    s_red[half][row] = sum;
    fence_proxy_async();
    tcgen05_fence_before();
    __syncthreads();
    ATTN_PROF(5 + (q0 >> 7) * 8);
    if (threadIdx.x == 0) {
      tcgen05_fence_after();
      const uint32_t idesc = make_idesc_rt(kHd, false, true);
      uint32_t acc = 0;
      const int vterms = lo ? 2 : 1;
      for (int term = 0; term < vterms; ++term) {  // P_hi V_hi, P_hi V_lo
        const uint32_t sb = term == 1 ? s_vl : s_vh;
#pragma unroll 2
        for (int j = 0; j < NK / UMMA_K; ++j) {
          const uint64_t da = make_smem_desc_kmajor_sw128(s_p + (uint32_t)(j >> 2) * (uint32_t)kTileBytes + (uint32_t)(j & 3) * 32u);
          const uint64_t db = make_smem_desc_mnmajor_sw128(sb + (uint32_t)j * 2048u, 8192);
          umma_bf16(t_o, da, db, idesc, acc);
          acc = 1;
        }
      }
      umma_commit(bar);
    }
    sum = (s_red[0][row] + s_red[1][row]) * keep;
    mbar_wait(bar, phase);
    phase ^= 1u;
    tcgen05_fence_after();
    ATTN_PROF(6 + (q0 >> 7) * 8);
    if (lo) {
      // second pass through the same P buffer: P_lo (a single bf16 P leaves ~3e-3 absolute error in P V)
      for (int ch = ch_lo; ch < ch_hi; ++ch) {
        float s[32];
        ld_row32(t_row + (uint32_t)(ch * 32), s);
        store_row32_bf16<false>(s_p, row, ch, s);
      }
      fence_proxy_async();
      tcgen05_fence_before();
      __syncthreads();
      ATTN_PROF(7 + (q0 >> 7) * 8);
      if (threadIdx.x == 0) {
        tcgen05_fence_after();
        const uint32_t idesc = make_idesc_rt(kHd, false, true);
        for (int j = 0; j < NK / UMMA_K; ++j) {  // + P_lo V_hi
          const uint64_t da = make_smem_desc_kmajor_sw128(s_p + (uint32_t)(j >> 2) * (uint32_t)kTileBytes + (uint32_t)(j & 3) * 32u);
          const uint64_t db = make_smem_desc_mnmajor_sw128(s_vh + (uint32_t)j * 2048u, 8192);
          umma_bf16(t_o, da, db, idesc, 1u);
        }
        umma_commit(bar);
      }
      mbar_wait(bar, phase);
      phase ^= 1u;
      tcgen05_fence_after();
      ATTN_PROF(8 + (q0 >> 7) * 8);
    }

    // ---- epilogue: normalise and store; this warp writes output columns [32 half, 32 half + 32) ----
    const float inv = sum > 0.f ? 1.f / sum : 0.f;
    const int q = q0 + row;
    {
      // staging tile: the Q tile (parity: hi + lo; fast: hi + the first P atom), dead once the last P V MMA has completed
      float o[32];
      ld_row32(t_o + lane_off + (uint32_t)(half * 32), o);
      store_tile_coalesced(s_qh, o, inv, row, half, p.out + ((int64_t)b * L + q0) * p.out_pitch + h * p.hd, p.out_pitch, q_valid, p.hd);
    }
    if (half == 0 && q < L && p.lse != nullptr) p.lse[(int64_t)bh * L + q] = sum > 0.f ? m * p.scale + logf(sum) : -INFINITY;
    tcgen05_fence_before();
    __syncthreads();  // TMEM, s_red and the Q / P tiles are reused by the next query tile
    tcgen05_fence_after();
    ATTN_PROF(9 + (q0 >> 7) * 8);
  }
  if (warp == 0) tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------
// shared memory (1024-aligned 16 KB tiles): K hi, K lo, V hi, V lo, Q hi, Q lo, dO hi, dO lo, then P, dS hi and
// dS lo (two 64-key atoms each; dS carries a lo part because dQ / dK are sums of dS-weighted rows and a single
// bf16 dS leaves ~3e-3 absolute error, above the 1e-3 tolerance).
// TMEM columns: S [0,128) dP [128,256) dQ tile 0 [256,320) dQ tile 1 [320,384) dK [384,448) dV [448,512).
__global__ void __launch_bounds__(kAttThreads, 1) attention_bwd_kernel(const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const bool lo = p.terms == 3;
  const uint32_t T = (uint32_t)kTileBytes;
  const uint32_t s_kh = base, s_kl = base + T, s_vh = base + 2 * T, s_vl = base + 3 * T;
  const uint32_t s_qh = base + 4 * T, s_ql = base + 5 * T, s_dh = base + 6 * T, s_dl = base + 7 * T;
  const uint32_t s_p = base + 8 * T, s_ds = base + 10 * T, s_dsl = base + 12 * T;
  __shared__ __align__(8) unsigned long long s_bar;
  __shared__ uint32_t s_tmem;
  __shared__ uint32_t s_valid[8];
  __shared__ float s_delta[2][128];  // delta of both query tiles (written on the first key tile)
  const uint32_t bar = smem_u32(&s_bar);
  const int warp = threadIdx.x >> 5;
  const int half = warp >> 2;
  const int row = threadIdx.x & 127;
  const int bh = blockIdx.x, b = bh / p.H, h = bh % p.H;
  const int L = p.L;

  ATTN_PROF(0);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(smem_u32(&s_tmem));
  build_valid_bits(p, b, s_valid);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(&s_tmem);
  // dQ of both query tiles stays in TMEM across the key tiles (no read-modify-write of dq in global memory)
  const uint32_t t_s = tmem, t_dp = tmem + 128u, t_dq0 = tmem + 256u, t_dk = tmem + 384u, t_dv = tmem + 448u;
  const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
  const unsigned long long seed = p.seed != nullptr ? *p.seed : 0ull;
  const float c2 = p.scale * 1.4426950408889634f;
  const float log2_inv_keep = p.drop_thresh != 0u ? __log2f(p.inv_keep) : 0.f;
  const float keep = p.drop_thresh != 0u ? 1.f / p.inv_keep : 1.f;
  const uint32_t thr_hi = p.drop_thresh << 16;
  uint32_t phase = 0;

  QdoRegs R;
  auto issue_q_do_loads = [&](int k0n, int q0n) {
    const int64_t r0g = (int64_t)b * L + q0n;
    load_q_do(R, p.q + r0g * p.q_pitch + h * p.hd, p.q_pitch, p.dout + r0g * p.do_pitch + h * p.hd, p.do_pitch,
              p.o + r0g * p.o_pitch + h * p.hd, p.o_pitch, k0n == 0, (L - q0n) < 128 ? (L - q0n) : 128, p.hd);
  };
  issue_q_do_loads(0, 0);
  for (int k0 = 0; k0 < L; k0 += 128) {
    const int k_valid = (L - k0) < 128 ? (L - k0) : 128;
    const float* kg = p.k + ((int64_t)b * L + k0) * p.k_pitch + h * p.hd;
    const float* vg = p.v + ((int64_t)b * L + k0) * p.v_pitch + h * p.hd;
    if (lo) stage_two_tiles<true>(kg, p.k_pitch, s_kh, s_kl, vg, p.v_pitch, s_vh, s_vl, k_valid, p.hd);
    else stage_two_tiles<false>(kg, p.k_pitch, s_kh, 0, vg, p.v_pitch, s_vh, 0, k_valid, p.hd);
    for (int q0 = 0; q0 < L; q0 += 128) {
      // the Q / dO (/ O) registers of this iteration were loaded during the previous iteration's MMA wait.
      // delta = sum_d dO[q,d] * O[q,d] over this head (row-sum of P .* dP) is computed on the first key tile and kept in
      // shared memory for the later ones
      const int q = q0 + row;
      const bool q_ok = q < L;
      const float lse_q = q_ok ? __ldg(p.lse + (int64_t)bh * L + q) : 0.f;
      {
        float* sd = k0 == 0 ? s_delta[q0 >> 7] : nullptr;
        if (lo) store_q_do<true>(R, s_qh, s_ql, s_dh, s_dl, sd);
        else store_q_do<false>(R, s_qh, 0, s_dh, 0, sd);
      }
      fence_proxy_async();
      tcgen05_fence_before();
      __syncthreads();
      ATTN_PROF(1 + ((k0 >> 7) * 2 + (q0 >> 7)) * 8);
      if (threadIdx.x == 0) {
        tcgen05_fence_after();
        const uint32_t idesc = make_idesc_rt(128, false, false);
        // S = Q K^T, dP = dO V^T
        for (int which = 0; which < 2; ++which) {
          const uint32_t ah = which ? s_dh : s_qh, al = which ? s_dl : s_ql, bh_ = which ? s_vh : s_kh, bl = which ? s_vl : s_kl;
          const uint32_t td = which ? t_dp : t_s;
          uint32_t acc = 0;
          for (int term = 0; term < p.terms; ++term) {
            const uint64_t da = make_smem_desc_kmajor_sw128(term == 1 ? al : ah), db = make_smem_desc_kmajor_sw128(term == 2 ? bl : bh_);
#pragma unroll
            for (int j = 0; j < kHd / UMMA_K; ++j) {
              umma_bf16(td, da + (uint64_t)(2 * j), db + (uint64_t)(2 * j), idesc, acc);
              acc = 1;
            }
          }
        }
        umma_commit(bar);
      }
      const float delta_keep = s_delta[q0 >> 7][row] * keep;
      const float lse2 = lse_q * 1.4426950408889634f - log2_inv_keep;
      const uint32_t row_key = drop_row_key((uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)(bh * L + q));
      mbar_wait(bar, phase);
      phase ^= 1u;
      tcgen05_fence_after();
      ATTN_PROF(2 + ((k0 >> 7) * 2 + (q0 >> 7)) * 8);

      // ---- query q0 + row, key chunks [2 half, 2 half + 2): P (after dropout) and dS ----
#pragma unroll 1
      for (int ch = 2 * half; ch < 2 * half + 2; ++ch) {
        float s[32], dp[32];
        ld_row32(t_s + lane_off + (uint32_t)(ch * 32), s);
        ld_row32(t_dp + lane_off + (uint32_t)(ch * 32), dp);
        const uint32_t vb = q_ok ? s_valid[(k0 >> 5) + ch] : 0u;
        // s <- p / (1 - rate) (the dropout factor rides in the exponent), dp <- p (dP_kept - delta) / scale-free dS: the
        // softmax scale is applied to the dQ / dK tiles when they are stored
        if (vb == 0xffffffffu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) s[j] = ex2_approx(fmaf(s[j], c2, -lse2));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) s[j] = ((vb >> j) & 1u) ? ex2_approx(fmaf(s[j], c2, -lse2)) : 0.f;
        }
        if (p.drop_thresh != 0u) {
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const uint32_t hsh = drop_hash_pair(row_key, (uint32_t)((k0 >> 1) + ch * 16 + (j >> 1)));
            const bool k0b = (hsh << 16) >= thr_hi, k1b = hsh >= thr_hi;
            dp[j] = s[j] * ((k0b ? dp[j] : 0.f) - delta_keep);         // dS uses the un-dropped probability
            dp[j + 1] = s[j + 1] * ((k1b ? dp[j + 1] : 0.f) - delta_keep);
            s[j] = k0b ? s[j] : 0.f;                                    // dropped probabilities (dV operand)
            s[j + 1] = k1b ? s[j + 1] : 0.f;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) dp[j] = s[j] * (dp[j] - delta_keep);
        }
        if (lo) {
          store_row32_bf16<true>(s_p, row, ch, s);
          st_row32(t_s + lane_off + (uint32_t)(ch * 32), s);  // P residual stays in TMEM for the lo pass
          store_row32_bf16<true>(s_ds, row, ch, dp);
          store_row32_bf16<false>(s_dsl, row, ch, dp);
        } else {
          store_row32_bf16<false>(s_p, row, ch, s);
          store_row32_bf16<false>(s_ds, row, ch, dp);
        }
      }
      if (lo) tmem_st_wait();
      fence_proxy_async();
      tcgen05_fence_before();
      __syncthreads();
      ATTN_PROF(3 + ((k0 >> 7) * 2 + (q0 >> 7)) * 8);
      if (threadIdx.x == 0) {
        tcgen05_fence_after();
        const uint32_t idesc_q = make_idesc_rt(kHd, false, true);  // A K-major (dS), B MN-major (K tile)
        const uint32_t idesc_t = make_idesc_rt(kHd, true, true);   // A MN-major (dS^T / P^T), B MN-major (Q / dO)
        const int nt = lo ? 2 : 1;
        // dQ(tile) += dS K      (contraction over the 128 keys); terms hi*hi, lo*hi, hi*lo
        const uint32_t t_dq = t_dq0 + (uint32_t)(q0 >> 7) * 64u;
        uint32_t acc = k0 > 0 ? 1u : 0u;
        for (int term = 0; term < p.terms; ++term) {
          const uint32_t sa = term == 1 ? s_dsl : s_ds, sb = term == 2 ? s_kl : s_kh;
#pragma unroll
          for (int j = 0; j < 128 / UMMA_K; ++j) {
            const uint64_t da = make_smem_desc_kmajor_sw128(sa + (uint32_t)(j >> 2) * T + (uint32_t)(j & 3) * 32u);
            const uint64_t db = make_smem_desc_mnmajor_sw128(sb + (uint32_t)j * 2048u, 8192);
            umma_bf16(t_dq, da, db, idesc_q, acc);
            acc = 1;
          }
        }
        // dK += dS^T Q (hi*hi, lo*hi, hi*lo), dV += P_hi^T dO (dO hi/lo; the P_lo term follows below)
        for (int which = 0; which < 2; ++which) {
          const uint32_t td = which ? t_dv : t_dk;
          uint32_t acc2 = q0 > 0 ? 1u : 0u;
          const int nterm = which ? nt : p.terms;
          for (int term = 0; term < nterm; ++term) {
            const uint32_t sa = which ? s_p : (term == 1 ? s_dsl : s_ds);
            const uint32_t sb = which ? (term ? s_dl : s_dh) : (term == 2 ? s_ql : s_qh);
  #pragma unroll
          for (int j = 0; j < 128 / UMMA_K; ++j) {
              const uint64_t da = make_smem_desc_mnmajor_sw128(sa + (uint32_t)j * 2048u, T);
              const uint64_t db = make_smem_desc_mnmajor_sw128(sb + (uint32_t)j * 2048u, 8192);
              umma_bf16(td, da, db, idesc_t, acc2);
              acc2 = 1;
            }
          }
        }
        umma_commit(bar);
      }
      // the tensor cores are busy for a few thousand cycles: put the next iteration's tile loads in flight now
      if (q0 + 128 < L) issue_q_do_loads(k0, q0 + 128);
      else if (k0 + 128 < L) issue_q_do_loads(k0 + 128, 0);
      mbar_wait(bar, phase);
      phase ^= 1u;
      tcgen05_fence_after();
      ATTN_PROF(4 + ((k0 >> 7) * 2 + (q0 >> 7)) * 8);
      ATTN_PROF(5 + ((k0 >> 7) * 2 + (q0 >> 7)) * 8);
      if (lo) {
        // dV += P_lo^T dO_hi through the same P buffer (all MMAs that read P_hi have completed)
#pragma unroll 1
        for (int ch = 2 * half; ch < 2 * half + 2; ++ch) {
          float s[32];
          ld_row32(t_s + lane_off + (uint32_t)(ch * 32), s);
          store_row32_bf16<false>(s_p, row, ch, s);
        }
        fence_proxy_async();
        tcgen05_fence_before();
        __syncthreads();
        ATTN_PROF(6 + ((k0 >> 7) * 2 + (q0 >> 7)) * 8);
        if (threadIdx.x == 0) {
          tcgen05_fence_after();
          const uint32_t idesc_t = make_idesc_rt(kHd, true, true);
#pragma unroll
          for (int j = 0; j < 128 / UMMA_K; ++j) {
            const uint64_t da = make_smem_desc_mnmajor_sw128(s_p + (uint32_t)j * 2048u, T);
            const uint64_t db = make_smem_desc_mnmajor_sw128(s_dh + (uint32_t)j * 2048u, 8192);
            umma_bf16(t_dv, da, db, idesc_t, 1u);
          }
          umma_commit(bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1u;
        tcgen05_fence_after();
        ATTN_PROF(7 + ((k0 >> 7) * 2 + (q0 >> 7)) * 8);
      }
      tcgen05_fence_before();
      __syncthreads();  // Q / dO / P / dS tiles and the S / dP columns are reused
      tcgen05_fence_after();
    }
    // ---- dK, dV rows of this key tile (thread = key, columns [32 half, +32)); staging tiles: the Q and dO tiles, dead
    // until the next iteration's registers are converted into them ----
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      float g[32];
      ld_row32((which ? t_dv : t_dk) + lane_off + (uint32_t)(half * 32), g);
      store_tile_coalesced(which ? s_dh : s_qh, g, which ? 1.f : p.scale, row, half,
                           (which ? p.dv : p.dk) + ((int64_t)b * L + k0) * (which ? p.dv_pitch : p.dk_pitch) + h * p.hd,
                           which ? p.dv_pitch : p.dk_pitch, k_valid, p.hd);
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    ATTN_PROF(40 + (k0 >> 7));
  }
  // ---- dQ rows (thread = query, columns [32 half, +32)) ----
  for (int q0 = 0; q0 < L; q0 += 128) {
    float g[32];
    ld_row32(t_dq0 + (uint32_t)(q0 >> 7) * 64u + lane_off + (uint32_t)(half * 32), g);
    store_tile_coalesced(q0 ? s_dh : s_qh, g, p.scale, row, half, p.dq + ((int64_t)b * L + q0) * p.dq_pitch + h * p.hd, p.dq_pitch,
                         (L - q0) < 128 ? (L - q0) : 128, p.hd);
  }
  ATTN_PROF(42);
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

static int check_attn_ptr(const void* ptr, int64_t pitch, const char* name) {
  if (ptr == nullptr || (reinterpret_cast<uintptr_t>(ptr) & 15u) != 0 || (pitch & 3) != 0) {
    set_error("atq_attention: %s must be non-null, 16-byte aligned, with pitch %% 4 == 0", name);
    return ATQ_EINVAL;
  }
  return ATQ_OK;
}

static int fill_common(AttnParams& p, int B, int H, int L, int head_dim, const float* q, int64_t q_pitch, const float* k, int64_t k_pitch,
                       const float* v, int64_t v_pitch, const uint8_t* key_padding, float scale, float dropout_p,
                       const unsigned long long* seed, int terms) {
  if (B <= 0 || H <= 0 || L <= 0 || L > 256) {
    set_error("atq_attention: needs B, H > 0 and 1 <= L <= 256 (got B=%d H=%d L=%d)", B, H, L);
    return ATQ_EINVAL;
  }
  if (terms != 1 && terms != 3) {
    set_error("atq_attention: terms must be 1 (bf16) or 3 (hi/lo split)");
    return ATQ_EINVAL;
  }
  if (!(dropout_p >= 0.f && dropout_p < 1.f)) {
    set_error("atq_attention: dropout_p must be in [0, 1)");
    return ATQ_EINVAL;
  }
  if (head_dim < 8 || head_dim > kHd || (head_dim % 8) != 0) {
    set_error("atq_attention: head_dim must be a multiple of 8 in [8, 64] (got %d)", head_dim);
    return ATQ_EINVAL;
  }
  const int64_t width = (int64_t)H * head_dim;
  if (q_pitch < width || k_pitch < width || v_pitch < width) {
    set_error("atq_attention: row pitch smaller than H * head_dim");
    return ATQ_EINVAL;
  }
  int r;
  if ((r = check_attn_ptr(q, q_pitch, "q")) != ATQ_OK) return r;
  if ((r = check_attn_ptr(k, k_pitch, "k")) != ATQ_OK) return r;
  if ((r = check_attn_ptr(v, v_pitch, "v")) != ATQ_OK) return r;
  memset(&p, 0, sizeof(p));
  p.q = q; p.k = k; p.v = v;
  p.q_pitch = q_pitch; p.k_pitch = k_pitch; p.v_pitch = v_pitch;
  p.key_pad = key_padding;
  p.B = B; p.H = H; p.L = L;
  p.hd = head_dim;
  p.scale = scale;
  p.terms = terms;
  p.seed = seed;
  dropout_threshold(dropout_p, &p.drop_thresh, &p.inv_keep);  // 16-bit threshold; effective rate thresh / 65536
  return ATQ_OK;
}

}  // namespace atq

using namespace atq;

extern "C" {

int atq_attention_fwd(int device, int B, int H, int L, int head_dim, const float* q, int64_t q_pitch, const float* k, int64_t k_pitch,
                      const float* v, int64_t v_pitch, const uint8_t* key_padding, float scale, float dropout_p,
                      const unsigned long long* seed, int terms, float* out, int64_t out_pitch, float* lse,
                      atq_stream_t stream_) {
  AttnParams p;
  int r;
  if ((r = fill_common(p, B, H, L, head_dim, q, q_pitch, k, k_pitch, v, v_pitch, key_padding, scale, dropout_p, seed, terms)) != ATQ_OK) return r;
  if ((r = check_attn_ptr(out, out_pitch, "out")) != ATQ_OK) return r;
  ATQ_CHECK_ARG(out_pitch >= (int64_t)H * head_dim, "out pitch smaller than H * head_dim");
  ATQ_ENSURE_DEVICE(device);
  p.out = out; p.out_pitch = out_pitch; p.lse = lse;
  p.NK = (L + 31) / 32 * 32;
  const int nt = terms == 3 ? 2 : 1;
  const int p_atoms = (p.NK + 63) / 64;
  const size_t smem = (size_t)2 * nt * p.NK * 128 + (size_t)nt * kTileBytes + (size_t)p_atoms * kTileBytes + 1024;
  static bool attr_done[64] = {false};
  if (!attr_done[device & 63]) {
    cudaError_t e = cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1280);
    if (e != cudaSuccess) {
      set_error("atq_attention_fwd: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return ATQ_ECUDA;
    }
    attr_done[device & 63] = true;
  }
  if (smem > (size_t)(227 * 1024 - 1280)) {
    set_error("atq_attention_fwd: shared memory %zu exceeds the per-CTA limit", smem);
    return ATQ_EINVAL;
  }
  attention_fwd_kernel<<<B * H, kAttThreads, smem, (cudaStream_t)stream_>>>(p);
  ATQ_LAUNCH_CHECK();
  return ATQ_OK;
}

int atq_attention_bwd(int device, int B, int H, int L, int head_dim, const float* q, int64_t q_pitch, const float* k, int64_t k_pitch,
                      const float* v, int64_t v_pitch, const uint8_t* key_padding, float scale, float dropout_p,
                      const unsigned long long* seed, int terms, const float* out, int64_t out_pitch, const float* dout,
                      int64_t dout_pitch, const float* lse, float* dq, int64_t dq_pitch, float* dk, int64_t dk_pitch,
                      float* dv, int64_t dv_pitch, atq_stream_t stream_) {
  AttnParams p;
  int r;
  if ((r = fill_common(p, B, H, L, head_dim, q, q_pitch, k, k_pitch, v, v_pitch, key_padding, scale, dropout_p, seed, terms)) != ATQ_OK) return r;
  if ((r = check_attn_ptr(out, out_pitch, "out")) != ATQ_OK) return r;
  if ((r = check_attn_ptr(dout, dout_pitch, "dout")) != ATQ_OK) return r;
  if ((r = check_attn_ptr(dq, dq_pitch, "dq")) != ATQ_OK) return r;
  if ((r = check_attn_ptr(dk, dk_pitch, "dk")) != ATQ_OK) return r;
  if ((r = check_attn_ptr(dv, dv_pitch, "dv")) != ATQ_OK) return r;
  ATQ_CHECK_ARG(lse != nullptr, "lse is null");
  const int64_t width = (int64_t)H * head_dim;
  ATQ_CHECK_ARG(out_pitch >= width && dout_pitch >= width && dq_pitch >= width && dk_pitch >= width && dv_pitch >= width,
                "row pitch smaller than H * head_dim");
  ATQ_ENSURE_DEVICE(device);
  p.o = out; p.o_pitch = out_pitch; p.dout = dout; p.do_pitch = dout_pitch;
  p.lse = const_cast<float*>(lse);
  p.dq = dq; p.dk = dk; p.dv = dv;
  p.dq_pitch = dq_pitch; p.dk_pitch = dk_pitch; p.dv_pitch = dv_pitch;
  p.NK = 128;
  const size_t smem = (size_t)14 * kTileBytes + 1024;
  static bool attr_done[64] = {false};
  if (!attr_done[device & 63]) {
    cudaError_t e = cudaFuncSetAttribute(attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("atq_attention_bwd: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return ATQ_ECUDA;
    }
    attr_done[device & 63] = true;
  }
  attention_bwd_kernel<<<B * H, kAttThreads, smem, (cudaStream_t)stream_>>>(p);
  ATQ_LAUNCH_CHECK();
  return ATQ_OK;
}

#ifdef ATQ_ATTN_PROF
int atq_debug_attn_prof(long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, g_attn_prof, sizeof(long long) * 64) == cudaSuccess ? 0 : 1;
}
#endif

}  // extern "C"
