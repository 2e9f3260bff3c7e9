// Ternary GEMMs for B200: TMA-fed tcgen05.mma with TMEM accumulators.
//
//   D[rows, cols] = sum_terms A_t[rows, kdim] . B_t[cols, kdim]^T      (fp32 accumulate in TMEM)
//
// Both operands are K-major bf16 tiles fetched by TMA into 128-byte-swizzled shared memory.
// fp32 activations / gradients arrive as a bf16 (hi, lo) pair (x = hi + lo to ~16 mantissa
// bits): the MMA warp issues one tcgen05.mma per term (hi*hi, lo*hi, hi*lo) into the SAME
// TMEM accumulator, which is what keeps the result inside rtol 1e-2 / atol 1e-3 of the
// reference's fp32 F.linear.  Ternary weights are exact in bf16.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer (one
// elected lane), warps 2..5 = epilogue (TMEM -> registers -> global) with the layer's
// epilogue fused: alpha scale, bias, routing/precision mask, d(alpha) reduction.
//
// Replaces: F.linear in atq/layers.py:43 and atq/precision_boost.py:74, the unpack-then-matmul
// of atq/bit_packing.py:165-176, and autograd's two backward GEMMs for those nodes.
#include <cuda.h>
#include "common.cuh"
#include "tc_sm100.cuh"

namespace atq {

constexpr int BLOCK_K = 64;  // 64 bf16 = 128 B = one SWIZZLE_128B row
constexpr int kGemmThreads = 192;
// packed-B kernels add BLOCK_N/32 converter warps (one B-tile row per thread)
constexpr int gemm_threads(bool packed, int block_n) { return kGemmThreads + (packed ? block_n : 0); }
constexpr int kSmemBudget = 227 * 1024 - 2048;

// operand layouts: 0 = A,B K-major ([rows|cols, k] row-major: forward), 1 = B MN-major (memory is
// [k, cols] row-major: dX straight from the [out, in] weight copy), 2 = A and B MN-major (memory is
// [k, rows] and [k, cols]: dW straight from dY [tokens, out] and X [tokens, in], no transposes)
enum { LAYOUT_KK = 0, LAYOUT_KM = 1, LAYOUT_MM = 2 };

// ------------------------------------------------------------------------------------------
// 2-bit codec -> bf16 in registers.  One 32-bit word holds 16 codes (code i at bits 2i..2i+1,
// value = code - 1).  Codes i and i+8 are 16 bits apart, so a single AND/OR drops both into the
// low mantissa bits of a bf16x2 whose exponent is 2^7: (0x4300 | c << sh) == 128 + c * 2^sh,
// and one bf16x2 FMA maps that to c - 1 exactly.  A byte permute restores K order.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t bf16x2_fma(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ uint32_t f16x2_fma(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
// fp16 flavour of the same trick: (0x6400 | c << sh) == 1024 + c * 2^sh exactly (10 mantissa bits), one f16x2 FMA -> c - 1
__device__ __forceinline__ void unpack16_to_f16(uint32_t w, uint4& lo8, uint4& hi8) {
  constexpr uint32_t kMagic = 0x64006400u;  // 1024.0, 1024.0
  constexpr uint32_t kS0 = 0x3C003C00u, kB0 = 0xE401E401u;  // * 1      - 1025
  constexpr uint32_t kS2 = 0x34003400u, kB2 = 0xDC04DC04u;  // * 0.25   - 257
  constexpr uint32_t kS4 = 0x2C002C00u, kB4 = 0xD410D410u;  // * 0.0625 - 65
  const uint32_t x1 = w >> 6, x2 = w >> 12;
  const uint32_t a0 = f16x2_fma((w & 0x00030003u) | kMagic, kS0, kB0);
  const uint32_t a1 = f16x2_fma((w & 0x000C000Cu) | kMagic, kS2, kB2);
  const uint32_t a2 = f16x2_fma((w & 0x00300030u) | kMagic, kS4, kB4);
  const uint32_t a3 = f16x2_fma((x1 & 0x00030003u) | kMagic, kS0, kB0);
  const uint32_t a4 = f16x2_fma((x1 & 0x000C000Cu) | kMagic, kS2, kB2);
  const uint32_t a5 = f16x2_fma((x1 & 0x00300030u) | kMagic, kS4, kB4);
  const uint32_t a6 = f16x2_fma((x2 & 0x00030003u) | kMagic, kS0, kB0);
  const uint32_t a7 = f16x2_fma((x2 & 0x000C000Cu) | kMagic, kS2, kB2);
  lo8 = make_uint4(__byte_perm(a0, a1, 0x5410), __byte_perm(a2, a3, 0x5410), __byte_perm(a4, a5, 0x5410),
                   __byte_perm(a6, a7, 0x5410));
  hi8 = make_uint4(__byte_perm(a0, a1, 0x7632), __byte_perm(a2, a3, 0x7632), __byte_perm(a4, a5, 0x7632),
                   __byte_perm(a6, a7, 0x7632));
}
__device__ __forceinline__ void unpack16_to_bf16(uint32_t w, uint4& lo8, uint4& hi8) {
  constexpr uint32_t kMagic = 0x43004300u;  // 128.0, 128.0
  constexpr uint32_t kS0 = 0x3F803F80u, kB0 = 0xC301C301u;  // * 1      - 129
  constexpr uint32_t kS2 = 0x3E803E80u, kB2 = 0xC204C204u;  // * 0.25   - 33
  constexpr uint32_t kS4 = 0x3D803D80u, kB4 = 0xC110C110u;  // * 0.0625 - 9
  const uint32_t x1 = w >> 6, x2 = w >> 12;
  const uint32_t a0 = bf16x2_fma((w & 0x00030003u) | kMagic, kS0, kB0);   // codes 0, 8
  const uint32_t a1 = bf16x2_fma((w & 0x000C000Cu) | kMagic, kS2, kB2);   // codes 1, 9
  const uint32_t a2 = bf16x2_fma((w & 0x00300030u) | kMagic, kS4, kB4);   // codes 2, 10
  const uint32_t a3 = bf16x2_fma((x1 & 0x00030003u) | kMagic, kS0, kB0);  // codes 3, 11
  const uint32_t a4 = bf16x2_fma((x1 & 0x000C000Cu) | kMagic, kS2, kB2);  // codes 4, 12
  const uint32_t a5 = bf16x2_fma((x1 & 0x00300030u) | kMagic, kS4, kB4);  // codes 5, 13
  const uint32_t a6 = bf16x2_fma((x2 & 0x00030003u) | kMagic, kS0, kB0);  // codes 6, 14
  const uint32_t a7 = bf16x2_fma((x2 & 0x000C000Cu) | kMagic, kS2, kB2);  // codes 7, 15
  lo8 = make_uint4(__byte_perm(a0, a1, 0x5410), __byte_perm(a2, a3, 0x5410), __byte_perm(a4, a5, 0x5410),
                   __byte_perm(a6, a7, 0x5410));  // values 0..7
  hi8 = make_uint4(__byte_perm(a0, a1, 0x7632), __byte_perm(a2, a3, 0x7632), __byte_perm(a4, a5, 0x7632),
                   __byte_perm(a6, a7, 0x7632));  // values 8..15
}

// ------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------
enum { EPI_LINEAR = 0, EPI_MASKED = 1 };
static bool g_cta_pairs = true;   // atq_set_cta_pairs(): A/B switch for the cta_group::2 kernels
static bool g_force_pairs = false;  // bit 3: pairs also below one wave of single-CTA tiles
static bool g_wide_pairs = true;  // bit 1 of the same switch: 256-wide single-buffered pair tiles for long contractions

struct GemmParams {
  int64_t rows, cols, kdim;
  float* out;
  int64_t out_pitch;
  const float* scale;     // device scalar, nullable           (LINEAR)
  const float* bias;      // [cols], nullable                   (LINEAR)
  const float* dot_ref;   // fp32 [rows, dot_ref_pitch], nullable (LINEAR): partial += acc * ref
  int64_t dot_ref_pitch;
  const float* mask;      // fp32 [rows, cols] contiguous, nullable (MASKED): out = acc * mask
  const uint8_t* tern;    // 2-bit codec bytes of T [rows*cols], nullable (MASKED): partial += acc*T*(1-mask)
  float* partials;        // [gridDim.x], nullable
  const uint8_t* b_packed;  // B_PACKED kernels: 2-bit codec bytes of T, [cols, kdim/4] row-major
  int64_t b_packed_pitch;   // bytes per row (= kdim / 4)
  int splits;               // split-K: work item = (tile, k-range); range s writes out + s * split_stride
  int64_t split_stride;     // elements
  const float* inv_a;       // nullable device scalars: 1/scale of a scaled-fp16 operand (exact powers of two);
  const float* inv_b;       //   the accumulator is multiplied by both before anything else in the epilogue
  int f16;                  // operands are fp16 (both), else bf16
  unsigned int* absmax_bits;  // nullable (LINEAR): atomicMax of the bit pattern of max|out| (slot[0] of a scale slot)
};

constexpr int kStagingBytes = 4 * 32 * 36 * 4;  // per-epilogue-warp [32][36] fp32 transposition buffers (rows 16-byte aligned)

template <int NUM_A, int NUM_B, int BLOCK_N, int BK = BLOCK_K, bool CTA2 = false>
struct GemmCfg {
  static constexpr int kABytes = BLOCK_M * BK * 2;
  static constexpr int kBRows = CTA2 ? BLOCK_N / 2 : BLOCK_N;  // a CTA of a pair holds half of the B tile
  static constexpr int kBBytes = kBRows * BK * 2;
  static constexpr int kStageBytes = NUM_A * kABytes + NUM_B * kBBytes;
  static constexpr int kStagesRaw = (kSmemBudget - kStagingBytes) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr bool kDual = (NUM_A + NUM_B > 2);
  static constexpr int kAccCols = (kDual ? 2 : 1) * BLOCK_N;
  // two tiles' accumulators (the MMA warp runs ahead of the epilogue) whenever they fit the 512 TMEM columns; the
  // 256-wide dual-accumulator tile fills TMEM by itself: single-buffered, worthwhile for long contractions
  static constexpr int kAccBufs = (2 * kAccCols <= 512) ? 2 : 1;
  static constexpr int kTmemCols = kAccBufs * kAccCols;
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + 1024 /*align slack*/ + 256 /*barriers*/;
  static_assert(kStages >= 2, "not enough shared memory for a 2-stage pipeline");
  static_assert(kTmemCols <= 512, "TMEM has 512 columns");
};

// Persistent, warp-specialised: one CTA per SM walks output tiles t = blockIdx.x + i*gridDim.x
// (column tile fastest, so the CTAs working at the same time share A tiles through L2 and the
// whole B operand stays L2-resident).  The accumulator is double-buffered in TMEM: the MMA warp
// starts tile i+1 while the epilogue warps drain tile i.
//
// CTA2 = true: the same kernel run by CTA PAIRS (cluster of 2, launched with a cluster dimension): a pair owns a
// 256 x BLOCK_N tile, each CTA loads its 128 rows of A and HALF of the B tile, the leader issues
// tcgen05.mma.cta_group::2 (M = 256), every CTA drains its own 128 accumulator rows.  B shared-memory reads per
// flop halve, which is what the 128-wide dual-accumulator tiles (hi/lo operands) are short of.
template <int NUM_A, int NUM_B, int BLOCK_N, int EPI, bool B_PACKED = false, int LAYOUT = LAYOUT_KK, int BK = BLOCK_K, bool CTA2 = false>
__global__ void __launch_bounds__(gemm_threads(B_PACKED, BLOCK_N), 1)
    tgemm_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                 const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                 const GemmParams p) {
  using Cfg = GemmCfg<NUM_A, NUM_B, BLOCK_N, BK, CTA2>;
  static_assert(!(CTA2 && (B_PACKED || BK != 64 || (BLOCK_N != 128 && BLOCK_N != 256))), "CTA pairs: 128 / 256-wide TMA-fed tiles only");
  constexpr int kAccBufs = Cfg::kAccBufs;
  constexpr int kPairM = CTA2 ? 2 * BLOCK_M : BLOCK_M;  // rows of the tile a work item covers
  const uint32_t cta_rank = CTA2 ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0u;
  const int work_first = CTA2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int work_step = CTA2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  constexpr int kStages = Cfg::kStages;
  constexpr bool A_MN = (LAYOUT == LAYOUT_MM), B_MN = (LAYOUT != LAYOUT_KK);
  static_assert(!(B_PACKED && LAYOUT != LAYOUT_KK), "packed B is K-major");
  static_assert(BK == 64 || BK == 32, "BLOCK_K is 64 (SWIZZLE_128B rows) or 32 (SWIZZLE_64B rows)");
  static_assert(!(B_PACKED && BK != 64), "the converter writes 128-byte rows");
  constexpr int kMnBlock = BK * 128;  // bytes of one [BK k x 64 mn] MN-major block
  constexpr bool kDual = Cfg::kDual;             // an operand has a lo part: second accumulator for the corrections
  constexpr int kAccCols = Cfg::kAccCols;        // TMEM columns per tile (1 or 2 accumulators)
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  constexpr int kBarOff = kStages * Cfg::kStageBytes + kStagingBytes;
  const uint32_t bars = smem_base + kBarOff;  // 8-byte aligned
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (kStages + s); };
  auto tmem_full_bar = [&](int a) { return bars + 8u * (2 * kStages + a); };
  auto tmem_empty_bar = [&](int a) { return bars + 8u * (2 * kStages + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * kStages + 4);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + kBarOff + 8 * (2 * kStages + 4));
  float* s_part = reinterpret_cast<float*>(smem_gen + kBarOff + 8 * (2 * kStages + 5));

  const int warp_idx = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int num_kb = (int)((p.kdim + BK - 1) / BK);
  const int tiles_n = (int)((p.cols + BLOCK_N - 1) / BLOCK_N);
  const int tiles_m = (int)((p.rows + kPairM - 1) / kPairM);
  const int num_tiles = tiles_n * tiles_m;
  const int splits = p.splits > 1 ? p.splits : 1;
  const int num_work = num_tiles * splits;
  // L2-aware rasterisation: tiles are walked in bands of kGroupM row tiles (column tile next, row-in-band fastest), so
  // the CTAs running at the same time cover ~kGroupM row tiles x (SMs / kGroupM) column tiles: both operands of a wave
  // stay in L2 and B is re-read tiles_m / kGroupM times instead of once per ~2 row tiles (ncu: 1.02 GB -> see profiles)
  constexpr int kGroupM = 8;
  auto tile_mn = [&](int t, int& tm, int& tn) {
    const int per_band = kGroupM * tiles_n;
    const int band = t / per_band;
    const int first_m = band * kGroupM;
    const int rows_in_band = (tiles_m - first_m) < kGroupM ? (tiles_m - first_m) : kGroupM;
    const int r = t - band * per_band;
    tm = first_m + r % rows_in_band;
    tn = r / rows_in_band;
  };
  // work item w -> tile w % num_tiles, k-blocks [kb_lo(w), kb_hi(w)) (balanced, never empty: splits <= num_kb)
  auto kb_lo = [&](int w) { return (int)(((long long)(w / num_tiles) * num_kb) / splits); };
  auto kb_hi = [&](int w) { return (int)(((long long)(w / num_tiles + 1) * num_kb) / splits); };

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&map_a_hi);
    if (!B_PACKED) tma_prefetch_desc(&map_b_hi);
    if (NUM_A == 2) tma_prefetch_desc(&map_a_lo);
    if (NUM_B == 2) tma_prefetch_desc(&map_b_lo);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), B_PACKED ? 1 + BLOCK_N / 32 : 1);  // TMA expect-tx arrival (+ one per converter warp)
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tmem_full_bar(a), 1);
      mbar_init(tmem_empty_bar(a), CTA2 ? 8 : 4);  // one arrival per epilogue warp (of both CTAs of a pair: the leader's copy counts)
    }
    fence_barrier_init();
  }
  if (warp_idx == 1) {
    if constexpr (CTA2) tmem_alloc_2sm<Cfg::kTmemCols>(tmem_slot);
    else tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  }
  tcgen05_fence_before();
  if constexpr (CTA2) cluster_sync_all();  // both CTAs' barriers initialised and TMEM allocated before any cross-CTA signal
  else __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  if (warp_idx == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      // (CTA pairs: the loads of both CTAs count on the leader's full barrier)
      auto tma_load_2d = [&](uint32_t dst, const CUtensorMap* map, int32_t c0, int32_t c1, uint32_t bar) {
        if constexpr (CTA2) atq::tma_load_2d_2sm(dst, map, c0, c1, bar);
        else atq::tma_load_2d(dst, map, c0, c1, bar);
      };
      for (int w = work_first; w < num_work; w += work_step) {
        int tm, tn;
        tile_mn(w % num_tiles, tm, tn);
        const int32_t n0 = tn * BLOCK_N + (int32_t)cta_rank * Cfg::kBRows;       // this CTA's part of the B tile
        const int32_t m0 = tm * kPairM + (int32_t)cta_rank * BLOCK_M;            // this CTA's rows of A
        for (int kb = kb_lo(w); kb < kb_hi(w); ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          if (leader) mbar_expect_tx(full_bar(stage), (B_PACKED ? NUM_A * Cfg::kABytes : Cfg::kStageBytes) * (CTA2 ? 2 : 1));
          const int32_t kc = kb * BK;
          if constexpr (!A_MN) {
            tma_load_2d(sa, &map_a_hi, kc, m0, full_bar(stage));
            if (NUM_A == 2) tma_load_2d(sa + Cfg::kABytes, &map_a_lo, kc, m0, full_bar(stage));
          } else {  // [64 k x 64 mn] boxes of the row-major [k, rows] tensor
#pragma unroll
            for (int j = 0; j < BLOCK_M / 64; ++j) {
              tma_load_2d(sa + j * kMnBlock, &map_a_hi, m0 + 64 * j, kc, full_bar(stage));
              if (NUM_A == 2) tma_load_2d(sa + Cfg::kABytes + j * kMnBlock, &map_a_lo, m0 + 64 * j, kc, full_bar(stage));
            }
          }
          if constexpr (!B_PACKED) {
            const uint32_t sb = sa + NUM_A * Cfg::kABytes;
            if constexpr (!B_MN) {
              tma_load_2d(sb, &map_b_hi, kc, n0, full_bar(stage));
              if (NUM_B == 2) tma_load_2d(sb + Cfg::kBBytes, &map_b_lo, kc, n0, full_bar(stage));
            } else {
#pragma unroll
              for (int j = 0; j < (Cfg::kBRows + 63) / 64; ++j) {
                tma_load_2d(sb + j * kMnBlock, &map_b_hi, n0 + 64 * j, kc, full_bar(stage));
                if (NUM_B == 2) tma_load_2d(sb + Cfg::kBBytes + j * kMnBlock, &map_b_lo, n0 + 64 * j, kc, full_bar(stage));
              }
            }
          }
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ================= MMA issuer (the leader CTA of a pair) =================
    if (lane == 0 && leader) {
      auto umma_bf16 = [&](uint32_t d, uint64_t da, uint64_t db, uint32_t id, uint32_t acc) {
        if constexpr (CTA2) atq::umma_bf16_2sm(d, da, db, id, acc);
        else atq::umma_bf16(d, da, db, id, acc);
      };
      auto umma_commit = [&](uint32_t bar) {
        if constexpr (CTA2) atq::umma_commit_2sm(bar);
        else atq::umma_commit(bar);
      };
      // a/b format fields (bits 7-9, 10-12): 1 = bf16, 0 = fp16
      uint32_t idesc = p.f16 ? (make_idesc<BLOCK_N, A_MN, B_MN>() & ~((7u << 7) | (7u << 10))) : make_idesc<BLOCK_N, A_MN, B_MN>();
      if constexpr (CTA2) idesc = (idesc & ~(31u << 24)) | ((uint32_t)(kPairM >> 4) << 24);  // M = 256
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int w = work_first; w < num_work; w += work_step, ++it) {
        const int a = it % kAccBufs;
        const uint32_t aphase = (uint32_t)(it / kAccBufs) & 1u;
        mbar_wait(tmem_empty_bar(a), aphase ^ 1u);  // epilogue has drained this accumulator buffer
        tcgen05_fence_after();
        // Two accumulators per tile when an operand has a lo part.  tcgen05 accumulates in fp32 with truncation:
        // every instruction rounds the WHOLE accumulator once (<= 1 ulp toward zero, however small the addend;
        // measured ~4e-8 relative per instruction, 1.7e-5 at k = 4096 with three interleaved terms).  The
        // correction terms (lo x hi, hi x lo; ~2^-11 of the result) therefore go to their own accumulator, whose
        // roundings are 2^-11 as large, and the hi x hi accumulator is rounded k/16 times instead of 2-3 k/16
        // times; the epilogue adds the two in fp32 (round to nearest).
        const uint32_t tmem_acc = tmem_base + (uint32_t)(a * kAccCols);
        const uint32_t tmem_cor = tmem_acc + (uint32_t)BLOCK_N;
        uint32_t accumulate = 0, accumulate_cor = 0;
        for (int kb = kb_lo(w); kb < kb_hi(w); ++kb) {
          mbar_wait(full_bar(stage), phase);
          tcgen05_fence_after();
          const uint32_t sa = smem_base + stage * Cfg::kStageBytes;
          const uint32_t sb = sa + NUM_A * Cfg::kABytes;
          auto kdesc = [](uint32_t addr) { return BK == 64 ? make_smem_desc_kmajor_sw128(addr) : make_smem_desc_kmajor_sw64(addr); };
          const uint64_t da_hi = A_MN ? make_smem_desc_mnmajor_sw128(sa, kMnBlock) : kdesc(sa);
          const uint64_t da_lo = A_MN ? make_smem_desc_mnmajor_sw128(sa + Cfg::kABytes, kMnBlock) : kdesc(sa + Cfg::kABytes);
          const uint64_t db_hi = B_MN ? make_smem_desc_mnmajor_sw128(sb, kMnBlock) : kdesc(sb);
          const uint64_t db_lo = B_MN ? make_smem_desc_mnmajor_sw128(sb + Cfg::kBBytes, kMnBlock) : kdesc(sb + Cfg::kBBytes);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // one K step (16 elements): 32 B inside the swizzled row when K-major, 16 rows of 128 B
            // when MN-major; in 16-byte units
            const uint64_t koa = (uint64_t)(A_MN ? (k * UMMA_K * 128) >> 4 : (k * UMMA_K * 2) >> 4);
            const uint64_t kob = (uint64_t)(B_MN ? (k * UMMA_K * 128) >> 4 : (k * UMMA_K * 2) >> 4);
            umma_bf16(tmem_acc, da_hi + koa, db_hi + kob, idesc, accumulate);
            accumulate = 1;
            if (NUM_A == 2) {
              umma_bf16(tmem_cor, da_lo + koa, db_hi + kob, idesc, accumulate_cor);
              accumulate_cor = 1;
            }
            if (NUM_B == 2) {
              umma_bf16(tmem_cor, da_hi + koa, db_lo + kob, idesc, accumulate_cor);
              accumulate_cor = 1;
            }
          }
          umma_commit(empty_bar(stage));  // frees this smem stage once the MMAs above retire
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(tmem_full_bar(a));  // accumulator of this tile complete
      }
    }
  } else if (B_PACKED && warp_idx >= 6) {
    // ================= converter warps 6..: packed 2-bit T -> bf16 B tile in shared memory =================
    // Each thread owns one row of the tile: one 16-byte load (64 codes) per
    // k-block, expanded to 128 bytes of bf16 and stored with the SWIZZLE_128B pattern the UMMA
    // descriptor expects (16-byte chunk c of row r lives at chunk c ^ (r & 7)).
    if constexpr (B_PACKED) {
    constexpr int kRows = 1;
    const int ct = threadIdx.x - 6 * 32;  // 0..BLOCK_N-1
    int stage = 0;
    uint32_t phase = 0;
    for (int w = blockIdx.x; w < num_work; w += gridDim.x) {  // packed kernels are never split (splits == 1)
      int tm, tn;
      tile_mn(w % num_tiles, tm, tn);
      const int64_t n0 = (int64_t)tn * BLOCK_N;
      uint4 cur[kRows];
      auto load_row = [&](int i, int kb) -> uint4 {
        const int64_t row = n0 + ct + i * BLOCK_N;
        if (row < p.cols) return __ldg(reinterpret_cast<const uint4*>(p.b_packed + row * p.b_packed_pitch + (int64_t)kb * 16));
        return make_uint4(0x55555555u, 0x55555555u, 0x55555555u, 0x55555555u);  // code 1 = 0.0
      };
#pragma unroll
      for (int i = 0; i < kRows; ++i) cur[i] = load_row(i, 0);
      for (int kb = 0; kb < num_kb; ++kb) {
        uint4 nxt[kRows];
        if (kb + 1 < num_kb) {
#pragma unroll
          for (int i = 0; i < kRows; ++i) nxt[i] = load_row(i, kb + 1);  // in flight while this block converts
        }
        mbar_wait(empty_bar(stage), phase ^ 1u);
        const uint32_t sb = smem_base + stage * Cfg::kStageBytes + NUM_A * Cfg::kABytes;
#pragma unroll
        for (int i = 0; i < kRows; ++i) {
          const int r = ct + i * BLOCK_N;
          const uint32_t row_addr = sb + (uint32_t)r * 128u;
          const uint32_t sw = (uint32_t)(r & 7);
          const uint32_t words[4] = {cur[i].x, cur[i].y, cur[i].z, cur[i].w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 lo8, hi8;
            if (p.f16) unpack16_to_f16(words[q], lo8, hi8);
            else unpack16_to_bf16(words[q], lo8, hi8);
            st_shared_v4(row_addr + (((uint32_t)(2 * q) ^ sw) << 4), lo8);
            st_shared_v4(row_addr + (((uint32_t)(2 * q + 1) ^ sw) << 4), hi8);
          }
        }
        fence_proxy_async();  // make these generic-proxy stores visible to tcgen05.mma (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(full_bar(stage));
        if (kb + 1 < num_kb) {
#pragma unroll
          for (int i = 0; i < kRows; ++i) cur[i] = nxt[i];
        }
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
    }
    }  // if constexpr (B_PACKED)
  } else {
    // ================= epilogue warps 2..5 =================
    // TMEM gives each thread one accumulator ROW (32 columns per tcgen05.ld).  Rows are staged
    // through a private shared-memory buffer so that every global access below is a warp-wide
    // 128-byte row segment: lane = column.
    const int quarter = warp_idx & 3;  // TMEM lane quarter this warp may access
    float* stg = reinterpret_cast<float*>(smem_gen + kStages * Cfg::kStageBytes) + (warp_idx - 2) * (32 * 36);
    const float scale = (EPI == EPI_LINEAR && p.scale != nullptr) ? __ldg(p.scale) : 1.f;
    // 1/(s_a s_b) of scaled-fp16 operands: exact power of two, applied to the raw accumulator first
    const float opscale = (p.inv_a != nullptr ? __ldg(p.inv_a) : 1.f) * (p.inv_b != nullptr ? __ldg(p.inv_b) : 1.f);
    float partial = 0.f;
    float amax = 0.f;  // max |stored output| of this thread (p.absmax_bits: the consumer's operand scale, no extra pass)
    // 128-bit path: every lane owns 4 consecutive columns of 8 rows per chunk (4x fewer shared/global
    // instructions than lane = column); needs 16-byte aligned rows on every tensor the epilogue touches
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
    const bool vec_ok = ((p.cols | p.out_pitch | p.split_stride) & 3) == 0 && al16(p.out) &&
                        (EPI == EPI_LINEAR ? ((p.bias == nullptr || al16(p.bias)) &&
                                              (p.dot_ref == nullptr || (al16(p.dot_ref) && (p.dot_ref_pitch & 3) == 0)))
                                           : (p.mask == nullptr || al16(p.mask)));
    const int vq = lane & 7, vrg = lane >> 3;
    int it = 0;
    // the MMA issuer waits on the LEADER's tmem_empty barrier: both CTAs of a pair arrive there
    auto release_acc = [&](int a) {
      if constexpr (CTA2) mbar_arrive_leader(tmem_empty_bar(a));
      else mbar_arrive(tmem_empty_bar(a));
    };
    for (int w = work_first; w < num_work; w += work_step, ++it) {
      const int a = it % kAccBufs;
      const uint32_t aphase = (uint32_t)(it / kAccBufs) & 1u;
      int tm, tn;
      tile_mn(w % num_tiles, tm, tn);
      const int64_t n0 = (int64_t)tn * BLOCK_N;
      const int64_t m0 = (int64_t)tm * kPairM + (int64_t)cta_rank * BLOCK_M;  // this CTA's accumulator rows
      float* const out = p.out + (int64_t)(w / num_tiles) * p.split_stride;  // split-K partial slab
      const int64_t r_base = m0 + quarter * 32;
      // number of 32-column chunks of this tile that hold real output (warp-uniform)
      int nch = (int)((p.cols - n0 + 31) / 32);
      if (nch > BLOCK_N / 32) nch = BLOCK_N / 32;
      if (r_base >= p.rows) nch = 0;
      const int rows_here = (int)((p.rows - r_base) < 32 ? (p.rows - r_base) : 32);
      // Side inputs of the epilogue (mask + codec bytes, or the dot-product reference) do not depend on the
      // accumulator: chunk 0 is fetched BEFORE waiting for the MMAs and chunk ch+1 while chunk ch is
      // processed, so short GEMMs do not pay one memory latency per chunk after the mainloop.
      const bool side = (EPI == EPI_MASKED) || p.dot_ref != nullptr;
      const uint32_t tmem_acc = tmem_base + (uint32_t)(a * kAccCols) + ((uint32_t)(quarter * 32) << 16);
      if (vec_ok) {
        // ---- 128-bit path ----
        auto load_side_v = [&](int ch, float4 (&f)[8], uint32_t (&cd)[8]) {
          const int64_t c = n0 + ch * 32 + 4 * vq;
          const bool col_ok = c < p.cols;
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const int row = 4 * g + vrg;
            const bool ok = col_ok && row < rows_here;
            if constexpr (EPI == EPI_LINEAR) {
              f[g] = ok ? __ldg(reinterpret_cast<const float4*>(p.dot_ref + (r_base + row) * p.dot_ref_pitch + c))
                        : make_float4(0.f, 0.f, 0.f, 0.f);
            } else {
              const int64_t i = (r_base + row) * p.cols + c;  // multiple of 4: one codec byte per group
              f[g] = (p.mask != nullptr && ok) ? __ldg(reinterpret_cast<const float4*>(p.mask + i)) : make_float4(1.f, 1.f, 1.f, 1.f);
              cd[g] = (p.tern != nullptr && ok) ? (uint32_t)__ldg(p.tern + (i >> 2)) : 0x55u;
            }
          }
        };
        float4 vf[8];
        uint32_t vc[8];
        if (side && nch > 0) load_side_v(0, vf, vc);
        mbar_wait(tmem_full_bar(a), aphase);
        tcgen05_fence_after();
        if (nch == 0) {
          tcgen05_fence_before();
          if (lane == 0) release_acc(a);
        }
#pragma unroll 1
        for (int ch = 0; ch < nch; ++ch) {
          uint32_t acc[32];
          tmem_ld_32x32b_x32(tmem_acc + (uint32_t)(ch * 32), acc);
          uint32_t cor[kDual ? 32 : 1];
          if constexpr (kDual) tmem_ld_32x32b_x32(tmem_acc + (uint32_t)(BLOCK_N + ch * 32), cor);
          float4 nf[8];
          uint32_t nc[8];
          if (side && ch + 1 < nch) load_side_v(ch + 1, nf, nc);
          tmem_ld_wait();
          if constexpr (kDual) {
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] = __float_as_uint(__uint_as_float(acc[j]) + __uint_as_float(cor[j]));
          }
          if (ch == nch - 1) {  // everything this warp needs has left TMEM: hand the buffer back
            tcgen05_fence_before();
            if (lane == 0) release_acc(a);
          }
          // thread = accumulator row: 8 x 16-byte stores, 144-byte row pitch (conflict-free per quarter-warp)
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(stg + lane * 36 + 4 * j) = make_uint4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
          __syncwarp();
          const int64_t c = n0 + ch * 32 + 4 * vq;
          const bool col_ok = c < p.cols;
          float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if constexpr (EPI == EPI_LINEAR) {
            if (p.bias != nullptr && col_ok) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + c));
          }
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const int row = 4 * g + vrg;
            if (col_ok && row < rows_here) {
              float4 v = *reinterpret_cast<const float4*>(stg + row * 36 + 4 * vq);
              v.x *= opscale; v.y *= opscale; v.z *= opscale; v.w *= opscale;
              float4 o;
              if constexpr (EPI == EPI_LINEAR) {
                if (p.dot_ref != nullptr) partial += (v.x * vf[g].x + v.y * vf[g].y) + (v.z * vf[g].z + v.w * vf[g].w);
                o = make_float4(v.x * scale + bias4.x, v.y * scale + bias4.y, v.z * scale + bias4.z, v.w * scale + bias4.w);
                amax = fmaxf(fmaxf(amax, fmaxf(fabsf(o.x), fabsf(o.y))), fmaxf(fabsf(o.z), fabsf(o.w)));
              } else {
                const uint32_t b = vc[g];
                const float t0 = (float)(b & 3u) - 1.f, t1 = (float)((b >> 2) & 3u) - 1.f;
                const float t2 = (float)((b >> 4) & 3u) - 1.f, t3 = (float)((b >> 6) & 3u) - 1.f;
                const float4 mk = vf[g];
                partial += (v.x * t0 * (1.f - mk.x) + v.y * t1 * (1.f - mk.y)) + (v.z * t2 * (1.f - mk.z) + v.w * t3 * (1.f - mk.w));
                o = make_float4(v.x * mk.x, v.y * mk.y, v.z * mk.z, v.w * mk.w);
              }
              *reinterpret_cast<float4*>(out + (r_base + row) * p.out_pitch + c) = o;
            }
          }
          if (side && ch + 1 < nch) {
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              vf[g] = nf[g];
              if constexpr (EPI == EPI_MASKED) vc[g] = nc[g];
            }
          }
          __syncwarp();  // staging buffer reuse + reconverge before the next warp-aligned tcgen05.ld
        }
        continue;
      }
      // ---- scalar path (columns / pitches that are not multiples of 4, unaligned views): lane = column ----
      mbar_wait(tmem_full_bar(a), aphase);
      tcgen05_fence_after();
      if (nch == 0) {
        tcgen05_fence_before();
        if (lane == 0) release_acc(a);
      }
#pragma unroll 1
      for (int ch = 0; ch < nch; ++ch) {
        uint32_t acc[32];
        tmem_ld_32x32b_x32(tmem_acc + (uint32_t)(ch * 32), acc);
        uint32_t cor[kDual ? 32 : 1];
        if constexpr (kDual) tmem_ld_32x32b_x32(tmem_acc + (uint32_t)(BLOCK_N + ch * 32), cor);
        tmem_ld_wait();
        if constexpr (kDual) {
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[j] = __float_as_uint(__uint_as_float(acc[j]) + __uint_as_float(cor[j]));
        }
        if (ch == nch - 1) {
          tcgen05_fence_before();
          if (lane == 0) release_acc(a);
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) stg[lane * 36 + j] = __uint_as_float(acc[j]);  // 4-way bank conflict accepted here
        __syncwarp();
        const int64_t c = n0 + ch * 32 + lane;
        if (c < p.cols) {
          const float bias = (EPI == EPI_LINEAR && p.bias != nullptr) ? __ldg(p.bias + c) : 0.f;
#pragma unroll 4
          for (int rr = 0; rr < rows_here; ++rr) {
            const float v = stg[rr * 36 + lane] * opscale;
            if constexpr (EPI == EPI_LINEAR) {
              if (p.dot_ref != nullptr) partial += v * __ldg(p.dot_ref + (r_base + rr) * p.dot_ref_pitch + c);
              const float o1 = v * scale + bias;
              amax = fmaxf(amax, fabsf(o1));
              out[(r_base + rr) * p.out_pitch + c] = o1;
            } else {
              const int64_t i = (r_base + rr) * p.cols + c;  // mask / codec bytes are contiguous [rows, cols]
              const float mk = p.mask != nullptr ? __ldg(p.mask + i) : 1.f;
              const uint32_t code = p.tern != nullptr ? (((uint32_t)__ldg(p.tern + (i >> 2)) >> (2 * (int)(i & 3))) & 3u) : 1u;
              partial += v * ((float)code - 1.f) * (1.f - mk);
              out[(r_base + rr) * p.out_pitch + c] = v * mk;
            }
          }
        }
        __syncwarp();
      }
    }
    if (p.partials != nullptr) {
      partial = warp_sum(partial);
      if (lane == 0) s_part[quarter] = partial;
    }
    if (EPI == EPI_LINEAR && p.absmax_bits != nullptr) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
      if (lane == 0) atomicMax(p.absmax_bits, __float_as_uint(amax));  // non-negative floats order as uints
    }
  }
  tcgen05_fence_before();
  if constexpr (CTA2) cluster_sync_all();  // the peer may still be reading / signalling: nobody leaves early
  else __syncthreads();
  if (warp_idx == 1) {
    tcgen05_fence_after();
    if constexpr (CTA2) tmem_dealloc_2sm<Cfg::kTmemCols>(tmem_base);
    else tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
  if (p.partials != nullptr && threadIdx.x == 0) {
    p.partials[blockIdx.x] = (s_part[0] + s_part[1]) + (s_part[2] + s_part[3]);
  }
}

// scale slot written by a GEMM epilogue (slot[0] = bits of max|out|): derive {scale, 1/scale} for bound = max * bound_mul
// and re-arm the slot (CUDA-graph replays)
__global__ void slot_finalize_kernel(float* slot, float bound_mul) {
  unsigned int* bits = reinterpret_cast<unsigned int*>(slot);
  const float b = __uint_as_float(bits[0]) * bound_mul;
  float sc, inv;
  pow2_scale_for(b, sc, inv);
  slot[1] = sc;
  slot[2] = inv;
  bits[0] = 0u;
}

// deterministic final sum of the per-CTA partials
__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ part, int64_t n, float* __restrict__ out) {
  __shared__ double s[256];
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += 256) acc += (double)part[i];
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = (float)s[0];
}

// split-K second stage for the masked dW GEMM: G = sum_s partial[s] (fixed order), dW = G .* mask,
// per-CTA partial of sum(G .* T .* (1-mask)) for d(alpha).  Coalesced, one element per thread step.
__global__ void __launch_bounds__(256)
    splitk_finalize_masked_kernel(const float* __restrict__ slabs, int splits, int64_t slab_stride, int64_t rows, int64_t cols,
                                  const float* __restrict__ mask, const uint8_t* __restrict__ tern, float* __restrict__ out,
                                  int64_t out_pitch, float* __restrict__ partials) {
  __shared__ float s_part[8];
  const int64_t n = rows * cols;
  float partial = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    float g = 0.f;
    for (int s = 0; s < splits; ++s) g += slabs[s * slab_stride + i];
    const float mk = mask ? __ldg(mask + i) : 1.f;
    if (tern != nullptr) {
      const uint32_t code = ((uint32_t)__ldg(tern + (i >> 2)) >> (2 * (int)(i & 3))) & 3u;
      partial += g * ((float)code - 1.f) * (1.f - mk);
    }
    const int64_t r = i / cols, c = i - r * cols;
    out[r * out_pitch + c] = g * mk;
  }
  if (partials != nullptr) {
    partial = warp_sum(partial);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = partial;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int i = 0; i < 8; ++i) t += s_part[i];
      partials[blockIdx.x] = t;
    }
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  }
  return fn;
}

// bf16 [rows, kdim] row-major with pitch; box = [BLOCK_K, box_rows]; OOB reads return zero
static int make_map(CUtensorMap* map, const uint16_t* ptr, int64_t rows, int64_t kdim, int64_t pitch, int box_rows, int bk = BLOCK_K) {
  PFN_encodeTiled enc = get_encode_fn();
  if (enc == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return ATQ_ECUDA;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)kdim, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)pitch * 2};
  cuuint32_t box[2] = {(cuuint32_t)bk, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)ptr, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) rows=%lld kdim=%lld pitch=%lld ptr=%p", (int)r, (long long)rows, (long long)kdim,
              (long long)pitch, (const void*)ptr);
    return ATQ_ECUDA;
  }
  return ATQ_OK;
}

// bf16 [kdim, mn] row-major with pitch (MN-major operand); box = [64 mn, 64 k]
static int make_map_mn(CUtensorMap* map, const uint16_t* ptr, int64_t mn, int64_t kdim, int64_t pitch, int bk = BLOCK_K) {
  PFN_encodeTiled enc = get_encode_fn();
  if (enc == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return ATQ_ECUDA;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)mn, (cuuint64_t)kdim};
  cuuint64_t gstride[1] = {(cuuint64_t)pitch * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)bk};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)ptr, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (MN-major) failed (%d) mn=%lld kdim=%lld pitch=%lld ptr=%p", (int)r, (long long)mn,
              (long long)kdim, (long long)pitch, (const void*)ptr);
    return ATQ_ECUDA;
  }
  return ATQ_OK;
}

template <int NUM_A, int NUM_B, int BLOCK_N, int EPI, bool B_PACKED = false, int LAYOUT = LAYOUT_KK, int BK = BLOCK_K, bool CTA2 = false>
static int launch_cfg(const atq_bf16_operand* a, const atq_bf16_operand* b, const GemmParams& p, cudaStream_t stream, int* grid_used) {
  using Cfg = GemmCfg<NUM_A, NUM_B, BLOCK_N, BK, CTA2>;
  CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
  int r;
  constexpr bool A_MN = (LAYOUT == LAYOUT_MM), B_MN = (LAYOUT != LAYOUT_KK);
  auto map_a = [&](CUtensorMap* m, const uint16_t* ptr) {
    return A_MN ? make_map_mn(m, ptr, p.rows, p.kdim, a->pitch, BK) : make_map(m, ptr, p.rows, p.kdim, a->pitch, BLOCK_M, BK);
  };
  auto map_b = [&](CUtensorMap* m, const uint16_t* ptr) {
    return B_MN ? make_map_mn(m, ptr, p.cols, p.kdim, b->pitch, BK) : make_map(m, ptr, p.cols, p.kdim, b->pitch, Cfg::kBRows, BK);
  };
  if ((r = map_a(&ma_hi, a->hi)) != ATQ_OK) return r;
  if constexpr (!B_PACKED) {
    if ((r = map_b(&mb_hi, b->hi)) != ATQ_OK) return r;
  } else {
    mb_hi = ma_hi;  // unused by the packed-B kernel
  }
  ma_lo = ma_hi;
  mb_lo = mb_hi;
  if (NUM_A == 2 && (r = map_a(&ma_lo, a->lo)) != ATQ_OK) return r;
  if (NUM_B == 2 && (r = map_b(&mb_lo, b->lo)) != ATQ_OK) return r;
  GemmParams pp = p;
  pp.f16 = a->format != 0;
  pp.inv_a = a->inv_scale;
  pp.inv_b = B_PACKED ? nullptr : b->inv_scale;
  auto kern = tgemm_kernel<NUM_A, NUM_B, BLOCK_N, EPI, B_PACKED, LAYOUT, BK, CTA2>;
  static bool attr_done_dev[64] = {false};  // per instantiation, per device
  int dev = 0;
  cudaGetDevice(&dev);
  bool& attr_done = attr_done_dev[dev & 63];
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(smem=%d) failed: %s", Cfg::kSmemBytes, cudaGetErrorString(e));
      return ATQ_ECUDA;
    }
    attr_done = true;
  }
  constexpr int kTileM = CTA2 ? 2 * BLOCK_M : BLOCK_M;
  const int64_t tiles = ((p.cols + BLOCK_N - 1) / BLOCK_N) * ((p.rows + kTileM - 1) / kTileM);
  if (tiles > 0x7fffffff) {
    set_error("tgemm: too many tiles (%lld)", (long long)tiles);
    return ATQ_EINVAL;
  }
  const int sms = sm_count(dev);
  const int64_t work = tiles * (p.splits > 1 ? p.splits : 1);
  int grid = (int)(work < sms ? work : sms);
  cudaError_t e;
  if constexpr (CTA2) {
    // one cluster of 2 CTAs (the two SMs of a TPC) per work item; persistent over min(work, SMs / 2) pairs
    const int pairs = (int)(work < sms / 2 ? work : sms / 2);
    grid = 2 * pairs;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)gemm_threads(B_PACKED, BLOCK_N));
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, kern, ma_hi, ma_lo, mb_hi, mb_lo, pp);
    if (e != cudaSuccess) {
      set_error("tgemm (CTA pair) launch failed: %s", cudaGetErrorString(e));
      return ATQ_ECUDA;
    }
  } else {
    kern<<<grid, gemm_threads(B_PACKED, BLOCK_N), Cfg::kSmemBytes, stream>>>(ma_hi, ma_lo, mb_hi, mb_lo, pp);
  }
  *grid_used = grid;
  e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("tgemm launch failed: %s", cudaGetErrorString(e));
    return ATQ_ECUDA;
  }
  note_launch();
  return ATQ_OK;
}

template <int EPI, int LAYOUT>
static int dispatch_layout(const atq_bf16_operand* a, const atq_bf16_operand* b, const GemmParams& p, cudaStream_t stream, int* grid_used) {
  const bool a2 = a->lo != nullptr, b2 = b->lo != nullptr;
  // tile width: 256 columns when the pipeline still has >= 3 stages, else 128; narrow outputs use 64/128
  const bool narrow = p.cols <= 64;
  const bool mid = p.cols <= 128;
#define ATQ_GO(NA, NB, BN) return launch_cfg<NA, NB, BN, EPI, false, LAYOUT>(a, b, p, stream, grid_used)
  if (narrow) {
    if (a2 && b2) ATQ_GO(2, 2, 64);
    if (a2) ATQ_GO(2, 1, 64);
    if (b2) ATQ_GO(1, 2, 64);
    ATQ_GO(1, 1, 64);
  }
  // operands with a lo part keep two accumulators per tile (hi x hi, corrections): 4 x BLOCK_N TMEM columns
  // double-buffered, so those kernels use 128-wide tiles; single-term GEMMs use 256-wide tiles when wide enough
  // CTA pairs (cta_group::2) once the output has at least one 256-row tile per pair-column: halves the B-operand
  // shared-memory reads per flop of the 128-wide dual-accumulator tiles
  // ... and only when the problem fills the machine at least once with single-CTA tiles: below that the GEMM is bound by
  // its fixed latency, and the cluster launch + the two cluster barriers of a pair cost more than the shared-memory
  // bandwidth they save (CUPTI, 800 x 192 x 192, 3 terms: 9.2 us as pairs, 8.0 us as single CTAs; dX 8.7 vs 7.6 us)
  int dev_p = 0;
  cudaGetDevice(&dev_p);
  const int64_t tiles128 = ((p.rows + BLOCK_M - 1) / BLOCK_M) * ((p.cols + 127) / 128) * (p.splits > 1 ? p.splits : 1);
  const bool pairs = g_cta_pairs && p.rows >= 256 && p.cols >= 128 && (g_force_pairs || tiles128 >= sm_count(dev_p));
#define ATQ_GO2(NA, NB, BN) return launch_cfg<NA, NB, BN, EPI, false, LAYOUT, BLOCK_K, true>(a, b, p, stream, grid_used)
  if (pairs) {
    // long contractions: 256 x 256 pair tiles (half the operand bytes per flop; the two 256-wide accumulators fill
    // TMEM, so the epilogue is not overlapped -- it is < 15 % of a tile from ~32 k-blocks on)
    int dev = 0;
    cudaGetDevice(&dev);
    const int64_t kb = (p.kdim + BLOCK_K - 1) / BLOCK_K / (p.splits > 1 ? p.splits : 1);
    const int64_t tiles256 = ((p.cols + 255) / 256) * ((p.rows + 255) / 256) * (p.splits > 1 ? p.splits : 1);
    const bool wide = g_wide_pairs && kb >= 32 && p.cols >= 256 && tiles256 * 2 >= sm_count(dev) / 2;
    if (wide) {
      if (a2 && b2) ATQ_GO2(2, 2, 256);
      if (b2) ATQ_GO2(1, 2, 256);
      if (a2) ATQ_GO2(2, 1, 256);
    }
    if (a2 && b2) ATQ_GO2(2, 2, 128);
    if (b2) ATQ_GO2(1, 2, 128);
    if (a2) ATQ_GO2(2, 1, 128);
  }
#undef ATQ_GO2
  if (a2 && b2) ATQ_GO(2, 2, 128);
  if (b2) ATQ_GO(1, 2, 128);
  if (a2) ATQ_GO(2, 1, 128);
  if (mid) ATQ_GO(1, 1, 128);
  ATQ_GO(1, 1, 256);
#undef ATQ_GO
}

template <int EPI>
static int dispatch(const atq_bf16_operand* a, const atq_bf16_operand* b, const GemmParams& p, cudaStream_t stream, int* grid_used) {
  const bool amn = a->mn_major != 0, bmn = b->mn_major != 0;
  if (!amn && !bmn) return dispatch_layout<EPI, LAYOUT_KK>(a, b, p, stream, grid_used);
  if (!amn && bmn) return dispatch_layout<EPI, LAYOUT_KM>(a, b, p, stream, grid_used);
  if (amn && bmn) return dispatch_layout<EPI, LAYOUT_MM>(a, b, p, stream, grid_used);
  set_error("tgemm: A MN-major with B K-major is not built");
  return ATQ_EINVAL;
}

static int check_operand(const atq_bf16_operand* o, const char* name) {
  if (o == nullptr || o->hi == nullptr) {
    set_error("tgemm: operand %s is null", name);
    return ATQ_EINVAL;
  }
  if ((o->pitch % 8) != 0 || (reinterpret_cast<uintptr_t>(o->hi) & 15u) || (o->lo && (reinterpret_cast<uintptr_t>(o->lo) & 15u))) {
    set_error("tgemm: operand %s needs pitch %% 8 == 0 and 16-byte aligned pointers", name);
    return ATQ_EINVAL;
  }
  if (o->format != 0 && o->format != 1) {
    set_error("tgemm: operand %s has unknown element format %d (0 = bf16, 1 = fp16)", name, (int)o->format);
    return ATQ_EINVAL;
  }
  return ATQ_OK;
}

static int64_t num_tiles_upper(int64_t rows, int64_t cols) {
  return ((rows + BLOCK_M - 1) / BLOCK_M) * ((cols + 63) / 64);
}

}  // namespace atq

using namespace atq;

extern "C" {

int atq_set_cta_pairs(int enabled) {
  const int old = (g_cta_pairs ? 1 : 0) | (g_wide_pairs ? 2 : 0) | (g_force_pairs ? 8 : 0);
  g_cta_pairs = (enabled & 1) != 0;
  g_wide_pairs = (enabled & 2) != 0 || enabled == 1;
  g_force_pairs = (enabled & 8) != 0;  // pairs also for problems smaller than one wave (tests of ragged shapes)
  return old;
}

size_t atq_workspace_bytes_tgemm(int64_t rows, int64_t cols) {
  return (size_t)(((num_tiles_upper(rows, cols) * sizeof(float)) + 255) & ~(size_t)255);
}

int atq_tgemm_absmax(int device, int64_t rows, int64_t cols, int64_t kdim, const atq_bf16_operand* a, const atq_bf16_operand* b,
                     const float* scale, const float* bias, float* out, int64_t out_pitch, float* out_scale_slot, float bound_mul,
                     atq_stream_t stream_) {
  ATQ_CHECK_ARG(rows > 0 && cols > 0 && kdim > 0 && out != nullptr && out_pitch >= cols, "bad shape or null output");
  ATQ_CHECK_ARG(out_scale_slot != nullptr && bound_mul > 0.f, "needs a scale slot and a positive bound_mul");
  int r;
  if ((r = check_operand(a, "a")) != ATQ_OK) return r;
  if ((r = check_operand(b, "b")) != ATQ_OK) return r;
  ATQ_CHECK_ARG(a->pitch >= (a->mn_major ? rows : kdim) && b->pitch >= (b->mn_major ? cols : kdim),
                "operand pitch smaller than its contiguous extent");
  ATQ_CHECK_ARG(a->format == b->format, "A and B operands must have the same element format (tcgen05 kind::f16)");
  ATQ_ENSURE_DEVICE(device);
  cudaStream_t stream = (cudaStream_t)stream_;
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.rows = rows; p.cols = cols; p.kdim = kdim;
  p.out = out; p.out_pitch = out_pitch;
  p.scale = scale; p.bias = bias;
  p.absmax_bits = reinterpret_cast<unsigned int*>(out_scale_slot);
  int grid = 0;
  if ((r = dispatch<EPI_LINEAR>(a, b, p, stream, &grid)) != ATQ_OK) return r;
  slot_finalize_kernel<<<1, 1, 0, stream>>>(out_scale_slot, bound_mul);
  ATQ_LAUNCH_CHECK();
  return ATQ_OK;
}

int atq_tgemm(int device, int64_t rows, int64_t cols, int64_t kdim, const atq_bf16_operand* a, const atq_bf16_operand* b,
              const float* scale, const float* bias, float* out, int64_t out_pitch, const float* dot_ref,
              int64_t dot_ref_pitch, float* dot_out, void* ws, size_t ws_bytes, atq_stream_t stream_) {
  ATQ_CHECK_ARG(rows > 0 && cols > 0 && kdim > 0 && out != nullptr && out_pitch >= cols, "bad shape or null output");
  int r;
  if ((r = check_operand(a, "a")) != ATQ_OK) return r;
  if ((r = check_operand(b, "b")) != ATQ_OK) return r;
  ATQ_CHECK_ARG(a->pitch >= (a->mn_major ? rows : kdim) && b->pitch >= (b->mn_major ? cols : kdim),
                "operand pitch smaller than its contiguous extent");
  ATQ_CHECK_ARG((dot_ref == nullptr) == (dot_out == nullptr), "dot_ref and dot_out go together");
  ATQ_CHECK_ARG(a->format == b->format, "A and B operands must have the same element format (tcgen05 kind::f16)");
  if (dot_out != nullptr && (ws == nullptr || ws_bytes < atq_workspace_bytes_tgemm(rows, cols))) {
    set_error("atq_tgemm: workspace too small");
    return ATQ_EWORKSPACE;
  }
  ATQ_ENSURE_DEVICE(device);
  cudaStream_t stream = (cudaStream_t)stream_;
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.rows = rows; p.cols = cols; p.kdim = kdim;
  p.out = out; p.out_pitch = out_pitch;
  p.scale = scale; p.bias = bias;
  p.dot_ref = dot_ref; p.dot_ref_pitch = dot_ref_pitch;
  p.partials = dot_out ? (float*)ws : nullptr;
  int grid = 0;
  if ((r = dispatch<EPI_LINEAR>(a, b, p, stream, &grid)) != ATQ_OK) return r;
  if (dot_out) {
    reduce_partials_kernel<<<1, 256, 0, stream>>>((const float*)ws, grid, dot_out);
    ATQ_LAUNCH_CHECK();
  }
  return ATQ_OK;
}

int atq_tgemm_packed(int device, int64_t rows, int64_t cols, int64_t kdim, const atq_bf16_operand* a,
                     const uint8_t* b_packed, const float* scale, const float* bias, float* out, int64_t out_pitch,
                     const float* dot_ref, int64_t dot_ref_pitch, float* dot_out, void* ws, size_t ws_bytes,
                     atq_stream_t stream_) {
  ATQ_CHECK_ARG(rows > 0 && cols > 0 && kdim > 0 && out != nullptr && out_pitch >= cols, "bad shape or null output");
  ATQ_CHECK_ARG(b_packed != nullptr && (reinterpret_cast<uintptr_t>(b_packed) & 15u) == 0, "packed B must be 16-byte aligned");
  ATQ_CHECK_ARG((kdim % 64) == 0, "packed-B GEMM needs kdim % 64 == 0 (16-byte codec rows per k-block)");
  int r;
  if ((r = check_operand(a, "a")) != ATQ_OK) return r;
  ATQ_CHECK_ARG(a->pitch >= kdim, "operand pitch smaller than kdim");
  ATQ_CHECK_ARG((dot_ref == nullptr) == (dot_out == nullptr), "dot_ref and dot_out go together");
  if (dot_out != nullptr && (ws == nullptr || ws_bytes < atq_workspace_bytes_tgemm(rows, cols))) {
    set_error("atq_tgemm_packed: workspace too small");
    return ATQ_EWORKSPACE;
  }
  ATQ_ENSURE_DEVICE(device);
  cudaStream_t stream = (cudaStream_t)stream_;
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.rows = rows; p.cols = cols; p.kdim = kdim;
  p.out = out; p.out_pitch = out_pitch;
  p.scale = scale; p.bias = bias;
  p.dot_ref = dot_ref; p.dot_ref_pitch = dot_ref_pitch;
  p.partials = dot_out ? (float*)ws : nullptr;
  p.b_packed = b_packed; p.b_packed_pitch = kdim / 4;
  int grid = 0;
  const bool a2 = a->lo != nullptr;
  if (a2) {
    r = launch_cfg<2, 1, 128, EPI_LINEAR, true>(a, nullptr, p, stream, &grid);  // two accumulators per tile
  } else if (cols <= 128) {
    r = launch_cfg<1, 1, 128, EPI_LINEAR, true>(a, nullptr, p, stream, &grid);
  } else {
    r = launch_cfg<1, 1, 256, EPI_LINEAR, true>(a, nullptr, p, stream, &grid);
  }
  if (r != ATQ_OK) return r;
  if (dot_out) {
    reduce_partials_kernel<<<1, 256, 0, stream>>>((const float*)ws, grid, dot_out);
    ATQ_LAUNCH_CHECK();
  }
  return ATQ_OK;
}

int atq_tgemm_fwd(int device, int64_t n_tokens, int64_t out_features, int64_t in_features, const atq_bf16_operand* x,
                  const atq_bf16_operand* w, const float* alpha, const float* bias, float* y, int64_t y_pitch, void* ws,
                  size_t ws_bytes, atq_stream_t stream) {
  return atq_tgemm(device, n_tokens, out_features, in_features, x, w, alpha, bias, y, y_pitch, nullptr, 0, nullptr, ws,
                   ws_bytes, stream);
}

int atq_tgemm_dx(int device, int64_t n_tokens, int64_t in_features, int64_t out_features, const atq_bf16_operand* dy,
                 const atq_bf16_operand* w_t, const float* alpha, float* dx, int64_t dx_pitch, const float* x_ref,
                 int64_t x_pitch, float* dalpha_out, void* ws, size_t ws_bytes, atq_stream_t stream) {
  return atq_tgemm(device, n_tokens, in_features, out_features, dy, w_t, alpha, nullptr, dx, dx_pitch, x_ref, x_pitch,
                   dalpha_out, ws, ws_bytes, stream);
}

// split-K plan for the dW GEMM: few output tiles, very long contraction (tokens)
static int dw_splits(int device, int64_t out_features, int64_t in_features, int64_t n_tokens) {
  const int64_t tiles = ((in_features + 127) / 128) * ((out_features + BLOCK_M - 1) / BLOCK_M);
  const int64_t num_kb = (n_tokens + BLOCK_K - 1) / BLOCK_K;
  const int sms = sm_count(device);
  if (tiles * 2 > sms || num_kb < 32) return 1;
  int64_t s = sms / tiles;
  if (s > 8) s = 8;
  if (s > num_kb / 8) s = num_kb / 8;
  return (int)(s < 1 ? 1 : s);
}

size_t atq_workspace_bytes_tgemm_dw(int64_t out_features, int64_t in_features, int64_t n_tokens) {
  // CTA partials (d(alpha)) + up to 8 split-K slabs of the [out, in] gradient
  const int64_t tiles = ((in_features + 127) / 128) * ((out_features + BLOCK_M - 1) / BLOCK_M);
  const int64_t num_kb = (n_tokens + BLOCK_K - 1) / BLOCK_K;
  size_t slabs = (tiles * 2 > 256 || num_kb < 32) ? 0 : (size_t)8 * (size_t)out_features * (size_t)in_features * sizeof(float);
  return atq_workspace_bytes_tgemm(out_features, in_features) + 4096 + ((slabs + 255) & ~(size_t)255);
}

int atq_tgemm_dw_masked(int device, int64_t out_features, int64_t in_features, int64_t n_tokens,
                        const atq_bf16_operand* dy_t, const atq_bf16_operand* x_t, const float* mask,
                        const uint8_t* packed, float* dw, int64_t dw_pitch, float* dalpha_out, void* ws,
                        size_t ws_bytes, atq_stream_t stream_) {
  ATQ_CHECK_ARG(out_features > 0 && in_features > 0 && n_tokens > 0 && dw != nullptr && dw_pitch >= in_features,
                "bad shape or null output");
  int r;
  if ((r = check_operand(dy_t, "dy_t")) != ATQ_OK) return r;
  if ((r = check_operand(x_t, "x_t")) != ATQ_OK) return r;
  ATQ_CHECK_ARG(dy_t->pitch >= (dy_t->mn_major ? out_features : n_tokens) && x_t->pitch >= (x_t->mn_major ? in_features : n_tokens),
                "operand pitch smaller than its contiguous extent");
  ATQ_CHECK_ARG((packed == nullptr) == (dalpha_out == nullptr), "packed and dalpha_out go together");
  ATQ_CHECK_ARG(dy_t->format == x_t->format, "dY and X operands must have the same element format");
  ATQ_ENSURE_DEVICE(device);
  cudaStream_t stream = (cudaStream_t)stream_;
  const size_t part_bytes = atq_workspace_bytes_tgemm(out_features, in_features) + 4096;
  int splits = dw_splits(device, out_features, in_features, n_tokens);
  const size_t slab_bytes = (size_t)out_features * (size_t)in_features * sizeof(float);
  if (splits > 1 && (ws == nullptr || ws_bytes < part_bytes + (size_t)splits * slab_bytes)) splits = 1;  // caller sized for no split
  if (dalpha_out != nullptr && (ws == nullptr || ws_bytes < part_bytes)) {
    set_error("atq_tgemm_dw_masked: workspace too small");
    return ATQ_EWORKSPACE;
  }
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.rows = out_features; p.cols = in_features; p.kdim = n_tokens;
  int grid = 0;
  if (splits <= 1) {
    p.out = dw; p.out_pitch = dw_pitch;
    p.mask = mask; p.tern = packed;
    p.partials = dalpha_out ? (float*)ws : nullptr;
    if ((r = dispatch<EPI_MASKED>(dy_t, x_t, p, stream, &grid)) != ATQ_OK) return r;
  } else {
    // split-K: every (tile, token-range) work item writes its raw partial product to its slab; the
    // finalize kernel adds the slabs in fixed order and applies the mask / d(alpha) epilogue
    float* slabs = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + part_bytes);
    p.out = slabs; p.out_pitch = in_features;
    p.splits = splits; p.split_stride = out_features * in_features;
    if ((r = dispatch<EPI_LINEAR>(dy_t, x_t, p, stream, &grid)) != ATQ_OK) return r;
    const int64_t n = out_features * in_features;
    int64_t fg = (n + 256 * 8 - 1) / (256 * 8);
    const int64_t cap = (int64_t)sm_count(device) * 4;
    if (fg > cap) fg = cap;
    if (fg < 1) fg = 1;
    grid = (int)fg;
    splitk_finalize_masked_kernel<<<grid, 256, 0, stream>>>(slabs, splits, p.split_stride, out_features, in_features, mask, packed,
                                                             dw, dw_pitch, dalpha_out ? (float*)ws : nullptr);
    ATQ_LAUNCH_CHECK();
  }
  if (dalpha_out) {
    reduce_partials_kernel<<<1, 256, 0, stream>>>((const float*)ws, grid, dalpha_out);
    ATQ_LAUNCH_CHECK();
  }
  return ATQ_OK;
}

}  // extern "C"
