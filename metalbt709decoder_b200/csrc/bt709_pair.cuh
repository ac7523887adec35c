// bt709_pair.cuh — the paired-column pass-2 walker (tables + one warp tile), shared by k_rescale_pair
// (bt709_rescale.cu) and the fused decode -> rescale kernel (bt709_fused.cu).
#pragma once
#include <type_traits>

#include "bt709_device.cuh"

namespace bt709 {

constexpr int kWalkTH = 64;       // output rows per warp tile
constexpr int kWalkWarps = 32;    // warps per CTA, one CTA per SM
constexpr int kWalkPD = 3;        // source rows fetched ahead of use
constexpr int kWalkRowInfo = (kWalkTH + 1) * 4;  // words of row info per warp (one sentinel entry)

// -------------------------------------------------------------- rescale, paired columns (round 2)
// The two-column walker again, rebuilt around three measurements of k_rescale_cols2 (profiles/r01_rescale_cols_*):
// 62 instructions per output pixel with issue slots 80 % busy, and 12 shared-memory wavefronts per 32 output pixels of
// which the encode gathers are ~8 (16 LUT replicas: two lanes per replica collide on random content).
//   * the output of pass 1 is opaque (A = 255, AAPLShaders.metal:243) unless an alpha stream was decoded, and a
//     blend of 1.0 with 1.0 is 1.0 in every float32 step of the oracle: the alpha channel is dropped (no table
//     gathers, no blends, no rounding) on a SPECULATIVE basis - every tap word is AND-ed into an accumulator, and if
//     a warp finds a tap with A != 255 it redoes its tile with the four-channel code and starts the next tiles
//     there (two LOP3 per source row is all the opaque case pays);
//   * the packed pairs are (column a, column b) of ONE channel instead of (R, G) / (B, A) of one column: the table
//     gathers land directly in pair registers, the horizontal blend takes the per-column fractions as one packed
//     operand, and column b's taps are chosen by selecting the two texel WORDS (2 SEL) instead of four blended
//     pairs (8 SEL);
//   * the sRGB encode uses the biased log-cell LUT (bt709_tables.h): 645 cells, ONE REPLICA PER LANE (82 KB): a
//     gather is a single conflict-free wavefront for any content, for one extra packed add per pair of lookups;
//   * the alpha byte -> float step is arithmetic (PRMT + packed add + packed multiply), not a table.
// Same row-info slices, unclamped row walk and three-row register prefetch queue as k_rescale_cols2.
// k_rescale_pair: 20 warps x 96 registers, taps fetched 2 rows ahead.  Measured on 1080p -> 1440p (k frames/s) after the
// inner-tile walker, the 9-conversion blend and the CTA-by-CTA tile order: 20 x 96 / 2 rows 158.5, 16 x 128 / 2 rows 155.9,
// 16 x 128 / 3 rows 155.3, 16 x 128 / 4 rows 152.6; before them 16 x 128 / 3 rows was the best (150.8 against 143.4 for
// 20 x 96 / 2 rows and 126 for 24 x 80).  The fused kernel also runs 20 warps, 2 rows ahead.
#ifndef BT709_PAIR_WARPS
#define BT709_PAIR_WARPS 20
#endif
#ifndef BT709_PAIR_PD
#define BT709_PAIR_PD 2
#endif
#ifndef BT709_PAIR_NP
#define BT709_PAIR_NP 2
#endif
#ifndef BT709_PAIR_TH
#define BT709_PAIR_TH 64
#endif
constexpr int kPairWarps = BT709_PAIR_WARPS;
constexpr int kPairTH = BT709_PAIR_TH;        // output rows per warp tile of k_rescale_pair (the fused kernel's bands stay kWalkTH)
constexpr int kPairRowInfo = (kPairTH + 1) * 4;
constexpr int kPairPD = BT709_PAIR_PD;        // source rows fetched ahead of use by pair_tile (standalone kernel)
constexpr int kPairNP = BT709_PAIR_NP;        // column pairs per thread on the opaque code: warp tile = 64 * kPairNP columns
constexpr int kFusedWarps = 20, kFusedPD = 2;
#ifdef BT709_NO_INNER
constexpr bool kPairInner = false;
#else
constexpr bool kPairInner = true;  // pair_tile: separate code for warp tiles without clamped taps (see `fetch`)
#endif

struct PairCtx {
  uint32_t lin_lane, rows_base, enc_base;
  uint32_t extra;  // shared address of the caller's extra region (pair_tables: extra_bytes), 128-byte aligned
  float clamp_lo;
  f32x2 one, neg1, bias2, inv255_2, nmagic2;
};

struct H3 {
  f32x2 r, g, b, a;  // (column a, column b) per channel, linear light; a only when ALPHA
};

// Layout (dynamic shared memory):
//   [encode LUT, biased log cells: cell i replica r at word i*32 + r][pad][lin table at the next 64 KB boundary:
//    word b*64 + r = sRGB byte b -> linear float, replica r; the upper 128 B of each 256-B entry are unused][row info]
// extra_bytes: room the caller wants between the encode LUT and the pad (the fused kernel's decode LUT)
__device__ __forceinline__ PairCtx pair_tables(const RescaleArgs &p, uint32_t *s_mem, uint32_t extra_bytes = 0,
                                               int row_words = kWalkRowInfo) {  // row info: one slice of row_words per warp of the CTA
  const uint32_t smem0 = (uint32_t) __cvta_generic_to_shared(s_mem);
  const int nenc = p.enc_blog_cells << 5;
  const uint32_t extra_at = ((uint32_t) (nenc + 4) * 4u + 127u) & ~127u;
  const uint32_t head = extra_at + extra_bytes;
  const uint32_t pad = head + ((0x10000u - ((smem0 + head) & 0xFFFFu)) & 0xFFFFu);
  float *s_lin = reinterpret_cast<float *>(s_mem + (pad >> 2));
  uint32_t *s_rows = s_mem + (pad >> 2) + 256 * 64;
  uint32_t *s_enc = s_mem + 4;
  fill_replicated<5>(reinterpret_cast<uint32_t *>(s_lin), 64, reinterpret_cast<const uint32_t *>(p.lin256), 256);
  fill_replicated<5>(s_enc, 32, p.enc_blog_lut, p.enc_blog_cells);
  // (index register << 7) + enc_base, the index register holding first + cell: the bias goes through shared memory so
  // that ptxas sees an opaque register and emits one LEA per lookup
  if (threadIdx.x == 0) s_enc[-1] = (uint32_t) __cvta_generic_to_shared(s_enc) - (p.enc_blog_first << 7);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  PairCtx c;
  c.lin_lane = (smem0 + pad) | (lane << 2);
  c.rows_base = (uint32_t) __cvta_generic_to_shared(s_rows + warp * row_words);
  c.enc_base = *reinterpret_cast<volatile uint32_t *>(s_enc - 1) + (lane << 2);
  c.extra = smem0 + extra_at;
  c.clamp_lo = p.enc_blog_clamp_lo;
  c.one = pk(p.one, p.one);
  c.neg1 = pk(-1.0f, -1.0f);
  c.bias2 = pk(p.enc_blog_bias, p.enc_blog_bias);
  c.inv255_2 = pk(1.0f / 255.0f, 1.0f / 255.0f);
  c.nmagic2 = pk(-8388608.0f, -8388608.0f);
  return c;
}

// One warp tile (64 * NP output columns x th output rows).  Returns true when every tap word it fetched had A = 255.
// src_frame / dst_frame: row 0 of the source and destination frame; pi: source pitch in bytes; oy0: first output row;
// ox0: first output column of the tile (64 * NP columns wide).
// NC: the source is read-only for the whole kernel (ld.global.nc); the fused kernel reads rows its own CTA has just
// written and must use plain loads (coherent within the SM after the CTA barrier).
// NP: column pairs per thread, 64 columns apart (pair q of lane l = columns ox0 + 64 q + 2 l, + 1).  A third of the
// instruction stream of the one-pair walker is per-row bookkeeping (row info, pointers, clamps, loop control); with two
// pairs per thread it is paid once per four pixels.
// SPEC: the opaque code is speculative (track the AND of the tap words); false when the caller knows the source is
// opaque (the fused kernel: pass 1 without an alpha stream writes A = 255).
// PD: source rows fetched ahead of use.  l2_policy (NC == false only): createpolicy handle for the tap loads (the fused
// kernel keeps its slice resident in L2 with evict_last).
template <bool ALPHA, bool NC = true, int NP = 1, bool SPEC = true, int PD = kPairPD, bool INNER_PATH = kPairInner>
__device__ __forceinline__ bool pair_tile(const RescaleArgs &p, const PairCtx &c, const uint8_t *src_frame, unsigned pi,
                                          uint8_t *dst_frame, int ox0, int oy0, int lane, unsigned long long l2_policy = 0,
                                          int th = kWalkTH) {
  uint32_t lin_lane = c.lin_lane, enc_base = c.enc_base;
  const uint32_t rows_base = c.rows_base;
  asm volatile("" : "+r"(lin_lane), "+r"(enc_base));  // keep them in registers (ptxas rebuilt them every source row)
  const float clamp_lo = c.clamp_lo;
  const f32x2 one = c.one, neg1 = c.neg1;
  auto lin_of = [&](uint32_t px, uint32_t sel) -> float {
    float v;
    asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(__byte_perm(px, lin_lane, sel)));  // table | byte << 8 | lane << 2
    return v;
  };
  auto lerp = [&](f32x2 t0, f32x2 t1, f32x2 f2) -> f32x2 {  // t0 + f * (t1 - t0): the oracle's three rounded steps
    return fma2_rn(mul2(f2, fma2_rn(t0, neg1, t1)), one, t0);
  };
  // two clamped values -> the two sums whose byte 3 is the sRGB code (biased log-cell LUT, one replica per lane)
  auto enc2 = [&](f32x2 v, uint32_t &sa, uint32_t &sb) {
    float va, vb, ua, ub;
    upk(v, va, vb);
    va = fmaxf(va, clamp_lo);
    vb = fmaxf(vb, clamp_lo);
    upk(add2(pk(va, vb), c.bias2), ua, ub);
    uint32_t ea, eb;
    asm("ld.shared.u32 %0, [%1];" : "=r"(ea) : "r"((__byte_perm(__float_as_uint(ua), 0, 0x4432) << 7) + enc_base));
    asm("ld.shared.u32 %0, [%1];" : "=r"(eb) : "r"((__byte_perm(__float_as_uint(ub), 0, 0x4432) << 7) + enc_base));
    sa = ea + __float_as_uint(va);
    sb = eb + __float_as_uint(vb);
  };

  __syncwarp();
  // th: output rows of this tile (<= the row-info slice the caller sized, a run-time value: small calls use
  // shorter tiles so that every warp of the grid has one)
  for (int r = lane; r < th; r += 32) {
    const int oy = oy0 + r;
    const float sy = __fadd_rn(__fmul_rn(__fadd_rn((float) oy, 0.5f), p.scale_y), -0.5f);
    const float fy0 = floorf(sy);
    asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(rows_base + r * 16), "r"(__float_as_uint(__fsub_rn(sy, fy0))),
                 "r"(oy < p.oh ? (uint32_t) (int) fy0 : 0x7fffffffu) : "memory");
  }
  if (lane == 0) asm volatile("st.shared.v2.u32 [%0], {%1,%1};" ::"r"(rows_base + th * 16), "r"(0x7fffffffu) : "memory");
  __syncwarp();

  const int ylast = p.ih - 1, xlast = p.iw - 1;
  const unsigned po = (unsigned) p.po;
  // ---- per pair: the two columns, their fractions, the three (clamped) source texels they tap
  bool bsel[NP], live[NP];  // live: the pair lies inside the image (ow is even on this path: a pair is never half in)
  f32x2 fx2[NP];
  unsigned d1[NP], d2[NP];
  const uint8_t *fp[NP];
#pragma unroll
  for (int q = 0; q < NP; q++) {
    const int oxa = ox0 + q * 64 + lane * 2;
    const float sxa = __fadd_rn(__fmul_rn(__fadd_rn((float) oxa, 0.5f), p.scale_x), -0.5f);
    const float sxb = __fadd_rn(__fmul_rn(__fadd_rn((float) (oxa + 1), 0.5f), p.scale_x), -0.5f);
    const float fxa0 = floorf(sxa), fxb0 = floorf(sxb);
    const int xia = (int) fxa0, xib = (int) fxb0;  // xib - xia is 0 or 1 for scales <= 1
    bsel[q] = xib != xia;                          // column b uses (T1, T2) instead of (T0, T1)
    fx2[q] = pk(__fsub_rn(sxa, fxa0), __fsub_rn(sxb, fxb0));
    const int p0 = min(max(xia, 0), xlast), p1 = min(max(xia + 1, 0), xlast), p2 = min(max(xia + 2, 0), xlast);
    d1[q] = (unsigned) (p1 - p0) * 4u;
    d2[q] = (unsigned) (p2 - p0) * 4u;
    fp[q] = src_frame + (size_t) p0 * 4;
    live[q] = oxa < p.ow;
    asm volatile("" : "+r"(d1[q]), "+r"(d2[q]));  // stay in registers (no rematerialisation)
  }
  // this thread's first pair in the row above the tile; stepped down before each row.  Pair q sits 256 q bytes to
  // the right (an immediate offset).  The launcher guarantees an even ow and 8-byte aligned rows: one 8-byte store
  // per pair, predicated by whether the pair is inside the image.
  uint8_t *dp = dst_frame + (long long) (oy0 - 1) * (long long) po + (size_t) (ox0 + lane * 2) * 4;

  uint32_t acc = 0xFFFFFFFFu;  // AND of every tap word: byte 3 stays 255 while the tile is opaque
  struct Taps { uint32_t w0, w1, w2; };
  // INNER (a compile-time flag, see below): no tap of this warp tile is clamped at the left or right image edge, so
  // the three texels of every pair are adjacent words: immediate offsets instead of two 64-bit adds per load
  auto fetch = [&](int row, Taps (&t)[NP], auto inner) {
    constexpr bool INNER = decltype(inner)::value;
    const unsigned step = (unsigned) row < (unsigned) ylast ? pi : 0u;
#pragma unroll
    for (int q = 0; q < NP; q++) {
      const uint8_t *f1 = INNER ? fp[q] + 4 : fp[q] + d1[q], *f2 = INNER ? fp[q] + 8 : fp[q] + d2[q];
      if (NC) {
        t[q].w0 = __ldg(reinterpret_cast<const uint32_t *>(fp[q]));
        t[q].w1 = __ldg(reinterpret_cast<const uint32_t *>(f1));
        t[q].w2 = __ldg(reinterpret_cast<const uint32_t *>(f2));
      } else {
        asm volatile("ld.global.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(t[q].w0) : "l"(fp[q]), "l"(l2_policy) : "memory");
        asm volatile("ld.global.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(t[q].w1) : "l"(f1), "l"(l2_policy) : "memory");
        asm volatile("ld.global.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(t[q].w2) : "l"(f2), "l"(l2_policy) : "memory");
      }
      fp[q] += step;
    }
  };
  auto blend = [&](const Taps &t, int q) -> H3 {
    const uint32_t w0 = t.w0, w1 = t.w1, w2 = t.w2;
    const uint32_t u0 = bsel[q] ? w1 : w0, u1 = bsel[q] ? w2 : w1;  // column b's taps (alpha; colour with BT709_PAIR_TAPS12)
    if (SPEC) {
      acc = acc & w0 & w1;
      acc &= w2;
    }
    H3 h;
#ifdef BT709_PAIR_TAPS12
    h.r = lerp(pk(lin_of(w0, 0x7624), lin_of(u0, 0x7624)), pk(lin_of(w1, 0x7624), lin_of(u1, 0x7624)), fx2[q]);
    h.g = lerp(pk(lin_of(w0, 0x7614), lin_of(u0, 0x7614)), pk(lin_of(w1, 0x7614), lin_of(u1, 0x7614)), fx2[q]);
    h.b = lerp(pk(lin_of(w0, 0x7604), lin_of(u0, 0x7604)), pk(lin_of(w1, 0x7604), lin_of(u1, 0x7604)), fx2[q]);
#else
    // each of the three texels is converted once (9 table reads per pair and row instead of 12: the shared-memory pipe
    // is as busy as the issue slots in this kernel), column b's taps are selected as floats
    auto chan = [&](uint32_t sel) -> f32x2 {
      const float l0 = lin_of(w0, sel), l1 = lin_of(w1, sel), l2 = lin_of(w2, sel);
      return lerp(pk(l0, bsel[q] ? l1 : l0), pk(l1, bsel[q] ? l2 : l1), fx2[q]);
    };
    h.r = chan(0x7624);
    h.g = chan(0x7614);
    h.b = chan(0x7604);
#endif
    if (ALPHA) {  // alpha is linear: byte * (1/255), exactly as the oracle's tap
      const f32x2 a0 = mul2(add2(pku(magic_byte<3>(w0), magic_byte<3>(u0)), c.nmagic2), c.inv255_2);
      const f32x2 a1 = mul2(add2(pku(magic_byte<3>(w1), magic_byte<3>(u1)), c.nmagic2), c.inv255_2);
      h.a = lerp(a0, a1, fx2[q]);
    }
    return h;
  };

  bool inner_ok = true;
#pragma unroll
  for (int q = 0; q < NP; q++) inner_ok = inner_ok && d1[q] == 4u && d2[q] == 8u;
  inner_ok = __all_sync(0xffffffffu, inner_ok);  // warp-uniform: one of the two instantiations below runs

  auto walk = [&](auto inner) {
  uint32_t ip = rows_base;
  uint32_t info_fy, info_y;
  asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(info_fy), "=r"(info_y) : "r"(ip));
  int y = (int) info_y;
  Taps qq[PD][NP], first[NP];
#pragma unroll
  for (int q = 0; q < NP; q++) fp[q] = ptr_mad(fp[q], (unsigned) max(y, 0), pi);
  fetch(y, first, inner);
#pragma unroll
  for (int j = 0; j < PD; j++) fetch(y + 1 + j, qq[j], inner);
  H3 H[2][NP];
#pragma unroll
  for (int q = 0; q < NP; q++) H[0][q] = blend(first[q], q);
  bool done = false;
  while (!done) {
#pragma unroll
    for (int j = 0; j < 2 * PD; j++) {
#pragma unroll
      for (int q = 0; q < NP; q++) H[(j + 1) & 1][q] = blend(qq[j % PD][q], q);
      fetch(y + 1 + PD, qq[j % PD], inner);
      while ((int) info_y <= y) {
        dp += po;
        const float fy = __uint_as_float(info_fy);
        const f32x2 fy2 = pk(fy, fy);
#pragma unroll
        for (int q = 0; q < NP; q++) {
          const H3 &A = H[j & 1][q], &B = H[(j + 1) & 1][q];
          uint32_t ra, rb, ga, gb, ba, bb;
          enc2(lerp(A.r, B.r, fy2), ra, rb);
          enc2(lerp(A.g, B.g, fy2), ga, gb);
          enc2(lerp(A.b, B.b, fy2), ba, bb);
          uint32_t wa, wb;
          if (ALPHA) {
            // round(255 * v), half away from zero: the integer part of RZ(x + 32768.5) sits in byte 1
            // (the product goes through an FMA with +0 so that no multiply feeds the round-toward-zero add directly)
            uint32_t ta, tb;
            upku(add2_rz_pair(fma2_rn(lerp(A.a, B.a, fy2), pk(255.0f, 255.0f), pk(0.0f, 0.0f))), ta, tb);
            wa = __byte_perm(__byte_perm(ba, ga, 0x0073), __byte_perm(ra, ta, 0x0053), 0x5410);
            wb = __byte_perm(__byte_perm(bb, gb, 0x0073), __byte_perm(rb, tb, 0x0053), 0x5410);
          } else {
            wa = __byte_perm(__byte_perm(ba, ga, 0x0073), ra, 0x7710) | 0xFF000000u;
            wb = __byte_perm(__byte_perm(bb, gb, 0x0073), rb, 0x7710) | 0xFF000000u;
          }
          // (an inner tile has no pair beyond the right edge: such a pair's taps would all be clamped)
          if (decltype(inner)::value || live[q]) st_stream_u64(dp + q * 256, make_uint2(wa, wb));  // streaming: never read back
        }
        ip += 16;
        asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(info_fy), "=r"(info_y) : "r"(ip));
      }
      y++;
      if (info_y == 0x7fffffffu) { done = true; break; }
    }
  }
  };
  if (INNER_PATH && inner_ok) walk(std::true_type{});
  else walk(std::false_type{});
  return __all_sync(0xffffffffu, (acc >> 24) == 0xFFu);
}

}  // namespace bt709
