// bt709_decode_row.cuh — the per-row decode arithmetic shared by the pass-1 kernels (bt709_decode.cu) and the fused
// decode -> rescale kernel (bt709_fused.cu): packed constants, the paired cell-LUT lookup and decode_row.
#pragma once
#include "bt709_device.cuh"

namespace bt709 {

// loop-invariant packed constants {k, k}
struct DecodeK2 {
  f32x2 kY, kC, inv255, ys, g_cb, b_cb, r_cr, g_cr, fN, magic, one;
  f32x2 bias;  // CL2 < 0 only (biased log cells)
};
// two LUT lookups at once: clamp (scalar), index (packed FFMA2.RZ), two LDS, two adds.
// returns the two sums whose byte 3 is the 8-bit code
template <int CL2, int RL2>
__device__ __forceinline__ void lut2(float xa, float xb, float clamp_lo, const DecodeK2 &k, uint32_t sbase,
                                     uint32_t &sa, uint32_t &sb) {
  float ca = fmaxf(xa, clamp_lo), cb = fmaxf(xb, clamp_lo);
  uint32_t ta, tb;
  if (CL2 <= 0) {  // log-spaced cells: the index is the upper half of the float (sign, exponent, 7 mantissa bits)
    float ua = ca, ub = cb;
    if (CL2 < 0) upk(add2(pk(ca, cb), k.bias), ua, ub);  // biased log cells: the cell of x + bias (one packed add per pair)
    // PRMT, not a shift: ptxas merges (x >> 16) << 7 into SHF + LOP3 + IADD, this stays PRMT + LEA
    ta = __byte_perm(__float_as_uint(ua), 0, 0x4432);
    tb = __byte_perm(__float_as_uint(ub), 0, 0x4432);
  } else {
    upku(fma2_rz(pk(ca, cb), k.fN, k.magic), ta, tb);
  }
  uint32_t ea, eb;
  asm("ld.shared.u32 %0, [%1];" : "=r"(ea) : "r"((ta << (2 + RL2)) + sbase));
  asm("ld.shared.u32 %0, [%1];" : "=r"(eb) : "r"((tb << (2 + RL2)) + sbase));
  sa = ea + __float_as_uint(ca);
  sb = eb + __float_as_uint(cb);
}

template <int CL2, int RL2, bool ALPHA>
__device__ __forceinline__ uint4 decode_row(uint32_t yw, uint32_t aw, f32x2 cb_g, f32x2 cb_b, f32x2 cr_r, f32x2 cr_g,
                                            const DecodeK2 &k, float clamp_lo, uint32_t sbase) {
  // luma terms, BT709.h:494 then :424 — pairs (px0, px2) and (px1, px3)
  f32x2 y02 = mul2(mul2(add2(pku(magic_byte<0>(yw), magic_byte<2>(yw)), k.kY), k.inv255), k.ys);
  f32x2 y13 = mul2(mul2(add2(pku(magic_byte<1>(yw), magic_byte<3>(yw)), k.kY), k.inv255), k.ys);
  // (Y + Cb) first, then + Cr (BT709.h:425).  ptxas contracts a packed mul.rn.f32x2 feeding an
  // add.rn.f32x2 into one FFMA2 (a single rounding: not bit-exact), so the packed add is written as
  // fma(y, one, cb_g) with `one` = 1.0f read from the kernel arguments: y*1 is exact, the sum is
  // rounded once exactly like an add, and neither multiply can be folded into it.
  f32x2 g02 = fma2_rn(y02, k.one, cb_g), g13 = fma2_rn(y13, k.one, cb_g);
  float y0, y1, y2, y3, g0, g1, g2, g3, crA, crB, cgA, cgB, cbA, cbB;
  upk(y02, y0, y2); upk(y13, y1, y3);
  upk(g02, g0, g2); upk(g13, g1, g3);
  upk(cr_r, crA, crB); upk(cr_g, cgA, cgB); upk(cb_b, cbA, cbB);
  uint32_t r0, r1, r2, r3, gg0, gg1, gg2, gg3, b0, b1, b2, b3;
  lut2<CL2, RL2>(add_sat(y0, crA), add_sat(y2, crB), clamp_lo, k, sbase, r0, r2);
  lut2<CL2, RL2>(add_sat(y1, crA), add_sat(y3, crB), clamp_lo, k, sbase, r1, r3);
  lut2<CL2, RL2>(add_sat(g0, cgA), add_sat(g2, cgB), clamp_lo, k, sbase, gg0, gg2);
  lut2<CL2, RL2>(add_sat(g1, cgA), add_sat(g3, cgB), clamp_lo, k, sbase, gg1, gg3);
  lut2<CL2, RL2>(add_sat(y0, cbA), add_sat(y2, cbB), clamp_lo, k, sbase, b0, b2);
  lut2<CL2, RL2>(add_sat(y1, cbA), add_sat(y3, cbB), clamp_lo, k, sbase, b1, b3);
  uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  if (ALPHA) {
    // AAPLShaders.metal:249-271 via the CPU twin BT709.h:466-513 with neutral chroma
    f32x2 a02 = mul2(mul2(add2(pku(magic_byte<0>(aw), magic_byte<2>(aw)), k.kY), k.inv255), k.ys);
    f32x2 a13 = mul2(mul2(add2(pku(magic_byte<1>(aw), magic_byte<3>(aw)), k.kY), k.inv255), k.ys);
    float f0, f1, f2, f3;
    upk(a02, f0, f2); upk(a13, f1, f3);
    a0 = (uint32_t) round_pos(__fmul_rn(add_sat(f0, 0.0f), 255.0f));
    a1 = (uint32_t) round_pos(__fmul_rn(add_sat(f1, 0.0f), 255.0f));
    a2 = (uint32_t) round_pos(__fmul_rn(add_sat(f2, 0.0f), 255.0f));
    a3 = (uint32_t) round_pos(__fmul_rn(add_sat(f3, 0.0f), 255.0f));
  }
  auto pack = [](uint32_t sb, uint32_t sg, uint32_t sr, uint32_t a) -> uint32_t {
    uint32_t bg = __byte_perm(sb, sg, 0x0073);  // byte0 = B code, byte1 = G code
    if (ALPHA) return __byte_perm(bg, __byte_perm(sr, a, 0x0043), 0x5410);
    return __byte_perm(bg, sr, 0x7710) | 0xFF000000u;  // A = 255 (AAPLShaders.metal:243)
  };
  return make_uint4(pack(b0, gg0, r0, a0), pack(b1, gg1, r1, a1), pack(b2, gg2, r2, a2), pack(b3, gg3, r3, a3));
}

}  // namespace bt709
