// bt709_kernels.cu — hand-written sm_100a kernels for the BT.709 4:2:0 <-> sRGB BGRA8 path.
//
// Replaces AAPLShaders.metal (both passes) and the CPU loops of CVPixelBufferUtils.h:241-399
// and BGRAToBT709Converter.m:89-198, 1042-1099.  All arithmetic that feeds a quantiser is done
// with explicitly rounded float32 operations (__fmul_rn/__fadd_rn: never contracted to FMA) in
// the order the reference's CPU code uses (BT709.h:389-397, 424-426, 494-500, 223-247), and every
// transfer curve is an exact table lookup (bt709_tables.h), so results are bit-identical to the
// reference's CPU path, not merely within +-1 LSB.
//
// No tensor cores: the path is a per-pixel map with 5.5 B/pixel of HBM traffic and no
// contraction.  What matters is (a) 16-byte coalesced stores / 4..16-byte coalesced loads with
// streaming cache hints, (b) enough loads in flight per SM, (c) keeping the per-pixel
// instruction count under the issue budget (about 37 warp-instructions per 32 pixels at 80 % of
// the HBM roofline) and (d) keeping the three curve lookups per pixel off the MUFU pipe and
// free of shared-memory bank conflicts (per-lane replicated LUT).
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>

#include "bt709_kernels.cuh"

namespace bt709 {

// ------------------------------------------------------------------ small device helpers
__device__ __forceinline__ float add_sat(float a, float b) {
  float r;  // clamps to [0,1] in the adder: saturatef of BT709.h:444-446 for free
  asm("add.rn.sat.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}

// byte k of `word` as an exact float minus `bias` (bias = 16 for Y/alpha, 128 for chroma):
// PRMT builds the float 2^23 + byte, one FADD removes 2^23 + bias.  No I2F (XU pipe).
template <int K>
__device__ __forceinline__ float byte_minus(uint32_t word, float bias_plus_2p23) {
  uint32_t bits = __byte_perm(word, 0x4B000000u, 0x7540 + K);
  return __fadd_rn(__uint_as_float(bits), -bias_plus_2p23);
}

__device__ __forceinline__ uint32_t ld_stream_u32(const void *p) {
  uint32_t v;
  asm volatile("ld.global.cs.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ uint4 ld_stream_u128(const void *p) {
  uint4 v;
  asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_stream_u128(void *p, uint4 v) {
  asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_stream_u32(void *p, uint32_t v) {
  asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// base + a * b with 32-bit unsigned a, b: one IMAD.WIDE.U32 (the compiler's own 64-bit address
// arithmetic spent ~7 instructions per pixel on sign extensions and carry chains)
__device__ __forceinline__ const uint8_t *ptr_mad(const uint8_t *base, unsigned a, unsigned b) {
  unsigned long long r;
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"((unsigned long long) base));
  return (const uint8_t *) r;
}

// round-half-away-from-zero of a non-negative float, exactly: the +0.5 is done with
// round-toward-zero so it can never be rounded up across an integer.
__device__ __forceinline__ int round_pos(float r) {
  float t;
  asm("add.rz.f32 %0, %1, 0f3F000000;" : "=f"(t) : "f"(r));
  return __float2int_rz(t);
}

// ---- cell LUT lookups (bt709_tables.cpp: build_cell_lut) --------------------------------
// shared, per-lane replicated: entry i of replica r lives at word (i << RL2) + r
template <int CL2, int RL2>
__device__ __forceinline__ uint32_t lut_sum_smem(float x, float clamp_lo, uint32_t sbase) {
  float xc = fmaxf(x, clamp_lo);
  float t;
  asm("fma.rz.f32 %0, %1, %2, 0f4B000000;" : "=f"(t) : "f"(xc), "f"((float) (1 << CL2)));
  uint32_t addr = (__float_as_uint(t) << (2 + RL2)) + sbase;
  uint32_t e;
  asm("ld.shared.u32 %0, [%1];" : "=r"(e) : "r"(addr));
  return e + __float_as_uint(xc);  // byte 3 is the 8-bit code
}

// generic: table pointer (global through the read-only path, or shared), runtime size
__device__ __forceinline__ uint32_t lut_code(const uint32_t *lut, float x, float clamp_lo, float fN) {
  float xc = fmaxf(x, clamp_lo);
  float t;
  asm("fma.rz.f32 %0, %1, %2, 0f4B000000;" : "=f"(t) : "f"(xc), "f"(fN));
  uint32_t idx = __float_as_uint(t) - 0x4B000000u;
  return (lut[idx] + __float_as_uint(xc)) >> 24;
}

// ============================================================================ decode, fast
// One warp owns a strip of consecutive 128-pixel x 2-row segments (one chroma row).  In a
// segment lane l handles pixels [4l, 4l+4) of both rows, so every warp-level access is contiguous:
//   Y     2 rows x LDG.32   (128 B per instruction)
//   CbCr  1 x LDG.32        (128 B per instruction; loaded once, used for both rows)
//   BGRA  2 rows x STG.128  (512 B per instruction)
// The next segment's loads are issued before the current one is computed (software prefetch).
// Grid: persistent, one CTA per SM (the replicated LUT takes 128 KB of shared memory).
//
// The float32 steps of the matrix are done two pixels at a time with the Blackwell packed
// instructions (add/mul/fma .f32x2 -> FADD2/FMUL2/FFMA2): identical IEEE results per lane, half
// the issue slots.  A pair always holds the same pixel position of chroma pair A (pixels 0,1) and
// chroma pair B (pixels 2,3), so chroma products pair up naturally with the luma terms.
typedef unsigned long long f32x2;

__device__ __forceinline__ f32x2 pk(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ f32x2 pku(uint32_t lo, uint32_t hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ void upk(f32x2 v, float &lo, float &hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ void upku(f32x2 v, uint32_t &lo, uint32_t &hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 fma2_rz(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rz.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

// loop-invariant packed constants {k, k}
struct DecodeK2 {
  f32x2 kY, kC, inv255, ys, g_cb, b_cb, r_cr, g_cr, fN, magic, one;
};
__device__ __forceinline__ f32x2 fma2_rn(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

template <int K>
__device__ __forceinline__ uint32_t magic_byte(uint32_t word) {  // bits of the float 2^23 + byte K
  return __byte_perm(word, 0x4B000000u, 0x7540 + K);
}

// two LUT lookups at once: clamp (scalar), index (packed FFMA2.RZ), two LDS, two adds.
// returns the two sums whose byte 3 is the 8-bit code
template <int CL2, int RL2>
__device__ __forceinline__ void lut2(float xa, float xb, float clamp_lo, const DecodeK2 &k, uint32_t sbase,
                                     uint32_t &sa, uint32_t &sb) {
  float ca = fmaxf(xa, clamp_lo), cb = fmaxf(xb, clamp_lo);
  uint32_t ta, tb;
  upku(fma2_rz(pk(ca, cb), k.fN, k.magic), ta, tb);
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

// Work layout.  A warp owns a 128-pixel x 2-row tile (lane l: pixels [4l, 4l+4) of both rows).
// The warps of the CTA sit side by side: they cover `rps` complete row pairs per step
// (rps = warps / tiles per row), so a CTA touches 2*rps whole contiguous rows of Y, rps rows of
// CbCr and 2*rps rows of BGRA at a time, and consecutive CTAs take consecutive bands.  That keeps
// DRAM pages hot (a warp streaming its own row did not: 4736 independent row streams measured
// 51 % long-scoreboard stalls at 4.2 TB/s).  Each warp's (tile, row-in-band) is fixed for the whole
// kernel; per step only (frame, row pair) advance, by a constant with a carry, and every address
// is two IMAD.WIDE away.  The loads of the next step are issued before the current step is
// computed; the loop is unrolled twice so the two load buffers ping-pong without moves.
// FULL = (w % 128 == 0 && h % 2 == 0): no per-lane or per-row predicates anywhere.
template <bool ALPHA>
struct TileLoad {
  uint32_t y0, y1, c, a0, a1;
};

// Addressing.  Every plane is addressed as base + 32-bit offset in load/store units (4 bytes
// for Y / CbCr / alpha, 16 bytes for BGRA), so an address is a single IMAD.WIDE.U32 and stepping
// to the next (frame, row pair) is one 32-bit add per plane plus a predicated add on a frame
// wrap.  The launcher splits batches whose offsets would not fit in 32 bits.
template <bool ALPHA, bool FULL>
__device__ __forceinline__ void tile_load(const DecodeArgs &p, unsigned oy, unsigned oc, unsigned oa, unsigned rp,
                                          bool valid, TileLoad<ALPHA> &t) {
  const bool has_row1 = FULL || rp * 2 + 1 < (unsigned) p.h;
  if (!FULL) t.y0 = t.y1 = t.c = t.a0 = t.a1 = 0;
  if (FULL || valid) {
    t.y0 = ld_stream_u32(ptr_mad(p.y, oy, 4));
    t.c = ld_stream_u32(ptr_mad(p.c, oc, 4));
    if (has_row1) t.y1 = ld_stream_u32(ptr_mad(p.y, oy + p.py4, 4));
    if (ALPHA) {
      t.a0 = ld_stream_u32(ptr_mad(p.a, oa, 4));
      if (has_row1) t.a1 = ld_stream_u32(ptr_mad(p.a, oa + p.pa4, 4));
    }
  }
}

template <int CL2, int RL2, bool ALPHA, bool FULL>
__device__ __forceinline__ void tile_compute_store(const DecodeArgs &p, unsigned oo, unsigned rp, bool valid,
                                                   const TileLoad<ALPHA> &t, const DecodeK2 &k, float clamp_lo,
                                                   uint32_t sbase) {
  if (!FULL && !valid) return;
  const bool has_row1 = FULL || rp * 2 + 1 < (unsigned) p.h;
  // chroma pairs A (bytes 0,1) and B (bytes 2,3): BT709.h:499-500 then the matrix columns
  f32x2 cbn = mul2(add2(pku(magic_byte<0>(t.c), magic_byte<2>(t.c)), k.kC), k.inv255);
  f32x2 crn = mul2(add2(pku(magic_byte<1>(t.c), magic_byte<3>(t.c)), k.kC), k.inv255);
  f32x2 cb_g = mul2(cbn, k.g_cb), cb_b = mul2(cbn, k.b_cb);
  f32x2 cr_r = mul2(crn, k.r_cr), cr_g = mul2(crn, k.g_cr);
  st_stream_u128((uint8_t *) ptr_mad(p.out, oo, 16),
                 decode_row<CL2, RL2, ALPHA>(t.y0, t.a0, cb_g, cb_b, cr_r, cr_g, k, clamp_lo, sbase));
  if (has_row1)
    st_stream_u128((uint8_t *) ptr_mad(p.out, oo + p.po16, 16),
                   decode_row<CL2, RL2, ALPHA>(t.y1, t.a1, cb_g, cb_b, cr_r, cr_g, k, clamp_lo, sbase));
}

// WRAP = false: the batch is one contiguous tall image (frame stride == rows * pitch for every
// plane), so stepping never crosses a frame seam and the carry logic disappears.
template <int CL2, int RL2, bool ALPHA, bool FULL, bool WRAP, int PD, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) k_decode_fast(const DecodeArgs p) {
  extern __shared__ __align__(128) uint32_t s_lut[];
  constexpr int N = 1 << CL2, R = 1 << RL2;
  for (int i = threadIdx.x; i < (N + 1) * R; i += THREADS) s_lut[i] = __ldg(&p.lut[i >> RL2]);
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const unsigned warp = threadIdx.x >> 5;
  const uint32_t sbase = (uint32_t) __cvta_generic_to_shared(s_lut) + ((lane & (R - 1)) << 2) -
                         (0x4B000000u << (2 + RL2));
  const float clamp_lo = p.clamp_lo;
  DecodeK2 k;
  k.kY = pk(-(8388608.0f + 16.0f), -(8388608.0f + 16.0f));
  k.kC = pk(-(8388608.0f + 128.0f), -(8388608.0f + 128.0f));
  k.inv255 = pk(p.m.inv255, p.m.inv255);
  k.ys = pk(p.m.ys, p.m.ys);
  k.g_cb = pk(p.m.g_cb, p.m.g_cb);
  k.b_cb = pk(p.m.b_cb, p.m.b_cb);
  k.r_cr = pk(p.m.r_cr, p.m.r_cr);
  k.g_cr = pk(p.m.g_cr, p.m.g_cr);
  k.fN = pk((float) N, (float) N);
  k.magic = pk(8388608.0f, 8388608.0f);
  k.one = pk(p.one, p.one);

  // this warp's place in the CTA's band: fixed for the whole kernel
  const unsigned r_local = warp / p.tiles_per_row;
  if (r_local >= (unsigned) p.rps) return;  // spare warps when the warp count is not a multiple of the tiles per row
  const unsigned tile0 = warp - r_local * p.tiles_per_row;
  const unsigned rows2 = p.rows2;
  const unsigned total_rp = (unsigned) p.nframes * rows2;
  const unsigned rpg0 = blockIdx.x * p.rps + r_local;
  if (rpg0 >= total_rp) return;
  const unsigned f0 = rpg0 / rows2, rp0 = rpg0 - f0 * rows2;
  const unsigned nsteps = (total_rp - rpg0 + p.step_rp - 1) / p.step_rp;  // >= 1

  for (unsigned tile = tile0; tile < (unsigned) p.tiles_per_row; tile += p.tile_stride) {
    const unsigned col4 = tile * 32 + lane;  // column in 4-pixel units
    const bool valid = FULL || (int) (col4 * 4) < p.w;
    // Two offset streams: the load stream runs PD steps ahead of the store stream.  Measured
    // (tools/probes/gran_bw.cu): with 1024 threads per SM the HBM pipe needs >= 24 B of loads in
    // flight per thread to reach 6 TB/s at this byte mix; 12 B (one step ahead) gave 4.4 TB/s.
    unsigned oy = f0 * p.sy4 + rp0 * 2 * p.py4 + col4;
    unsigned oc = f0 * p.sc4 + rp0 * p.pc4 + col4;
    unsigned oa = ALPHA ? f0 * p.sa4 + rp0 * 2 * p.pa4 + col4 : 0;
    unsigned oo = f0 * p.so16 + rp0 * 2 * p.po16 + col4;
    unsigned rpl = rp0, rps_ = rp0;  // row pair of the next load / of the next store
    constexpr int NB = PD + 1;
    TileLoad<ALPHA> buf[NB];
    auto advance_load = [&]() {
      rpl += p.d_rp;
      oy += p.dy; oc += p.dc;
      if (ALPHA) oa += p.da;
      if (WRAP && rpl >= rows2) { rpl -= rows2; oy += p.wy; oc += p.wc; if (ALPHA) oa += p.wa; }
    };
#pragma unroll
    for (int j = 0; j < PD; j++) {
      if ((unsigned) j < nsteps) {
        tile_load<ALPHA, FULL>(p, oy, oc, oa, rpl, valid, buf[j]);
        advance_load();
      }
    }
    unsigned kk = 0;
    while (true) {
#pragma unroll
      for (int j = 0; j < NB; j++) {
        if (kk + PD < nsteps) {
          tile_load<ALPHA, FULL>(p, oy, oc, oa, rpl, valid, buf[(j + PD) % NB]);
          advance_load();
        }
        tile_compute_store<CL2, RL2, ALPHA, FULL>(p, oo, rps_, valid, buf[j], k, clamp_lo, sbase);
        rps_ += p.d_rp;
        oo += p.dout;
        if (WRAP && rps_ >= rows2) { rps_ -= rows2; oo += p.wout; }
        if (++kk >= nsteps) goto tile_done;
      }
    }
  tile_done:;
  }
}

__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

// ================================================================= decode, cp.async warp ring
// Same arithmetic and CTA-band layout as k_decode_fast, but each warp stages its own inputs through
// a private shared-memory ring: per step ONE cp.async instruction (16 bytes per lane; lanes 0-7
// fetch the 128 B of Y row 0, lanes 8-15 Y row 1, lanes 16-23 the CbCr row) replaces three
// LDG.32 and their address arithmetic, and the bytes in flight live in shared memory instead of
// registers (D - 1 steps ahead).  Needs FULL geometry and 16-byte aligned Y / CbCr planes.
template <int CL2, int RL2, bool ALPHA, bool WRAP, int D, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) k_decode_ring(const DecodeArgs p) {
  extern __shared__ __align__(128) uint32_t s_lut[];
  constexpr int N = 1 << CL2, R = 1 << RL2;
  constexpr uint32_t kLutBytes = (uint32_t) (N + 1) * R * 4;
  constexpr uint32_t kStage = ALPHA ? 640 : 384;  // Y0 | Y1 | CbCr (| A0 | A1), 128 B each
  for (int i = threadIdx.x; i < (N + 1) * R; i += THREADS) s_lut[i] = __ldg(&p.lut[i >> RL2]);
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const unsigned warp = threadIdx.x >> 5;
  const uint32_t smem0 = (uint32_t) __cvta_generic_to_shared(s_lut);
  const uint32_t sbase = smem0 + ((lane & (R - 1)) << 2) - (0x4B000000u << (2 + RL2));
  const uint32_t ring = smem0 + kLutBytes + warp * (D * kStage);
  const float clamp_lo = p.clamp_lo;
  DecodeK2 k;
  k.kY = pk(-(8388608.0f + 16.0f), -(8388608.0f + 16.0f));
  k.kC = pk(-(8388608.0f + 128.0f), -(8388608.0f + 128.0f));
  k.inv255 = pk(p.m.inv255, p.m.inv255);
  k.ys = pk(p.m.ys, p.m.ys);
  k.g_cb = pk(p.m.g_cb, p.m.g_cb);
  k.b_cb = pk(p.m.b_cb, p.m.b_cb);
  k.r_cr = pk(p.m.r_cr, p.m.r_cr);
  k.g_cr = pk(p.m.g_cr, p.m.g_cr);
  k.fN = pk((float) N, (float) N);
  k.magic = pk(8388608.0f, 8388608.0f);
  k.one = pk(p.one, p.one);

  const unsigned r_local = warp / p.tiles_per_row;
  if (r_local >= (unsigned) p.rps) return;
  const unsigned tile = warp - r_local * p.tiles_per_row;
  const unsigned rows2 = p.rows2;
  const unsigned total_rp = (unsigned) p.nframes * rows2;
  const unsigned rpg0 = blockIdx.x * p.rps + r_local;
  if (rpg0 >= total_rp) return;
  const unsigned f0 = rpg0 / rows2, rp0 = rpg0 - f0 * rows2;
  const unsigned nsteps = (total_rp - rpg0 + p.step_rp - 1) / p.step_rp;

  // ---- copy role of this lane: which plane row it fetches, as base + 32-bit offset in 4-byte units
  const int role = lane >> 3, sub = lane & 7;  // 0: Y row 0, 1: Y row 1, 2: CbCr, 3: idle (alpha: second pass)
  const uint8_t *cbase = role < 2 ? p.y : p.c;
  unsigned coff = role < 2 ? f0 * p.sy4 + (rp0 * 2 + role) * p.py4 : f0 * p.sc4 + rp0 * p.pc4;
  coff += tile * 32 + sub * 4;
  const unsigned cinc = role < 2 ? p.dy : p.dc, cwrap = role < 2 ? p.wy : p.wc;
  const uint32_t cdst = role * 128 + sub * 16;
  // alpha rows (ALPHA only): lanes 0-7 row 0, lanes 8-15 row 1, via a second cp.async
  unsigned aoff = ALPHA ? f0 * p.sa4 + (rp0 * 2 + (role & 1)) * p.pa4 + tile * 32 + sub * 4 : 0;
  const uint32_t adst = 384 + (role & 1) * 128 + sub * 16;
  unsigned rpl = rp0;

  auto issue = [&](uint32_t slot) {
    if (role < 3)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ring + slot * kStage + cdst),
                   "l"(ptr_mad(cbase, coff, 4))
                   : "memory");
    if (ALPHA && role < 2)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ring + slot * kStage + adst),
                   "l"(ptr_mad(p.a, aoff, 4))
                   : "memory");
    // step this lane's source to the warp's row pair in the CTA's next band
    rpl += p.d_rp;
    coff += cinc;
    if (ALPHA) aoff += p.da;
    if (WRAP && rpl >= rows2) { rpl -= rows2; coff += cwrap; if (ALPHA) aoff += p.wa; }
  };

  const unsigned col4 = tile * 32 + lane;
  unsigned oo = f0 * p.so16 + rp0 * 2 * p.po16 + col4;
  unsigned rps_ = rp0;
#pragma unroll
  for (int j = 0; j < D - 1; j++) {
    if ((unsigned) j < nsteps) issue(j);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  uint32_t slot = 0, fill = D - 1;
  for (unsigned kk = 0; kk < nsteps; kk++) {
    if (kk + D - 1 < nsteps) issue(fill);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group %0;" ::"n"(D - 1) : "memory");
    __syncwarp();
    const uint32_t st = ring + slot * kStage + lane * 4;
    TileLoad<ALPHA> t;
    t.y0 = lds_u32(st);
    t.y1 = lds_u32(st + 128);
    t.c = lds_u32(st + 256);
    if (ALPHA) {
      t.a0 = lds_u32(st + 384);
      t.a1 = lds_u32(st + 512);
    }
    __syncwarp();  // every lane has its words before this slot is refilled
    tile_compute_store<CL2, RL2, ALPHA, true>(p, oo, rps_, true, t, k, clamp_lo, sbase);
    rps_ += p.d_rp;
    oo += p.dout;
    if (WRAP && rps_ >= rows2) { rps_ -= rows2; oo += p.wout; }
    slot = slot + 1 == D ? 0 : slot + 1;
    fill = fill + 1 == D ? 0 : fill + 1;
  }
}

// ===================================================================== decode, bulk-staged
// Same arithmetic and the same CTA-band work layout as k_decode_fast, but the input planes reach
// the SM through the bulk async-copy engine (cp.async.bulk -> UBLKCP, the 1-D TMA path) into a
// shared-memory ring instead of through per-thread LDG into registers.
//
// Why: with per-thread loads the bytes in flight per SM are bounded by registers
// (32 warps x 3 x 128 B = 12 KB), and the measured throughput was the same 4.46 TB/s with the
// arithmetic compiled out - a latency (Little's law) bound, not an issue bound.  A ring of up to 8
// stages x 11.5 KB keeps ~90 KB of reads in flight per SM, costs no registers, and removes every
// load address computation from the consumer warps.
//
//   stage  = one CTA band: 2*rps rows of Y and rps rows of CbCr (and 2*rps rows of alpha),
//            each row one bulk copy of w bytes, issued by lane 0 of warp j for row j
//   full[s]   mbarrier, count = rows per stage, completes when all bytes of the stage have landed
//   empty[s]  mbarrier, count = consumer warps, completes when every warp has read its words
// Requires w % 128 == 0, h even, 16-byte aligned planes and pitches (else k_decode_fast).
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

template <int CL2, int RL2, bool ALPHA, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) k_decode_bulk(const DecodeArgs p) {
  extern __shared__ __align__(128) uint32_t s_lut[];
  constexpr int N = 1 << CL2, R = 1 << RL2;
  constexpr uint32_t kLutBytes = (uint32_t) (N + 1) * R * 4;
  for (int i = threadIdx.x; i < (N + 1) * R; i += THREADS) s_lut[i] = __ldg(&p.lut[i >> RL2]);

  const uint32_t smem0 = (uint32_t) __cvta_generic_to_shared(s_lut);
  const uint32_t ring0 = smem0 + kLutBytes;
  const uint32_t S = p.nstages;
  const uint32_t bar_full = ring0 + S * p.stage_bytes, bar_empty = bar_full + 8 * S;
  const unsigned nwarps_active = p.rps * p.tiles_per_row;
  if (threadIdx.x == 0) {
    for (uint32_t i = 0; i < S; i++) {
      mbar_init(bar_full + 8 * i, p.nrows);
      mbar_init(bar_empty + 8 * i, nwarps_active);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const unsigned warp = threadIdx.x >> 5;
  const uint32_t sbase = smem0 + ((lane & (R - 1)) << 2) - (0x4B000000u << (2 + RL2));
  const float clamp_lo = p.clamp_lo;
  DecodeK2 k;
  k.kY = pk(-(8388608.0f + 16.0f), -(8388608.0f + 16.0f));
  k.kC = pk(-(8388608.0f + 128.0f), -(8388608.0f + 128.0f));
  k.inv255 = pk(p.m.inv255, p.m.inv255);
  k.ys = pk(p.m.ys, p.m.ys);
  k.g_cb = pk(p.m.g_cb, p.m.g_cb);
  k.b_cb = pk(p.m.b_cb, p.m.b_cb);
  k.r_cr = pk(p.m.r_cr, p.m.r_cr);
  k.g_cr = pk(p.m.g_cr, p.m.g_cr);
  k.fN = pk((float) N, (float) N);
  k.magic = pk(8388608.0f, 8388608.0f);
  k.one = pk(p.one, p.one);

  if (warp >= nwarps_active) return;  // spare warps
  const unsigned rows2 = p.rows2, rps = p.rps, w = p.w;
  const unsigned total_rp = (unsigned) p.nframes * rows2;
  const unsigned nbands = (total_rp + rps - 1) / rps;
  if (blockIdx.x >= nbands) return;
  const unsigned nsteps = (nbands - blockIdx.x + gridDim.x - 1) / gridDim.x;
  const unsigned r_local = warp / p.tiles_per_row;
  const unsigned tile = warp - r_local * p.tiles_per_row;

  // ---- producer role: lane 0 of warp j feeds row j of every stage
  const bool producer = lane == 0 && warp < (unsigned) p.nrows;
  unsigned prow_rr = 0, prow_dst = 0;          // row pair within the band; byte offset within a stage
  const uint8_t *pbase = nullptr;              // plane base
  unsigned pstride = 0, ppitch = 0, prow_mul = 0, prow_add = 0;  // frame stride; pitch; row = rp*mul + add
  if (producer) {
    const unsigned j = warp;
    if (j < 2 * rps) {
      prow_rr = j >> 1; pbase = p.y; pstride = (unsigned) p.sy; ppitch = (unsigned) p.py; prow_mul = 2; prow_add = j & 1;
    } else if (j < 3 * rps) {
      prow_rr = j - 2 * rps; pbase = p.c; pstride = (unsigned) p.sc; ppitch = (unsigned) p.pc; prow_mul = 1; prow_add = 0;
    } else {
      const unsigned q = j - 3 * rps;
      prow_rr = q >> 1; pbase = p.a; pstride = (unsigned) p.sa; ppitch = (unsigned) p.pa; prow_mul = 2; prow_add = q & 1;
    }
    prow_dst = j * w;
  }
  auto produce = [&](unsigned kp) {
    if (!producer || kp >= nsteps) return;
    const uint32_t slot = kp % S;
    if (kp >= S) mbar_wait(bar_empty + 8 * slot, ((kp / S) - 1) & 1);
    const unsigned rpg = (blockIdx.x + kp * gridDim.x) * rps + prow_rr;
    const uint32_t bar = bar_full + 8 * slot;
    if (rpg < total_rp) {
      const unsigned f = rpg / rows2, rp = rpg - f * rows2;
      const uint8_t *src = ptr_mad(ptr_mad(pbase, f, pstride), rp * prow_mul + prow_add, ppitch);
      mbar_expect_tx(bar, w);
      bulk_g2s(ring0 + slot * p.stage_bytes + prow_dst, src, w, bar);
    } else {
      mbar_arrive(bar);  // band runs past the last row pair: nothing to copy for this row
    }
  };

  // ---- consumer role
  const unsigned col4 = tile * 32 + lane;
  const uint32_t my_y = 2 * r_local * w + col4 * 4;              // byte offsets within a stage
  const uint32_t my_c = (2 * rps + r_local) * w + col4 * 4;
  const uint32_t my_a = (3 * rps + 2 * r_local) * w + col4 * 4;
  unsigned rpg = blockIdx.x * rps + r_local;
  unsigned f0 = rpg / rows2, rp = rpg - f0 * rows2;
  unsigned oo = f0 * p.so16 + rp * 2 * p.po16 + col4;

  for (unsigned kp = 0; kp + 1 < S; kp++) produce(kp);
  for (unsigned kk = 0; kk < nsteps; kk++) {
    produce(kk + S - 1);
    const uint32_t slot = kk % S;
    mbar_wait(bar_full + 8 * slot, (kk / S) & 1);
    const uint32_t st = ring0 + slot * p.stage_bytes;
    TileLoad<ALPHA> t;
    t.y0 = lds_u32(st + my_y);
    t.y1 = lds_u32(st + my_y + w);
    t.c = lds_u32(st + my_c);
    if (ALPHA) {
      t.a0 = lds_u32(st + my_a);
      t.a1 = lds_u32(st + my_a + w);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_empty + 8 * slot);
    if (rpg < total_rp) tile_compute_store<CL2, RL2, ALPHA, true>(p, oo, rp, true, t, k, clamp_lo, sbase);
    // advance to this warp's row pair in the CTA's next band
    rpg += p.step_rp;
    rp += p.d_rp;
    oo += p.dout;
    if (rp >= rows2) { rp -= rows2; oo += p.wout; }
  }
}

// ========================================================================= decode, generic
// Any width/height/pitch/alignment (odd sizes, unaligned planes): one thread per 2x2 block,
// byte accesses, LUT read through the read-only path.  Same arithmetic, same results.
template <bool ALPHA>
__global__ void __launch_bounds__(256) k_decode_generic(const DecodeArgs p) {
  const int cw = (p.w + 1) >> 1;
  const long long blocks = (long long) p.nframes * p.rows2 * cw;
  const float fN = (float) (1 << p.cells_log2);
  const float kY = 16.0f, kC = 128.0f;
  for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < blocks;
       i += (long long) gridDim.x * blockDim.x) {
    const int bx = (int) (i % cw);
    const long long t = i / cw;
    const int rp = (int) (t % p.rows2);
    const long long f = t / p.rows2;
    const uint8_t *cp = p.c + f * p.sc + (long long) rp * p.pc + bx * 2;
    float cbn = __fmul_rn((float) cp[0] - kC, p.m.inv255);
    float crn = __fmul_rn((float) cp[1] - kC, p.m.inv255);
    float cb_g = __fmul_rn(cbn, p.m.g_cb), cb_b = __fmul_rn(cbn, p.m.b_cb);
    float cr_r = __fmul_rn(crn, p.m.r_cr), cr_g = __fmul_rn(crn, p.m.g_cr);
    for (int dy = 0; dy < 2; dy++) {
      const int row = rp * 2 + dy;
      if (row >= p.h) break;
      for (int dx = 0; dx < 2; dx++) {
        const int col = bx * 2 + dx;
        if (col >= p.w) break;
        float yf = (float) p.y[f * p.sy + (long long) row * p.py + col] - kY;
        float yt = __fmul_rn(__fmul_rn(yf, p.m.inv255), p.m.ys);
        float xr = add_sat(yt, cr_r);
        float xg = add_sat(__fadd_rn(yt, cb_g), cr_g);
        float xb = add_sat(yt, cb_b);
        uint32_t R = lut_code(p.lut, xr, p.clamp_lo, fN);
        uint32_t G = lut_code(p.lut, xg, p.clamp_lo, fN);
        uint32_t B = lut_code(p.lut, xb, p.clamp_lo, fN);
        uint32_t A = 255;
        if (ALPHA) {
          float af = (float) p.a[f * p.sa + (long long) row * p.pa + col] - kY;
          float an = add_sat(__fmul_rn(__fmul_rn(af, p.m.inv255), p.m.ys), 0.0f);
          A = (uint32_t) round_pos(__fmul_rn(an, 255.0f));
        }
        uint8_t *o = p.out + f * p.so + (long long) row * p.po + (long long) col * 4;
        o[0] = (uint8_t) B; o[1] = (uint8_t) G; o[2] = (uint8_t) R; o[3] = (uint8_t) A;
      }
    }
  }
}

static inline bool aligned(const void *p, size_t a) { return ((uintptr_t) p & (a - 1)) == 0; }
static inline bool aligned(long long v, size_t a) { return ((unsigned long long) v & (a - 1)) == 0; }

#ifndef BT709_DECODE_THREADS
#define BT709_DECODE_THREADS 1024
#endif
constexpr int kDecodeThreads = BT709_DECODE_THREADS;
#ifndef BT709_PREFETCH_DEPTH
#define BT709_PREFETCH_DEPTH 3
#endif
constexpr int kPrefetchDepth = BT709_PREFETCH_DEPTH;  // steps the loads run ahead of the math
#ifndef BT709_RING_DEPTH
#define BT709_RING_DEPTH 2
#endif
constexpr int kRingDepth = BT709_RING_DEPTH;  // stages of the per-warp cp.async ring

template <int CL2, int RL2, bool ALPHA, bool FULL, bool WRAP>
static int launch_decode_fast2(const DecodeArgs &a0, int sm_count, cudaStream_t s) {
  constexpr size_t smem = (size_t) ((1 << CL2) + 1) * (1 << RL2) * 4;
  auto k = k_decode_fast<CL2, RL2, ALPHA, FULL, WRAP, kPrefetchDepth, kDecodeThreads>;
  auto kring = k_decode_ring<CL2, RL2, ALPHA, WRAP, kRingDepth, kDecodeThreads>;
  constexpr size_t smem_ring = smem + (size_t) (kDecodeThreads / 32) * kRingDepth * (ALPHA ? 640 : 384);
  static const bool no_ring = getenv("BT709_NO_RING") != nullptr;
  const bool use_ring = FULL && !no_ring && aligned(a0.y, 16) && aligned(a0.py, 16) && aligned(a0.sy, 16) &&
                        aligned(a0.c, 16) && aligned(a0.pc, 16) && aligned(a0.sc, 16) &&
                        (!ALPHA || (aligned(a0.a, 16) && aligned(a0.pa, 16) && aligned(a0.sa, 16)));
  // frames per launch such that every 32-bit unit offset stays below 2^32
  const unsigned long long lim = 0xFFFF0000ull;
  unsigned long long per = lim;
  auto fit = [&](long long stride, unsigned unit, long long frame_bytes) {
    if (stride <= 0) return;
    unsigned long long n = (lim * unit - (unsigned long long) frame_bytes) / (unsigned long long) stride + 1;
    if (n < per) per = n;
  };
  fit(a0.sy, 4, (long long) a0.h * a0.py);
  fit(a0.sc, 4, (long long) a0.rows2 * a0.pc);
  fit(a0.so, 16, (long long) a0.h * a0.po);
  if (ALPHA) fit(a0.sa, 4, (long long) a0.h * a0.pa);
  if (per < 1) per = 1;
  if (a0.nframes == 1) {
    // one (possibly very tall) image: split by rows instead, in chunks whose offsets fit in 32 bits
    const unsigned long long out_bytes = (unsigned long long) a0.h * a0.po;
    if (out_bytes > lim * 16 || (unsigned long long) a0.h * a0.py > lim * 4) {
      long long rows_per = (long long) ((lim * 16) / (unsigned long long) a0.po);
      if ((unsigned long long) rows_per * a0.py > lim * 4) rows_per = (long long) ((lim * 4) / (unsigned long long) a0.py);
      rows_per &= ~1ll;
      for (long long r0 = 0; r0 < a0.h; r0 += rows_per) {
        DecodeArgs a = a0;
        a.h = (int) (a0.h - r0 < rows_per ? a0.h - r0 : rows_per);
        a.rows2 = (a.h + 1) / 2;
        a.y = a0.y + r0 * a0.py;
        a.c = a0.c + (r0 / 2) * a0.pc;
        a.out = a0.out + r0 * a0.po;
        if (ALPHA) a.a = a0.a + r0 * a0.pa;
        int e = launch_decode_fast2<CL2, RL2, ALPHA, FULL, WRAP>(a, sm_count, s);
        if (e) return e;
      }
      return 0;
    }
  }
  for (int f0 = 0; f0 < a0.nframes; f0 += (int) per) {
    DecodeArgs a = a0;
    a.nframes = (int) ((unsigned long long) (a0.nframes - f0) < per ? (unsigned long long) (a0.nframes - f0) : per);
    a.y = a0.y + (long long) f0 * a0.sy;
    a.c = a0.c + (long long) f0 * a0.sc;
    a.out = a0.out + (long long) f0 * a0.so;
    if (ALPHA) a.a = a0.a + (long long) f0 * a0.sa;
    const unsigned total_rp = (unsigned) a.nframes * a.rows2;
    unsigned grid = (total_rp + a.rps - 1) / a.rps;
    if (grid > (unsigned) sm_count) grid = sm_count;
    // per-step increments of the unit offsets, and the extra increment on a frame wrap (mod 2^32)
    a.step_rp = grid * a.rps;
    const unsigned d_f = a.step_rp / a.rows2;
    a.d_rp = a.step_rp - d_f * a.rows2;
    a.py4 = (unsigned) (a.py / 4); a.pc4 = (unsigned) (a.pc / 4); a.pa4 = (unsigned) (a.pa / 4); a.po16 = (unsigned) (a.po / 16);
    a.sy4 = (unsigned) (a.sy / 4); a.sc4 = (unsigned) (a.sc / 4); a.sa4 = (unsigned) (a.sa / 4); a.so16 = (unsigned) (a.so / 16);
    a.dy = d_f * a.sy4 + a.d_rp * 2 * a.py4;
    a.dc = d_f * a.sc4 + a.d_rp * a.pc4;
    a.da = d_f * a.sa4 + a.d_rp * 2 * a.pa4;
    a.dout = d_f * a.so16 + a.d_rp * 2 * a.po16;
    a.wy = a.sy4 - a.rows2 * 2 * a.py4;
    a.wc = a.sc4 - a.rows2 * a.pc4;
    a.wa = a.sa4 - a.rows2 * 2 * a.pa4;
    a.wout = a.so16 - a.rows2 * 2 * a.po16;
    if (use_ring)
      kring<<<grid, kDecodeThreads, smem_ring, s>>>(a);
    else
      k<<<grid, kDecodeThreads, smem, s>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int) e;
  }
  return 0;
}

constexpr size_t kMaxDynSmem = 232448;  // 227 KB per CTA on sm_100

template <int CL2, int RL2, bool ALPHA>
static bool bulk_ok(const DecodeArgs &a, int *nstages, int *stage_bytes, int *nrows) {
  constexpr size_t lut = (size_t) ((1 << CL2) + 1) * (1 << RL2) * 4;
  const int wpc = kDecodeThreads / 32;
  if (!(a.w % 128 == 0 && a.h % 2 == 0 && a.tiles_per_row <= wpc)) return false;
  if (!(aligned(a.y, 16) && aligned(a.py, 16) && aligned(a.sy, 16) && aligned(a.c, 16) && aligned(a.pc, 16) &&
        aligned(a.sc, 16)))
    return false;
  if (ALPHA && !(aligned(a.a, 16) && aligned(a.pa, 16) && aligned(a.sa, 16))) return false;
  *nrows = (ALPHA ? 5 : 3) * a.rps;
  if (*nrows > a.rps * a.tiles_per_row) return false;  // one producer lane per row
  *stage_bytes = *nrows * a.w;
  long long room = (long long) kMaxDynSmem - (long long) lut - 16 * 8 * 2;
  int n = (int) (room / *stage_bytes);
  if (n > 8) n = 8;
  if (n < 2) return false;
  *nstages = n;
  return true;
}

template <int CL2, int RL2, bool ALPHA>
static int launch_decode_fast(const DecodeArgs &a0, int sm_count, cudaStream_t s) {
  DecodeArgs a = a0;
  int nstages = 0, stage_bytes = 0, nrows = 0;
  // The bulk-staged variant is correct (tests/test_gpu_decode.py runs it) but measured slower than
  // per-thread loads (774 vs 811 Gpix/s on 600 x 1080p): the path is bound by HBM *write*
  // bandwidth (profiles/r01_membw_probe.json), not by read latency.  Opt-in for reproduction.
  static const bool use_bulk = getenv("BT709_BULK") != nullptr;
  if (use_bulk && bulk_ok<CL2, RL2, ALPHA>(a, &nstages, &stage_bytes, &nrows)) {
    constexpr size_t lut = (size_t) ((1 << CL2) + 1) * (1 << RL2) * 4;
    a.nstages = nstages; a.stage_bytes = stage_bytes; a.nrows = nrows;
    const size_t smem = lut + (size_t) nstages * stage_bytes + (size_t) nstages * 16;
    // the output offsets are 32-bit in 16-byte units: split batches that would overflow
    const unsigned long long lim = 0xFFFF0000ull;
    unsigned long long per = a.so > 0 ? (lim * 16 - (unsigned long long) a.h * a.po) / (unsigned long long) a.so + 1 : lim;
    if (per < 1) per = 1;
    for (int f0 = 0; f0 < a0.nframes; f0 += (int) per) {
      a.nframes = (int) ((unsigned long long) (a0.nframes - f0) < per ? (unsigned long long) (a0.nframes - f0) : per);
      a.y = a0.y + (long long) f0 * a0.sy;
      a.c = a0.c + (long long) f0 * a0.sc;
      a.out = a0.out + (long long) f0 * a0.so;
      if (ALPHA) a.a = a0.a + (long long) f0 * a0.sa;
      const unsigned total_rp = (unsigned) a.nframes * a.rows2;
      unsigned grid = (total_rp + a.rps - 1) / a.rps;
      if (grid > (unsigned) sm_count) grid = sm_count;
      a.step_rp = grid * a.rps;
      const unsigned d_f = a.step_rp / a.rows2;
      a.d_rp = a.step_rp - d_f * a.rows2;
      a.po16 = (unsigned) (a.po / 16); a.so16 = (unsigned) (a.so / 16);
      a.dout = d_f * a.so16 + a.d_rp * 2 * a.po16;
      a.wout = a.so16 - a.rows2 * 2 * a.po16;
      k_decode_bulk<CL2, RL2, ALPHA, kDecodeThreads><<<grid, kDecodeThreads, smem, s>>>(a);
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) return (int) e;
    }
    return 0;
  }
  const bool full = (a.w % 128 == 0) && (a.h % 2 == 0);
  if (!full) return launch_decode_fast2<CL2, RL2, ALPHA, false, true>(a, sm_count, s);
  // contiguous frames: decode the batch as one tall image (no frame seams to step across)
  const bool contiguous = a.nframes == 1 || (a.sy == (long long) a.h * a.py && a.sc == (long long) a.rows2 * a.pc &&
                                             a.so == (long long) a.h * a.po && (!ALPHA || a.sa == (long long) a.h * a.pa));
  if (contiguous && (long long) a.nframes * a.h < (1ll << 30)) {
    a.h *= a.nframes;
    a.rows2 *= a.nframes;
    a.nframes = 1;
    return launch_decode_fast2<CL2, RL2, ALPHA, true, false>(a, sm_count, s);
  }
  return launch_decode_fast2<CL2, RL2, ALPHA, true, true>(a, sm_count, s);
}

int launch_decode(const DecodeArgs &a0, int sm_count, void *stream) {
  cudaStream_t s = (cudaStream_t) stream;
  DecodeArgs a = a0;
  a.rows2 = (a.h + 1) / 2;
  // 128-pixel tiles; the CTA's warps cover rps whole row pairs per step (see k_decode_fast)
  a.tiles_per_row = (a.w + 127) / 128;
  const int wpc = kDecodeThreads / 32;
  a.rps = a.tiles_per_row >= wpc ? 1 : wpc / a.tiles_per_row;
  a.tile_stride = a.tiles_per_row >= wpc ? wpc : a.tiles_per_row;  // > 4096 px wide: warps loop over tiles
  const bool alpha = a.a != nullptr;
  const unsigned long long units = (unsigned long long) a.nframes * a.rows2 + 2ull * sm_count * 32;
  bool fast = (a.w % 4 == 0) && aligned(a.y, 4) && aligned(a.py, 4) && aligned(a.sy, 4) && aligned(a.c, 4) &&
              aligned(a.pc, 4) && aligned(a.sc, 4) && aligned(a.out, 16) && aligned(a.po, 16) &&
              aligned(a.so, 16) && units < 0x7FFFFFFFull && a.sy < (1ll << 32) && a.sc < (1ll << 32) &&
              a.so < (1ll << 34) && a.sa < (1ll << 32) && a.py < (1ll << 28) && a.pc < (1ll << 28) && a.po < (1ll << 30) &&
              a.pa < (1ll << 28) && a.sy >= 0 && a.sc >= 0 && a.so >= 0 && a.sa >= 0 &&
              (a.cells_log2 == 10 || a.cells_log2 == 12);
  if (alpha) fast = fast && aligned(a.a, 4) && aligned(a.pa, 4) && aligned(a.sa, 4);
  if (fast) {
    if (a.cells_log2 == 10)
      return alpha ? launch_decode_fast<10, 5, true>(a, sm_count, s) : launch_decode_fast<10, 5, false>(a, sm_count, s);
    return alpha ? launch_decode_fast<12, 3, true>(a, sm_count, s) : launch_decode_fast<12, 3, false>(a, sm_count, s);
  }
  const long long blocks = (long long) a.nframes * a.rows2 * ((a.w + 1) / 2);
  long long grid = (blocks + 255) / 256;
  if (grid > (long long) sm_count * 16) grid = (long long) sm_count * 16;
  if (alpha)
    k_decode_generic<true><<<(unsigned) grid, 256, 0, s>>>(a);
  else
    k_decode_generic<false><<<(unsigned) grid, 256, 0, s>>>(a);
  return (int) cudaGetLastError();
}

// ================================================================================= rescale
// Pass 2 (samplingShader, AAPLShaders.metal:73-85): one thread per output pixel.  Source texels
// come through the read-only path (each is reused by ~1.8 output pixels for 1080p -> 1440p and
// the four taps of neighbouring threads overlap, so L1/L2 absorb the reuse and HBM sees each
// source row about once).  sRGB->linear is a 256-float shared table (exact: the input is 8-bit);
// linear->sRGB is the cell LUT in shared memory.  float32 steps mirror oracle_rescale_bgra.
__global__ void __launch_bounds__(256) k_rescale(const RescaleArgs p) {
  extern __shared__ __align__(16) uint32_t s_mem[];
  float *s_lin = reinterpret_cast<float *>(s_mem);
  uint32_t *s_enc = s_mem + 256;
  const int nenc = (1 << p.enc_cells_log2) + 1;
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lin[i] = __ldg(&p.lin256[i]);
  for (int i = threadIdx.x; i < nenc; i += blockDim.x) s_enc[i] = __ldg(&p.enc_lut[i]);
  __syncthreads();
  const float fN = (float) (1 << p.enc_cells_log2);

  const int tiles_x = (p.ow + 31) / 32, tiles_y = (p.oh + 7) / 8;
  const long long tiles = (long long) p.nframes * tiles_x * tiles_y;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
    const int bx = (int) (t % tiles_x);
    const long long r = t / tiles_x;
    const int by = (int) (r % tiles_y);
    const long long f = r / tiles_y;
    const int ox = bx * 32 + tx, oy = by * 8 + ty;
    if (ox >= p.ow || oy >= p.oh) continue;
    float sx = __fadd_rn(__fmul_rn(__fadd_rn((float) ox, 0.5f), p.scale_x), -0.5f);
    float sy = __fadd_rn(__fmul_rn(__fadd_rn((float) oy, 0.5f), p.scale_y), -0.5f);
    float fx0 = floorf(sx), fy0 = floorf(sy);
    float fx = __fsub_rn(sx, fx0), fy = __fsub_rn(sy, fy0);
    int x0 = (int) fx0, y0 = (int) fy0;
    int x1 = min(max(x0 + 1, 0), p.iw - 1), y1 = min(max(y0 + 1, 0), p.ih - 1);
    x0 = min(max(x0, 0), p.iw - 1);
    y0 = min(max(y0, 0), p.ih - 1);
    const uint8_t *src = p.in + f * p.si;
    const uint32_t p00 = __ldg(reinterpret_cast<const uint32_t *>(src + (long long) y0 * p.pi) + x0);
    const uint32_t p10 = __ldg(reinterpret_cast<const uint32_t *>(src + (long long) y0 * p.pi) + x1);
    const uint32_t p01 = __ldg(reinterpret_cast<const uint32_t *>(src + (long long) y1 * p.pi) + x0);
    const uint32_t p11 = __ldg(reinterpret_cast<const uint32_t *>(src + (long long) y1 * p.pi) + x1);
    uint32_t outw = 0;
#pragma unroll
    for (int ch = 0; ch < 4; ch++) {
      const int sh = ch * 8;
      float t00, t10, t01, t11;
      if (ch < 3) {
        t00 = s_lin[(p00 >> sh) & 255]; t10 = s_lin[(p10 >> sh) & 255];
        t01 = s_lin[(p01 >> sh) & 255]; t11 = s_lin[(p11 >> sh) & 255];
      } else {
        const float k = 1.0f / 255.0f;
        t00 = __fmul_rn((float) (p00 >> 24), k); t10 = __fmul_rn((float) (p10 >> 24), k);
        t01 = __fmul_rn((float) (p01 >> 24), k); t11 = __fmul_rn((float) (p11 >> 24), k);
      }
      float top = __fadd_rn(t00, __fmul_rn(fx, __fsub_rn(t10, t00)));
      float bot = __fadd_rn(t01, __fmul_rn(fx, __fsub_rn(t11, t01)));
      float v = add_sat(top, __fmul_rn(fy, __fsub_rn(bot, top)));
      uint32_t code = ch < 3 ? lut_code(s_enc, v, p.enc_clamp_lo, fN) : (uint32_t) round_pos(__fmul_rn(v, 255.0f));
      outw |= code << sh;
    }
    *reinterpret_cast<uint32_t *>(p.out + f * p.so + (long long) oy * p.po + (long long) ox * 4) = outw;
  }
}

// -------------------------------------------------------------------- rescale, column walker
// Upscale / same-size fast path (scale <= 1 source pixel per output pixel in both directions).
// One thread owns one output column over a strip of TH output rows and walks down the source
// rows it needs: for each source row it converts its two horizontal taps to linear light
// (256-entry table replicated per lane: conflict-free) and blends them (the oracle's `top`/`bot`,
// same float32 steps), keeps the two most recent blended rows in registers, and emits every output
// row that lies between them.  Each blended row is reused by the ~1/scale output rows that touch
// it, nothing is staged in shared memory, and a warp reads ~96-128 contiguous source bytes and
// writes 128 contiguous output bytes per row.  The bilinear steps run on packed FADD2/FMUL2/FFMA2.
constexpr int kRescaleTH = 32;

struct Lin4 {
  f32x2 rg, ba;  // (R, G), (B, A) in linear light (A is linear already)
};

__global__ void __launch_bounds__(512, 2) k_rescale_cols(const RescaleArgs p) {
  extern __shared__ __align__(16) uint32_t s_mem[];
  // sRGB byte -> linear float, one replica per lane; entry stride 256 B so that the table offset
  // (byte << 8) | (lane << 2) is a single PRMT of the texel word and the lane register
  float *s_lin = reinterpret_cast<float *>(s_mem);  // entry b replica r at word b*64 + r
  uint32_t *s_enc = s_mem + 256 * 64;
  const int nenc = (1 << p.enc_cells_log2) + 1;
  for (int i = threadIdx.x; i < 256 * 32; i += blockDim.x) s_lin[(i >> 5) * 64 + (i & 31)] = __ldg(&p.lin256[i >> 5]);
  for (int i = threadIdx.x; i < nenc; i += blockDim.x) s_enc[i] = __ldg(&p.enc_lut[i]);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const uint32_t lin_base = (uint32_t) __cvta_generic_to_shared(s_lin);  // warp-uniform
  const uint32_t lane4 = lane << 2;
  const uint32_t enc_base = (uint32_t) __cvta_generic_to_shared(s_enc) - (0x4B000000u << 2);
  const float fN = (float) (1 << p.enc_cells_log2);
  const float clamp_lo = p.enc_clamp_lo;
  const f32x2 one = pk(p.one, p.one), inv255_1 = pk(1.0f / 255.0f, 1.0f);

  auto lin_of = [&](uint32_t px, int shift) -> float {
    const uint32_t off = __byte_perm(px, lane4, shift == 16 ? 0x7724 : shift == 8 ? 0x7714 : 0x7704);
    float v;
    asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(lin_base + off));
    return v;
  };
  auto texel = [&](uint32_t px) -> Lin4 {  // BGRA8 sRGB texel -> linear (R,G,B), A/255
    Lin4 t;
    t.rg = pk(lin_of(px, 16), lin_of(px, 8));
    const float a = __fadd_rn(__uint_as_float(__byte_perm(px, 0x4B000000u, 0x7543)), -8388608.0f);
    t.ba = mul2(pk(lin_of(px, 0), a), pk(1.0f, 1.0f / 255.0f));
    return t;
  };
  auto enc = [&](float v) -> uint32_t {  // linear -> sRGB code sum (byte 3)
    float xc = fmaxf(v, clamp_lo), t;
    asm("fma.rz.f32 %0, %1, %2, 0f4B000000;" : "=f"(t) : "f"(xc), "f"(fN));
    uint32_t e;
    asm("ld.shared.u32 %0, [%1];" : "=r"(e) : "r"((__float_as_uint(t) << 2) + enc_base));
    return e + __float_as_uint(xc);
  };

  const int tiles_x = (p.ow + 511) / 512, tiles_y = (p.oh + kRescaleTH - 1) / kRescaleTH;
  const unsigned tiles_per_frame = tiles_x * tiles_y;
  const unsigned tiles = (unsigned) p.nframes * tiles_per_frame;
  for (unsigned t = blockIdx.x; t < tiles; t += gridDim.x) {
    const unsigned f = t / tiles_per_frame, tr = t - f * tiles_per_frame;
    const int ty = tr / tiles_x, tx = tr - ty * tiles_x;
    const int ox = tx * 512 + threadIdx.x;
    if (ox >= p.ow) continue;
    // horizontal sample position of this column: same float32 steps as oracle_rescale_bgra
    const float sx = __fadd_rn(__fmul_rn(__fadd_rn((float) ox, 0.5f), p.scale_x), -0.5f);
    const float fx0 = floorf(sx);
    const float fx = __fsub_rn(sx, fx0);
    const int xi = (int) fx0;
    const int x0 = min(max(xi, 0), p.iw - 1), x1 = min(max(xi + 1, 0), p.iw - 1);
    const f32x2 fx2 = pk(fx, fx), neg1 = pk(-1.0f, -1.0f);
    const uint32_t *src0 = reinterpret_cast<const uint32_t *>(p.in + (long long) f * p.si) + x0;
    const uint32_t *src1 = reinterpret_cast<const uint32_t *>(p.in + (long long) f * p.si) + x1;
    uint8_t *dst = p.out + (long long) f * p.so + (long long) ox * 4;
    const unsigned pi4 = (unsigned) (p.pi >> 2), po = (unsigned) p.po;
    auto blend = [&](uint32_t w0, uint32_t w1) -> Lin4 {  // top/bot of the oracle: t0 + fx * (t1 - t0)
      const Lin4 t0 = texel(w0), t1 = texel(w1);
      Lin4 h;
      h.rg = fma2_rn(mul2(fx2, fma2_rn(t0.rg, neg1, t1.rg)), one, t0.rg);
      h.ba = fma2_rn(mul2(fx2, fma2_rn(t0.ba, neg1, t1.ba)), one, t0.ba);
      return h;
    };
    const int ylast = p.ih - 1;
    auto fetch = [&](int y, uint32_t &w0, uint32_t &w1) {  // taps of source row min(y, ih-1)
      const size_t o = (size_t) ((unsigned) min(y, ylast) * pi4);
      w0 = __ldg(src0 + o);
      w1 = __ldg(src1 + o);
    };
    auto emit = [&](int oy, float fy, const Lin4 &A, const Lin4 &B) {
      // v = top + fy * (bot - top), clamp01 (the saturating add), per channel
      const f32x2 fy2 = pk(fy, fy);
      float prg0, prg1, pba0, pba1, a0, a1, a2, a3;
      upk(mul2(fy2, fma2_rn(A.rg, neg1, B.rg)), prg0, prg1);
      upk(mul2(fy2, fma2_rn(A.ba, neg1, B.ba)), pba0, pba1);
      upk(A.rg, a0, a1); upk(A.ba, a2, a3);
      const uint32_t sr = enc(add_sat(a0, prg0)), sg = enc(add_sat(a1, prg1)), sb = enc(add_sat(a2, pba0));
      const uint32_t ac = (uint32_t) round_pos(__fmul_rn(add_sat(a3, pba1), 255.0f));
      const uint32_t bg = __byte_perm(sb, sg, 0x0073);
      const uint32_t w = __byte_perm(bg, __byte_perm(sr, ac, 0x0043), 0x5410);
      *reinterpret_cast<uint32_t *>(dst + (size_t) ((unsigned) oy * po)) = w;
    };
    // Walk the source rows this strip needs (they are consecutive for scale <= 1).  A = blended row
    // y, B = blended row y + 1; every output row whose top tap is y is emitted from (A, B).  The raw
    // taps of the next kPD rows sit in registers, fetched kPD rows ahead of use (one DRAM round trip
    // per source row per warp otherwise bounds the kernel: measured 196 Gpix/s out, issue 78 % idle
    // on the first use of the loaded texels).
    constexpr int kPD = 3;
    int oy = ty * kRescaleTH;
    const int oy1 = min(oy + kRescaleTH, p.oh);
    float oyf = (float) oy;
    float sy = __fadd_rn(__fmul_rn(__fadd_rn(oyf, 0.5f), p.scale_y), -0.5f);
    float fy0 = floorf(sy);
    int y = min(max((int) fy0, 0), ylast);  // first top row
    uint32_t q0[kPD], q1[kPD], a0w, a1w;
    fetch(y, a0w, a1w);
#pragma unroll
    for (int j = 0; j < kPD; j++) fetch(y + 1 + j, q0[j], q1[j]);
    Lin4 A = blend(a0w, a1w);
    while (oy < oy1) {  // every condition below depends on oy only: warp-uniform
#pragma unroll
      for (int j = 0; j < kPD; j++) {
        const Lin4 B = blend(q0[j], q1[j]);   // row min(y + 1, ih - 1); equals A bit for bit at the bottom edge
        fetch(y + 1 + kPD, q0[j], q1[j]);     // refill this slot with the row kPD further down
        while (oy < oy1) {                    // emit the output rows whose clamped top tap is y
          const int yi = (int) fy0;
          if (min(max(yi, 0), ylast) != y) break;
          // at the clamped image edges both taps are the same row and the blend degenerates to it
          const bool edge = min(max(yi + 1, 0), ylast) == y;
          emit(oy, __fsub_rn(sy, fy0), A, edge ? A : B);
          oy++;
          oyf += 1.0f;
          sy = __fadd_rn(__fmul_rn(__fadd_rn(oyf, 0.5f), p.scale_y), -0.5f);
          fy0 = floorf(sy);
        }
        A = B;
        y = min(y + 1, ylast);
        if (oy >= oy1) break;
      }
    }
  }
}

int launch_rescale(const RescaleArgs &a, int sm_count, void *stream) {
  if (a.iw <= a.ow && a.ih <= a.oh && a.enc_cells_log2 <= 13) {  // upscale or same size: column walker
    const size_t smem2 = (256 * 64 + (size_t) (1 << a.enc_cells_log2) + 1) * 4;
    const long long tiles2 = (long long) a.nframes * ((a.ow + 511) / 512) * ((a.oh + kRescaleTH - 1) / kRescaleTH);
    long long grid2 = tiles2 < (long long) sm_count * 2 ? tiles2 : (long long) sm_count * 2;
    k_rescale_cols<<<(unsigned) grid2, 512, smem2, (cudaStream_t) stream>>>(a);
    return (int) cudaGetLastError();
  }
  const size_t smem = (256 + (size_t) (1 << a.enc_cells_log2) + 1) * 4;
  const long long tiles = (long long) a.nframes * ((a.ow + 31) / 32) * ((a.oh + 7) / 8);
  long long grid = tiles < (long long) sm_count * 8 ? tiles : (long long) sm_count * 8;
  k_rescale<<<(unsigned) grid, 256, smem, (cudaStream_t) stream>>>(a);
  return (int) cudaGetLastError();
}

// ================================================================================== encode
// cvpbu_ycbcr_subsample + BT709_average_pixel_values (CVPixelBufferUtils.h:284-363,
// BT709.h:1349-1509).  VEC: one thread = 4 px x 2 rows (two 2x2 blocks), 16-byte loads,
// 4-byte Y and CbCr stores, so a warp reads 512 B and writes 128 B per row contiguously.
// !VEC: one thread per 2x2 block, 4-byte loads and byte stores (any even size / pitch).
struct EncTables {
  const float *lin;      // shared
  const float *compn;    // shared
  const uint32_t *olut;  // shared
  float clamp_lo, fN;
};

__device__ __forceinline__ uint32_t enc_luma(uint32_t px, const EncTables &t) {
  // per-pixel: requantised output-gamma RGB (BT709.h:1472-1476) -> matrix (BT709.h:223, 234, 245)
  float rn = t.compn[(px >> 16) & 255], gn = t.compn[(px >> 8) & 255], bn = t.compn[px & 255];
  float ey = __fadd_rn(__fadd_rn(__fmul_rn(0.2126f, rn), __fmul_rn(0.7152f, gn)), __fmul_rn(0.0722f, bn));
  return (uint32_t) round_pos(__fadd_rn(__fmul_rn(ey, 219.0f), 16.0f));
}

__device__ __forceinline__ uint32_t enc_chroma(uint32_t p1, uint32_t p2, uint32_t p3, uint32_t p4, const EncTables &t) {
  // mean in linear light (BT709.h:1400-1409), requantise in the output gamma (:1418-1420),
  // matrix on the 8-bit values (:1426 -> :223-247).  Returns Cb | Cr << 8.
  float q[3];
#pragma unroll
  for (int ch = 0; ch < 3; ch++) {
    const int sh = 16 - ch * 8;  // R, G, B
    float s = __fadd_rn(__fadd_rn(__fadd_rn(t.lin[(p1 >> sh) & 255], t.lin[(p2 >> sh) & 255]), t.lin[(p3 >> sh) & 255]),
                        t.lin[(p4 >> sh) & 255]);
    float mean = __fmul_rn(s, 0.25f);  // "/ 4.0f" is exact
    uint32_t code = lut_code(t.olut, mean, t.clamp_lo, t.fN);
    q[ch] = __fmul_rn((float) code, 1.0f / 255.0f);
  }
  float ey = __fadd_rn(__fadd_rn(__fmul_rn(0.2126f, q[0]), __fmul_rn(0.7152f, q[1])), __fmul_rn(0.0722f, q[2]));
  float eb = __fdiv_rn(__fsub_rn(q[2], ey), 1.8556f);
  float er = __fdiv_rn(__fsub_rn(q[0], ey), 1.5748f);
  uint32_t cb = (uint32_t) round_pos(__fadd_rn(__fmul_rn(eb, 224.0f), 128.0f));
  uint32_t cr = (uint32_t) round_pos(__fadd_rn(__fmul_rn(er, 224.0f), 128.0f));
  return cb | (cr << 8);
}

template <bool VEC>
__global__ void __launch_bounds__(256) k_encode(const EncodeArgs p) {
  extern __shared__ __align__(16) uint32_t s_mem[];
  float *s_lin = reinterpret_cast<float *>(s_mem);
  float *s_compn = s_lin + 256;
  uint32_t *s_olut = s_mem + 512;
  const int nl = (1 << p.out_cells_log2) + 1;
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    s_lin[i] = __ldg(&p.lin256[i]);
    s_compn[i] = __ldg(&p.compn256[i]);
  }
  for (int i = threadIdx.x; i < nl; i += blockDim.x) s_olut[i] = __ldg(&p.out_lut[i]);
  __syncthreads();
  EncTables t{s_lin, s_compn, s_olut, p.out_clamp_lo, (float) (1 << p.out_cells_log2)};

  const int bw = VEC ? p.w / 4 : p.w / 2;  // thread tiles per row pair
  const int rows2 = p.h / 2;
  const long long n = (long long) p.nframes * rows2 * bw;
  for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x) {
    const int bx = (int) (i % bw);
    const long long r = i / bw;
    const int rp = (int) (r % rows2);
    const long long f = r / rows2;
    const uint8_t *in0 = p.in + f * p.si + (long long) (2 * rp) * p.pi;
    const uint8_t *in1 = in0 + p.pi;
    uint8_t *y0 = p.y + f * p.sy + (long long) (2 * rp) * p.py;
    uint8_t *y1 = y0 + p.py;
    uint8_t *c = p.c + f * p.sc + (long long) rp * p.pc;
    if (VEC) {
      const uint4 a = ld_stream_u128(in0 + (long long) bx * 16);
      const uint4 b = ld_stream_u128(in1 + (long long) bx * 16);
      uint32_t yw0 = enc_luma(a.x, t) | (enc_luma(a.y, t) << 8) | (enc_luma(a.z, t) << 16) | (enc_luma(a.w, t) << 24);
      uint32_t yw1 = enc_luma(b.x, t) | (enc_luma(b.y, t) << 8) | (enc_luma(b.z, t) << 16) | (enc_luma(b.w, t) << 24);
      uint32_t cw = enc_chroma(a.x, a.y, b.x, b.y, t) | (enc_chroma(a.z, a.w, b.z, b.w, t) << 16);
      st_stream_u32(y0 + bx * 4, yw0);
      st_stream_u32(y1 + bx * 4, yw1);
      st_stream_u32(c + bx * 4, cw);
    } else {
      const uint32_t *r0 = reinterpret_cast<const uint32_t *>(in0) + bx * 2;
      const uint32_t *r1 = reinterpret_cast<const uint32_t *>(in1) + bx * 2;
      const uint32_t p1 = __ldg(r0), p2 = __ldg(r0 + 1), p3 = __ldg(r1), p4 = __ldg(r1 + 1);
      y0[bx * 2] = (uint8_t) enc_luma(p1, t);
      y0[bx * 2 + 1] = (uint8_t) enc_luma(p2, t);
      y1[bx * 2] = (uint8_t) enc_luma(p3, t);
      y1[bx * 2 + 1] = (uint8_t) enc_luma(p4, t);
      const uint32_t cc = enc_chroma(p1, p2, p3, p4, t);
      c[bx * 2] = (uint8_t) (cc & 255);
      c[bx * 2 + 1] = (uint8_t) (cc >> 8);
    }
  }
}

// ------------------------------------------------------------------------- encode, fast
// Same arithmetic as k_encode, organised like the decode: persistent CTAs of 1024 threads, a warp
// owns a 128-pixel x 2-row tile (lane l: pixels 4l..4l+3 of both rows = two 2x2 blocks), the warps
// of a CTA cover whole row pairs (bands), addresses are 32-bit unit offsets.  Per warp step:
// 2 x LDG.128 (512 B per row), 2 x STG.32 of Y (128 B per row), 1 x STG.32 of CbCr (128 B).
//   * both 256-entry byte tables are merged into one {linear, requantised/255} float2 table with one
//     8-byte replica per lane (entry stride 256 B), so the LDS.64 gathers are conflict-free and the
//     address (byte << 8) | (lane << 3) is a single PRMT of the pixel word and the lane register;
//   * float32 steps run two at a time on the packed FADD2/FMUL2/FFMA2 pipe.  ptxas contracts a
//     packed multiply feeding a packed add into one FFMA2 (one rounding instead of two), so every
//     add that follows a multiply is written fma(product, one, addend) with a run-time 1.0;
//   * x / 1.8556f and x / 1.5748f are q0 = x*rc; r = fma(-q0, c, x); q = fma(r, rc, q0), which equals
//     the IEEE quotient for every value this path can produce (all 2^24 8-bit RGB triples checked
//     on the host: tests/test_tables_cpu.py) and costs 3 instructions instead of ~10;
//   * round() of the positive results is add.rz(+0.5) then add.rz(+2^23): the code is the low byte.
struct EncK {
  f32x2 one, kr, kg, kb, s219, o16, s224, o128, rcb, ncb, rcr, ncr, quarter, half, magic, inv255, nmagic;
};

__device__ __forceinline__ f32x2 add2_rz(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// exact a + b where a (or b) may be the result of a packed multiply: never contracted
__device__ __forceinline__ f32x2 add2_nc(f32x2 a, f32x2 b, const EncK &k) { return fma2_rn(a, k.one, b); }

// Ey of BT709.h:223 for two pixels, then round(Ey*219 + 16) (BT709.h:234, 245): magic floats 2^23 + Y
__device__ __forceinline__ f32x2 enc_luma2(f32x2 rn, f32x2 gn, f32x2 bn, const EncK &k) {
  f32x2 ey = add2_nc(add2_nc(mul2(k.kr, rn), mul2(k.kg, gn), k), mul2(k.kb, bn), k);
  f32x2 ay = add2_nc(mul2(ey, k.s219), k.o16, k);
  return add2_rz(add2_rz(ay, k.half), k.magic);
}

template <bool WRAP>
__global__ void __launch_bounds__(1024, 1) k_encode_fast(const EncodeArgs p) {
  extern __shared__ __align__(16) uint32_t s_mem[];
  // [0, 64 KB): float2 table, entry b replica r at byte b*256 + r*8;  then the chroma cell LUT
  float2 *s_tab = reinterpret_cast<float2 *>(s_mem);
  uint32_t *s_olut = s_mem + 256 * 64;
  const int nl = (1 << p.out_cells_log2) + 1;
  for (int i = threadIdx.x; i < 256 * 32; i += blockDim.x) {
    const int b = i >> 5;
    s_tab[i] = make_float2(__ldg(&p.lin256[b]), __ldg(&p.compn256[b]));
  }
  // chroma cell LUT replicated 8x (entry i replica r at word i*8 + r): random means otherwise
  // cost ~3 wavefronts per lookup and the kernel is bound by shared-memory bandwidth
  for (int i = threadIdx.x; i < nl * 8; i += blockDim.x) s_olut[i] = __ldg(&p.out_lut[i >> 3]);
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const unsigned warp = threadIdx.x >> 5;
  const uint32_t tab = (uint32_t) __cvta_generic_to_shared(s_tab);  // warp-uniform
  const uint32_t lane8 = lane << 3;
  const uint32_t olut = (uint32_t) __cvta_generic_to_shared(s_olut) + ((lane & 7) << 2) - (0x4B000000u << 5);
  const float clamp_lo = p.out_clamp_lo;
  const float fN = (float) (1 << p.out_cells_log2);
  EncK k;
  auto bc = [](float v) { return pk(v, v); };
  k.one = bc(p.one); k.kr = bc(0.2126f); k.kg = bc(0.7152f); k.kb = bc(0.0722f);
  k.s219 = bc(219.0f); k.o16 = bc(16.0f); k.s224 = bc(224.0f); k.o128 = bc(128.0f);
  k.rcb = bc(1.0f / 1.8556f); k.ncb = bc(-1.8556f); k.rcr = bc(1.0f / 1.5748f); k.ncr = bc(-1.5748f);
  k.quarter = bc(0.25f); k.half = bc(0.5f); k.magic = bc(8388608.0f); k.nmagic = bc(-8388608.0f);
  k.inv255 = bc(1.0f / 255.0f);

  const unsigned r_local = warp / p.tiles_per_row;
  if (r_local >= (unsigned) p.rps) return;
  const unsigned tile = warp - r_local * p.tiles_per_row;
  const unsigned rows2 = p.h / 2;
  const unsigned total_rp = (unsigned) p.nframes * rows2;
  const unsigned rpg0 = blockIdx.x * p.rps + r_local;
  if (rpg0 >= total_rp) return;
  const unsigned f0 = rpg0 / rows2, rp0 = rpg0 - f0 * rows2;
  const unsigned nsteps = (total_rp - rpg0 + p.step_rp - 1) / p.step_rp;
  const unsigned col4 = tile * 32 + lane;
  unsigned oi = f0 * p.si16 + rp0 * 2 * p.pi16 + col4;  // BGRA in, 16-byte units
  unsigned oy = f0 * p.sy4 + rp0 * 2 * p.py4 + col4;    // Y out, 4-byte units
  unsigned oc = f0 * p.sc4 + rp0 * p.pc4 + col4;        // CbCr out, 4-byte units
  unsigned rp = rp0;

  auto fetch = [&](uint32_t px, int shift, float &lin, float &cn) {
    // table offset (byte << 8) | (lane << 3): byte 1 = the pixel's channel byte, byte 0 = lane * 8
    const uint32_t off = __byte_perm(px, lane8, shift == 16 ? 0x7724 : shift == 8 ? 0x7714 : 0x7704);  // bytes 2,3 = 0 (upper bytes of lane8)
    float2 v;
    asm("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(tab + off));
    lin = v.x;
    cn = v.y;
  };
  auto code_of = [&](float mean) -> float {  // BT709_from_linear code as a magic float 2^23 + code
    float xc = fmaxf(mean, clamp_lo), t;
    asm("fma.rz.f32 %0, %1, %2, 0f4B000000;" : "=f"(t) : "f"(xc), "f"(fN));
    uint32_t e;
    asm("ld.shared.u32 %0, [%1];" : "=r"(e) : "r"((__float_as_uint(t) << 5) + olut));
    return __uint_as_float(__byte_perm(e + __float_as_uint(xc), 0x4B000000u, 0x7543));
  };

  // one-step register prefetch: the loads of step k+1 are issued before the math of step k
  // (without it 40 % of the warp samples sat on the first use of the loaded pixels)
  unsigned rpl = rp0;
  uint4 na = ld_stream_u128(ptr_mad(p.in, oi, 16));
  uint4 nb = ld_stream_u128(ptr_mad(p.in, oi + p.pi16, 16));
  for (unsigned kk = 0; kk < nsteps; kk++) {
    const uint4 a = na, b = nb;
    rpl += p.d_rp;
    oi += p.di;
    if (WRAP && rpl >= rows2) { rpl -= rows2; oi += p.wi; }
    if (kk + 1 < nsteps) {
      na = ld_stream_u128(ptr_mad(p.in, oi, 16));
      nb = ld_stream_u128(ptr_mad(p.in, oi + p.pi16, 16));
    }
    // Walk the four pixel pairs in the order NW, NE (row 0), SW, SE (row 1): pair (0,2) holds the
    // west pixels of blocks A and B, pair (1,3) the east pixels.  Luma is finished per pair; the
    // linear-light sums accumulate in the reference's order ((NW + NE) + SW) + SE (BT709.h:1182-1183).
    uint32_t yw[2];
    f32x2 sum[3];
#pragma unroll
    for (int r = 0; r < 2; r++) {
      const uint32_t p0 = r ? b.x : a.x, p1 = r ? b.y : a.y, p2 = r ? b.z : a.z, p3 = r ? b.w : a.w;
      f32x2 m[2];
#pragma unroll
      for (int e = 0; e < 2; e++) {  // e = 0: pixels (0,2); e = 1: pixels (1,3)
        const uint32_t pw = e ? p1 : p0, pe = e ? p3 : p2;
        float lw[3], cw_[3], le[3], ce[3];
        fetch(pw, 16, lw[0], cw_[0]); fetch(pw, 8, lw[1], cw_[1]); fetch(pw, 0, lw[2], cw_[2]);
        fetch(pe, 16, le[0], ce[0]); fetch(pe, 8, le[1], ce[1]); fetch(pe, 0, le[2], ce[2]);
        m[e] = enc_luma2(pk(cw_[0], ce[0]), pk(cw_[1], ce[1]), pk(cw_[2], ce[2]), k);
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
          const f32x2 v = pk(lw[ch], le[ch]);
          sum[ch] = (r == 0 && e == 0) ? v : add2(sum[ch], v);
        }
      }
      uint32_t y0, y1, y2, y3;
      upku(m[0], y0, y2); upku(m[1], y1, y3);
      yw[r] = __byte_perm(__byte_perm(y0, y1, 0x0040), __byte_perm(y2, y3, 0x0040), 0x5410);
    }
    // ---- chroma of blocks A and B: mean, requantise to 8 bits in the output gamma
    // (BT709.h:1418-1420), back to unorm (sRGB.h:32-36)
    f32x2 q[3];
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
      float mA, mB;
      upk(mul2(sum[ch], k.quarter), mA, mB);
      q[ch] = mul2(add2(pk(code_of(mA), code_of(mB)), k.nmagic), k.inv255);
    }
    f32x2 ey = add2_nc(add2_nc(mul2(k.kr, q[0]), mul2(k.kg, q[1]), k), mul2(k.kb, q[2]), k);
    f32x2 xb = add2(q[2], mul2(ey, bc(-1.0f))), xr = add2(q[0], mul2(ey, bc(-1.0f)));
    f32x2 qb = mul2(xb, k.rcb), qr = mul2(xr, k.rcr);
    qb = fma2_rn(fma2_rn(qb, k.ncb, xb), k.rcb, qb);
    qr = fma2_rn(fma2_rn(qr, k.ncr, xr), k.rcr, qr);
    f32x2 mb = add2_rz(add2_rz(add2_nc(mul2(qb, k.s224), k.o128, k), k.half), k.magic);
    f32x2 mr = add2_rz(add2_rz(add2_nc(mul2(qr, k.s224), k.o128, k), k.half), k.magic);
    uint32_t cbA, cbB, crA, crB;
    upku(mb, cbA, cbB); upku(mr, crA, crB);
    const uint32_t cw = __byte_perm(__byte_perm(cbA, crA, 0x0040), __byte_perm(cbB, crB, 0x0040), 0x5410);
    st_stream_u32((void *) ptr_mad(p.y, oy, 4), yw[0]);
    st_stream_u32((void *) ptr_mad(p.y, oy + p.py4, 4), yw[1]);
    st_stream_u32((void *) ptr_mad(p.c, oc, 4), cw);
    rp += p.d_rp;
    oy += p.dy; oc += p.dc;
    if (WRAP && rp >= rows2) { rp -= rows2; oy += p.wy; oc += p.wc; }
  }
}

static int launch_encode_fast(const EncodeArgs &a0, int sm_count, cudaStream_t s) {
  EncodeArgs a = a0;
  const int rows2 = a.h / 2;
  const int wpc = 32;
  a.tiles_per_row = a.w / 128;
  a.rps = wpc / a.tiles_per_row;
  const bool contiguous = a.nframes == 1 || (a.si == (long long) a.h * a.pi && a.sy == (long long) a.h * a.py &&
                                             a.sc == (long long) rows2 * a.pc);
  bool wrap = true;
  if (contiguous) { a.h *= a.nframes; a.nframes = 1; wrap = false; }
  const unsigned total_rp = (unsigned) a.nframes * (a.h / 2);
  unsigned grid = (total_rp + a.rps - 1) / a.rps;
  if (grid > (unsigned) sm_count) grid = sm_count;
  a.step_rp = grid * a.rps;
  const unsigned r2 = a.h / 2, d_f = a.step_rp / r2;
  a.d_rp = a.step_rp - d_f * r2;
  a.pi16 = (unsigned) (a.pi / 16); a.py4 = (unsigned) (a.py / 4); a.pc4 = (unsigned) (a.pc / 4);
  a.si16 = (unsigned) (a.si / 16); a.sy4 = (unsigned) (a.sy / 4); a.sc4 = (unsigned) (a.sc / 4);
  a.di = d_f * a.si16 + a.d_rp * 2 * a.pi16; a.dy = d_f * a.sy4 + a.d_rp * 2 * a.py4; a.dc = d_f * a.sc4 + a.d_rp * a.pc4;
  a.wi = a.si16 - r2 * 2 * a.pi16; a.wy = a.sy4 - r2 * 2 * a.py4; a.wc = a.sc4 - r2 * a.pc4;
  a.one = 1.0f;
  const size_t smem = 65536 + ((size_t) (1 << a.out_cells_log2) + 1) * 32;
  if (wrap)
    k_encode_fast<true><<<grid, 1024, smem, s>>>(a);
  else
    k_encode_fast<false><<<grid, 1024, smem, s>>>(a);
  return (int) cudaGetLastError();
}

int launch_encode(const EncodeArgs &a, int sm_count, void *stream) {
  cudaStream_t s = (cudaStream_t) stream;
  const size_t smem = (512 + (size_t) (1 << a.out_cells_log2) + 1) * 4;
  const bool vec = (a.w % 4 == 0) && aligned(a.in, 16) && aligned(a.pi, 16) && aligned(a.si, 16) && aligned(a.y, 4) &&
                   aligned(a.py, 4) && aligned(a.sy, 4) && aligned(a.c, 4) && aligned(a.pc, 4) && aligned(a.sc, 4);
  // fast path: whole 128-pixel tiles, at most 32 tiles per row, offsets that fit 32-bit units
  const unsigned long long in_bytes = (unsigned long long) a.nframes * (a.si > 0 ? a.si : (long long) a.h * a.pi);
  if (vec && a.w % 128 == 0 && a.w <= 4096 && in_bytes < 0xFFFF0000ull * 16 && a.si >= 0 && a.sy >= 0 && a.sc >= 0 &&
      a.out_cells_log2 <= 12)
    return launch_encode_fast(a, sm_count, s);
  const long long n = (long long) a.nframes * (a.h / 2) * (vec ? a.w / 4 : a.w / 2);
  long long grid = (n + 255) / 256;
  if (grid > (long long) sm_count * 8) grid = (long long) sm_count * 8;
  if (vec)
    k_encode<true><<<(unsigned) grid, 256, smem, s>>>(a);
  else
    k_encode<false><<<(unsigned) grid, 256, smem, s>>>(a);
  return (int) cudaGetLastError();
}

// ===================================================================== packed 4:4:4 harness
// +convertSoftware: (BGRAToBT709Converter.m:89-144): Apple196_from_sRGB_convertRGBToYCbCr per pixel
__global__ void __launch_bounds__(256) k_convert444(const Packed444Args p) {
  __shared__ float s_enc[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_enc[i] = __ldg(&p.enc256[i]);
  __syncthreads();
  for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (long long) gridDim.x * blockDim.x) {
    const uint32_t px = p.in[i];
    float rn = s_enc[(px >> 16) & 255], gn = s_enc[(px >> 8) & 255], bn = s_enc[px & 255];
    float ey = __fadd_rn(__fadd_rn(__fmul_rn(0.2126f, rn), __fmul_rn(0.7152f, gn)), __fmul_rn(0.0722f, bn));
    float eb = __fdiv_rn(__fsub_rn(bn, ey), 1.8556f);
    float er = __fdiv_rn(__fsub_rn(rn, ey), 1.5748f);
    uint32_t Y = (uint32_t) round_pos(__fadd_rn(__fmul_rn(ey, 219.0f), 16.0f));
    uint32_t Cb = (uint32_t) round_pos(__fadd_rn(__fmul_rn(eb, 224.0f), 128.0f));
    uint32_t Cr = (uint32_t) round_pos(__fadd_rn(__fmul_rn(er, 224.0f), 128.0f));
    p.out[i] = (Cr << 16) | (Cb << 8) | Y;
  }
}

// +unconvertSoftware: (BGRAToBT709Converter.m:146-198): Apple196_to_sRGB per pixel, A = 0
__global__ void __launch_bounds__(256) k_unconvert444(const Packed444Args p) {
  const float fN = (float) (1 << p.cells_log2);
  for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (long long) gridDim.x * blockDim.x) {
    const uint32_t px = p.in[i];
    float yt = __fmul_rn(__fmul_rn((float) (px & 255) - 16.0f, p.m.inv255), p.m.ys);
    float cbn = __fmul_rn((float) ((px >> 8) & 255) - 128.0f, p.m.inv255);
    float crn = __fmul_rn((float) ((px >> 16) & 255) - 128.0f, p.m.inv255);
    float xr = add_sat(yt, __fmul_rn(crn, p.m.r_cr));
    float xg = add_sat(__fadd_rn(yt, __fmul_rn(cbn, p.m.g_cb)), __fmul_rn(crn, p.m.g_cr));
    float xb = add_sat(yt, __fmul_rn(cbn, p.m.b_cb));
    uint32_t R = lut_code(p.lut, xr, p.clamp_lo, fN), G = lut_code(p.lut, xg, p.clamp_lo, fN),
             B = lut_code(p.lut, xb, p.clamp_lo, fN);
    p.out[i] = (R << 16) | (G << 8) | B;
  }
}

// +copyBT709ToCoreVideo: (BGRAToBT709Converter.m:1042-1099): the chroma of a 2x2 block ends up
// being that of its last row's even column (the odd row overwrites the even row's write).
__global__ void __launch_bounds__(256) k_copy_packed(const CopyPackedArgs p) {
  const long long n = (long long) p.w * p.h;
  for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x) {
    const int col = (int) (i % p.w), row = (int) (i / p.w);
    const uint32_t px = p.in[i];
    p.y[(long long) row * p.py + col] = (uint8_t) (px & 255);
    const bool last_row_of_block = (row & 1) || row == p.h - 1;
    if ((col & 1) == 0 && last_row_of_block) {
      uint8_t *c = p.c + (long long) (row >> 1) * p.pc + col;
      c[0] = (uint8_t) ((px >> 8) & 255);
      c[1] = (uint8_t) ((px >> 16) & 255);
    }
  }
}

// ======================================================================= NV12 <-> planar chroma
// +copyYCBCr: (BGRAToBT709Converter.m:688-790) splits CbCr pairs into U and V planes for the .y4m
// writer; the inverse feeds .y4m frames to the decoder.  Pure byte shuffles, 2 B/sample each way.
template <bool SPLIT>
__global__ void __launch_bounds__(256) k_cbcr_planar(const PlanarArgs p) {
  const long long n = (long long) p.nframes * p.ch * p.cw;
  for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x) {
    const int x = (int) (i % p.cw);
    const long long r = i / p.cw;
    const int y = (int) (r % p.ch);
    const long long f = r / p.ch;
    uint8_t *c = p.c + f * p.sc + (long long) y * p.pc + 2 * x;
    uint8_t *u = p.u + f * p.su + (long long) y * p.pu + x;
    uint8_t *v = p.v + f * p.sv + (long long) y * p.pv + x;
    if (SPLIT) { *u = c[0]; *v = c[1]; } else { c[0] = *u; c[1] = *v; }
  }
}

int launch_cbcr_planar(const PlanarArgs &a, bool split, void *stream) {
  const long long n = (long long) a.nframes * a.ch * a.cw;
  long long g = (n + 255) / 256;
  g = g > 148 * 16 ? 148 * 16 : (g < 1 ? 1 : g);
  if (split)
    k_cbcr_planar<true><<<(unsigned) g, 256, 0, (cudaStream_t) stream>>>(a);
  else
    k_cbcr_planar<false><<<(unsigned) g, 256, 0, (cudaStream_t) stream>>>(a);
  return (int) cudaGetLastError();
}

static unsigned grid_1d(long long n) {
  long long g = (n + 255) / 256;
  return (unsigned) (g > 148 * 16 ? 148 * 16 : (g < 1 ? 1 : g));
}

int launch_convert444(const Packed444Args &a, void *stream) {
  k_convert444<<<grid_1d(a.n), 256, 0, (cudaStream_t) stream>>>(a);
  return (int) cudaGetLastError();
}
int launch_unconvert444(const Packed444Args &a, void *stream) {
  k_unconvert444<<<grid_1d(a.n), 256, 0, (cudaStream_t) stream>>>(a);
  return (int) cudaGetLastError();
}
int launch_copy_packed(const CopyPackedArgs &a, void *stream) {
  k_copy_packed<<<grid_1d((long long) a.w * a.h), 256, 0, (cudaStream_t) stream>>>(a);
  return (int) cudaGetLastError();
}

template <int CL2, int RL2, bool ALPHA>
static int optin_decode_smem() {
  constexpr size_t smem = (size_t) ((1 << CL2) + 1) * (1 << RL2) * 4;
  int e = (int) cudaFuncSetAttribute(k_decode_fast<CL2, RL2, ALPHA, true, true, kPrefetchDepth, kDecodeThreads>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
  if (e) return e;
  e = (int) cudaFuncSetAttribute(k_decode_fast<CL2, RL2, ALPHA, true, false, kPrefetchDepth, kDecodeThreads>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
  if (e) return e;
  e = (int) cudaFuncSetAttribute(k_decode_ring<CL2, RL2, ALPHA, true, kRingDepth, kDecodeThreads>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kMaxDynSmem);
  if (e) return e;
  e = (int) cudaFuncSetAttribute(k_decode_ring<CL2, RL2, ALPHA, false, kRingDepth, kDecodeThreads>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kMaxDynSmem);
  if (e) return e;
  e = (int) cudaFuncSetAttribute(k_decode_bulk<CL2, RL2, ALPHA, kDecodeThreads>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kMaxDynSmem);
  if (e) return e;
  return (int) cudaFuncSetAttribute(k_decode_fast<CL2, RL2, ALPHA, false, true, kPrefetchDepth, kDecodeThreads>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
}

// per device (called from bt709_ctx_create with the device current): opt in to > 48 KB of
// dynamic shared memory for the replicated-LUT decode kernels
int kernels_init() {
  int e;
  if ((e = (int) cudaFuncSetAttribute(k_rescale_cols, cudaFuncAttributeMaxDynamicSharedMemorySize, (256 * 64 + 8193) * 4))) return e;
  if ((e = (int) cudaFuncSetAttribute(k_encode_fast<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 4097 * 32))) return e;
  if ((e = (int) cudaFuncSetAttribute(k_encode_fast<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 4097 * 32))) return e;
  if ((e = optin_decode_smem<10, 5, false>())) return e;
  if ((e = optin_decode_smem<10, 5, true>())) return e;
  if ((e = optin_decode_smem<12, 3, false>())) return e;
  if ((e = optin_decode_smem<12, 3, true>())) return e;
  return 0;
}

}  // namespace bt709
