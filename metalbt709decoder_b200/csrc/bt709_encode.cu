// bt709_encode.cu — inverse path: sRGB BGRA8 -> BT.709 4:2:0 NV12 with 2x2 linear-light chroma
// averaging (cvpbu_ycbcr_subsample + BT709_average_pixel_values).
#include "bt709_device.cuh"

namespace bt709 {

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
//   * the chroma steps of the lane's two blocks run two at a time on the packed FADD2/FMUL2/FFMA2 pipe.
//     ptxas contracts a packed multiply feeding a packed add into one FFMA2 (one rounding instead of
//     two), so every add that follows a multiply is written fma(product, one, addend) with a run-time 1.0;
//   * x / 1.8556f and x / 1.5748f are q0 = x*rc; r = fma(-q0, c, x); q = fma(r, rc, q0), which equals
//     the IEEE quotient for every value this path can produce (all 2^24 8-bit RGB triples checked
//     on the host: tests/test_tables_cpu.py) and costs 3 instructions instead of ~10;
//   * round() of the positive results is one add.rz(+32768.5): the code is byte 1 of the float.
#ifndef BT709_ENC_FENCE
#define BT709_ENC_FENCE 2
#endif
constexpr int kEncFence = BT709_ENC_FENCE;  // pixels per dependence group

struct EncK {
  f32x2 one, kr, kg, kb, s224, o128, rcb, ncb, rcr, ncr, quarter, rnd, inv255, nmagic;
};

__device__ __forceinline__ f32x2 add2_rz(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// exact a + b where a (or b) may be the result of a packed multiply: never contracted
__device__ __forceinline__ f32x2 add2_nc(f32x2 a, f32x2 b, const EncK &k) { return fma2_rn(a, k.one, b); }

// FULL = (w % 128 == 0): no per-lane predicates; otherwise the lanes past the row end load nothing and store nothing
// WIDE: rows of more than 32 tiles (4096 pixels): the warps loop over the tile columns, 32 apart
// OMODE: form of the chroma LUT (BT709_from_linear as a cell LUT).  ncu on 385 x 1080p (session 54): 30 shared loads per
// warp-step but 65.9 shared wavefronts, 11.9 of them bank conflicts - all in the six chroma lookups, whose 4096-cell
// uniform LUT only fits with 8 replicas (random means: ~3 wavefronts per lookup), on a kernel that keeps the L1 / shared
// pipe 96.7 % busy.  0: uniform cells, 8 replicas (the fallback); 1: uniform cells, one replica per lane (curves that need
// <= 256 cells: the Linear output); 2: biased log cells (645 for the Apple and sRGB curves), one replica per lane - one
// PRMT more per lookup, no conflicts for any content.
template <bool WRAP, bool FULL, bool WIDE, int OMODE>
__global__ void __launch_bounds__(1024, 1) k_encode_fast(const EncodeArgs p) {
  extern __shared__ __align__(16) uint32_t s_mem[];
  // [0, 64 KB): float2 table, entry b replica r at byte b*256 + r*8;  then the chroma cell LUT
  float2 *s_tab = reinterpret_cast<float2 *>(s_mem);
  uint32_t *s_olut = s_mem + 256 * 64;
  const int nl = OMODE == 2 ? p.out_blog_cells : (1 << p.out_cells_log2) + 1;
  pdl_launch_dependents();
  // every table entry is read from global memory once and its replicas are written with 16-byte stores (the old
  // loops read an entry once per replica: ~10 dependent L2 round trips per thread, a third of a one-frame call)
  for (int b = threadIdx.x >> 4; b < 256; b += blockDim.x >> 4) {  // 16 lanes per entry: 32 x {lin, compn} = 256 bytes
    const float l = __ldg(&p.lin256[b]), cn = __ldg(&p.compn256[b]);
    reinterpret_cast<float4 *>(s_tab + b * 32)[threadIdx.x & 15] = make_float4(l, cn, l, cn);
  }
  // chroma cell LUT: entry i replica r at word i*R + r (R = 8 or 32, see OMODE)
  if (OMODE == 0) fill_replicated<3>(s_olut, 8, p.out_lut, nl);
  else fill_replicated<5>(s_olut, 32, OMODE == 2 ? p.out_blog_lut : p.out_lut, nl);
  __syncthreads();
  pdl_wait();  // small calls are launched with programmatic stream serialization: the tables above are constant

  const int lane = threadIdx.x & 31;
  const unsigned warp = threadIdx.x >> 5;
  const uint32_t tab = (uint32_t) __cvta_generic_to_shared(s_tab);  // warp-uniform
  const uint32_t lane8 = lane << 3;
  // lookup base: replica of this lane, minus the bias of the index register (2^23 magic for uniform cells, the first
  // cell for log cells), so that a lookup address is (index register << shift) + olut
  const uint32_t olut = (uint32_t) __cvta_generic_to_shared(s_olut) +
                        (OMODE == 0 ? ((lane & 7) << 2) - (0x4B000000u << 5)
                                    : (lane << 2) - ((OMODE == 2 ? p.out_blog_first : 0x4B000000u) << 7));
  const float clamp_lo = OMODE == 2 ? p.out_blog_clamp_lo : p.out_clamp_lo;
  const float obias = p.out_blog_bias;
  const float fN = (float) (1 << p.out_cells_log2);
  EncK k;
  auto bc = [](float v) { return pk(v, v); };
  k.one = bc(p.one); k.kr = bc(0.2126f); k.kg = bc(0.7152f); k.kb = bc(0.0722f);
  k.s224 = bc(224.0f); k.o128 = bc(128.0f);
  k.rcb = bc(1.0f / 1.8556f); k.ncb = bc(-1.8556f); k.rcr = bc(1.0f / 1.5748f); k.ncr = bc(-1.5748f);
  k.quarter = bc(0.25f); k.rnd = bc(32768.5f); k.nmagic = bc(-8388608.0f);
  k.inv255 = bc(1.0f / 255.0f);

  unsigned tile0, rpg0;
  if (p.lin_warps) {  // linear tile order: every warp of the CTA has a tile column
    const unsigned g = blockIdx.x * (blockDim.x >> 5) + warp;
    if (g >= (unsigned) p.lin_warps) return;
    rpg0 = g / p.tiles_per_row;
    tile0 = g - rpg0 * p.tiles_per_row;
  } else {
    const unsigned r_local = warp / p.tiles_per_row;
    if (r_local >= (unsigned) p.rps) return;
    tile0 = warp - r_local * p.tiles_per_row;
    rpg0 = blockIdx.x * p.rps + r_local;
  }
  const unsigned rows2 = p.h / 2;
  const unsigned total_rp = (unsigned) p.nframes * rows2;
  if (rpg0 >= total_rp) return;
  const unsigned f0 = rpg0 / rows2, rp0 = rpg0 - f0 * rows2;
  const unsigned nsteps = (total_rp - rpg0 + p.step_rp - 1) / p.step_rp;
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
    uint32_t e;
    if (OMODE == 2) {  // cell of xc + bias: the upper half of the float (PRMT, not a shift: see lut2 in bt709_decode_row.cuh)
      const uint32_t cell = __byte_perm(__float_as_uint(__fadd_rn(xc, obias)), 0, 0x4432);
      asm("ld.shared.u32 %0, [%1];" : "=r"(e) : "r"((cell << 7) + olut));
    } else {
      asm("fma.rz.f32 %0, %1, %2, 0f4B000000;" : "=f"(t) : "f"(xc), "f"(fN));
      asm("ld.shared.u32 %0, [%1];" : "=r"(e) : "r"((__float_as_uint(t) << (OMODE ? 7 : 5)) + olut));
    }
    return __uint_as_float(__byte_perm(e + __float_as_uint(xc), 0x4B000000u, 0x7543));
  };

  for (unsigned tile = tile0; WIDE ? tile < (unsigned) p.tiles_per_row : tile == tile0; tile += WIDE ? p.tile_stride : 1u) {
  const unsigned col4 = tile * 32 + lane;
  const bool valid = FULL || (int) (col4 * 4) < p.w;
  // w % 4 == 2: the last lane of a row owns one 2x2 block only and stores 2 bytes per plane row; its 16-byte
  // loads stay inside the row pitch (a multiple of 16 on this path)
  const bool half = !FULL && (int) (col4 * 4 + 2) == p.w;
  unsigned oi = f0 * p.si16 + rp0 * 2 * p.pi16 + col4;  // BGRA in, 16-byte units
  unsigned oy = f0 * p.sy4 + rp0 * 2 * p.py4 + col4;    // Y out, 4-byte units
  unsigned oc = f0 * p.sc4 + rp0 * p.pc4 + col4;        // CbCr out, 4-byte units
  unsigned rp = rp0;

  // one-step register prefetch: the loads of step k+1 are issued before the math of step k
  // (without it 40 % of the warp samples sat on the first use of the loaded pixels)
  unsigned rpl = rp0;
  uint4 na = make_uint4(0, 0, 0, 0), nb = na;
  if (valid) {
    na = ld_stream_u128(ptr_mad(p.in, oi, 16));
    nb = ld_stream_u128(ptr_mad(p.in, oi + p.pi16, 16));
  }
  for (unsigned kk = 0; kk < nsteps; kk++) {
    const uint4 a = na, b = nb;
    rpl += p.d_rp;
    oi += p.di;
    if (WRAP && rpl >= rows2) { rpl -= rows2; oi += p.wi; }
    if (kk + 1 < nsteps && valid) {
      na = ld_stream_u128(ptr_mad(p.in, oi, 16));
      nb = ld_stream_u128(ptr_mad(p.in, oi + p.pi16, 16));
    }
    // Per pixel: three LDS.64 {linear, requantised/255}; the luma runs on scalar float ops straight
    // from the loaded registers (packing two pixels into f32x2 pairs cost two MOVs per pair because
    // an LDS.64 dictates where its results land: 23 % of the loop was register moves).  The
    // linear-light sums accumulate in the reference's order ((NW + NE) + SW) + SE (BT709.h:1182-1183).
    uint32_t yw[2];
    float sum[2][3];  // [block A/B][R,G,B]
    uint32_t fence = 0;  // always 0; carries a data dependence from one pixel group to the next
#pragma unroll
    for (int r = 0; r < 2; r++) {
      float ty[4];
      const uint4 row = r ? b : a;
#pragma unroll
      for (int i = 0; i < 4; i++) {
        // Unrolled, ptxas hoists all 24 LDS.64 of the step (48 result registers) and spills the loop state; a
        // rolled row loop costs ~35 instructions of control per step.  Instead every FENCE-th pixel word is
        // OR-ed with (bits of the previous luma & zero), zero being a run-time 0: one LOP3 that changes
        // nothing but makes the pixel's table addresses depend on the previous group's result.
        uint32_t px = i == 0 ? row.x : i == 1 ? row.y : i == 2 ? row.z : row.w;
        if ((r * 4 + i) % kEncFence == 0 && (r | i)) px |= fence;
        float l[3], c[3];
        fetch(px, 16, l[0], c[0]); fetch(px, 8, l[1], c[1]); fetch(px, 0, l[2], c[2]);
        // Ey of BT709.h:223, then round(Ey*219 + 16) (BT709.h:234, 245): the integer part of
        // RZ(x + 32768.5) sits in byte 1 of the float
        const float ey = __fadd_rn(__fadd_rn(__fmul_rn(0.2126f, c[0]), __fmul_rn(0.7152f, c[1])), __fmul_rn(0.0722f, c[2]));
        asm("add.rz.f32 %0, %1, 0f47000080;" : "=f"(ty[i]) : "f"(__fadd_rn(__fmul_rn(ey, 219.0f), 16.0f)));
        if ((r * 4 + i + 1) % kEncFence == 0) fence = __float_as_uint(ty[i]) & p.zero;
#pragma unroll
        for (int ch = 0; ch < 3; ch++)
          sum[i >> 1][ch] = (r == 0 && (i & 1) == 0) ? l[ch] : __fadd_rn(sum[i >> 1][ch], l[ch]);
      }
      yw[r] = __byte_perm(__byte_perm(__float_as_uint(ty[0]), __float_as_uint(ty[1]), 0x0051),
                          __byte_perm(__float_as_uint(ty[2]), __float_as_uint(ty[3]), 0x0051), 0x5410);
    }
    // ---- chroma of blocks A and B (packed A|B): mean, requantise to 8 bits in the output gamma
    // (BT709.h:1418-1420), back to unorm (sRGB.h:32-36)
    f32x2 q[3];
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
      float mA, mB;
      upk(mul2(pk(sum[0][ch], sum[1][ch]), k.quarter), mA, mB);
      q[ch] = mul2(add2(pk(code_of(mA), code_of(mB)), k.nmagic), k.inv255);
    }
    f32x2 ey = add2_nc(add2_nc(mul2(k.kr, q[0]), mul2(k.kg, q[1]), k), mul2(k.kb, q[2]), k);
    f32x2 xb = add2(q[2], mul2(ey, bc(-1.0f))), xr = add2(q[0], mul2(ey, bc(-1.0f)));
    f32x2 qb = mul2(xb, k.rcb), qr = mul2(xr, k.rcr);
    qb = fma2_rn(fma2_rn(qb, k.ncb, xb), k.rcb, qb);
    qr = fma2_rn(fma2_rn(qr, k.ncr, xr), k.rcr, qr);
    f32x2 mb = add2_rz(add2_nc(mul2(qb, k.s224), k.o128, k), k.rnd);
    f32x2 mr = add2_rz(add2_nc(mul2(qr, k.s224), k.o128, k), k.rnd);
    uint32_t cbA, cbB, crA, crB;
    upku(mb, cbA, cbB); upku(mr, crA, crB);
    const uint32_t cw = __byte_perm(__byte_perm(cbA, crA, 0x0051), __byte_perm(cbB, crB, 0x0051), 0x5410);
    if (!FULL && half) {
      *reinterpret_cast<uint16_t *>(const_cast<uint8_t *>(ptr_mad(p.y, oy, 4))) = (uint16_t) yw[0];
      *reinterpret_cast<uint16_t *>(const_cast<uint8_t *>(ptr_mad(p.y, oy + p.py4, 4))) = (uint16_t) yw[1];
      *reinterpret_cast<uint16_t *>(const_cast<uint8_t *>(ptr_mad(p.c, oc, 4))) = (uint16_t) cw;
    } else if (valid) {
      st_stream_u32((void *) ptr_mad(p.y, oy, 4), yw[0]);
      st_stream_u32((void *) ptr_mad(p.y, oy + p.py4, 4), yw[1]);
      st_stream_u32((void *) ptr_mad(p.c, oc, 4), cw);
    }
    rp += p.d_rp;
    oy += p.dy; oc += p.dc;
    if (WRAP && rp >= rows2) { rp -= rows2; oy += p.wy; oc += p.wc; }
  }
  }
}

static int launch_encode_fast1(const EncodeArgs &a0, int sm_count, cudaStream_t s);

// Every plane is addressed with 32-bit unit offsets (16 B for BGRA, 4 B for Y / CbCr): split the batch by frames so
// that no offset of any of the three planes reaches 2^32 (padded output frame strides can get there long before
// the input does), like the decode launcher.
static int launch_encode_fast(const EncodeArgs &a0, int sm_count, cudaStream_t s) {
  const unsigned long long lim = 0xFFFF0000ull;
  unsigned long long per = lim;
  auto fit = [&](long long stride, unsigned unit, long long frame_bytes) {
    if (stride <= 0 || a0.nframes == 1) return;
    if ((unsigned long long) frame_bytes >= lim * unit) { per = 1; return; }
    unsigned long long n = (lim * unit - (unsigned long long) frame_bytes) / (unsigned long long) stride + 1;
    if (n < per) per = n;
  };
  fit(a0.si, 16, (long long) a0.h * a0.pi);
  fit(a0.sy, 4, (long long) a0.h * a0.py);
  fit(a0.sc, 4, (long long) (a0.h / 2) * a0.pc);
  if (per < 1) per = 1;
  for (long long f0 = 0; f0 < a0.nframes; f0 += (long long) per) {
    EncodeArgs a = a0;
    a.nframes = (int) ((unsigned long long) (a0.nframes - f0) < per ? (unsigned long long) (a0.nframes - f0) : per);
    a.in = a0.in + f0 * a0.si;
    a.y = a0.y + f0 * a0.sy;
    a.c = a0.c + f0 * a0.sc;
    int e = launch_encode_fast1(a, sm_count, s);
    if (e) return e;
  }
  return 0;
}

static int launch_encode_fast1(const EncodeArgs &a0, int sm_count, cudaStream_t s) {
  EncodeArgs a = a0;
  const int rows2 = a.h / 2;
  const int wpc = 32;
  a.tiles_per_row = (a.w + 127) / 128;
  a.rps = a.tiles_per_row >= wpc ? 1 : wpc / a.tiles_per_row;
  a.tile_stride = a.tiles_per_row >= wpc ? wpc : a.tiles_per_row;
  const bool full = a.w % 128 == 0;
  const bool contiguous = a.nframes == 1 || (a.si == (long long) a.h * a.pi && a.sy == (long long) a.h * a.py &&
                                             a.sc == (long long) rows2 * a.pc);
  bool wrap = true;
  if (contiguous) { a.h *= a.nframes; a.nframes = 1; wrap = false; }
  const unsigned total_rp = (unsigned) a.nframes * (a.h / 2);
  unsigned grid = (total_rp + a.rps - 1) / a.rps;
  if (grid > (unsigned) sm_count) grid = sm_count;
  a.step_rp = grid * a.rps;
  // rows of 15 or 30 tiles (1920, 3840 wide) leave 2 of the 32 warps idle in the band layout; in a contiguous batch the
  // tiles can be taken in linear order instead, like the decode (session 40: 1920-wide 862 -> 912 Gpix/s expected from
  // the 32-warp geometries).  BT709_NO_ENC_LINEAR=1: band layout always.
  static const bool no_enc_linear = getenv("BT709_NO_ENC_LINEAR") != nullptr;
  a.lin_warps = 0;
  if (!wrap && !no_enc_linear && a.tiles_per_row < wpc && wpc % a.tiles_per_row != 0) {
    const unsigned long long tiles = (unsigned long long) total_rp * a.tiles_per_row;
    unsigned long long g = (tiles + wpc - 1) / wpc;
    if (g > (unsigned long long) sm_count) g = sm_count;
    const unsigned W = (unsigned) g * wpc;
    if (W >= (unsigned) a.tiles_per_row) {
      grid = (unsigned) g;
      a.lin_warps = (int) (W / a.tiles_per_row * a.tiles_per_row);
      a.step_rp = (unsigned) a.lin_warps / a.tiles_per_row;
      a.tile_stride = a.tiles_per_row;  // (the looping instantiation runs its loop once)
    }
  }
  const unsigned r2 = a.h / 2, d_f = a.step_rp / r2;
  a.d_rp = a.step_rp - d_f * r2;
  a.pi16 = (unsigned) (a.pi / 16); a.py4 = (unsigned) (a.py / 4); a.pc4 = (unsigned) (a.pc / 4);
  a.si16 = (unsigned) (a.si / 16); a.sy4 = (unsigned) (a.sy / 4); a.sc4 = (unsigned) (a.sc / 4);
  a.di = d_f * a.si16 + a.d_rp * 2 * a.pi16; a.dy = d_f * a.sy4 + a.d_rp * 2 * a.py4; a.dc = d_f * a.sc4 + a.d_rp * a.pc4;
  a.wi = a.si16 - r2 * 2 * a.pi16; a.wy = a.sy4 - r2 * 2 * a.py4; a.wc = a.sc4 - r2 * a.pc4;
  a.one = 1.0f;
  a.zero = 0;
  // chroma LUT form (see k_encode_fast): biased log cells with a replica per lane when the context has them and they fit,
  // else uniform cells with a replica per lane when the curve needs <= 256 of them, else uniform cells with 8 replicas.
  // BT709_ENC_OMODE=0|1|2 forces a form that is available (reproduction).
  static const char *omode_env = getenv("BT709_ENC_OMODE");
  int omode = 0;
  if (a.out_blog_lut && a.out_blog_cells > 0 && 65536 + (size_t) a.out_blog_cells * 128 <= kMaxDynSmem) omode = 2;
  else if (a.out_cells_log2 <= 8) omode = 1;
  if (omode_env) {
    const int want = atoi(omode_env);
    if (want == 0 || (want == 1 && a.out_cells_log2 <= 8) || (want == 2 && omode == 2)) omode = want;
  }
  const size_t smem = 65536 + (omode == 2 ? (size_t) a.out_blog_cells * 128
                                          : ((size_t) (1 << a.out_cells_log2) + 1) * (omode == 1 ? 128 : 32));
  // The looping instantiation is also valid for narrower rows (its loop runs once) and its code schedule measured
  // faster when all 32 warps have a tile (2048 wide: 905 vs 883 Gpix/s), slower with 30 (1920 wide: 881 vs 916).
  const bool wide = a.tiles_per_row > wpc || wpc % a.tiles_per_row == 0 || a.lin_warps > 0;
  // calls of a few steps per warp (one or two frames) are launched with programmatic stream serialization
  const bool pdl = pdl_enabled((unsigned long long) total_rp <= 4ull * a.step_rp);
  cudaError_t le = cudaSuccess;
#define BT709_ENC_LAUNCH1(W, F, X, O) le = launch_pdl(k_encode_fast<W, F, X, O>, grid, 1024, smem, s, pdl, a), note_launch()
#define BT709_ENC_LAUNCH(W, F)                                                      \
  do {                                                                              \
    if (wide) {                                                                     \
      if (omode == 2) BT709_ENC_LAUNCH1(W, F, true, 2); else if (omode == 1) BT709_ENC_LAUNCH1(W, F, true, 1); else BT709_ENC_LAUNCH1(W, F, true, 0);    \
    } else {                                                                        \
      if (omode == 2) BT709_ENC_LAUNCH1(W, F, false, 2); else if (omode == 1) BT709_ENC_LAUNCH1(W, F, false, 1); else BT709_ENC_LAUNCH1(W, F, false, 0); \
    }                                                                               \
  } while (0)
  if (wrap) {
    if (full) BT709_ENC_LAUNCH(true, true); else BT709_ENC_LAUNCH(true, false);
  } else {
    if (full) BT709_ENC_LAUNCH(false, true); else BT709_ENC_LAUNCH(false, false);
  }
#undef BT709_ENC_LAUNCH
#undef BT709_ENC_LAUNCH1
  if (le != cudaSuccess) return (int) le;
  return (int) cudaGetLastError();
}

int launch_encode(const EncodeArgs &a, int sm_count, void *stream) {
  cudaStream_t s = (cudaStream_t) stream;
  const size_t smem = (512 + (size_t) (1 << a.out_cells_log2) + 1) * 4;
  const bool al = aligned(a.in, 16) && aligned(a.pi, 16) && aligned(a.si, 16) && aligned(a.y, 4) && aligned(a.py, 4) &&
                  aligned(a.sy, 4) && aligned(a.c, 4) && aligned(a.pc, 4) && aligned(a.sc, 4);
  const bool vec = (a.w % 4 == 0) && aligned(a.in, 16) && aligned(a.pi, 16) && aligned(a.si, 16) && aligned(a.y, 4) &&
                   aligned(a.py, 4) && aligned(a.sy, 4) && aligned(a.c, 4) && aligned(a.pc, 4) && aligned(a.sc, 4);
  // fast path: aligned planes (any even width), offsets that fit 32-bit units
  // (one frame must fit the 32-bit unit offsets of every plane; batches are split by frames in launch_encode_fast)
  const unsigned long long lim = 0xFFFF0000ull;
  const bool frame_fits = (unsigned long long) a.h * a.pi < lim * 16 && (unsigned long long) a.h * a.py < lim * 4 &&
                          (unsigned long long) (a.h / 2) * a.pc < lim * 4;
  if (al && frame_fits && a.si >= 0 && a.sy >= 0 && a.sc >= 0 && a.out_cells_log2 <= 12)
    return launch_encode_fast(a, sm_count, s);
  const long long n = (long long) a.nframes * (a.h / 2) * (vec ? a.w / 4 : a.w / 2);
  long long grid = (n + 255) / 256;
  if (grid > (long long) sm_count * 8) grid = (long long) sm_count * 8;
  if (vec)
    k_encode<true><<<(unsigned) grid, 256, smem, s>>>(a), note_launch();
  else
    k_encode<false><<<(unsigned) grid, 256, smem, s>>>(a), note_launch();
  return (int) cudaGetLastError();
}

template <bool W, bool F, bool X, int O>
static int optin_encode_smem1() {
  return (int) cudaFuncSetAttribute(k_encode_fast<W, F, X, O>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kMaxDynSmem);
}
template <bool W, bool F>
static int optin_encode_smem() {
  int e;
  if ((e = optin_encode_smem1<W, F, false, 0>())) return e;
  if ((e = optin_encode_smem1<W, F, false, 1>())) return e;
  if ((e = optin_encode_smem1<W, F, false, 2>())) return e;
  if ((e = optin_encode_smem1<W, F, true, 0>())) return e;
  if ((e = optin_encode_smem1<W, F, true, 1>())) return e;
  return optin_encode_smem1<W, F, true, 2>();
}

int encode_kernels_init() {
  int e;
  if ((e = optin_encode_smem<true, true>())) return e;
  if ((e = optin_encode_smem<true, false>())) return e;
  if ((e = optin_encode_smem<false, true>())) return e;
  return optin_encode_smem<false, false>();
}

}  // namespace bt709
