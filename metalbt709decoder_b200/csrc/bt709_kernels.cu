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
// One warp owns a segment of 256 pixels x 2 rows (one chroma row).  Lane l handles pixels
// [4l, 4l+4) and [128+4l, 128+4l+4) of both rows, so every warp-level access is contiguous:
//   Y     2 rows x 2 x LDG.32   (128 B per instruction)
//   CbCr  2 x LDG.32            (128 B per instruction; loaded once, used for both rows)
//   BGRA  2 rows x 2 x STG.128  (512 B per instruction)
// Grid: persistent, one CTA per SM (the replicated LUT takes 128 KB of shared memory).
template <int CL2, int RL2, bool ALPHA>
__device__ __forceinline__ void decode_quad(uint32_t yw0, uint32_t yw1, uint32_t cw, uint32_t aw0, uint32_t aw1,
                                            const MatrixConsts &m, float clamp_lo, uint32_t sbase, uint4 &o0,
                                            uint4 &o1) {
  const float kY = 8388608.0f + 16.0f, kC = 8388608.0f + 128.0f;
  uint32_t r0[4], r1[4];
#pragma unroll
  for (int pr = 0; pr < 2; pr++) {
    // chroma pair pr: bytes 2pr (Cb), 2pr+1 (Cr).  BT709.h:499-500 then the matrix columns.
    float cbn = __fmul_rn(pr == 0 ? byte_minus<0>(cw, kC) : byte_minus<2>(cw, kC), m.inv255);
    float crn = __fmul_rn(pr == 0 ? byte_minus<1>(cw, kC) : byte_minus<3>(cw, kC), m.inv255);
    float cb_g = __fmul_rn(cbn, m.g_cb), cb_b = __fmul_rn(cbn, m.b_cb);
    float cr_r = __fmul_rn(crn, m.r_cr), cr_g = __fmul_rn(crn, m.g_cr);
#pragma unroll
    for (int row = 0; row < 2; row++) {
      uint32_t yw = row ? yw1 : yw0;
      uint32_t aw = row ? aw1 : aw0;
#pragma unroll
      for (int px = 0; px < 2; px++) {
        const int k = pr * 2 + px;
        float yf = k == 0 ? byte_minus<0>(yw, kY) : k == 1 ? byte_minus<1>(yw, kY)
                 : k == 2 ? byte_minus<2>(yw, kY) : byte_minus<3>(yw, kY);
        float yt = __fmul_rn(__fmul_rn(yf, m.inv255), m.ys);            // BT709.h:494, 424
        float xr = add_sat(yt, cr_r);                                   // + Cb*0 is a no-op
        float xg = add_sat(__fadd_rn(yt, cb_g), cr_g);                  // (Y + Cb) + Cr order
        float xb = add_sat(yt, cb_b);
        uint32_t sr = lut_sum_smem<CL2, RL2>(xr, clamp_lo, sbase);
        uint32_t sg = lut_sum_smem<CL2, RL2>(xg, clamp_lo, sbase);
        uint32_t sb = lut_sum_smem<CL2, RL2>(xb, clamp_lo, sbase);
        uint32_t bg = __byte_perm(sb, sg, 0x0073);  // byte0 = B code, byte1 = G code
        uint32_t w;
        if (ALPHA) {
          // AAPLShaders.metal:249-271 via the CPU twin BT709.h:466-513 with neutral chroma
          float af = k == 0 ? byte_minus<0>(aw, kY) : k == 1 ? byte_minus<1>(aw, kY)
                   : k == 2 ? byte_minus<2>(aw, kY) : byte_minus<3>(aw, kY);
          float an = add_sat(__fmul_rn(__fmul_rn(af, m.inv255), m.ys), 0.0f);
          uint32_t ac = (uint32_t) round_pos(__fmul_rn(an, 255.0f));
          uint32_t ra = __byte_perm(sr, ac, 0x0043);  // byte0 = R code, byte1 = alpha
          w = __byte_perm(bg, ra, 0x5410);
        } else {
          w = __byte_perm(bg, sr, 0x7710) | 0xFF000000u;  // A = 255 (AAPLShaders.metal:243)
        }
        if (row) r1[k] = w; else r0[k] = w;
      }
    }
  }
  o0 = make_uint4(r0[0], r0[1], r0[2], r0[3]);
  o1 = make_uint4(r1[0], r1[1], r1[2], r1[3]);
}

template <int CL2, int RL2, bool ALPHA, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) k_decode_fast(const DecodeArgs p) {
  extern __shared__ __align__(16) uint32_t s_lut[];
  constexpr int N = 1 << CL2, R = 1 << RL2;
  for (int i = threadIdx.x; i < (N + 1) * R; i += THREADS) s_lut[i] = __ldg(&p.lut[i >> RL2]);
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const uint32_t sbase = (uint32_t) __cvta_generic_to_shared(s_lut) + ((lane & (R - 1)) << 2) -
                         (0x4B000000u << (2 + RL2));
  constexpr int WPC = THREADS / 32;
  const unsigned nwarps = gridDim.x * WPC;
  const unsigned total = (unsigned) p.nframes * p.rows2 * p.segs;
  unsigned unit = blockIdx.x * WPC + (threadIdx.x >> 5);
  if (unit >= total) return;
  // unit -> (frame, row pair, segment), then stepped incrementally (no division in the loop)
  unsigned seg = unit % p.segs, rpg = unit / p.segs;
  unsigned rp = rpg % p.rows2, f = rpg / p.rows2;
  const unsigned d_seg = nwarps % p.segs, d_rpg = nwarps / p.segs;
  const unsigned d_rp = d_rpg % p.rows2, d_f = d_rpg / p.rows2;
  const MatrixConsts m = p.m;
  const float clamp_lo = p.clamp_lo;

  for (; unit < total; unit += nwarps) {
    const int col0 = seg * 256 + lane * 4;
    const int row0 = rp * 2;
    const bool has_row1 = row0 + 1 < p.h;
    const uint8_t *yp = p.y + f * p.sy + (long long) row0 * p.py + col0;
    const uint8_t *cp = p.c + f * p.sc + (long long) rp * p.pc + col0;
    uint8_t *op = p.out + f * p.so + (long long) row0 * p.po + (long long) col0 * 4;
    const uint8_t *ap = ALPHA ? p.a + f * p.sa + (long long) row0 * p.pa + col0 : nullptr;
    const bool v0 = col0 < p.w, v1 = col0 + 128 < p.w;

    uint32_t y00 = 0, y01 = 0, y10 = 0, y11 = 0, c0 = 0, c1 = 0, a00 = 0, a01 = 0, a10 = 0, a11 = 0;
    if (v0) {
      y00 = ld_stream_u32(yp);
      c0 = ld_stream_u32(cp);
      if (has_row1) y10 = ld_stream_u32(yp + p.py);
      if (ALPHA) {
        a00 = ld_stream_u32(ap);
        if (has_row1) a10 = ld_stream_u32(ap + p.pa);
      }
    }
    if (v1) {
      y01 = ld_stream_u32(yp + 128);
      c1 = ld_stream_u32(cp + 128);
      if (has_row1) y11 = ld_stream_u32(yp + p.py + 128);
      if (ALPHA) {
        a01 = ld_stream_u32(ap + 128);
        if (has_row1) a11 = ld_stream_u32(ap + p.pa + 128);
      }
    }
    if (v0) {
      uint4 o0, o1;
      decode_quad<CL2, RL2, ALPHA>(y00, y10, c0, a00, a10, m, clamp_lo, sbase, o0, o1);
      st_stream_u128(op, o0);
      if (has_row1) st_stream_u128(op + p.po, o1);
    }
    if (v1) {
      uint4 o0, o1;
      decode_quad<CL2, RL2, ALPHA>(y01, y11, c1, a01, a11, m, clamp_lo, sbase, o0, o1);
      st_stream_u128(op + 512, o0);
      if (has_row1) st_stream_u128(op + p.po + 512, o1);
    }

    seg += d_seg;
    rp += d_rp;
    f += d_f;
    if (seg >= (unsigned) p.segs) { seg -= p.segs; rp++; }
    if (rp >= (unsigned) p.rows2) { rp -= p.rows2; f++; }
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

constexpr int kDecodeThreads = 1024;

template <int CL2, int RL2, bool ALPHA>
static int launch_decode_fast(const DecodeArgs &a, int sm_count, cudaStream_t s) {
  constexpr size_t smem = (size_t) ((1 << CL2) + 1) * (1 << RL2) * 4;
  auto k = k_decode_fast<CL2, RL2, ALPHA, kDecodeThreads>;
  const unsigned total = (unsigned) a.nframes * a.rows2 * a.segs;
  const unsigned wpc = kDecodeThreads / 32;
  unsigned grid = (total + wpc - 1) / wpc;
  if (grid > (unsigned) sm_count) grid = sm_count;
  k<<<grid, kDecodeThreads, smem, s>>>(a);
  return (int) cudaGetLastError();
}

int launch_decode(const DecodeArgs &a0, int sm_count, void *stream) {
  cudaStream_t s = (cudaStream_t) stream;
  DecodeArgs a = a0;
  a.rows2 = (a.h + 1) / 2;
  a.segs = (a.w + 255) / 256;
  const bool alpha = a.a != nullptr;
  const unsigned long long units = (unsigned long long) a.nframes * a.rows2 * a.segs;
  bool fast = (a.w % 4 == 0) && aligned(a.y, 4) && aligned(a.py, 4) && aligned(a.sy, 4) && aligned(a.c, 4) &&
              aligned(a.pc, 4) && aligned(a.sc, 4) && aligned(a.out, 16) && aligned(a.po, 16) &&
              aligned(a.so, 16) && units < 0x7FFFFFFFull && (a.cells_log2 == 10 || a.cells_log2 == 12);
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

int launch_rescale(const RescaleArgs &a, int sm_count, void *stream) {
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

int launch_encode(const EncodeArgs &a, int sm_count, void *stream) {
  cudaStream_t s = (cudaStream_t) stream;
  const size_t smem = (512 + (size_t) (1 << a.out_cells_log2) + 1) * 4;
  const bool vec = (a.w % 4 == 0) && aligned(a.in, 16) && aligned(a.pi, 16) && aligned(a.si, 16) && aligned(a.y, 4) &&
                   aligned(a.py, 4) && aligned(a.sy, 4) && aligned(a.c, 4) && aligned(a.pc, 4) && aligned(a.sc, 4);
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
  return (int) cudaFuncSetAttribute(k_decode_fast<CL2, RL2, ALPHA, kDecodeThreads>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
}

// per device (called from bt709_ctx_create with the device current): opt in to > 48 KB of
// dynamic shared memory for the replicated-LUT decode kernels
int kernels_init() {
  int e;
  if ((e = optin_decode_smem<10, 5, false>())) return e;
  if ((e = optin_decode_smem<10, 5, true>())) return e;
  if ((e = optin_decode_smem<12, 3, false>())) return e;
  if ((e = optin_decode_smem<12, 3, true>())) return e;
  return 0;
}

}  // namespace bt709
