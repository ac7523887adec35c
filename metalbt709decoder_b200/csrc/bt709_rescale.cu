// bt709_rescale.cu — pass 2: bilinear resample of an sRGB BGRA8 image in linear light
// (samplingShader, AAPLShaders.metal:73-85 + MetalScaleRenderContext.m:55-105).
#include "bt709_device.cuh"

namespace bt709 {

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

int rescale_kernels_init() {
  return (int) cudaFuncSetAttribute(k_rescale_cols, cudaFuncAttributeMaxDynamicSharedMemorySize, (256 * 64 + 8193) * 4);
}

}  // namespace bt709
