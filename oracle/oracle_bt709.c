/*
 * oracle_bt709.c — plain-C restatement of the reference's CPU math for the
 * BT.709 4:2:0 <-> sRGB BGRA8 pixel path.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle_bt709.h).  Nothing under
 * metalbt709decoder_b200/ may depend on this file.
 *
 * Arithmetic contract, taken from the reference (paths relative to
 * /root/reference/Renderer):
 *   - every intermediate is stored as float32; pow() and round() are the C
 *     double versions, so transfer curves are evaluated in double and rounded
 *     to float once per assignment (sRGB.h:53,70; BT709.h:77,99,133,147);
 *   - round() is half-away-from-zero (BT709.h:245-247, 711-713);
 *   - no FMA contraction (the reference is built for plain x86-64/arm64 without
 *     -ffast-math); this file is compiled with -ffp-contract=off.
 *
 * Pinned against the reference's own golden vectors (tests/golden/ JSON files) and,
 * where /root/reference is present, bit-for-bit against oracle/_ref (the
 * unmodified headers compiled where they lie): tests/test_oracle_*.py.
 *
 * Pass-2 rescale and the alpha plane have no CPU twin in the reference:
 * PARITY UNPINNED for those two (restated from the Metal shader).
 */
#include "oracle_bt709.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------- constants */
/* BT709.h:40-45 */
static const float kKr = 0.2126f;
static const float kKg = 0.7152f;
static const float kKb = 0.0722f;
static const float kCrSpan = 1.5748f; /* 2(1-Kr) */
static const float kCbSpan = 1.8556f; /* 2(1-Kb) */

typedef struct {
  float r, g, b;
} rgbf;

/* sRGB.h:18-27 */
static float clamp01(float v) { return v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v); }

/* sRGB.h:32-36 */
static float unorm8(int v) { return v * (1.0f / 255.0f); }

/* BT709.h:711-713 and friends: float product, double round, truncate to int */
static int quant255(float v) { return (int) round(v * 255.0f); }

/* ------------------------------------------------------ transfer functions */
/* sRGB.h:43-57 */
static float srgb_eotf(float v) {
  if (v <= 0.04045f) return v * (1.0f / 12.92f);
  float s = (v + 0.055f) * (1.0f / (1.0f + 0.055f));
  return (float) pow(s, 2.4f);
}

/* sRGB.h:62-74.  The whole right-hand side is a double expression. */
static float srgb_oetf(float v) {
  if (v <= 0.0031308f) return v * 12.92f;
  const float a = 0.055f;
  const float g = 1.0f / 2.4f;
  return (float) ((1.0f + a) * pow(v, g) - a);
}

/* BT709.h:68-81 */
static float bt709_eotf(float v) {
  if (v < 0.081f) return v * (1.0f / 4.5f);
  const float a = 0.099f;
  const float g = 1.0f / 0.45f;
  float s = (v + a) * (1.0f / (1.0f + a));
  return (float) pow(s, g);
}

/* BT709.h:90-103 */
static float bt709_oetf(float v) {
  if (v < 0.018f) return v * 4.5f;
  const float a = 0.099f;
  const float g = 0.45f;
  return (float) ((1.0f + a) * pow(v, g) - a);
}

#define APPLE_GAMMA 1.960938f /* BT709.h:123 */

/* BT709.h:125-137 */
static float apple_eotf(float v) {
  if (v < 0.05583828f) return v * (1.0f / 16.0f);
  return (float) pow(v, APPLE_GAMMA);
}

/* BT709.h:139-151 */
static float apple_oetf(float v) {
  if (v < 0.00349f) return v * 16.0f;
  const float g = 1.0f / APPLE_GAMMA;
  return (float) pow(v, g);
}

float oracle_srgb_to_linear(float v) { return srgb_eotf(v); }
float oracle_linear_to_srgb(float v) { return srgb_oetf(v); }

/* ------------------------------------------------------------------ matrix */
/* BT709.h:348-460.  video_range != 0 rescales 219/224 code ranges to 255. */
static rgbf ycbcr_to_rgb_norm(float yn, float cbn, float crn, int video_range) {
  const float ys = 255.0f / (video_range ? (float) (235 - 16) : 255.0f);
  const float cs = 255.0f / (video_range ? (float) (240 - 16) : 255.0f);
  const float kb_kg = kKb / kKg; /* BT709.h:53 */
  const float kr_kg = kKr / kKg; /* BT709.h:52 */
  const float m_r_cr = cs * kCrSpan;
  const float m_g_cb = -1.0f * cs * kCbSpan * kb_kg;
  const float m_g_cr = -1.0f * cs * kCrSpan * kr_kg;
  const float m_b_cb = cs * kCbSpan;
  rgbf o;
  o.r = (yn * ys) + (cbn * 0.000f) + (crn * m_r_cr);
  o.g = (yn * ys) + (cbn * m_g_cb) + (crn * m_g_cr);
  o.b = (yn * ys) + (cbn * m_b_cb) + (crn * 0.000f);
  o.r = clamp01(o.r);
  o.g = clamp01(o.g);
  o.b = clamp01(o.b);
  return o;
}

/* BT709.h:466-513: integer offsets first, then one float multiply each */
static rgbf ycbcr8_to_rgb_nonlinear(int Y, int Cb, int Cr) {
  float yn = (Y - 16) * (1.0f / 255.0f);
  float cbn = (Cb - 128) * (1.0f / 255.0f);
  float crn = (Cr - 128) * (1.0f / 255.0f);
  return ycbcr_to_rgb_norm(yn, cbn, crn, 1);
}

/* BT709.h:199-269 */
static void rgb_nonlinear_to_ycbcr8(float r, float g, float b, int *Y, int *Cb, int *Cr) {
  float ey = (kKr * r) + (kKg * g) + (kKb * b);
  float eb = (b - ey) / kCbSpan;
  float er = (r - ey) / kCrSpan;
  float ay = (ey * (235 - 16)) + 16;
  float ab = (eb * (240 - 16)) + 128;
  float ar = (er * (240 - 16)) + 128;
  *Y = (int) round(ay);
  *Cb = (int) round(ab);
  *Cr = (int) round(ar);
}

/* ------------------------------------------------------------------ decode */
/* BT.709 spec  BT709.h:668-738;  Apple 1.961  BT709.h:821-908;
 * sRGB-tagged  BT709.h:948-1003; Linear = BT709.h:517-556 (gamma off) followed
 * by the sRGB store conversion of the render target (sRGB.h:62-74). */
void oracle_decode_pixel(int gamma, int Y, int Cb, int Cr, int *R, int *G, int *B) {
  rgbf p = ycbcr8_to_rgb_nonlinear(Y, Cb, Cr);
  switch (gamma) {
    case ORACLE_GAMMA_BT709:
      p.r = srgb_oetf(bt709_eotf(p.r));
      p.g = srgb_oetf(bt709_eotf(p.g));
      p.b = srgb_oetf(bt709_eotf(p.b));
      break;
    case ORACLE_GAMMA_APPLE:
      p.r = srgb_oetf(apple_eotf(p.r));
      p.g = srgb_oetf(apple_eotf(p.g));
      p.b = srgb_oetf(apple_eotf(p.b));
      break;
    case ORACLE_GAMMA_LINEAR:
      p.r = srgb_oetf(p.r);
      p.g = srgb_oetf(p.g);
      p.b = srgb_oetf(p.b);
      break;
    default: /* ORACLE_GAMMA_SRGB: values are already sRGB-encoded */
      break;
  }
  *R = quant255(p.r);
  *G = quant255(p.g);
  *B = quant255(p.b);
}

/* AAPLShaders.metal:249-271; CPU near-twin BT709.h:466-513 with neutral chroma.
 * Alpha is stored linearly (not sRGB encoded) in the render target. */
int oracle_alpha_pixel(int A) {
  rgbf p = ycbcr8_to_rgb_nonlinear(A, 128, 128);
  return quant255(p.r);
}

/* BT709.h:558-605 */
void oracle_ycbcr_to_linear_rgb(int Y, int Cb, int Cr, int applyGamma, int *R, int *G, int *B) {
  rgbf p = ycbcr8_to_rgb_nonlinear(Y, Cb, Cr);
  if (applyGamma) {
    p.r = bt709_eotf(p.r);
    p.g = bt709_eotf(p.g);
    p.b = bt709_eotf(p.b);
  }
  *R = quant255(p.r);
  *G = quant255(p.g);
  *B = quant255(p.b);
}

/* ------------------------------------------------------------ encode 4:4:4 */
/* BT.709: BT709.h:610-664 (+276-322); Apple: :743-817; sRGB: :914-944;
 * Linear: the "previous API" BT709.h:326-341 with the gamma map off. */
void oracle_encode444_pixel(int gamma, int R, int G, int B, int *Y, int *Cb, int *Cr) {
  float r = unorm8(R), g = unorm8(G), b = unorm8(B);
  switch (gamma) {
    case ORACLE_GAMMA_BT709:
      r = bt709_oetf(srgb_eotf(r));
      g = bt709_oetf(srgb_eotf(g));
      b = bt709_oetf(srgb_eotf(b));
      break;
    case ORACLE_GAMMA_APPLE:
      r = apple_oetf(srgb_eotf(r));
      g = apple_oetf(srgb_eotf(g));
      b = apple_oetf(srgb_eotf(b));
      break;
    default: /* SRGB and LINEAR: bytes go through the matrix unchanged */
      break;
  }
  rgb_nonlinear_to_ycbcr8(r, g, b, Y, Cb, Cr);
}

/* BT709.h:326-341 */
void oracle_linear_rgb_to_ycbcr(int R, int G, int B, int applyGamma, int *Y, int *Cb, int *Cr) {
  float r = unorm8(R), g = unorm8(G), b = unorm8(B);
  if (applyGamma) {
    r = bt709_oetf(r);
    g = bt709_oetf(g);
    b = bt709_oetf(b);
  }
  rgb_nonlinear_to_ycbcr8(r, g, b, Y, Cb, Cr);
}

/* ------------------------------------------------------------ encode 4:2:0 */
/* BT709.h:1098-1148 (only Srgb / Apple / Linear are defined there) */
static float to_linear(int v, int gamma) {
  float n = unorm8(v);
  if (gamma == ORACLE_GAMMA_SRGB) return srgb_eotf(n);
  if (gamma == ORACLE_GAMMA_APPLE) return apple_eotf(n);
  return n;
}

/* BT709.h:1150-1167: linear float -> output-gamma *integer* */
static int from_linear(float v, int gamma) {
  float nl;
  if (gamma == ORACLE_GAMMA_SRGB)
    nl = srgb_oetf(v);
  else if (gamma == ORACLE_GAMMA_APPLE)
    nl = apple_oetf(v);
  else
    nl = v;
  return quant255(nl);
}

/* BT709.h:1172-1190 */
static float mean4(float a, float b, float c, float d) {
  float s = (a + b + c + d);
  return s / 4.0f;
}

/* BT709.h:1349-1509.  Chroma: mean of the four pixels in linear light,
 * requantised to 8-bit in the output gamma, then the matrix.  Luma: each pixel
 * requantised to 8-bit in the output gamma, then the matrix. */
void oracle_average_block(const int *rgb12, int gin, int gout, int *y4, int *cb, int *cr) {
  float lin[12];
  for (int i = 0; i < 12; i++) lin[i] = to_linear(rgb12[i], gin);

  float mr = mean4(lin[0], lin[3], lin[6], lin[9]);
  float mg = mean4(lin[1], lin[4], lin[7], lin[10]);
  float mb = mean4(lin[2], lin[5], lin[8], lin[11]);

  int qr = from_linear(mr, gout);
  int qg = from_linear(mg, gout);
  int qb = from_linear(mb, gout);

  int ymean;
  rgb_nonlinear_to_ycbcr8(unorm8(qr), unorm8(qg), unorm8(qb), &ymean, cb, cr);

  for (int k = 0; k < 4; k++) {
    int pr = from_linear(lin[k * 3 + 0], gout);
    int pg = from_linear(lin[k * 3 + 1], gout);
    int pb = from_linear(lin[k * 3 + 2], gout);
    int pcb, pcr;
    rgb_nonlinear_to_ycbcr8(unorm8(pr), unorm8(pg), unorm8(pb), &y4[k], &pcb, &pcr);
  }
}

/* BT709.h:1041-1093 + :1200-1318 (pinned by CoreImageMetalFilterTests.m:1421-1592) */
void oracle_srgb_average_cbcr(int *rgb12) {
  float yl[4], cbl[4], crl[4];
  for (int k = 0; k < 4; k++) {
    float r = srgb_eotf(unorm8(rgb12[k * 3 + 0]));
    float g = srgb_eotf(unorm8(rgb12[k * 3 + 1]));
    float b = srgb_eotf(unorm8(rgb12[k * 3 + 2]));
    yl[k] = (kKr * r) + (kKg * g) + (kKb * b);
    cbl[k] = (b - yl[k]) / kCbSpan;
    crl[k] = (r - yl[k]) / kCrSpan;
  }
  float cbm = mean4(cbl[0], cbl[1], cbl[2], cbl[3]);
  float crm = mean4(crl[0], crl[1], crl[2], crl[3]);
  for (int k = 0; k < 4; k++) {
    rgbf p = ycbcr_to_rgb_norm(yl[k], cbm, crm, 0);
    rgb12[k * 3 + 0] = from_linear(p.r, ORACLE_GAMMA_SRGB);
    rgb12[k * 3 + 1] = from_linear(p.g, ORACLE_GAMMA_SRGB);
    rgb12[k * 3 + 2] = from_linear(p.b, ORACLE_GAMMA_SRGB);
  }
}

/* --------------------------------------------------------- row-band threads */
typedef void (*band_fn)(void *ctx, int row0, int row1);
typedef struct {
  band_fn fn;
  void *ctx;
  int row0, row1;
} band_job;

static void *band_entry(void *p) {
  band_job *j = (band_job *) p;
  j->fn(j->ctx, j->row0, j->row1);
  return NULL;
}

/* rows are split in units of `gran` rows (2 for 4:2:0 work) */
static void run_bands(band_fn fn, void *ctx, int rows, int gran, int nthreads) {
  int units = (rows + gran - 1) / gran;
  if (nthreads > units) nthreads = units;
  if (nthreads <= 1) {
    fn(ctx, 0, rows);
    return;
  }
  pthread_t *th = (pthread_t *) malloc(sizeof(pthread_t) * nthreads);
  band_job *jobs = (band_job *) malloc(sizeof(band_job) * nthreads);
  for (int t = 0; t < nthreads; t++) {
    int u0 = (int) ((long long) units * t / nthreads);
    int u1 = (int) ((long long) units * (t + 1) / nthreads);
    jobs[t].fn = fn;
    jobs[t].ctx = ctx;
    jobs[t].row0 = u0 * gran;
    jobs[t].row1 = u1 * gran > rows ? rows : u1 * gran;
    pthread_create(&th[t], NULL, band_entry, &jobs[t]);
  }
  for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  free(th);
  free(jobs);
}

/* ------------------------------------------------------------ frame: decode */
typedef struct {
  int gamma, w, h;
  const uint8_t *y, *c, *a;
  long py, pc, pa, po;
  uint8_t *out;
} dec_ctx;

/* AAPLShaders.metal:336-357 (Y at gid, CbCr at gid/2 = nearest-neighbour 4:2:0
 * upsample); output bytes B,G,R,A (MetalRenderContext.m:188-198); alpha from the
 * second stream's Y plane, no premultiply (AAPLShaders.metal:411-438). */
static void dec_band(void *p, int row0, int row1) {
  dec_ctx *d = (dec_ctx *) p;
  for (int row = row0; row < row1; row++) {
    const uint8_t *yr = d->y + (long) row * d->py;
    const uint8_t *cr = d->c + (long) (row >> 1) * d->pc;
    const uint8_t *ar = d->a ? d->a + (long) row * d->pa : NULL;
    uint8_t *o = d->out + (long) row * d->po;
    for (int col = 0; col < d->w; col++) {
      int R, G, B;
      oracle_decode_pixel(d->gamma, yr[col], cr[(col >> 1) * 2], cr[(col >> 1) * 2 + 1], &R, &G, &B);
      o[col * 4 + 0] = (uint8_t) B;
      o[col * 4 + 1] = (uint8_t) G;
      o[col * 4 + 2] = (uint8_t) R;
      o[col * 4 + 3] = ar ? (uint8_t) oracle_alpha_pixel(ar[col]) : 255;
    }
  }
}

int oracle_decode_nv12(int gamma, const uint8_t *y, long pitch_y, const uint8_t *cbcr, long pitch_c,
                       const uint8_t *alpha_y, long pitch_a, uint8_t *bgra, long pitch_out,
                       int w, int h, int nthreads) {
  if (!y || !cbcr || !bgra || w <= 0 || h <= 0) return -1;
  if (gamma < 0 || gamma > 3) return -1;
  dec_ctx d = {gamma, w, h, y, cbcr, alpha_y, pitch_y, pitch_c, pitch_a, pitch_out, bgra};
  run_bands(dec_band, &d, h, 2, nthreads);
  return 0;
}

/* ------------------------------------------------------------ frame: encode */
typedef struct {
  int gin, gout, w, h;
  const uint8_t *in;
  long pi, py, pc;
  uint8_t *y, *c;
} enc_ctx;

/* CVPixelBufferUtils.h:284-363: 2x2 blocks; input word is B | G<<8 | R<<16 | A<<24
 * (A ignored); chroma stored as Cb | Cr<<8 (:357-361). */
static void enc_band(void *p, int row0, int row1) {
  enc_ctx *e = (enc_ctx *) p;
  for (int row = row0; row < row1; row += 2) {
    for (int col = 0; col < e->w; col += 2) {
      int rgb[12];
      for (int k = 0; k < 4; k++) {
        const uint8_t *px = e->in + (long) (row + (k >> 1)) * e->pi + (long) (col + (k & 1)) * 4;
        rgb[k * 3 + 0] = px[2];
        rgb[k * 3 + 1] = px[1];
        rgb[k * 3 + 2] = px[0];
      }
      int y4[4], cb, cr;
      oracle_average_block(rgb, e->gin, e->gout, y4, &cb, &cr);
      e->y[(long) row * e->py + col] = (uint8_t) y4[0];
      e->y[(long) row * e->py + col + 1] = (uint8_t) y4[1];
      e->y[(long) (row + 1) * e->py + col] = (uint8_t) y4[2];
      e->y[(long) (row + 1) * e->py + col + 1] = (uint8_t) y4[3];
      e->c[(long) (row >> 1) * e->pc + col] = (uint8_t) cb;
      e->c[(long) (row >> 1) * e->pc + col + 1] = (uint8_t) cr;
    }
  }
}

int oracle_encode_bgra_nv12(int gin, int gout, const uint8_t *bgra, long pitch_in, uint8_t *y,
                            long pitch_y, uint8_t *cbcr, long pitch_c, int w, int h, int nthreads) {
  if (!bgra || !y || !cbcr || w <= 0 || h <= 0) return -1;
  if ((w & 1) || (h & 1)) return -2; /* CVPixelBufferUtils.h:275-276 */
  if (gin < 0 || gin > 2 || gout < 0 || gout > 2) return -3; /* BT709.h:1135-1137,1162-1164 */
  enc_ctx e = {gin, gout, w, h, bgra, pitch_in, pitch_y, pitch_c, y, cbcr};
  run_bands(enc_band, &e, h, 2, nthreads);
  return 0;
}

/* ----------------------------------------------------------- frame: rescale */
/* PARITY UNPINNED: restated from samplingShader (AAPLShaders.metal:73-85),
 * identityVertexShader (:33-66) and -renderScaled: (MetalScaleRenderContext.m:55-105).
 *   - the source is an sRGB-format BGRA8 texture and the sampler is
 *     mag=min=linear with the default clamp-to-edge addressing, so each tap is
 *     sRGB->linear decoded (alpha is linear) before the bilinear blend;
 *   - the destination is an sRGB drawable (AAPLRenderer.m:130), so the blended
 *     linear value is sRGB encoded and rounded to unorm8 on store;
 *   - sample position = pixel centre: src = (dst + 0.5) * in/out - 0.5; the
 *     vertex shader's V flip and the top-left texture origin cancel.
 * float32 steps in exactly this order (the CUDA kernel mirrors them op by op):
 *   s  = (d + 0.5f) * scale - 0.5f        scale = (float)in / (float)out
 *   i0 = floorf(s), f = s - i0, taps clamp(i0), clamp(i0+1)
 *   top = t00 + fx*(t10 - t00); bot = t01 + fx*(t11 - t01); v = top + fy*(bot - top)
 */
typedef struct {
  const uint8_t *in;
  uint8_t *out;
  long pi, po;
  int iw, ih, ow, oh;
  float lin[256];
} rs_ctx;

static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

static void rs_band(void *p, int row0, int row1) {
  rs_ctx *r = (rs_ctx *) p;
  const float sx = (float) r->iw / (float) r->ow;
  const float sy = (float) r->ih / (float) r->oh;
  for (int oy = row0; oy < row1; oy++) {
    float fy_s = ((float) oy + 0.5f) * sy - 0.5f;
    float fy0 = floorf(fy_s);
    float fy = fy_s - fy0;
    int y0 = clampi((int) fy0, 0, r->ih - 1);
    int y1 = clampi((int) fy0 + 1, 0, r->ih - 1);
    const uint8_t *r0 = r->in + (long) y0 * r->pi;
    const uint8_t *r1 = r->in + (long) y1 * r->pi;
    uint8_t *o = r->out + (long) oy * r->po;
    for (int ox = 0; ox < r->ow; ox++) {
      float fx_s = ((float) ox + 0.5f) * sx - 0.5f;
      float fx0 = floorf(fx_s);
      float fx = fx_s - fx0;
      int x0 = clampi((int) fx0, 0, r->iw - 1);
      int x1 = clampi((int) fx0 + 1, 0, r->iw - 1);
      for (int ch = 0; ch < 4; ch++) {
        float t00, t10, t01, t11;
        if (ch < 3) {
          t00 = r->lin[r0[x0 * 4 + ch]];
          t10 = r->lin[r0[x1 * 4 + ch]];
          t01 = r->lin[r1[x0 * 4 + ch]];
          t11 = r->lin[r1[x1 * 4 + ch]];
        } else {
          t00 = unorm8(r0[x0 * 4 + ch]);
          t10 = unorm8(r0[x1 * 4 + ch]);
          t01 = unorm8(r1[x0 * 4 + ch]);
          t11 = unorm8(r1[x1 * 4 + ch]);
        }
        float top = t00 + fx * (t10 - t00);
        float bot = t01 + fx * (t11 - t01);
        float v = top + fy * (bot - top);
        v = clamp01(v);
        if (ch < 3) v = srgb_oetf(v);
        o[ox * 4 + ch] = (uint8_t) quant255(v);
      }
    }
  }
}

int oracle_rescale_bgra(const uint8_t *in, long pitch_in, int iw, int ih, uint8_t *out,
                        long pitch_out, int ow, int oh, int nthreads) {
  if (!in || !out || iw <= 0 || ih <= 0 || ow <= 0 || oh <= 0) return -1;
  rs_ctx r;
  r.in = in;
  r.out = out;
  r.pi = pitch_in;
  r.po = pitch_out;
  r.iw = iw;
  r.ih = ih;
  r.ow = ow;
  r.oh = oh;
  for (int i = 0; i < 256; i++) r.lin[i] = srgb_eotf(unorm8(i));
  run_bands(rs_band, &r, oh, 1, nthreads);
  return 0;
}

/* ---------------------------------------------------- packed 4:4:4 harness */
/* BGRAToBT709Converter.m:89-144: Apple curve only; out word Cr<<16 | Cb<<8 | Y */
int oracle_convert_packed444(const uint32_t *bgra, uint32_t *ycbcr, int w, int h) {
  if ((w & 1) || (h & 1)) return -2; /* BGRAToBT709Converter.m:41-46 */
  for (long i = 0; i < (long) w * h; i++) {
    uint32_t p = bgra[i];
    int Y, Cb, Cr;
    oracle_encode444_pixel(ORACLE_GAMMA_APPLE, (p >> 16) & 0xFF, (p >> 8) & 0xFF, p & 0xFF, &Y, &Cb, &Cr);
    ycbcr[i] = ((uint32_t) Cr << 16) | ((uint32_t) Cb << 8) | (uint32_t) Y;
  }
  return 0;
}

/* BGRAToBT709Converter.m:146-198: out word R<<16 | G<<8 | B with A = 0 (:191) */
int oracle_unconvert_packed444(const uint32_t *ycbcr, uint32_t *bgra, int w, int h) {
  if ((w & 1) || (h & 1)) return -2; /* BGRAToBT709Converter.m:69-74 */
  for (long i = 0; i < (long) w * h; i++) {
    uint32_t p = ycbcr[i];
    int R, G, B;
    oracle_decode_pixel(ORACLE_GAMMA_APPLE, p & 0xFF, (p >> 8) & 0xFF, (p >> 16) & 0xFF, &R, &G, &B);
    bgra[i] = ((uint32_t) R << 16) | ((uint32_t) G << 8) | (uint32_t) B;
  }
  return 0;
}

/* BGRAToBT709Converter.m:1042-1099: every row writes the chroma of its even
 * columns into row/2, so the odd row's value is the one that survives. */
int oracle_copy_packed444_to_nv12(const uint32_t *ycbcr, uint8_t *y, long pitch_y, uint8_t *cbcr,
                                  long pitch_c, int w, int h) {
  for (int row = 0; row < h; row++) {
    for (int col = 0; col < w; col++) {
      uint32_t p = ycbcr[(long) row * w + col];
      y[(long) row * pitch_y + col] = (uint8_t) (p & 0xFF);
      if ((col & 1) == 0) {
        cbcr[(long) (row >> 1) * pitch_c + col] = (uint8_t) ((p >> 8) & 0xFF);
        cbcr[(long) (row >> 1) * pitch_c + col + 1] = (uint8_t) ((p >> 16) & 0xFF);
      }
    }
  }
  return 0;
}

/* ---------------------------------------------------------------- histogram */
/* CoreImageMetalFilterTests.m:420-537 (BT.709), :557-674 (Apple), :696-813 (sRGB) */
typedef struct {
  int gamma;
  long long b[8][11];
} hist_ctx;

typedef struct {
  hist_ctx *h;
  int slot, r0, r1;
} hist_job;

static void *hist_entry(void *p) {
  hist_job *j = (hist_job *) p;
  long long *b = j->h->b[j->slot];
  for (int R = j->r0; R < j->r1; R++)
    for (int G = 0; G < 256; G++)
      for (int B = 0; B < 256; B++) {
        int Y, Cb, Cr, r, g, bb;
        oracle_encode444_pixel(j->h->gamma, R, G, B, &Y, &Cb, &Cr);
        oracle_decode_pixel(j->h->gamma, Y, Cb, Cr, &r, &g, &bb);
        int d = abs(r - R);
        if (abs(g - G) > d) d = abs(g - G);
        if (abs(bb - B) > d) d = abs(bb - B);
        b[d > 10 ? 10 : d]++;
      }
  return NULL;
}

void oracle_roundtrip_histogram(int gamma, long long *buckets11) {
  hist_ctx h;
  memset(&h, 0, sizeof(h));
  h.gamma = gamma;
  pthread_t th[8];
  hist_job jobs[8];
  for (int t = 0; t < 8; t++) {
    jobs[t].h = &h;
    jobs[t].slot = t;
    jobs[t].r0 = t * 32;
    jobs[t].r1 = (t + 1) * 32;
    pthread_create(&th[t], NULL, hist_entry, &jobs[t]);
  }
  for (int i = 0; i < 11; i++) buckets11[i] = 0;
  for (int t = 0; t < 8; t++) {
    pthread_join(th[t], NULL);
    for (int i = 0; i < 11; i++) buckets11[i] += h.b[t][i];
  }
}
