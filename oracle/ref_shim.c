/*
 * oracle/_ref shim — TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Compiles the reference's own CPU math *unmodified, where it lies*:
 *   /root/reference/Renderer/BT709.h  and  /root/reference/Renderer/sRGB.h
 * are pulled in through -I (see oracle/Makefile); no reference source is copied
 * into this repository.  The only accommodations are the Objective-C BOOL
 * typedef the header expects (BT709.h:1051,1180) and silencing printf, because
 * BT709_average_pixel_values has `const int debug = 1` (BT709.h:1373) and would
 * print 7 lines per 2x2 block.  Arithmetic is untouched.
 *
 * The exported symbols (prefix ref_) take plain ints so that tests can compare
 * them 1:1 against the restatement in oracle_bt709.c.
 */
#include <stdio.h>
#include <stdint.h>
#include <math.h>
#include <assert.h>
#include <string.h>
#include <stdlib.h>

typedef signed char BOOL;
#ifndef TRUE
#define TRUE 1
#endif
#ifndef FALSE
#define FALSE 0
#endif

#define printf(...) ((void)0)
#include "BT709.h"
#undef printf

/* gamma numbering shared with include/bt709_b200.h */
enum { G_APPLE = 0, G_SRGB = 1, G_LINEAR = 2, G_BT709 = 3 };

static BT709Gamma to_ref_gamma(int g) {
  switch (g) {
    case G_APPLE:  return BT709GammaApple;
    case G_SRGB:   return BT709GammaSrgb;
    default:       return BT709GammaLinear;
  }
}

/* (Y,Cb,Cr) -> sRGB (R,G,B).  One function per decode mode of the reference. */
void ref_decode_pixel(int gamma, int Y, int Cb, int Cr, int *R, int *G, int *B) {
  if (gamma == G_BT709) {
    BT709_to_sRGB_convertYCbCrToRGB(Y, Cb, Cr, R, G, B, 1);
  } else if (gamma == G_APPLE) {
    Apple196_to_sRGB_convertYCbCrToRGB(Y, Cb, Cr, R, G, B, 1);
  } else if (gamma == G_SRGB) {
    sRGB_to_sRGB_convertYCbCrToRGB(Y, Cb, Cr, R, G, B, 1);
  } else {
    /* Metal "Linear" mode: matrix only, then the sRGB texture store encodes
       (AAPLShaders.metal:387-407 + MetalRenderContext.m:88-92). */
    float r, g, b;
    BT709_convertYCbCrToLinearRGB(Y, Cb, Cr, &r, &g, &b, 0);
    r = sRGB_linearNormToNonLinear(r);
    g = sRGB_linearNormToNonLinear(g);
    b = sRGB_linearNormToNonLinear(b);
    *R = (int) round(r * 255.0f);
    *G = (int) round(g * 255.0f);
    *B = (int) round(b * 255.0f);
  }
}

/* sRGB (R,G,B) -> (Y,Cb,Cr), per-pixel 4:4:4 encoders. */
void ref_encode444_pixel(int gamma, int R, int G, int B, int *Y, int *Cb, int *Cr) {
  if (gamma == G_BT709) {
    BT709_from_sRGB_convertRGBToYCbCr(R, G, B, Y, Cb, Cr, 1);
  } else if (gamma == G_APPLE) {
    Apple196_from_sRGB_convertRGBToYCbCr(R, G, B, Y, Cb, Cr);
  } else if (gamma == G_SRGB) {
    sRGB_from_sRGB_convertRGBToYCbCr(R, G, B, Y, Cb, Cr);
  } else {
    BT709_convertRGBToYCbCr(R, G, B, Y, Cb, Cr, 0);
  }
}

/* "previous API" used by CoreImageMetalFilterTests.m:94-360 */
void ref_linear_rgb_to_ycbcr(int R, int G, int B, int applyGamma, int *Y, int *Cb, int *Cr) {
  BT709_convertRGBToYCbCr(R, G, B, Y, Cb, Cr, applyGamma);
}
void ref_ycbcr_to_linear_rgb(int Y, int Cb, int Cr, int applyGamma, int *R, int *G, int *B) {
  BT709_convertYCbCrToRGB(Y, Cb, Cr, R, G, B, applyGamma);
}

/* 2x2 block, 12 ints R1 G1 B1 ... R4 G4 B4 -> 4 Y + Cb + Cr */
void ref_average_block(const int *rgb, int gin, int gout, int *y4, int *cb, int *cr) {
  BT709_average_pixel_values(rgb[0], rgb[1], rgb[2], rgb[3], rgb[4], rgb[5],
                             rgb[6], rgb[7], rgb[8], rgb[9], rgb[10], rgb[11],
                             &y4[0], &y4[1], &y4[2], &y4[3], cb, cr,
                             to_ref_gamma(gin), to_ref_gamma(gout));
}

void ref_srgb_average_cbcr(int *rgb12) {
  sRGB_from_sRGB_average_cbcr(&rgb12[0], &rgb12[1], &rgb12[2], &rgb12[3], &rgb12[4], &rgb12[5],
                              &rgb12[6], &rgb12[7], &rgb12[8], &rgb12[9], &rgb12[10], &rgb12[11]);
}

/* alpha near-twin: Y plane of the alpha stream through the matrix with neutral chroma */
int ref_alpha_pixel(int A) {
  float r, g, b;
  BT709_convertYCbCrToNonLinearRGB(A, 128, 128, &r, &g, &b);
  return (int) round(r * 255.0f);
}

float ref_srgb_to_linear(float v) { return sRGB_nonLinearNormToLinear(v); }
float ref_linear_to_srgb(float v) { return sRGB_linearNormToNonLinear(v); }

/* frame loops (plain planes; nearest chroma as AAPLShaders.metal:350) */
void ref_decode_nv12(int gamma, const uint8_t *y, long py, const uint8_t *c, long pc,
                     uint8_t *bgra, long po, int w, int h) {
  for (int row = 0; row < h; row++) {
    for (int col = 0; col < w; col++) {
      int Y = y[row * py + col];
      int Cb = c[(row / 2) * pc + (col / 2) * 2];
      int Cr = c[(row / 2) * pc + (col / 2) * 2 + 1];
      int R, G, B;
      ref_decode_pixel(gamma, Y, Cb, Cr, &R, &G, &B);
      uint8_t *o = bgra + row * po + col * 4;
      o[0] = (uint8_t) B; o[1] = (uint8_t) G; o[2] = (uint8_t) R; o[3] = 255;
    }
  }
}

/* the 4:2:0 encoder loop of CVPixelBufferUtils.h:284-363 over plain planes */
void ref_encode_bgra_nv12(int gin, int gout, const uint8_t *bgra, long pi,
                          uint8_t *y, long py, uint8_t *c, long pc, int w, int h) {
  for (int row = 0; row < h; row += 2) {
    for (int col = 0; col < w; col += 2) {
      int rgb[12];
      for (int k = 0; k < 4; k++) {
        const uint8_t *p = bgra + (row + (k >> 1)) * pi + (col + (k & 1)) * 4;
        rgb[k * 3 + 0] = p[2]; rgb[k * 3 + 1] = p[1]; rgb[k * 3 + 2] = p[0];
      }
      int y4[4], cb, cr;
      ref_average_block(rgb, gin, gout, y4, &cb, &cr);
      y[row * py + col] = (uint8_t) y4[0];
      y[row * py + col + 1] = (uint8_t) y4[1];
      y[(row + 1) * py + col] = (uint8_t) y4[2];
      y[(row + 1) * py + col + 1] = (uint8_t) y4[3];
      c[(row / 2) * pc + col] = (uint8_t) cb;
      c[(row / 2) * pc + col + 1] = (uint8_t) cr;
    }
  }
}

/* exhaustive encode->decode histogram (CoreImageMetalFilterTests.m:420-813);
   buckets[0]=exact, [1..9]=max-channel |d| <= k, [10]=more */
void ref_roundtrip_histogram(int gamma, long long *buckets) {
  for (int i = 0; i < 11; i++) buckets[i] = 0;
  for (int R = 0; R < 256; R++)
    for (int G = 0; G < 256; G++)
      for (int B = 0; B < 256; B++) {
        int Y, Cb, Cr, r, g, b;
        ref_encode444_pixel(gamma, R, G, B, &Y, &Cb, &Cr);
        ref_decode_pixel(gamma, Y, Cb, Cr, &r, &g, &b);
        int d = abs(r - R);
        if (abs(g - G) > d) d = abs(g - G);
        if (abs(b - B) > d) d = abs(b - B);
        buckets[d > 10 ? 10 : d]++;
      }
}
