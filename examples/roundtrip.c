/*
 * roundtrip.c — plain C99 against include/metal_bt709_decoder.h: what a call site of the reference's
 * Objective-C classes looks like after the switch.
 *
 *   sRGB BGRA image --createYCbCrFromBGRA--> NV12 (BT.709 video range, Apple or sRGB gamma)
 *                   --decodeBT709---------->  sRGB BGRA texture
 *
 * mirrors the encode -> decode round trips of the reference's tests (AppleEncodeDecodeBT709Tests.m,
 * MetalBT709DecoderTests.m:223-255) and prints the max-channel error histogram of the round trip.
 *
 *   usage: roundtrip [width height [srgb|apple]]       exit 0 when every |delta| <= 5
 *   build: make -C metalbt709decoder_b200/csrc examples
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "metal_bt709_decoder.h"

int main(int argc, char **argv) {
  int w = argc > 2 ? atoi(argv[1]) : 256, h = argc > 2 ? atoi(argv[2]) : 128;
  int as_srgb = argc > 3 && strcmp(argv[3], "srgb") == 0;
  if (w <= 0 || h <= 0 || (w & 1) || (h & 1)) {
    fprintf(stderr, "width and height must be even and positive\n");
    return 2;
  }

  MetalRenderContext *mrc = MetalRenderContext_new(0);
  if (!mrc || !MetalRenderContext_setupMetal(mrc)) {
    fprintf(stderr, "no CUDA device: this library has no CPU fallback\n");
    return 3;
  }

  /* a deterministic test card: every 8-bit gray plus colour ramps */
  uint32_t *src = (uint32_t *) malloc((size_t) w * h * 4);
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      /* 2x2 blocks of one colour so that the 4:2:0 subsampling itself loses nothing */
      uint32_t bx = (uint32_t) x / 2, by = (uint32_t) y / 2;
      uint32_t r = (bx * 7 + by * 3) & 255, g = (bx * 2) & 255, b = (by * 5 + bx) & 255;
      if (by & 1) r = g = b = bx & 255;
      src[(size_t) y * w + x] = 0xFF000000u | (r << 16) | (g << 8) | b;
    }

  /* +createYCbCrFromCGImage:isLinear:asSRGBGamma: */
  MBDPixelBuffer *pb = BGRAToBT709Converter_createYCbCrFromBGRA(mrc, src, w, h, /*isLinear*/ 0, as_srgb);
  if (!pb) {
    fprintf(stderr, "encode failed\n");
    return 4;
  }

  MetalBT709Decoder *dec = MetalBT709Decoder_new();
  MetalBT709Decoder_setMetalRenderContext(dec, mrc);
  MetalBT709Decoder_setGamma(dec, as_srgb ? MetalBT709GammaSRGB : MetalBT709GammaApple);
  if (!MetalBT709Decoder_setupMetal(dec)) return 5;
  MBDTexture *tex = MetalRenderContext_makeBGRATexture(mrc, w, h, NULL, /*isSRGB*/ 1);
  if (!MetalBT709Decoder_decodeBT709(dec, pb, NULL, tex, NULL, w, h, /*waitUntilCompleted*/ 1)) {
    fprintf(stderr, "decodeBT709 returned FALSE\n");
    return 6;
  }

  long hist[256] = {0};
  int worst = 0;
  for (int y = 0; y < h; y++) {
    const uint32_t *row = (const uint32_t *) (tex->pixels + (size_t) y * tex->bytesPerRow);
    for (int x = 0; x < w; x++) {
      uint32_t a = src[(size_t) y * w + x], b = row[x];
      int d = 0;
      for (int sh = 0; sh < 24; sh += 8) {
        int e = abs((int) ((a >> sh) & 255) - (int) ((b >> sh) & 255));
        if (e > d) d = e;
      }
      hist[d]++;
      if (d > worst) worst = d;
    }
  }
  printf("%dx%d %s round trip, max-channel |delta| histogram:", w, h, as_srgb ? "sRGB" : "Apple-1.961");
  for (int d = 0; d <= worst; d++) printf(" %d:%ld", d, hist[d]);
  printf("\nkernels launched: %llu\n", MetalRenderContext_launchCount(mrc));

  MBDTextureRelease(tex);
  MBDPixelBufferRelease(pb);
  MetalBT709Decoder_release(dec);
  MetalRenderContext_release(mrc);
  free(src);
  return worst <= 5 ? 0 : 1;
}
