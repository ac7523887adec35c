/*
 * oracle_bt709.h — CPU oracle for the BT.709 4:2:0 <-> sRGB BGRA8 pixel path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library, and
 * only as the checker.  The product (metalbt709decoder_b200/) never links,
 * imports or calls it.
 *
 * Parity status:
 *   decode (all four gammas), 4:4:4 encode, 4:2:0 encode   PINNED — checked against
 *       the reference's own golden vectors (tests/golden/) and bit-for-bit against
 *       the reference headers compiled where they lie (oracle/_ref).
 *   pass-2 bilinear rescale, alpha decode                   PARITY UNPINNED — the
 *       reference has no CPU twin and no test for them; restated from
 *       AAPLShaders.metal:73-85 / :249-271 and Metal's sampling rules.
 */
#ifndef ORACLE_BT709_H
#define ORACLE_BT709_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* same numbering as bt709_gamma in include/bt709_b200.h */
enum {
  ORACLE_GAMMA_APPLE = 0,  /* MetalBT709GammaApple  (MetalBT709Decoder.h:16) */
  ORACLE_GAMMA_SRGB = 1,   /* MetalBT709GammaSRGB   (MetalBT709Decoder.h:17) */
  ORACLE_GAMMA_LINEAR = 2, /* MetalBT709GammaLinear (MetalBT709Decoder.h:18) */
  ORACLE_GAMMA_BT709 = 3   /* CPU-only BT.709 spec curve (BT709.h:668-738)   */
};

/* ---- per-pixel --------------------------------------------------------- */
void oracle_decode_pixel(int gamma, int Y, int Cb, int Cr, int *R, int *G, int *B);
void oracle_encode444_pixel(int gamma, int R, int G, int B, int *Y, int *Cb, int *Cr);
void oracle_linear_rgb_to_ycbcr(int R, int G, int B, int applyGamma, int *Y, int *Cb, int *Cr);
void oracle_ycbcr_to_linear_rgb(int Y, int Cb, int Cr, int applyGamma, int *R, int *G, int *B);
void oracle_average_block(const int *rgb12, int gin, int gout, int *y4, int *cb, int *cr);
void oracle_srgb_average_cbcr(int *rgb12);
int oracle_alpha_pixel(int A);
float oracle_srgb_to_linear(float v);
float oracle_linear_to_srgb(float v);

/* ---- frames (host memory, byte pitches) -------------------------------- */
/* nthreads <= 1: scalar loop.  >1: static row-band split over pthreads. */
int oracle_decode_nv12(int gamma, const uint8_t *y, long pitch_y, const uint8_t *cbcr, long pitch_c,
                       const uint8_t *alpha_y, long pitch_a, uint8_t *bgra, long pitch_out,
                       int w, int h, int nthreads);
int oracle_encode_bgra_nv12(int gin, int gout, const uint8_t *bgra, long pitch_in, uint8_t *y,
                            long pitch_y, uint8_t *cbcr, long pitch_c, int w, int h, int nthreads);
int oracle_rescale_bgra(const uint8_t *in, long pitch_in, int iw, int ih, uint8_t *out,
                        long pitch_out, int ow, int oh, int nthreads);
int oracle_convert_packed444(const uint32_t *bgra, uint32_t *ycbcr, int w, int h);
int oracle_unconvert_packed444(const uint32_t *ycbcr, uint32_t *bgra, int w, int h);
int oracle_copy_packed444_to_nv12(const uint32_t *ycbcr, uint8_t *y, long pitch_y, uint8_t *cbcr,
                                  long pitch_c, int w, int h);

/* exhaustive 2^24 encode->decode histogram: buckets[0]=exact, [k]=max|d|==k, [10]=more */
void oracle_roundtrip_histogram(int gamma, long long *buckets11);

#ifdef __cplusplus
}
#endif
#endif
