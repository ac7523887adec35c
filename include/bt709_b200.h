/*
 * bt709_b200.h — C ABI of the B200-native BT.709 4:2:0 <-> sRGB BGRA8 pixel path.
 *
 * Drop-in boundary for the per-pixel hot path of mdejong/MetalBT709Decoder.
 * Every entry point takes plain pitched DEVICE pointers (what CVPixelBuffer
 * planes / MTLTextures were in the reference) and a cudaStream_t passed as
 * void* (what id<MTLCommandBuffer> was).  No CUDA, torch or C++ types appear in
 * the signatures, so the header binds from C, cgo, JNI, ctypes...
 * Reference citations are relative to /root/reference/Renderer.
 *
 * Contract:
 *   - the caller owns every buffer; the library allocates nothing per call;
 *     a context owns only small constant tables uploaded once;
 *   - calls are asynchronous and stream-ordered unless `wait` is non-zero
 *     (MetalBT709Decoder.m:486-489 waitUntilCompleted);
 *   - inputs must stay valid until the stream reaches the kernel (the reference
 *     retains its wrapped input textures for the same reason,
 *     MetalBT709Decoder.m:449-452);
 *   - a context is bound to one device; calls on one (ctx, stream) pair may come
 *     from any thread; there is no global state;
 *   - errors are returned, never asserted; there is no CPU fallback: without a
 *     CUDA device bt709_ctx_create fails with BT709_E_CUDA.
 *
 * Pixel layouts:
 *   Y plane     w x h bytes, video range [16,235] (out-of-range bytes saturate)
 *   CbCr plane  ceil(w/2) x ceil(h/2) byte pairs, Cb at the even byte
 *               (CVPixelBufferUtils.h:358-361)
 *   BGRA        bytes B,G,R,A; as a little-endian word A<<24 | R<<16 | G<<8 | B
 *               (CVPixelBufferUtils.h:304-318, MetalRenderContext.m:188-198)
 * A batch of `nframes` equally sized frames is described per plane by
 * {ptr, pitch, frame_stride}: frame f, row r starts at ptr + f*frame_stride + r*pitch.
 */
#ifndef BT709_B200_H
#define BT709_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define BT709_API __attribute__((visibility("default")))
#else
#define BT709_API
#endif

/* Transfer function of the Y'CbCr data.  0..2 follow MetalBT709Gamma
 * (MetalBT709Decoder.h:15-19); 3 is the BT.709 spec curve that only the CPU
 * path of the reference applies (BT709.h:668-738). */
typedef enum {
  BT709_GAMMA_APPLE = 0,  /* Apple 1.961 curve   AAPLShaders.metal:336-357, BT709.h:821-908 */
  BT709_GAMMA_SRGB = 1,   /* sRGB curve          AAPLShaders.metal:361-382, BT709.h:948-1003 */
  BT709_GAMMA_LINEAR = 2, /* no curve            AAPLShaders.metal:387-407, BT709.h:517-556 */
  BT709_GAMMA_BT709 = 3,  /* BT.709 spec curve   BT709.h:668-738 (decode and the packed 4:4:4 pair) */
  /* packed 4:4:4 entries only: LINEAR-light RGB bytes <-> BT.709 with the spec curve, the reference's
   * "previous API" BT709_convertRGBToYCbCr / BT709_convertYCbCrToRGB with applyGammaMap = 1
   * (BT709.h:320-345, 560-607; EmptyiOSTests/CoreImageMetalFilterTests.m:159-360) */
  BT709_GAMMA_BT709_LINEAR_RGB = 4
} bt709_gamma;

typedef enum {
  BT709_OK = 0,
  BT709_E_NULL = -1,       /* null ctx / plane / pointer          (MetalBT709Decoder.m:48-54) */
  BT709_E_DIMS = -2,       /* w,h <= 0; odd w/h on the 4:2:0 encode (BGRAToBT709Converter.m:41-46,
                              CVPixelBufferUtils.h:275-276); nframes <= 0 */
  BT709_E_PITCH = -3,      /* pitch smaller than a row, or frame_stride smaller than a frame */
  BT709_E_GAMMA = -4,      /* unknown gamma (MetalBT709Decoder.m:125,177 assert(0)), or the
                              BT.709 spec curve on the 4:2:0 encode (BT709.h:1135-1137) */
  BT709_E_ALPHA_MODE = -5, /* alpha plane with a gamma other than sRGB
                              (MetalBT709Decoder.m:165-169 forces sRGB; here it is explicit) */
  BT709_E_CUDA = -6,       /* CUDA error; see bt709_last_cuda_error */
  BT709_E_SCRATCH = -7,    /* two-pass decode+rescale needs an intermediate plane */
  BT709_E_UNSUPPORTED = -8 /* bt709_decode_rescale_nv12_fused: the fused kernel does not apply to this geometry */
} bt709_status;

/* One plane (or packed image) of a batch of equally sized frames, in device memory. */
typedef struct {
  void *ptr;           /* device pointer to frame 0, row 0 */
  size_t pitch;        /* bytes between rows */
  size_t frame_stride; /* bytes between frames (ignored when nframes == 1) */
} bt709_plane;

typedef struct bt709_ctx bt709_ctx;

/* Replaces MetalRenderContext -setupMetal + MetalBT709Decoder -setupMetal
 * (MetalRenderContext.m:36-74, MetalBT709Decoder.m:46-104): binds a device,
 * builds and uploads the transfer-curve tables. */
BT709_API int bt709_ctx_create(int device, bt709_ctx **out);
BT709_API void bt709_ctx_destroy(bt709_ctx *ctx);
BT709_API int bt709_ctx_device(const bt709_ctx *ctx);
/* cudaError_t of the last failing CUDA call made through this context (0 if none). */
BT709_API int bt709_last_cuda_error(const bt709_ctx *ctx);
BT709_API const char *bt709_status_string(int status);
/* Number of kernels launched through this context since creation (bench accounting). */
BT709_API unsigned long long bt709_launch_count(const bt709_ctx *ctx);

/* Host-only self-check of the constant tables (needs no GPU; used by tests/test_tables_cpu.py):
 *   1. for every gamma mode and every one of the 2^24 (Y,Cb,Cr) inputs, the cell-LUT lookup of the
 *      G-channel matrix output equals the direct double-precision evaluation of the curve;
 *   1b. the biased log-cell LUT of the sRGB encode equals the step function around every threshold and cell
 *      boundary and on 4 M floats spread over [0, 1];
 *   2. the three-instruction constant division of the encode path equals IEEE division for every
 *      8-bit RGB triple.
 * Returns 0 on success, else the 1-based index of the first failing check. */
BT709_API int bt709_selftest_host(void);

/* Pass 1.  Replaces -[MetalBT709Decoder decodeBT709:alphaPixelBuffer:bgraSRGBTexture:
 * commandBuffer:renderPassDescriptor:renderWidth:renderHeight:waitUntilCompleted:]
 * (MetalBT709Decoder.h:65-72, .m:217-545) and the kernels/fragments of
 * AAPLShaders.metal:275-438: nearest 4:2:0 chroma upsample, BT.709 matrix,
 * transfer function removal, sRGB encode, BGRA8 store.
 *   alpha_y  nullable; Y plane of the separate alpha stream (MetalBT709Decoder.m:376-378),
 *            written to the A byte without premultiplying (AAPLShaders.metal:411-438);
 *            requires gamma == BT709_GAMMA_SRGB.  Without it A = 255 (AAPLShaders.metal:243).
 *   w, h     frame size; output has the same size (MetalBT709Decoder.m:272-290). */
BT709_API int bt709_decode_nv12(const bt709_ctx *ctx, bt709_gamma gamma, const bt709_plane *y,
                                const bt709_plane *cbcr, const bt709_plane *alpha_y,
                                const bt709_plane *bgra_out, int w, int h, int nframes, void *stream,
                                int wait);

/* Pass 2.  Replaces -[MetalScaleRenderContext renderScaled:...bgraTexture:]
 * (MetalScaleRenderContext.h:34-40, .m:55-105) + samplingShader (AAPLShaders.metal:73-85):
 * bilinear resample of an sRGB BGRA8 image in linear light, clamp-to-edge,
 * pixel-centre sampling, sRGB re-encode on store. */
BT709_API int bt709_rescale_bgra(const bt709_ctx *ctx, const bt709_plane *bgra_in, int in_w, int in_h,
                                 const bt709_plane *bgra_out, int out_w, int out_h, int nframes,
                                 void *stream, int wait);

/* The per-frame policy of -[AAPLRenderer drawInMTKView:] (AAPLRenderer.m:891-977):
 * output size == frame size -> pass 1 straight into bgra_out; otherwise pass 1 into the same-size sRGB
 * intermediate `scratch` (the reference's _resizeTexture) and pass 2 into bgra_out (two kernels).
 * `scratch` may be NULL only in the same-size case. */
BT709_API int bt709_decode_rescale_nv12(const bt709_ctx *ctx, bt709_gamma gamma, const bt709_plane *y,
                                        const bt709_plane *cbcr, const bt709_plane *alpha_y,
                                        const bt709_plane *scratch, const bt709_plane *bgra_out, int w,
                                        int h, int out_w, int out_h, int nframes, void *stream,
                                        int wait);

/* The same operation FUSED into one kernel (MetalBT709Decoder.m:612-618 is the author's note
 * that a single pass would halve the traffic): each CTA decodes the source rows of one band of 64 output rows into
 * a slice of `scratch` that it overwrites band after band (~400 KB per SM, resident in L2), then resamples them, so
 * the intermediate image is never written out as a whole and `scratch` does NOT hold it afterwards.  Bit-identical
 * to the two-pass result.  Needs: no alpha stream, w a multiple of 128, even h, even out_w >= w, 8-byte aligned
 * destination rows, tightly packed scratch rows (pitch == 4 w) and frames, at least one slice of scratch
 * ((64 h / out_h + 12) rows; frames back to back: the whole batch-sized area is used, one slice per SM; any other
 * frame stride: only the first frame's area); returns BT709_E_UNSUPPORTED otherwise.  bt709_decode_rescale_nv12 itself
 * takes this path when the environment variable BT709_FUSED is set.
 * Measured on B200 (1080p -> 1440p): 39 % less DRAM traffic, 12 % lower frame rate than the two kernels (both
 * halves are instruction-bound), see DESIGN.md section 4.4. */
BT709_API int bt709_decode_rescale_nv12_fused(const bt709_ctx *ctx, bt709_gamma gamma, const bt709_plane *y,
                                              const bt709_plane *cbcr, const bt709_plane *scratch,
                                              const bt709_plane *bgra_out, int w, int h, int out_w, int out_h,
                                              int nframes, void *stream, int wait);

/* Inverse path.  Replaces cvpbu_ycbcr_subsample (CVPixelBufferUtils.h:241-399) +
 * BT709_average_pixel_values (BT709.h:1349-1509) as driven by
 * +[BGRAToBT709Converter createYCbCrFromCGImage:isLinear:asSRGBGamma:]
 * (BGRAToBT709Converter.m:875-1037): BGRA8 -> video-range NV12, chroma = 2x2 mean
 * in linear light, requantised to 8 bits in the output gamma before the matrix.
 * gamma_in/gamma_out in {APPLE, SRGB, LINEAR}; w and h must be even. */
BT709_API int bt709_encode_bgra_to_nv12(const bt709_ctx *ctx, bt709_gamma gamma_in, bt709_gamma gamma_out,
                                        const bt709_plane *bgra_in, const bt709_plane *y_out,
                                        const bt709_plane *cbcr_out, int w, int h, int nframes,
                                        void *stream, int wait);

/* Test-harness 4:4:4 packed path of the reference, Apple curve only.
 * convert:   +[BGRAToBT709Converter convert:...type:Software]  (BGRAToBT709Converter.m:33-57, 89-144)
 *            BGRA word -> Cr<<16 | Cb<<8 | Y
 * unconvert: +[BGRAToBT709Converter unconvert:...type:Software] (BGRAToBT709Converter.m:61-85, 146-198)
 *            word -> R<<16 | G<<8 | B with A = 0
 * Both require even w and h and tightly packed w*h uint32 device arrays. */
BT709_API int bt709_convert_packed444(const bt709_ctx *ctx, const uint32_t *bgra_in, uint32_t *ycbcr_out,
                                      int w, int h, void *stream, int wait);
BT709_API int bt709_unconvert_packed444(const bt709_ctx *ctx, const uint32_t *ycbcr_in, uint32_t *bgra_out,
                                        int w, int h, void *stream, int wait);
/* The same two kernels with the curve as an argument: the three per-pixel 4:4:4 encoder/decoder pairs of the
 * reference's CPU path, which its exhaustive 2^24-colour round-trip tests exercise
 * (EmptyiOSTests/CoreImageMetalFilterTests.m:420-537, 557-674, 696-813):
 *   BT709_GAMMA_BT709  BT709_from_sRGB_convertRGBToYCbCr / BT709_to_sRGB_convertYCbCrToRGB      (BT709.h:610-738)
 *   BT709_GAMMA_APPLE  Apple196_from_sRGB_convertRGBToYCbCr / Apple196_to_sRGB_convertYCbCrToRGB (BT709.h:743-908)
 *   BT709_GAMMA_SRGB   sRGB_from_sRGB_convertRGBToYCbCr / sRGB_to_sRGB_convertYCbCrToRGB         (BT709.h:914-1003)
 *   BT709_GAMMA_BT709_LINEAR_RGB  BT709_convertRGBToYCbCr / BT709_convertYCbCrToRGB, applyGammaMap = 1
 *                      (with applyGammaMap = 0 that pair is the BT709_GAMMA_SRGB one: byteNorm, matrix, round)
 * BT709_GAMMA_LINEAR has no such pair in the reference: BT709_E_GAMMA. */
BT709_API int bt709_convert_packed444_gamma(const bt709_ctx *ctx, bt709_gamma gamma, const uint32_t *bgra_in,
                                            uint32_t *ycbcr_out, int w, int h, void *stream, int wait);
BT709_API int bt709_unconvert_packed444_gamma(const bt709_ctx *ctx, bt709_gamma gamma, const uint32_t *ycbcr_in,
                                              uint32_t *bgra_out, int w, int h, void *stream, int wait);
/* +[BGRAToBT709Converter copyBT709ToCoreVideo:cvPixelBuffer:] (BGRAToBT709Converter.m:1042-1099):
 * packed words -> NV12 planes; chroma is taken from the even column of every row, so the
 * odd row's value is the one that survives. */
BT709_API int bt709_copy_packed444_to_nv12(const bt709_ctx *ctx, const uint32_t *ycbcr_in,
                                           const bt709_plane *y_out, const bt709_plane *cbcr_out, int w,
                                           int h, void *stream, int wait);

#ifdef __cplusplus
}
#endif
#endif /* BT709_B200_H */
