/*
 * metal_bt709_decoder.h — host-side mirror of the reference's operator objects.
 *
 * The reference's Objective-C classes take CVPixelBufferRef / id<MTLTexture> /
 * id<MTLCommandBuffer>.  On Linux there is no Objective-C runtime, CoreVideo or
 * Metal, so this layer restates the same objects, method names, argument meaning
 * and BOOL/FALSE error behaviour in C over HOST buffers, and drives the CUDA path
 * of bt709_b200.h underneath (upload -> kernel -> read-back on a CUDA stream).
 * A maintainer porting a call site changes `[obj method:arg ...]` into
 * `Class_method(obj, arg, ...)`; see INTEGRATION.md.
 *
 *   reference class / function                      here
 *   ------------------------------------------------------------------------------------
 *   CVPixelBufferRef (420YpCbCr8BiPlanarVideoRange)  MBDPixelBuffer
 *   id<MTLTexture> BGRA8Unorm[_sRGB]                 MBDTexture
 *   MetalRenderContext        (.h:20-105)            MetalRenderContext_*
 *   MetalBT709Decoder         (.h:21-74)             MetalBT709Decoder_*
 *   MetalScaleRenderContext   (.h:20-42)             MetalScaleRenderContext_*
 *   BGRAToBT709Converter      (.h:30-133)            BGRAToBT709Converter_*
 *   cvpbu_ycbcr_subsample     (CVPixelBufferUtils.h:241)  cvpbu_ycbcr_subsample
 *   AAPLRenderer -drawInMTKView: (.m:874-987)        AAPLRenderer_drawFrame
 *
 * Return convention: BOOL-returning reference methods return 1 (TRUE) / 0 (FALSE).
 */
#ifndef METAL_BT709_DECODER_H
#define METAL_BT709_DECODER_H

#include <stddef.h>
#include <stdint.h>

#include "bt709_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* MetalBT709Decoder.h:15-19 */
typedef enum {
  MetalBT709GammaApple = 0, /* default */
  MetalBT709GammaSRGB = 1,
  MetalBT709GammaLinear = 2,
  /* Extension (not in the reference's enum): the BT.709 SPEC curve.  The reference's Metal kernels never apply it
   * (AAPLShaders.metal:94-107 is defined, its caller :144-152 is commented out); its CPU path does
   * (BT709_to_sRGB_convertYCbCrToRGB, BT709.h:668-738).  Checked against the same transfer tag as the Apple curve. */
  MetalBT709GammaBT709Spec = 3
} MetalBT709Gamma;

/* BT709.h:20-25 (note the different numbering) */
typedef enum { BT709GammaSrgb = 1, BT709GammaApple = 2, BT709GammaLinear = 3 } BT709Gamma;

/* BGRAToBT709Converter.h:24-28 */
typedef enum {
  BGRAToBT709ConverterSoftware = 0,
  BGRAToBT709ConverterVImage = 1,
  BGRAToBT709ConverterMetal = 2
} BGRAToBT709ConverterTypeEnum;

/* CoreVideo colour tags the decoder checks (MetalBT709Decoder.m:311-368) */
typedef enum {
  kMBDYCbCrMatrix_Untagged = 0,
  kMBDYCbCrMatrix_ITU_R_709_2 = 1,
  kMBDYCbCrMatrix_ITU_R_601_4 = 2
} MBDYCbCrMatrix;
typedef enum {
  kMBDTransferFunction_Untagged = 0,
  kMBDTransferFunction_ITU_R_709_2 = 1,
  kMBDTransferFunction_sRGB = 2,
  kMBDTransferFunction_Linear = 3
} MBDTransferFunction;

/* CVPixelBufferRef stand-in: NV12 video range in (pinned) host memory. */
typedef struct MBDPixelBuffer {
  int width, height;
  uint8_t *y;
  size_t yBytesPerRow;
  uint8_t *cbcr; /* (width+1)/2 pairs x (height+1)/2 rows, Cb first */
  size_t cbcrBytesPerRow;
  int yCbCrMatrix;      /* MBDYCbCrMatrix */
  int transferFunction; /* MBDTransferFunction */
  int ownsMemory;
} MBDPixelBuffer;

/* id<MTLTexture> stand-in: BGRA8 in (pinned) host memory; isSRGB as in
 * -makeBGRATexture:pixels:usage:isSRGB: (MetalRenderContext.m:79-120). */
typedef struct MBDTexture {
  int width, height;
  uint8_t *pixels;
  size_t bytesPerRow;
  int isSRGB;
  int ownsMemory;
} MBDTexture;

typedef struct MetalRenderContext MetalRenderContext;
typedef struct MetalBT709Decoder MetalBT709Decoder;
typedef struct MetalScaleRenderContext MetalScaleRenderContext;
typedef struct AAPLRenderer AAPLRenderer;

/* ---- MetalRenderContext (device, queue -> CUDA device, stream ring, staging) ---- */
BT709_API MetalRenderContext *MetalRenderContext_new(int device);
BT709_API void MetalRenderContext_release(MetalRenderContext *mrc);
BT709_API int MetalRenderContext_setupMetal(MetalRenderContext *mrc); /* MetalRenderContext.m:36-74 */
BT709_API MBDTexture *MetalRenderContext_makeBGRATexture(MetalRenderContext *mrc, int width, int height,
                                                         const uint32_t *pixels, int isSRGB);
BT709_API void MBDTextureRelease(MBDTexture *t);
/* wait for every command queued through this context (commit + waitUntilCompleted) */
BT709_API int MetalRenderContext_waitUntilCompleted(MetalRenderContext *mrc);
BT709_API bt709_ctx *MetalRenderContext_ctx(MetalRenderContext *mrc);
/* kernels launched through this context so far */
BT709_API unsigned long long MetalRenderContext_launchCount(MetalRenderContext *mrc);

/* ---- MetalBT709Decoder ---- */
BT709_API MetalBT709Decoder *MetalBT709Decoder_new(void);
BT709_API void MetalBT709Decoder_release(MetalBT709Decoder *d);
BT709_API void MetalBT709Decoder_setMetalRenderContext(MetalBT709Decoder *d, MetalRenderContext *mrc);
BT709_API void MetalBT709Decoder_setGamma(MetalBT709Decoder *d, MetalBT709Gamma gamma);
BT709_API MetalBT709Gamma MetalBT709Decoder_gamma(const MetalBT709Decoder *d);
BT709_API void MetalBT709Decoder_setUseComputeRenderer(MetalBT709Decoder *d, int flag);
BT709_API void MetalBT709Decoder_setHasAlphaChannel(MetalBT709Decoder *d, int flag);
BT709_API int MetalBT709Decoder_setupMetal(MetalBT709Decoder *d); /* MetalBT709Decoder.m:46-104 */
/* -decodeBT709:alphaPixelBuffer:bgraSRGBTexture:commandBuffer:renderPassDescriptor:
 *  renderWidth:renderHeight:waitUntilCompleted:  (MetalBT709Decoder.h:65-72).
 * commandBuffer: a cudaStream_t, or NULL to use the context's own stream ring.
 * The renderPassDescriptor (render straight into a view) case is AAPLRenderer_drawFrame. */
BT709_API int MetalBT709Decoder_decodeBT709(MetalBT709Decoder *d, const MBDPixelBuffer *yCbCrPixelBuffer,
                                            const MBDPixelBuffer *alphaPixelBuffer, MBDTexture *bgraSRGBTexture,
                                            void *commandBuffer, int renderWidth, int renderHeight,
                                            int waitUntilCompleted);

/* ---- MetalScaleRenderContext ---- */
BT709_API MetalScaleRenderContext *MetalScaleRenderContext_new(void);
BT709_API void MetalScaleRenderContext_release(MetalScaleRenderContext *s);
/* -renderScaled:mtkView:renderWidth:renderHeight:commandBuffer:renderPassDescriptor:bgraTexture:
 * (MetalScaleRenderContext.h:34-40); `drawable` stands for the MTKView's current drawable. */
BT709_API int MetalScaleRenderContext_renderScaled(MetalScaleRenderContext *s, MetalRenderContext *mrc,
                                                   MBDTexture *drawable, int renderWidth, int renderHeight,
                                                   void *commandBuffer, const MBDTexture *bgraTexture,
                                                   int waitUntilCompleted);

/* ---- BGRAToBT709Converter ---- */
BT709_API MBDPixelBuffer *BGRAToBT709Converter_createCoreVideoYCbCrBuffer(int width, int height); /* .m:471-494 */
BT709_API void MBDPixelBufferRelease(MBDPixelBuffer *pb);
BT709_API int BGRAToBT709Converter_setBT709Attributes(MBDPixelBuffer *pb); /* .m:387-445 */
/* +convert:outBT709Pixels:width:height:type: / +unconvert: (BGRAToBT709Converter.h:34-46).
 * Software = the CPU arithmetic of the reference, executed by the CUDA kernels;
 * VImage is a closed Apple library and Metal is FALSE in the reference too (.m:52-54). */
BT709_API int BGRAToBT709Converter_convert(MetalRenderContext *mrc, const uint32_t *inBGRAPixels,
                                           uint32_t *outBT709Pixels, int width, int height,
                                           BGRAToBT709ConverterTypeEnum type);
BT709_API int BGRAToBT709Converter_unconvert(MetalRenderContext *mrc, const uint32_t *inBT709Pixels,
                                             uint32_t *outBGRAPixels, int width, int height,
                                             BGRAToBT709ConverterTypeEnum type);
/* +copyBT709ToCoreVideo:cvPixelBuffer: (.m:1042-1099) */
BT709_API int BGRAToBT709Converter_copyBT709ToCoreVideo(MetalRenderContext *mrc, const uint32_t *inBT709Pixels,
                                                        MBDPixelBuffer *cvPixelBuffer);
/* +createYCbCrFromCGImage:isLinear:asSRGBGamma: (.m:875-1037) over a BGRA8 image instead of a
 * CGImage: picks (Linear,Linear) / (Srgb,Srgb) / (Srgb,Apple) and tags the buffer. */
BT709_API MBDPixelBuffer *BGRAToBT709Converter_createYCbCrFromBGRA(MetalRenderContext *mrc, const uint32_t *bgraPixels,
                                                                  int width, int height, int isLinear,
                                                                  int asSRGBGamma);
/* +copyYCBCr:Y:Cb:Cr:dump: (BGRAToBT709Converter.m:688-790): NV12 buffer -> tightly packed planar
 * Y (w*h), Cb and Cr ((w/2)*(h/2)) as the .y4m writer wants them; the CbCr split runs on the GPU. */
BT709_API int BGRAToBT709Converter_copyYCBCr(MetalRenderContext *mrc, const MBDPixelBuffer *cvPixelBuffer, uint8_t *Y,
                                             uint8_t *Cb, uint8_t *Cr);
/* cvpbu_ycbcr_subsample (CVPixelBufferUtils.h:241-242) */
BT709_API int cvpbu_ycbcr_subsample(MetalRenderContext *mrc, const uint32_t *inPixelsPtr, int width, int height,
                                    MBDPixelBuffer *dst, BT709Gamma inputGamma, BT709Gamma outputGamma);

/* ---- AAPLRenderer: the per-frame 1-pass / 2-pass policy (AAPLRenderer.m:874-987) ---- */
BT709_API AAPLRenderer *AAPLRenderer_new(MetalRenderContext *mrc, MetalBT709Gamma decodeGamma, int hasAlphaChannel);
BT709_API void AAPLRenderer_release(AAPLRenderer *r);
/* drawable size == frame size: decode straight into the drawable; otherwise decode into the
 * same-size intermediate (_resizeTexture, kept on the device) and rescale into the drawable. */
BT709_API int AAPLRenderer_drawFrame(AAPLRenderer *r, const MBDPixelBuffer *yCbCr, const MBDPixelBuffer *alpha,
                                     MBDTexture *drawable, int waitUntilCompleted);

#ifdef __cplusplus
}
#endif
#endif /* METAL_BT709_DECODER_H */
