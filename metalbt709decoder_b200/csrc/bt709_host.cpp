// bt709_host.cpp — reference-facing host layer: the reference's operator objects restated in C
// over host buffers (include/metal_bt709_decoder.h), driving the device C ABI (bt709_b200.h).
//
// Per call: pinned host planes -> cudaMemcpy2DAsync -> kernel -> cudaMemcpy2DAsync -> pinned
// host texture, all on one stream of a small ring, so consecutive asynchronous calls overlap
// upload, compute and read-back.  No arithmetic happens on the host.
#include "bt709_host.h"

#include "../../include/y4m_b200.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

using namespace bt709host;

// ------------------------------------------------------------------------------ helpers
bool DevBuf::ensure(size_t bytes) {
  if (bytes <= cap) return true;
  if (p) cudaFree(p);
  p = nullptr;
  cap = 0;
  if (cudaMalloc(&p, bytes) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  cap = bytes;
  return true;
}

void DevBuf::release() {
  if (p) cudaFree(p);
  p = nullptr;
  cap = 0;
}

namespace {

void *host_alloc(size_t bytes, int *pinned) {
  void *p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) == cudaSuccess) {
    *pinned = 1;
    return p;
  }
  cudaGetLastError();  // no device: plain memory still lets callers build buffers
  *pinned = 0;
  if (posix_memalign(&p, 256, bytes ? bytes : 1) != 0) return nullptr;
  return p;
}

void host_free(void *p, int pinned) {
  if (!p) return;
  if (pinned)
    cudaFreeHost(p);
  else
    free(p);
}

struct Guard {
  int prev = -1;
  explicit Guard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) cudaSetDevice(dev); else prev = -1;
  }
  ~Guard() { if (prev >= 0) cudaSetDevice(prev); }
};

// pick the stream + staging slot for one call.  With a caller stream the slot's buffers are
// still used, ordered after the slot's previous work by an event wait on the device.
Slot *acquire(MetalRenderContext *mrc, void *commandBuffer, cudaStream_t *stream) {
  Slot *s;
  {
    std::lock_guard<std::mutex> lk(mrc->mu);
    s = &mrc->slots[mrc->next];
    mrc->next = (mrc->next + 1) % kSlots;
  }
  s->busy.lock();  // released by finish(); more than kSlots concurrent callers queue up here
  *stream = commandBuffer ? (cudaStream_t) commandBuffer : s->stream;
  // Always order this call after the slot's previous work: the previous call may have run on a caller stream
  // (or this one does), and the staging buffers are shared.  A wait on an event recorded on the same stream, or
  // never recorded, is a no-op.
  cudaStreamWaitEvent(*stream, s->done, 0);
  return s;
}

bool upload2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t row_bytes, size_t rows, cudaStream_t st) {
  return cudaMemcpy2DAsync(dst, dpitch, src, spitch, row_bytes, rows, cudaMemcpyHostToDevice, st) == cudaSuccess;
}
bool download2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t row_bytes, size_t rows, cudaStream_t st) {
  if (dpitch == spitch && dpitch == row_bytes)  // tightly packed on both sides: one 1-D copy
    return cudaMemcpyAsync(dst, src, row_bytes * rows, cudaMemcpyDeviceToHost, st) == cudaSuccess;
  return cudaMemcpy2DAsync(dst, dpitch, src, spitch, row_bytes, rows, cudaMemcpyDeviceToHost, st) == cudaSuccess;
}

bool finish(Slot *s, cudaStream_t st, bool ok, int wait) {
  cudaEventRecord(s->done, st);
  s->busy.unlock();
  if (!ok) {
    cudaGetLastError();
    return false;
  }
  if (wait && cudaStreamSynchronize(st) != cudaSuccess) return false;
  return true;
}

bt709_gamma to_abi_gamma(MetalBT709Gamma g) {
  switch (g) {
    case MetalBT709GammaSRGB: return BT709_GAMMA_SRGB;
    case MetalBT709GammaLinear: return BT709_GAMMA_LINEAR;
    case MetalBT709GammaBT709Spec: return BT709_GAMMA_BT709;
    default: return BT709_GAMMA_APPLE;
  }
}

int to_abi_gamma(BT709Gamma g, bt709_gamma *out) {
  switch (g) {
    case BT709GammaSrgb: *out = BT709_GAMMA_SRGB; return 1;
    case BT709GammaApple: *out = BT709_GAMMA_APPLE; return 1;
    case BT709GammaLinear: *out = BT709_GAMMA_LINEAR; return 1;
  }
  return 0;  // BT709.h:1135-1137 assert(0)
}

// stage the NV12 planes of `pb` (and the alpha stream's Y plane) into the slot
struct Staged {
  bt709_plane y, c, a;
  bool has_a = false;
};

bool stage_nv12(Slot *s, const MBDPixelBuffer *pb, const MBDPixelBuffer *alpha, cudaStream_t st, Staged *out) {
  const int w = pb->width, h = pb->height;
  const size_t cw = (size_t) ((w + 1) / 2) * 2, ch = (size_t) (h + 1) / 2;
  // Buffers made by BGRAToBT709Converter_createCoreVideoYCbCrBuffer hold both planes in ONE pinned block (CbCr right
  // behind Y) with 64-byte row pitches: mirror that layout on the device and the whole frame is a single 1-D copy
  // (one DMA descriptor instead of two pitched copies of 1080 + 540 rows).
  const size_t ybytes = pb->yBytesPerRow * (size_t) h, cbytes = pb->cbcrBytesPerRow * ch;
  if (pb->cbcr == pb->y + ybytes && pb->yBytesPerRow % 16 == 0 && pb->cbcrBytesPerRow % 16 == 0 && ybytes % 256 == 0) {
    if (!s->y.ensure(ybytes + cbytes)) return false;
    if (cudaMemcpyAsync(s->y.p, pb->y, ybytes + cbytes, cudaMemcpyHostToDevice, st) != cudaSuccess) return false;
    out->y = {s->y.p, pb->yBytesPerRow, ybytes};
    out->c = {(uint8_t *) s->y.p + ybytes, pb->cbcrBytesPerRow, cbytes};
  } else {
    const size_t py = align_up(w, 256), pc = align_up(cw, 256);
    if (!s->y.ensure(py * h) || !s->c.ensure(pc * ch)) return false;
    if (!upload2d(s->y.p, py, pb->y, pb->yBytesPerRow, w, h, st)) return false;
    if (!upload2d(s->c.p, pc, pb->cbcr, pb->cbcrBytesPerRow, cw, ch, st)) return false;
    out->y = {s->y.p, py, py * h};
    out->c = {s->c.p, pc, pc * ch};
  }
  const size_t py = align_up(w, 256);
  out->has_a = alpha != nullptr;
  if (alpha) {
    // only the Y plane of the alpha stream is read (MetalBT709Decoder.m:376-378)
    if (!s->a.ensure(py * h)) return false;
    if (!upload2d(s->a.p, py, alpha->y, alpha->yBytesPerRow, w, h, st)) return false;
    out->a = {s->a.p, py, py * h};
  }
  return true;
}

// the guard clauses of -processBT709ToSRGB: (MetalBT709Decoder.m:272-368), in order
bool decode_args_ok(const MetalBT709Decoder *d, const MBDPixelBuffer *pb, const MBDPixelBuffer *alpha, int outW,
                    int outH, int renderWidth, int renderHeight, MetalBT709Gamma gamma) {
  if (!pb || !pb->y || !pb->cbcr) return false;
  const int w = pb->width, h = pb->height;
  if (outW >= 0 && (w != outW || h != outH)) return false;                 // :278-290
  if (w != renderWidth || h != renderHeight) return false;                 // :292-300
  if (alpha) {
    if (!alpha->y) return false;
    if (alpha->width != w || alpha->height != h) return false;             // :303-305
  }
  if (pb->yCbCrMatrix != kMBDYCbCrMatrix_ITU_R_709_2) return false;        // :311-318
  const int tf = pb->transferFunction;                                     // :320-353
  if ((gamma == MetalBT709GammaApple || gamma == MetalBT709GammaBT709Spec) && tf != kMBDTransferFunction_ITU_R_709_2) return false;
  if (gamma == MetalBT709GammaSRGB && tf != kMBDTransferFunction_sRGB) return false;
  if (gamma == MetalBT709GammaLinear && tf != kMBDTransferFunction_Linear) return false;
  if (alpha && alpha->transferFunction != kMBDTransferFunction_Linear) return false;  // :357-368
  (void) d;
  return true;
}

}  // namespace

extern "C" {

// ================================================================= MetalRenderContext
MetalRenderContext *MetalRenderContext_new(int device) {
  MetalRenderContext *m = new (std::nothrow) MetalRenderContext();
  if (m) m->device = device;
  return m;
}

int MetalRenderContext_setupMetal(MetalRenderContext *mrc) {
  if (!mrc) return 0;
  std::lock_guard<std::mutex> lk(mrc->mu);  // may be called from any thread
  if (mrc->ctx) return 1;  // idempotent, as in the reference
  bt709_ctx *ctx = nullptr;
  if (bt709_ctx_create(mrc->device, &ctx) != BT709_OK) return 0;
  Guard g(mrc->device);
  bool ok = true;
  for (int i = 0; i < kSlots && ok; i++) {
    ok = cudaStreamCreateWithFlags(&mrc->slots[i].stream, cudaStreamNonBlocking) == cudaSuccess &&
         cudaEventCreateWithFlags(&mrc->slots[i].done, cudaEventDisableTiming) == cudaSuccess;
  }
  if (!ok) {  // tear down the partial setup; the context is published only when everything exists
    for (int i = 0; i < kSlots; i++) {
      if (mrc->slots[i].done) cudaEventDestroy(mrc->slots[i].done);
      if (mrc->slots[i].stream) cudaStreamDestroy(mrc->slots[i].stream);
      mrc->slots[i].done = nullptr;
      mrc->slots[i].stream = nullptr;
    }
    cudaGetLastError();
    bt709_ctx_destroy(ctx);
    return 0;
  }
  mrc->ctx = ctx;
  return 1;
}

void MetalRenderContext_release(MetalRenderContext *mrc) {
  if (!mrc) return;
  if (mrc->ctx) {
    Guard g(mrc->device);
    for (int i = 0; i < kSlots; i++) {
      Slot &s = mrc->slots[i];
      if (s.stream) cudaStreamSynchronize(s.stream);
      s.y.release(); s.c.release(); s.a.release(); s.out.release(); s.mid.release(); s.in.release();
      if (s.done) cudaEventDestroy(s.done);
      if (s.stream) cudaStreamDestroy(s.stream);
    }
    bt709_ctx_destroy(mrc->ctx);
  }
  delete mrc;
}

int MetalRenderContext_waitUntilCompleted(MetalRenderContext *mrc) {
  if (!mrc || !mrc->ctx) return 0;
  Guard g(mrc->device);
  int ok = 1;
  for (int i = 0; i < kSlots; i++)
    if (cudaStreamSynchronize(mrc->slots[i].stream) != cudaSuccess) ok = 0;
  return ok;
}

bt709_ctx *MetalRenderContext_ctx(MetalRenderContext *mrc) { return mrc ? mrc->ctx : nullptr; }

unsigned long long MetalRenderContext_launchCount(MetalRenderContext *mrc) {
  return (mrc && mrc->ctx) ? bt709_launch_count(mrc->ctx) : 0ull;
}

MBDTexture *MetalRenderContext_makeBGRATexture(MetalRenderContext *mrc, int width, int height, const uint32_t *pixels,
                                               int isSRGB) {
  (void) mrc;
  if (width <= 0 || height <= 0) return nullptr;
  MBDTexture *t = new (std::nothrow) MBDTexture();
  if (!t) return nullptr;
  t->width = width;
  t->height = height;
  t->bytesPerRow = (size_t) width * 4;
  t->isSRGB = isSRGB;
  t->pixels = (uint8_t *) host_alloc(t->bytesPerRow * height, &t->ownsMemory);
  if (!t->pixels) { delete t; return nullptr; }
  t->ownsMemory += 1;  // 1 = malloc, 2 = pinned
  if (pixels)
    memcpy(t->pixels, pixels, t->bytesPerRow * height);
  else
    memset(t->pixels, 0, t->bytesPerRow * height);
  return t;
}

void MBDTextureRelease(MBDTexture *t) {
  if (!t) return;
  if (t->ownsMemory) host_free(t->pixels, t->ownsMemory == 2);
  delete t;
}

// ================================================================== MetalBT709Decoder
MetalBT709Decoder *MetalBT709Decoder_new(void) { return new (std::nothrow) MetalBT709Decoder(); }

void MetalBT709Decoder_release(MetalBT709Decoder *d) {
  if (!d) return;
  if (d->owns_mrc) MetalRenderContext_release(d->mrc);
  delete d;
}

void MetalBT709Decoder_setMetalRenderContext(MetalBT709Decoder *d, MetalRenderContext *mrc) {
  if (!d) return;
  if (d->owns_mrc && d->mrc != mrc) MetalRenderContext_release(d->mrc);
  d->mrc = mrc;
  d->owns_mrc = false;
}
void MetalBT709Decoder_setGamma(MetalBT709Decoder *d, MetalBT709Gamma g) { if (d) d->gamma = g; }
MetalBT709Gamma MetalBT709Decoder_gamma(const MetalBT709Decoder *d) { return d ? d->gamma : MetalBT709GammaApple; }
void MetalBT709Decoder_setUseComputeRenderer(MetalBT709Decoder *d, int f) { if (d) d->useComputeRenderer = f; }
void MetalBT709Decoder_setHasAlphaChannel(MetalBT709Decoder *d, int f) { if (d) d->hasAlphaChannel = f; }

int MetalBT709Decoder_setupMetal(MetalBT709Decoder *d) {
  if (!d) return 0;
  if (d->setup) return 1;
  if (!d->mrc) {  // MetalBT709Decoder.m:57-66: allocate a context when none was set
    d->mrc = MetalRenderContext_new(0);
    d->owns_mrc = true;
  }
  if (!d->mrc || !MetalRenderContext_setupMetal(d->mrc)) return 0;
  if ((unsigned) d->gamma > (unsigned) MetalBT709GammaBT709Spec) return 0;  // :125,177 assert(0)
  // :108 the compute pipeline does not support an alpha channel
  if (d->useComputeRenderer && d->hasAlphaChannel) return 0;
  // :165-169 an alpha channel forces the sRGB gamma
  if (d->hasAlphaChannel) d->gamma = MetalBT709GammaSRGB;
  d->setup = true;
  return 1;
}

int MetalBT709Decoder_decodeBT709(MetalBT709Decoder *d, const MBDPixelBuffer *pb, const MBDPixelBuffer *alpha,
                                  MBDTexture *tex, void *commandBuffer, int renderWidth, int renderHeight,
                                  int waitUntilCompleted) {
  if (!d || !MetalBT709Decoder_setupMetal(d)) return 0;  // MetalBT709Decoder.m:228-231
  if (!tex || !tex->pixels) return 0;
  if (!decode_args_ok(d, pb, alpha, tex->width, tex->height, renderWidth, renderHeight, d->gamma)) return 0;
  if ((alpha != nullptr) != (d->hasAlphaChannel != 0)) return 0;  // pipeline was built for one or the other
  MetalRenderContext *mrc = d->mrc;
  Guard g(mrc->device);
  cudaStream_t st;
  Slot *s = acquire(mrc, commandBuffer, &st);
  const int w = pb->width, h = pb->height;
  Staged in;
  bool ok = stage_nv12(s, pb, alpha, st, &in);
  const size_t po = align_up((size_t) w * 4, 256);
  ok = ok && s->out.ensure(po * h);
  if (ok) {
    bt709_plane out = {s->out.p, po, po * h};
    ok = bt709_decode_nv12(mrc->ctx, to_abi_gamma(d->gamma), &in.y, &in.c, in.has_a ? &in.a : nullptr, &out, w, h, 1,
                           st, 0) == BT709_OK;
    ok = ok && download2d(tex->pixels, tex->bytesPerRow, s->out.p, po, (size_t) w * 4, h, st);
  }
  return finish(s, st, ok, waitUntilCompleted) ? 1 : 0;
}

// ============================================================ MetalScaleRenderContext
MetalScaleRenderContext *MetalScaleRenderContext_new(void) { return new (std::nothrow) MetalScaleRenderContext(); }
void MetalScaleRenderContext_release(MetalScaleRenderContext *s) { delete s; }

int MetalScaleRenderContext_renderScaled(MetalScaleRenderContext *sc, MetalRenderContext *mrc, MBDTexture *drawable,
                                         int renderWidth, int renderHeight, void *commandBuffer,
                                         const MBDTexture *bgraTexture, int waitUntilCompleted) {
  if (!sc || !mrc || !drawable || !drawable->pixels || !bgraTexture || !bgraTexture->pixels) return 0;
  if (renderWidth <= 0 || renderHeight <= 0) return 0;  // MetalScaleRenderContext.m:63-64
  if (drawable->width < renderWidth || drawable->height < renderHeight) return 0;
  if (!MetalRenderContext_setupMetal(mrc)) return 0;
  Guard g(mrc->device);
  cudaStream_t st;
  Slot *s = acquire(mrc, commandBuffer, &st);
  const int iw = bgraTexture->width, ih = bgraTexture->height;
  const size_t pi = align_up((size_t) iw * 4, 256), po = align_up((size_t) renderWidth * 4, 256);
  bool ok = s->in.ensure(pi * ih) && s->out.ensure(po * renderHeight);
  ok = ok && upload2d(s->in.p, pi, bgraTexture->pixels, bgraTexture->bytesPerRow, (size_t) iw * 4, ih, st);
  if (ok) {
    bt709_plane in = {s->in.p, pi, pi * ih}, out = {s->out.p, po, po * renderHeight};
    ok = bt709_rescale_bgra(mrc->ctx, &in, iw, ih, &out, renderWidth, renderHeight, 1, st, 0) == BT709_OK;
    ok = ok && download2d(drawable->pixels, drawable->bytesPerRow, s->out.p, po, (size_t) renderWidth * 4, renderHeight, st);
  }
  return finish(s, st, ok, waitUntilCompleted) ? 1 : 0;
}

// =============================================================== BGRAToBT709Converter
MBDPixelBuffer *BGRAToBT709Converter_createCoreVideoYCbCrBuffer(int width, int height) {
  if (width <= 0 || height <= 0) return nullptr;
  MBDPixelBuffer *pb = new (std::nothrow) MBDPixelBuffer();
  if (!pb) return nullptr;
  pb->width = width;
  pb->height = height;
  // CoreVideo pads rows; 64-byte row alignment keeps both planes vector-friendly
  pb->yBytesPerRow = align_up(width, 64);
  pb->cbcrBytesPerRow = align_up((size_t) ((width + 1) / 2) * 2, 64);
  const size_t ybytes = pb->yBytesPerRow * height, cbytes = pb->cbcrBytesPerRow * ((height + 1) / 2);
  int pinned = 0;
  pb->y = (uint8_t *) host_alloc(ybytes + cbytes, &pinned);
  if (!pb->y) { delete pb; return nullptr; }
  pb->cbcr = pb->y + ybytes;
  pb->ownsMemory = pinned ? 2 : 1;
  memset(pb->y, 0, ybytes + cbytes);
  return pb;
}

void MBDPixelBufferRelease(MBDPixelBuffer *pb) {
  if (!pb) return;
  if (pb->ownsMemory) host_free(pb->y, pb->ownsMemory == 2);
  delete pb;
}

int BGRAToBT709Converter_setBT709Attributes(MBDPixelBuffer *pb) {
  if (!pb) return 0;
  pb->yCbCrMatrix = kMBDYCbCrMatrix_ITU_R_709_2;
  pb->transferFunction = kMBDTransferFunction_ITU_R_709_2;
  return 1;
}

static int packed444(MetalRenderContext *mrc, const uint32_t *in, uint32_t *out, int w, int h, bool forward) {
  if (!mrc || !in || !out) return 0;
  if (!MetalRenderContext_setupMetal(mrc)) return 0;
  Guard g(mrc->device);
  cudaStream_t st;
  Slot *s = acquire(mrc, nullptr, &st);
  const size_t bytes = (size_t) w * h * 4;
  bool ok = s->in.ensure(bytes) && s->out.ensure(bytes);
  ok = ok && cudaMemcpyAsync(s->in.p, in, bytes, cudaMemcpyHostToDevice, st) == cudaSuccess;
  if (ok) {
    int rc = forward ? bt709_convert_packed444(mrc->ctx, (const uint32_t *) s->in.p, (uint32_t *) s->out.p, w, h, st, 0)
                     : bt709_unconvert_packed444(mrc->ctx, (const uint32_t *) s->in.p, (uint32_t *) s->out.p, w, h, st, 0);
    ok = rc == BT709_OK;
    ok = ok && cudaMemcpyAsync(out, s->out.p, bytes, cudaMemcpyDeviceToHost, st) == cudaSuccess;
  }
  return finish(s, st, ok, 1) ? 1 : 0;
}

int BGRAToBT709Converter_convert(MetalRenderContext *mrc, const uint32_t *in, uint32_t *out, int width, int height,
                                 BGRAToBT709ConverterTypeEnum type) {
  if ((width % 2) != 0 || (height % 2) != 0) return 0;  // BGRAToBT709Converter.m:41-46
  if (type != BGRAToBT709ConverterSoftware) return 0;   // vImage: Apple-only; Metal: FALSE in the reference (:52-54)
  return packed444(mrc, in, out, width, height, true);
}

int BGRAToBT709Converter_unconvert(MetalRenderContext *mrc, const uint32_t *in, uint32_t *out, int width, int height,
                                   BGRAToBT709ConverterTypeEnum type) {
  if ((width % 2) != 0 || (height % 2) != 0) return 0;  // BGRAToBT709Converter.m:69-74
  if (type != BGRAToBT709ConverterSoftware) return 0;   // :80-82
  return packed444(mrc, in, out, width, height, false);
}

int BGRAToBT709Converter_copyBT709ToCoreVideo(MetalRenderContext *mrc, const uint32_t *in, MBDPixelBuffer *pb) {
  if (!mrc || !in || !pb || !pb->y || !pb->cbcr) return 0;
  if (!MetalRenderContext_setupMetal(mrc)) return 0;
  Guard g(mrc->device);
  cudaStream_t st;
  Slot *s = acquire(mrc, nullptr, &st);
  const int w = pb->width, h = pb->height;
  const size_t cw = (size_t) ((w + 1) / 2) * 2, ch = (size_t) (h + 1) / 2;
  const size_t py = align_up(w, 256), pc = align_up(cw, 256);
  bool ok = s->in.ensure((size_t) w * h * 4) && s->y.ensure(py * h) && s->c.ensure(pc * ch);
  ok = ok && cudaMemcpyAsync(s->in.p, in, (size_t) w * h * 4, cudaMemcpyHostToDevice, st) == cudaSuccess;
  if (ok) {
    bt709_plane y = {s->y.p, py, py * h}, c = {s->c.p, pc, pc * ch};
    ok = bt709_copy_packed444_to_nv12(mrc->ctx, (const uint32_t *) s->in.p, &y, &c, w, h, st, 0) == BT709_OK;
    ok = ok && download2d(pb->y, pb->yBytesPerRow, s->y.p, py, w, h, st);
    ok = ok && download2d(pb->cbcr, pb->cbcrBytesPerRow, s->c.p, pc, cw, ch, st);
  }
  return finish(s, st, ok, 1) ? 1 : 0;
}

int BGRAToBT709Converter_copyYCBCr(MetalRenderContext *mrc, const MBDPixelBuffer *pb, uint8_t *Y, uint8_t *Cb, uint8_t *Cr) {
  if (!mrc || !pb || !pb->y || !pb->cbcr || !Y || !Cb || !Cr) return 0;
  const int w = pb->width, h = pb->height;
  if ((w % 2) != 0 || (h % 2) != 0) return 0;  // BGRAToBT709Converter.m:697-698
  if (!MetalRenderContext_setupMetal(mrc)) return 0;
  for (int row = 0; row < h; row++) memcpy(Y + (size_t) row * w, pb->y + (size_t) row * pb->yBytesPerRow, w);
  Guard g(mrc->device);
  cudaStream_t st;
  Slot *s = acquire(mrc, nullptr, &st);
  const int cw = w / 2, ch = h / 2;
  const size_t pc = align_up((size_t) w, 256), pu = align_up((size_t) cw, 256);
  bool ok = s->c.ensure(pc * ch) && s->y.ensure(2 * pu * ch);
  ok = ok && upload2d(s->c.p, pc, pb->cbcr, pb->cbcrBytesPerRow, (size_t) w, ch, st);
  if (ok) {
    bt709_plane c = {s->c.p, pc, pc * ch}, u = {s->y.p, pu, pu * ch}, v = {(uint8_t *) s->y.p + pu * ch, pu, pu * ch};
    ok = bt709_cbcr_to_planar(mrc->ctx, &c, &u, &v, cw, ch, 1, st, 0) == BT709_OK;
    ok = ok && download2d(Cb, cw, u.ptr, pu, cw, ch, st);
    ok = ok && download2d(Cr, cw, v.ptr, pu, cw, ch, st);
  }
  return finish(s, st, ok, 1) ? 1 : 0;
}

int cvpbu_ycbcr_subsample(MetalRenderContext *mrc, const uint32_t *inPixelsPtr, int width, int height,
                          MBDPixelBuffer *dst, BT709Gamma inputGamma, BT709Gamma outputGamma) {
  if (!mrc || !inPixelsPtr || !dst || !dst->y || !dst->cbcr) return 0;
  if ((width % 2) != 0 || (height % 2) != 0) return 0;  // CVPixelBufferUtils.h:275-276
  if (dst->width != width || dst->height != height) return 0;
  bt709_gamma gi, go;
  if (!to_abi_gamma(inputGamma, &gi) || !to_abi_gamma(outputGamma, &go)) return 0;
  if (!MetalRenderContext_setupMetal(mrc)) return 0;
  Guard g(mrc->device);
  cudaStream_t st;
  Slot *s = acquire(mrc, nullptr, &st);
  const size_t pi = align_up((size_t) width * 4, 256), py = align_up(width, 256), pc = align_up(width, 256);
  bool ok = s->in.ensure(pi * height) && s->y.ensure(py * height) && s->c.ensure(pc * (height / 2));
  // the reference reads inPixelsPtr as tightly packed rows (CVPixelBufferUtils.h:295-299)
  ok = ok && upload2d(s->in.p, pi, inPixelsPtr, (size_t) width * 4, (size_t) width * 4, height, st);
  if (ok) {
    bt709_plane in = {s->in.p, pi, pi * height}, y = {s->y.p, py, py * height}, c = {s->c.p, pc, pc * (height / 2)};
    ok = bt709_encode_bgra_to_nv12(mrc->ctx, gi, go, &in, &y, &c, width, height, 1, st, 0) == BT709_OK;
    ok = ok && download2d(dst->y, dst->yBytesPerRow, s->y.p, py, width, height, st);
    ok = ok && download2d(dst->cbcr, dst->cbcrBytesPerRow, s->c.p, pc, width, height / 2, st);
  }
  return finish(s, st, ok, 1) ? 1 : 0;
}

MBDPixelBuffer *BGRAToBT709Converter_createYCbCrFromBGRA(MetalRenderContext *mrc, const uint32_t *bgra, int width,
                                                         int height, int isLinear, int asSRGBGamma) {
  if (isLinear && asSRGBGamma) return nullptr;  // BGRAToBT709Converter.m:879-881 assert
  if ((width % 2) != 0 || (height % 2) != 0) return nullptr;  // :540-541
  MBDPixelBuffer *pb = BGRAToBT709Converter_createCoreVideoYCbCrBuffer(width, height);
  if (!pb) return nullptr;
  BGRAToBT709Converter_setBT709Attributes(pb);
  BT709Gamma gi, go;  // :919-937
  if (isLinear) {
    gi = BT709GammaLinear; go = BT709GammaLinear;
    pb->transferFunction = kMBDTransferFunction_Linear;
  } else if (asSRGBGamma) {
    gi = BT709GammaSrgb; go = BT709GammaSrgb;
    pb->transferFunction = kMBDTransferFunction_sRGB;
  } else {
    gi = BT709GammaSrgb; go = BT709GammaApple;
    pb->transferFunction = kMBDTransferFunction_ITU_R_709_2;
  }
  if (!cvpbu_ycbcr_subsample(mrc, bgra, width, height, pb, gi, go)) {
    MBDPixelBufferRelease(pb);
    return nullptr;
  }
  return pb;
}

// ======================================================================== AAPLRenderer
AAPLRenderer *AAPLRenderer_new(MetalRenderContext *mrc, MetalBT709Gamma decodeGamma, int hasAlphaChannel) {
  if (!mrc) return nullptr;
  AAPLRenderer *r = new (std::nothrow) AAPLRenderer();
  if (!r) return nullptr;
  r->mrc = mrc;
  r->decoder = MetalBT709Decoder_new();
  r->scaler = MetalScaleRenderContext_new();
  MetalBT709Decoder_setMetalRenderContext(r->decoder, mrc);
  MetalBT709Decoder_setGamma(r->decoder, decodeGamma);
  MetalBT709Decoder_setHasAlphaChannel(r->decoder, hasAlphaChannel);
  if (!MetalBT709Decoder_setupMetal(r->decoder)) {
    AAPLRenderer_release(r);
    return nullptr;
  }
  return r;
}

void AAPLRenderer_release(AAPLRenderer *r) {
  if (!r) return;
  MetalBT709Decoder_release(r->decoder);
  MetalScaleRenderContext_release(r->scaler);
  delete r;
}

int AAPLRenderer_drawFrame(AAPLRenderer *r, const MBDPixelBuffer *pb, const MBDPixelBuffer *alpha, MBDTexture *drawable,
                           int waitUntilCompleted) {
  if (!r || !pb || !drawable || !drawable->pixels) return 0;
  MetalBT709Decoder *d = r->decoder;
  const int w = pb->width, h = pb->height;
  // AAPLRenderer.m:891-894
  const bool isExactlySameSize = (drawable->width == w) && (drawable->height == h);
  if (isExactlySameSize)  // :905-939 one pass, straight into the drawable
    return MetalBT709Decoder_decodeBT709(d, pb, alpha, drawable, nullptr, w, h, waitUntilCompleted);
  // :940-977 two passes through the same-size sRGB intermediate, which never leaves the device
  if (!decode_args_ok(d, pb, alpha, -1, -1, w, h, d->gamma)) return 0;
  if ((alpha != nullptr) != (d->hasAlphaChannel != 0)) return 0;
  MetalRenderContext *mrc = r->mrc;
  Guard g(mrc->device);
  cudaStream_t st;
  Slot *s = acquire(mrc, nullptr, &st);
  Staged in;
  bool ok = stage_nv12(s, pb, alpha, st, &in);
  const int ow = drawable->width, oh = drawable->height;
  const size_t pm = align_up((size_t) w * 4, 256), po = align_up((size_t) ow * 4, 256);
  ok = ok && s->mid.ensure(pm * h) && s->out.ensure(po * oh);
  if (ok) {
    bt709_plane mid = {s->mid.p, pm, pm * h}, out = {s->out.p, po, po * oh};
    ok = bt709_decode_rescale_nv12(mrc->ctx, to_abi_gamma(d->gamma), &in.y, &in.c, in.has_a ? &in.a : nullptr, &mid,
                                   &out, w, h, ow, oh, 1, st, 0) == BT709_OK;
    ok = ok && download2d(drawable->pixels, drawable->bytesPerRow, s->out.p, po, (size_t) ow * 4, oh, st);
  }
  return finish(s, st, ok, waitUntilCompleted) ? 1 : 0;
}

}  // extern "C"
