// bt709_api.cu — the C ABI of include/bt709_b200.h: context, argument validation, dispatch.
//
// Mirrors the guard clauses of -[MetalBT709Decoder processBT709ToSRGB:] (MetalBT709Decoder.m:
// 252-492), +[BGRAToBT709Converter convert:] (BGRAToBT709Converter.m:33-57) and
// cvpbu_ycbcr_subsample (CVPixelBufferUtils.h:275-276) as returned status codes.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/bt709_b200.h"
#include "../../include/y4m_b200.h"
#include "bt709_kernels.cuh"
#include "bt709_tables.h"

using namespace bt709;

thread_local unsigned long long bt709::tl_kernel_launches = 0;

struct bt709_ctx {
  int device = -1;
  int sm_count = 148;
  mutable std::atomic<int> last_cuda_error{0};
  mutable std::atomic<unsigned long long> launches{0};
  void *dev_blob = nullptr;  // one allocation holding every table below
  // decode: matrix output -> sRGB byte, per bt709_gamma
  const uint32_t *dec_lut[5] = {};  // [4] = kBt709LinearRgb, packed 4:4:4 entries only
  float dec_clamp[5] = {};
  int dec_cells[5] = {};
  // log-cell form of the Linear-mode step function (conflict-free when replicated per lane)
  const uint32_t *dec_log_lut = nullptr;
  float dec_log_clamp = 0.f;
  int dec_log_cells = 0;
  unsigned dec_log_first = 0;
  // biased log-cell form of the same step function (645 cells: a replica per lane fits next to other tables)
  const uint32_t *dec_blog_lut = nullptr;
  float dec_blog_clamp = 0.f, dec_blog_bias = 0.f;
  int dec_blog_cells = 0;
  unsigned dec_blog_first = 0;
  // BT709_from_linear: linear float -> byte in gamma g (g = APPLE, SRGB, LINEAR)
  const uint32_t *fl_lut[3] = {};
  float fl_clamp[3] = {};
  int fl_cells[3] = {};
  // biased log-cell form of the same step functions (optional: the 4:2:0 encode's chroma lookup, one replica per lane)
  const uint32_t *fl_blog_lut[3] = {};
  float fl_blog_clamp[3] = {}, fl_blog_bias[3] = {};
  int fl_blog_cells[3] = {};
  unsigned fl_blog_first[3] = {};
  const float *lin256[3] = {};       // BT709_tolinearNorm per input gamma
  const float *compn256[3][3] = {};  // [gin][gout] unorm8(from_linear(to_linear(b)))
  // per-pixel 4:4:4 encoders: sRGB byte -> non-linear value fed to the matrix, per bt709_gamma
  // (BT709.h:610-664 BT.709 curve, :743-817 Apple curve, :914-944 sRGB = byteNorm; Linear has no such encoder)
  const float *enc444[5] = {};
  MatrixConsts m;
};

namespace {

struct BlobBuilder {
  std::vector<uint32_t> words;
  size_t add(const void *src, size_t nwords) {
    size_t off = (words.size() + 3) & ~size_t(3);  // 16-byte aligned sub-tables
    words.resize(off + nwords);
    std::memcpy(&words[off], src, nwords * 4);
    return off;
  }
};

int cuda_fail(const bt709_ctx *ctx, cudaError_t e) {
  if (ctx) ctx->last_cuda_error.store((int) e);
  return BT709_E_CUDA;
}

bool plane_ok(const bt709_plane *p) { return p && p->ptr; }

// pitch >= row bytes; frame_stride >= rows * pitch when more than one frame
int check_plane(const bt709_plane *p, size_t row_bytes, size_t rows, int nframes) {
  if (p->pitch < row_bytes) return BT709_E_PITCH;
  if (nframes > 1 && p->frame_stride < (rows - 1) * p->pitch + row_bytes) return BT709_E_PITCH;
  return BT709_OK;
}

// every launcher call is bracketed by launches_begin() ... finish(): the context counts the kernels really launched
// (a batch may be split across several launches), not the API calls
unsigned long long launches_begin() { return tl_kernel_launches; }

int finish(const bt709_ctx *ctx, unsigned long long l0, int launch_err, void *stream, int wait) {
  ctx->launches.fetch_add(tl_kernel_launches - l0);
  if (launch_err) return cuda_fail(ctx, (cudaError_t) launch_err);
  if (wait) {
    cudaError_t e = cudaStreamSynchronize((cudaStream_t) stream);
    if (e != cudaSuccess) return cuda_fail(ctx, e);
  }
  return BT709_OK;
}

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    if (prev == dev) prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

}  // namespace

extern "C" {

int bt709_ctx_create(int device, bt709_ctx **out) {
  if (!out) return BT709_E_NULL;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
    // no CPU fallback: without a CUDA device the library refuses to work
    cudaGetLastError();
    return BT709_E_CUDA;
  }
  bt709_ctx *ctx = new (std::nothrow) bt709_ctx();
  if (!ctx) return BT709_E_NULL;
  ctx->device = device;
  ctx->m = matrix_consts();
  DeviceGuard guard(device);
  if (!guard.ok) { delete ctx; return BT709_E_CUDA; }
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) { delete ctx; return BT709_E_CUDA; }
  ctx->sm_count = prop.multiProcessorCount;

  // ---- build every table on the host, once
  BlobBuilder blob;
  size_t off_dec[5], off_fl[3], off_lin[3], off_comp[3][3], off_e444[5] = {}, off_log = 0, off_blog = 0, off_flb[3] = {};
  for (int g = 0; g < 5; g++) {
    StepFn fn;
    CellLut lut;
    if (!build_step_fn(decode_code, g, &fn)) { delete ctx; return BT709_E_GAMMA; }
    // fixed sizes so that the fast kernel has two instantiations: 2^10 cells x 32 replicas for
    // the Apple/sRGB/BT.709 curves, 2^12 cells x 8 replicas for the steep Linear (pure sRGB
    // encode) curve; both take 128 KB of shared memory
    const int n = (g == kLinear) ? 12 : 10;
    if (!build_cell_lut(fn, n, &lut)) { delete ctx; return BT709_E_GAMMA; }
    ctx->dec_clamp[g] = lut.clamp_lo;
    ctx->dec_cells[g] = n;
    off_dec[g] = blob.add(lut.entries.data(), lut.entries.size());
    LogLut llut;
    if (g == kLinear && build_log_lut(fn, &llut)) {  // optional: the uniform table above stays the fallback
      ctx->dec_log_clamp = llut.clamp_lo;
      ctx->dec_log_cells = (int) llut.entries.size();
      ctx->dec_log_first = llut.first;
      off_log = blob.add(llut.entries.data(), llut.entries.size());
    }
    BiasedLogLut blut;
    if (g == kLinear && build_biased_log_lut(fn, 1.0f / 32.0f, &blut)) {  // optional as well
      ctx->dec_blog_clamp = blut.clamp_lo;
      ctx->dec_blog_bias = blut.bias;
      ctx->dec_blog_cells = (int) blut.entries.size();
      ctx->dec_blog_first = blut.first;
      off_blog = blob.add(blut.entries.data(), blut.entries.size());
    }
  }
  for (int g = 0; g < 3; g++) {
    StepFn fn;
    CellLut lut;
    if (!build_step_fn(from_linear_code, g, &fn)) { delete ctx; return BT709_E_GAMMA; }
    int n = build_cell_lut_auto(fn, 8, 14, &lut);
    if (n < 0) { delete ctx; return BT709_E_GAMMA; }
    ctx->fl_clamp[g] = lut.clamp_lo;
    ctx->fl_cells[g] = n;
    off_fl[g] = blob.add(lut.entries.data(), lut.entries.size());
    BiasedLogLut fb;
    if (n > 8 && build_biased_log_lut(fn, 1.0f / 32.0f, &fb)) {  // steep curves only (Apple, sRGB): 645 cells instead of 4096
      ctx->fl_blog_clamp[g] = fb.clamp_lo;
      ctx->fl_blog_bias[g] = fb.bias;
      ctx->fl_blog_cells[g] = (int) fb.entries.size();
      ctx->fl_blog_first[g] = fb.first;
      off_flb[g] = blob.add(fb.entries.data(), fb.entries.size());
    }
    float lin[256];
    for (int b = 0; b < 256; b++) lin[b] = to_linear(g, b);
    off_lin[g] = blob.add(lin, 256);
  }
  for (int gi = 0; gi < 3; gi++)
    for (int go = 0; go < 3; go++) {
      float t[256];
      for (int b = 0; b < 256; b++) t[b] = from_linear_code(go, to_linear(gi, b)) * (1.0f / 255.0f);
      off_comp[gi][go] = blob.add(t, 256);
    }
  for (int g = 0; g < 5; g++) {
    if (g == kLinear) continue;
    float t[256];  // front half of the per-pixel encoders: byteNorm, sRGB -> linear, linear -> gamma g
    for (int b = 0; b < 256; b++) {
      const float n = b * (1.0f / 255.0f);
      t[b] = g == kSrgb ? n : g == kApple ? apple_oetf(srgb_eotf(n)) : g == kBt709 ? bt709_oetf(srgb_eotf(n)) : bt709_oetf(n);
    }
    off_e444[g] = blob.add(t, 256);
  }

  if ((e = cudaMalloc(&ctx->dev_blob, blob.words.size() * 4)) != cudaSuccess) { delete ctx; return BT709_E_CUDA; }
  if ((e = cudaMemcpy(ctx->dev_blob, blob.words.data(), blob.words.size() * 4, cudaMemcpyHostToDevice)) != cudaSuccess) {
    cudaFree(ctx->dev_blob);
    delete ctx;
    return BT709_E_CUDA;
  }
  const uint32_t *base = (const uint32_t *) ctx->dev_blob;
  for (int g = 0; g < 5; g++) ctx->dec_lut[g] = base + off_dec[g];
  if (ctx->dec_log_cells) ctx->dec_log_lut = base + off_log;
  if (ctx->dec_blog_cells) ctx->dec_blog_lut = base + off_blog;
  for (int g = 0; g < 3; g++) {
    ctx->fl_lut[g] = base + off_fl[g];
    if (ctx->fl_blog_cells[g] > 0) ctx->fl_blog_lut[g] = base + off_flb[g];
    ctx->lin256[g] = (const float *) (base + off_lin[g]);
    for (int go = 0; go < 3; go++) ctx->compn256[g][go] = (const float *) (base + off_comp[g][go]);
  }
  for (int g = 0; g < 5; g++)
    if (g != kLinear) ctx->enc444[g] = (const float *) (base + off_e444[g]);
  int ke = kernels_init();
  if (ke) {
    cudaFree(ctx->dev_blob);
    delete ctx;
    return BT709_E_CUDA;
  }
  *out = ctx;
  return BT709_OK;
}

void bt709_ctx_destroy(bt709_ctx *ctx) {
  if (!ctx) return;
  {
    DeviceGuard guard(ctx->device);
    if (ctx->dev_blob) cudaFree(ctx->dev_blob);
  }
  delete ctx;
}

int bt709_selftest_host(void) { return selftest_host(); }

int bt709_ctx_device(const bt709_ctx *ctx) { return ctx ? ctx->device : -1; }
int bt709_last_cuda_error(const bt709_ctx *ctx) { return ctx ? ctx->last_cuda_error.load() : 0; }
unsigned long long bt709_launch_count(const bt709_ctx *ctx) { return ctx ? ctx->launches.load() : 0; }

const char *bt709_status_string(int s) {
  switch (s) {
    case BT709_OK: return "ok";
    case BT709_E_NULL: return "null context, plane or pointer";
    case BT709_E_DIMS: return "bad dimensions (non-positive, or odd on the 4:2:0 encode)";
    case BT709_E_PITCH: return "pitch or frame stride too small, or BGRA plane not 4-byte aligned";
    case BT709_E_GAMMA: return "unsupported gamma for this entry point";
    case BT709_E_ALPHA_MODE: return "alpha stream requires BT709_GAMMA_SRGB";
    case BT709_E_CUDA: return "CUDA error (see bt709_last_cuda_error)";
    case BT709_E_SCRATCH: return "two-pass decode+rescale needs an intermediate plane";
    case BT709_E_UNSUPPORTED: return "the fused decode+rescale kernel does not apply to this geometry";
    default: return "unknown status";
  }
}

static void fill_decode_args(const bt709_ctx *ctx, bt709_gamma gamma, const bt709_plane *y, const bt709_plane *cbcr,
                             const bt709_plane *alpha_y, const bt709_plane *bgra_out, int w, int h, int nframes, DecodeArgs *pa);
static void fill_rescale_args(const bt709_ctx *ctx, const bt709_plane *in, int in_w, int in_h, const bt709_plane *out, int out_w,
                              int out_h, int nframes, RescaleArgs *pa);

int bt709_decode_nv12(const bt709_ctx *ctx, bt709_gamma gamma, const bt709_plane *y, const bt709_plane *cbcr,
                      const bt709_plane *alpha_y, const bt709_plane *bgra_out, int w, int h, int nframes,
                      void *stream, int wait) {
  if (!ctx || !plane_ok(y) || !plane_ok(cbcr) || !plane_ok(bgra_out)) return BT709_E_NULL;
  if (alpha_y && !alpha_y->ptr) return BT709_E_NULL;
  if (w <= 0 || h <= 0 || nframes <= 0) return BT709_E_DIMS;
  if ((int) gamma < 0 || (int) gamma > 3) return BT709_E_GAMMA;
  if (alpha_y && gamma != BT709_GAMMA_SRGB) return BT709_E_ALPHA_MODE;
  const size_t cw2 = (size_t) ((w + 1) / 2) * 2, ch = (size_t) (h + 1) / 2;
  int rc;
  if ((rc = check_plane(y, w, h, nframes))) return rc;
  if ((rc = check_plane(cbcr, cw2, ch, nframes))) return rc;
  if ((rc = check_plane(bgra_out, (size_t) w * 4, h, nframes))) return rc;
  if (alpha_y && (rc = check_plane(alpha_y, w, h, nframes))) return rc;
  DeviceGuard guard(ctx->device);
  if (!guard.ok) return cuda_fail(ctx, cudaErrorInvalidDevice);
  DecodeArgs a{};
  fill_decode_args(ctx, gamma, y, cbcr, alpha_y, bgra_out, w, h, nframes, &a);
  const unsigned long long l0 = launches_begin();
  return finish(ctx, l0, launch_decode(a, ctx->sm_count, stream), stream, wait);
}

static int rescale_impl(const bt709_ctx *ctx, const bt709_plane *in, int in_w, int in_h, const bt709_plane *out,
                        int out_w, int out_h, int nframes, void *stream, int wait, int known_opaque);

int bt709_rescale_bgra(const bt709_ctx *ctx, const bt709_plane *in, int in_w, int in_h, const bt709_plane *out,
                       int out_w, int out_h, int nframes, void *stream, int wait) {
  return rescale_impl(ctx, in, in_w, in_h, out, out_w, out_h, nframes, stream, wait, 0);
}

static int rescale_impl(const bt709_ctx *ctx, const bt709_plane *in, int in_w, int in_h, const bt709_plane *out,
                        int out_w, int out_h, int nframes, void *stream, int wait, int known_opaque) {
  if (!ctx || !plane_ok(in) || !plane_ok(out)) return BT709_E_NULL;
  if (in_w <= 0 || in_h <= 0 || out_w <= 0 || out_h <= 0 || nframes <= 0) return BT709_E_DIMS;
  int rc;
  if ((rc = check_plane(in, (size_t) in_w * 4, in_h, nframes))) return rc;
  if ((rc = check_plane(out, (size_t) out_w * 4, out_h, nframes))) return rc;
  // BGRA pixels are read and written as 32-bit words
  if (((uintptr_t) in->ptr | in->pitch | in->frame_stride | (uintptr_t) out->ptr | out->pitch | out->frame_stride) & 3)
    return BT709_E_PITCH;
  DeviceGuard guard(ctx->device);
  if (!guard.ok) return cuda_fail(ctx, cudaErrorInvalidDevice);
  RescaleArgs a{};
  fill_rescale_args(ctx, in, in_w, in_h, out, out_w, out_h, nframes, &a);
  a.known_opaque = known_opaque;
  const unsigned long long l0 = launches_begin();
  return finish(ctx, l0, launch_rescale(a, ctx->sm_count, stream), stream, wait);
}

// fills the kernel argument structs shared by the one-pass, two-pass and fused entries
static void fill_decode_args(const bt709_ctx *ctx, bt709_gamma gamma, const bt709_plane *y, const bt709_plane *cbcr,
                             const bt709_plane *alpha_y, const bt709_plane *bgra_out, int w, int h, int nframes, DecodeArgs *pa) {
  DecodeArgs &a = *pa;
  a.y = (const uint8_t *) y->ptr; a.py = (long long) y->pitch; a.sy = (long long) y->frame_stride;
  a.c = (const uint8_t *) cbcr->ptr; a.pc = (long long) cbcr->pitch; a.sc = (long long) cbcr->frame_stride;
  if (bgra_out) { a.out = (uint8_t *) bgra_out->ptr; a.po = (long long) bgra_out->pitch; a.so = (long long) bgra_out->frame_stride; }
  if (alpha_y) { a.a = (const uint8_t *) alpha_y->ptr; a.pa = (long long) alpha_y->pitch; a.sa = (long long) alpha_y->frame_stride; }
  a.w = w; a.h = h; a.nframes = nframes;
  a.lut = ctx->dec_lut[gamma]; a.clamp_lo = ctx->dec_clamp[gamma]; a.cells_log2 = ctx->dec_cells[gamma];
  if ((int) gamma == (int) kLinear && ctx->dec_log_lut) {
    a.log_lut = ctx->dec_log_lut; a.log_clamp_lo = ctx->dec_log_clamp;
    a.log_cells = ctx->dec_log_cells; a.log_first = ctx->dec_log_first;
  }
  if ((int) gamma == (int) kLinear && ctx->dec_blog_lut) {
    a.blog_lut = ctx->dec_blog_lut; a.blog_clamp_lo = ctx->dec_blog_clamp; a.blog_bias = ctx->dec_blog_bias;
    a.blog_cells = ctx->dec_blog_cells; a.blog_first = ctx->dec_blog_first;
  }
  a.m = ctx->m;
  a.one = 1.0f;
}

static void fill_rescale_args(const bt709_ctx *ctx, const bt709_plane *in, int in_w, int in_h, const bt709_plane *out, int out_w,
                              int out_h, int nframes, RescaleArgs *pa) {
  RescaleArgs &a = *pa;
  if (in) { a.in = (const uint8_t *) in->ptr; a.pi = (long long) in->pitch; a.si = (long long) in->frame_stride; }
  a.out = (uint8_t *) out->ptr; a.po = (long long) out->pitch; a.so = (long long) out->frame_stride;
  a.iw = in_w; a.ih = in_h; a.ow = out_w; a.oh = out_h; a.nframes = nframes;
  a.scale_x = (float) in_w / (float) out_w;
  a.scale_y = (float) in_h / (float) out_h;
  a.lin256 = ctx->lin256[kSrgb];
  a.enc_lut = ctx->fl_lut[kSrgb]; a.enc_clamp_lo = ctx->fl_clamp[kSrgb]; a.enc_cells_log2 = ctx->fl_cells[kSrgb];
  // the decoder's Linear-mode step function IS the sRGB encode (decode_code(kLinear) == from_linear_code(kSrgb))
  a.enc_log_lut = ctx->dec_log_lut; a.enc_log_clamp_lo = ctx->dec_log_clamp;
  a.enc_log_cells = ctx->dec_log_cells; a.enc_log_first = ctx->dec_log_first;
  a.enc_blog_lut = ctx->dec_blog_lut; a.enc_blog_clamp_lo = ctx->dec_blog_clamp; a.enc_blog_bias = ctx->dec_blog_bias;
  a.enc_blog_cells = ctx->dec_blog_cells; a.enc_blog_first = ctx->dec_blog_first;
  a.one = 1.0f;
}

// 1 = the fused kernel ran (status in *rc), 0 = it does not apply to these arguments
static int try_fused(const bt709_ctx *ctx, bt709_gamma gamma, const bt709_plane *y, const bt709_plane *cbcr,
                     const bt709_plane *alpha_y, const bt709_plane *scratch, const bt709_plane *bgra_out, int w, int h,
                     int out_w, int out_h, int nframes, void *stream, int wait, int *rc) {
  // `scratch` is used as a raw work area, so only when its rows and frames are tightly packed
  if (!(ctx && plane_ok(y) && plane_ok(cbcr) && plane_ok(bgra_out) && scratch && scratch->ptr && !alpha_y && w > 0 && h > 0 &&
        out_w > 0 && out_h > 0 && nframes > 0 && (int) gamma >= 0 && (int) gamma <= 3 && scratch->pitch == (size_t) w * 4))
    return 0;
  // frames back to back: the whole batch-sized area; any other frame stride: only the first frame's area is assumed
  const size_t frame_bytes = scratch->pitch * (size_t) h;
  const size_t scratch_bytes = (nframes > 1 && scratch->frame_stride == frame_bytes) ? frame_bytes * (size_t) nframes : frame_bytes;
  const size_t cw2 = (size_t) ((w + 1) / 2) * 2, ch = (size_t) (h + 1) / 2;
  if ((*rc = check_plane(y, w, h, nframes))) return 1;
  if ((*rc = check_plane(cbcr, cw2, ch, nframes))) return 1;
  if ((*rc = check_plane(bgra_out, (size_t) out_w * 4, out_h, nframes))) return 1;
  DeviceGuard guard(ctx->device);
  if (!guard.ok) { *rc = cuda_fail(ctx, cudaErrorInvalidDevice); return 1; }
  DecodeArgs d{};
  RescaleArgs r{};
  fill_decode_args(ctx, gamma, y, cbcr, nullptr, nullptr, w, h, nframes, &d);
  fill_rescale_args(ctx, nullptr, w, h, bgra_out, out_w, out_h, nframes, &r);
  const unsigned long long l0 = launches_begin();
  const int e = launch_fused(d, r, (uint8_t *) scratch->ptr, scratch_bytes, ctx->sm_count, stream);
  if (e == -1) return 0;
  *rc = finish(ctx, l0, e, stream, wait);
  return 1;
}

int bt709_decode_rescale_nv12_fused(const bt709_ctx *ctx, bt709_gamma gamma, const bt709_plane *y, const bt709_plane *cbcr,
                                    const bt709_plane *scratch, const bt709_plane *bgra_out, int w, int h, int out_w,
                                    int out_h, int nframes, void *stream, int wait) {
  if (!ctx) return BT709_E_NULL;
  if (!scratch || !scratch->ptr) return BT709_E_SCRATCH;
  int rc = BT709_OK;
  if (try_fused(ctx, gamma, y, cbcr, nullptr, scratch, bgra_out, w, h, out_w, out_h, nframes, stream, wait, &rc)) return rc;
  return BT709_E_UNSUPPORTED;
}

int bt709_decode_rescale_nv12(const bt709_ctx *ctx, bt709_gamma gamma, const bt709_plane *y, const bt709_plane *cbcr,
                              const bt709_plane *alpha_y, const bt709_plane *scratch, const bt709_plane *bgra_out,
                              int w, int h, int out_w, int out_h, int nframes, void *stream, int wait) {
  // AAPLRenderer.m:891-894: exact size -> one pass straight into the destination
  if (out_w == w && out_h == h)
    return bt709_decode_nv12(ctx, gamma, y, cbcr, alpha_y, bgra_out, w, h, nframes, stream, wait);
  if (!ctx) return BT709_E_NULL;
  if (!scratch || !scratch->ptr) return BT709_E_SCRATCH;
  // The fused kernel (one launch, the intermediate stays in L2: 39 % less HBM traffic) is bit-identical but measured
  // 12 % slower than the two launches on 1080p -> 1440p, because both halves are instruction-bound, not HBM-bound
  // (DESIGN.md section 4.4): it runs when asked for (the _fused entry, or BT709_FUSED=1 in the environment).
  static const bool prefer_fused = getenv("BT709_FUSED") != nullptr;
  if (prefer_fused) {
    int rc = BT709_OK;
    if (try_fused(ctx, gamma, y, cbcr, alpha_y, scratch, bgra_out, w, h, out_w, out_h, nframes, stream, wait, &rc)) return rc;
  }
  // AAPLRenderer.m:940-977: decode into the same-size sRGB intermediate, then rescale
  int rc = bt709_decode_nv12(ctx, gamma, y, cbcr, alpha_y, scratch, w, h, nframes, stream, 0);
  if (rc) return rc;
  // without an alpha stream pass 1 has just written A = 255 everywhere (AAPLShaders.metal:243): pass 2 need not watch for alpha
  return rescale_impl(ctx, scratch, w, h, bgra_out, out_w, out_h, nframes, stream, wait, alpha_y == nullptr);
}

int bt709_encode_bgra_to_nv12(const bt709_ctx *ctx, bt709_gamma gin, bt709_gamma gout, const bt709_plane *in,
                              const bt709_plane *y, const bt709_plane *cbcr, int w, int h, int nframes, void *stream,
                              int wait) {
  if (!ctx || !plane_ok(in) || !plane_ok(y) || !plane_ok(cbcr)) return BT709_E_NULL;
  if (w <= 0 || h <= 0 || nframes <= 0 || (w & 1) || (h & 1)) return BT709_E_DIMS;
  if ((int) gin < 0 || (int) gin > 2 || (int) gout < 0 || (int) gout > 2) return BT709_E_GAMMA;
  int rc;
  if ((rc = check_plane(in, (size_t) w * 4, h, nframes))) return rc;
  if ((rc = check_plane(y, w, h, nframes))) return rc;
  if ((rc = check_plane(cbcr, w, h / 2, nframes))) return rc;
  if (((uintptr_t) in->ptr | in->pitch | in->frame_stride) & 3) return BT709_E_PITCH;
  DeviceGuard guard(ctx->device);
  if (!guard.ok) return cuda_fail(ctx, cudaErrorInvalidDevice);
  EncodeArgs a{};
  a.in = (const uint8_t *) in->ptr; a.pi = (long long) in->pitch; a.si = (long long) in->frame_stride;
  a.y = (uint8_t *) y->ptr; a.py = (long long) y->pitch; a.sy = (long long) y->frame_stride;
  a.c = (uint8_t *) cbcr->ptr; a.pc = (long long) cbcr->pitch; a.sc = (long long) cbcr->frame_stride;
  a.w = w; a.h = h; a.nframes = nframes;
  a.lin256 = ctx->lin256[gin];
  a.compn256 = ctx->compn256[gin][gout];
  a.out_lut = ctx->fl_lut[gout]; a.out_clamp_lo = ctx->fl_clamp[gout]; a.out_cells_log2 = ctx->fl_cells[gout];
  a.out_blog_lut = ctx->fl_blog_lut[gout]; a.out_blog_clamp_lo = ctx->fl_blog_clamp[gout]; a.out_blog_bias = ctx->fl_blog_bias[gout];
  a.out_blog_cells = ctx->fl_blog_cells[gout]; a.out_blog_first = ctx->fl_blog_first[gout];
  const unsigned long long l0 = launches_begin();
  return finish(ctx, l0, launch_encode(a, ctx->sm_count, stream), stream, wait);
}

int bt709_convert_packed444_gamma(const bt709_ctx *ctx, bt709_gamma gamma, const uint32_t *in, uint32_t *out, int w, int h,
                                  void *stream, int wait) {
  if (!ctx || !in || !out) return BT709_E_NULL;
  if (w <= 0 || h <= 0 || (w & 1) || (h & 1)) return BT709_E_DIMS;  // BGRAToBT709Converter.m:41-46
  if ((int) gamma < 0 || (int) gamma > 4 || !ctx->enc444[gamma]) return BT709_E_GAMMA;
  DeviceGuard guard(ctx->device);
  if (!guard.ok) return cuda_fail(ctx, cudaErrorInvalidDevice);
  Packed444Args a{};
  a.in = in; a.out = out; a.n = (long long) w * h; a.enc256 = ctx->enc444[gamma]; a.m = ctx->m;
  const unsigned long long l0 = launches_begin();
  return finish(ctx, l0, launch_convert444(a, stream), stream, wait);
}

int bt709_unconvert_packed444_gamma(const bt709_ctx *ctx, bt709_gamma gamma, const uint32_t *in, uint32_t *out, int w,
                                    int h, void *stream, int wait) {
  if (!ctx || !in || !out) return BT709_E_NULL;
  if (w <= 0 || h <= 0 || (w & 1) || (h & 1)) return BT709_E_DIMS;  // BGRAToBT709Converter.m:69-74
  if ((int) gamma < 0 || (int) gamma > 4 || (int) gamma == (int) kLinear) return BT709_E_GAMMA;
  DeviceGuard guard(ctx->device);
  if (!guard.ok) return cuda_fail(ctx, cudaErrorInvalidDevice);
  Packed444Args a{};
  a.in = in; a.out = out; a.n = (long long) w * h; a.m = ctx->m;
  a.lut = ctx->dec_lut[gamma]; a.clamp_lo = ctx->dec_clamp[gamma]; a.cells_log2 = ctx->dec_cells[gamma];
  const unsigned long long l0 = launches_begin();
  return finish(ctx, l0, launch_unconvert444(a, stream), stream, wait);
}

int bt709_convert_packed444(const bt709_ctx *ctx, const uint32_t *in, uint32_t *out, int w, int h, void *stream,
                            int wait) {
  return bt709_convert_packed444_gamma(ctx, BT709_GAMMA_APPLE, in, out, w, h, stream, wait);
}

int bt709_unconvert_packed444(const bt709_ctx *ctx, const uint32_t *in, uint32_t *out, int w, int h, void *stream,
                              int wait) {
  return bt709_unconvert_packed444_gamma(ctx, BT709_GAMMA_APPLE, in, out, w, h, stream, wait);
}

int bt709_copy_packed444_to_nv12(const bt709_ctx *ctx, const uint32_t *in, const bt709_plane *y,
                                 const bt709_plane *cbcr, int w, int h, void *stream, int wait) {
  if (!ctx || !in || !plane_ok(y) || !plane_ok(cbcr)) return BT709_E_NULL;
  if (w <= 0 || h <= 0) return BT709_E_DIMS;
  int rc;
  if ((rc = check_plane(y, w, h, 1))) return rc;
  if ((rc = check_plane(cbcr, (size_t) ((w + 1) / 2) * 2, (h + 1) / 2, 1))) return rc;
  DeviceGuard guard(ctx->device);
  if (!guard.ok) return cuda_fail(ctx, cudaErrorInvalidDevice);
  CopyPackedArgs a{};
  a.in = in; a.y = (uint8_t *) y->ptr; a.py = (long long) y->pitch; a.c = (uint8_t *) cbcr->ptr;
  a.pc = (long long) cbcr->pitch; a.w = w; a.h = h;
  const unsigned long long l0 = launches_begin();
  return finish(ctx, l0, launch_copy_packed(a, stream), stream, wait);
}

static int cbcr_planar(const bt709_ctx *ctx, const bt709_plane *cbcr, const bt709_plane *u, const bt709_plane *v, int cw,
                       int ch, int nframes, void *stream, int wait, bool split) {
  if (!ctx || !plane_ok(cbcr) || !plane_ok(u) || !plane_ok(v)) return BT709_E_NULL;
  if (cw <= 0 || ch <= 0 || nframes <= 0) return BT709_E_DIMS;
  int rc;
  if ((rc = check_plane(cbcr, (size_t) cw * 2, ch, nframes))) return rc;
  if ((rc = check_plane(u, cw, ch, nframes))) return rc;
  if ((rc = check_plane(v, cw, ch, nframes))) return rc;
  DeviceGuard guard(ctx->device);
  if (!guard.ok) return cuda_fail(ctx, cudaErrorInvalidDevice);
  PlanarArgs a{};
  a.c = (uint8_t *) cbcr->ptr; a.pc = (long long) cbcr->pitch; a.sc = (long long) cbcr->frame_stride;
  a.u = (uint8_t *) u->ptr; a.pu = (long long) u->pitch; a.su = (long long) u->frame_stride;
  a.v = (uint8_t *) v->ptr; a.pv = (long long) v->pitch; a.sv = (long long) v->frame_stride;
  a.cw = cw; a.ch = ch; a.nframes = nframes;
  const unsigned long long l0 = launches_begin();
  return finish(ctx, l0, launch_cbcr_planar(a, split, stream), stream, wait);
}

int bt709_cbcr_to_planar(const bt709_ctx *ctx, const bt709_plane *cbcr, const bt709_plane *u, const bt709_plane *v,
                         int cw, int ch, int nframes, void *stream, int wait) {
  return cbcr_planar(ctx, cbcr, u, v, cw, ch, nframes, stream, wait, true);
}

int bt709_planar_to_cbcr(const bt709_ctx *ctx, const bt709_plane *u, const bt709_plane *v, const bt709_plane *cbcr,
                         int cw, int ch, int nframes, void *stream, int wait) {
  return cbcr_planar(ctx, cbcr, u, v, cw, ch, nframes, stream, wait, false);
}

}  // extern "C"
