// bt709_host.h — internals of the reference-facing host layer (include/metal_bt709_decoder.h).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <mutex>

#include "../../include/metal_bt709_decoder.h"

namespace bt709host {

// grow-only device buffer (the reference keeps a CVMetalTextureCache for the same purpose,
// MetalBT709Decoder.m:15,95): nothing is allocated per frame once sizes are stable
struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  bool ensure(size_t bytes);
  void release();
};

constexpr int kSlots = 4;  // frames in flight: upload of frame i+1 overlaps read-back of frame i

struct Slot {
  cudaStream_t stream = nullptr;
  cudaEvent_t done = nullptr;
  std::mutex busy;  // held from acquire() to finish(): one thread at a time enqueues on a slot's buffers
  DevBuf y, c, a, out, mid, in;
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace bt709host

struct MetalRenderContext {
  int device = 0;
  bt709_ctx *ctx = nullptr;
  bt709host::Slot slots[bt709host::kSlots];
  int next = 0;
  std::mutex mu;
};

struct MetalBT709Decoder {
  MetalRenderContext *mrc = nullptr;
  bool owns_mrc = false;
  MetalBT709Gamma gamma = MetalBT709GammaApple;
  int useComputeRenderer = 0;
  int hasAlphaChannel = 0;
  bool setup = false;
};

struct MetalScaleRenderContext {
  int unused = 0;
};

struct AAPLRenderer {
  MetalRenderContext *mrc = nullptr;
  MetalBT709Decoder *decoder = nullptr;
  MetalScaleRenderContext *scaler = nullptr;
};
