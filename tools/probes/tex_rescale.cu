// tex_rescale.cu — "a hardware-semantics pass 2": the B200 texture unit doing what the Metal sampler does in
// samplingShader (AAPLShaders.metal:73-85): an sRGB-format BGRA8 texture read with a linear filter, clamp-to-edge and
// normalised coordinates, i.e. sRGB -> linear conversion and the bilinear blend in fixed-function hardware.  The probe
// compares it with the exact float32 walker of the library (bt709_rescale_bgra, bit-identical to
// oracle_rescale_bgra) on the same frames: fraction of identical bytes, |delta| <= 1 and <= 2 fractions, max |delta|,
// and the kernel times of both.  The filter weights of the texture unit have 8 fractional bits (CUDA programming
// guide, linear filtering: 9-bit fixed point), so on high-contrast content near black the re-encoded result can
// differ by many codes; the probe quantifies that on white noise and on a smooth gradient.
//
//   build: make -C tools/probes tex_rescale      run: tools/probes/tex_rescale [repetitions]
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../include/bt709_b200.h"

#define CK(x)                                                          \
  do {                                                                 \
    cudaError_t e_ = (x);                                              \
    if (e_ != cudaSuccess) {                                           \
      fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));         \
      exit(2);                                                         \
    }                                                                  \
  } while (0)

// linear float -> sRGB code with the reference's formula (sRGB.h:62-74, double pow) and round half away from zero
__device__ __forceinline__ uint32_t srgb_code(float v) {
  v = fminf(fmaxf(v, 0.0f), 1.0f);
  float nl = v <= 0.0031308f ? v * 12.92f
                             : (float) ((double) 1.055f * pow((double) v, (double) (1.0f / 2.4f)) - (double) 0.055f);
  return (uint32_t) (int) round((double) (nl * 255.0f));
}

// one thread per output pixel; tex2D returns linear R,G,B (sRGB conversion in the texture unit) and linear A
__global__ void k_tex_rescale(cudaTextureObject_t tex, uint8_t *out, size_t po, int ow, int oh) {
  const int ox = blockIdx.x * blockDim.x + threadIdx.x, oy = blockIdx.y * blockDim.y + threadIdx.y;
  if (ox >= ow || oy >= oh) return;
  // normalised coordinates of the pixel centre, as the rasteriser hands them to the fragment shader
  const float u = ((float) ox + 0.5f) / (float) ow, v = ((float) oy + 0.5f) / (float) oh;
  const float4 t = tex2D<float4>(tex, u, v);  // BGRA texture read as uchar4: x = B, y = G, z = R, w = A
  const uint32_t B = srgb_code(t.x), G = srgb_code(t.y), R = srgb_code(t.z);
  const uint32_t A = (uint32_t) (int) round((double) (fminf(fmaxf(t.w, 0.f), 1.f) * 255.0f));
  *reinterpret_cast<uint32_t *>(out + (size_t) oy * po + (size_t) ox * 4) = B | (G << 8) | (R << 16) | (A << 24);
}

static void fill(std::vector<uint8_t> &px, int w, int h, int mode, unsigned seed) {
  px.resize((size_t) w * h * 4);
  unsigned s = seed;
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++)
      for (int c = 0; c < 4; c++) {
        s = s * 1664525u + 1013904223u;
        uint8_t v;
        if (mode == 0)
          v = (uint8_t) (s >> 24);  // white noise
        else
          v = (uint8_t) (((x * (c + 1)) / 8 + (y * (4 - c)) / 6 + ((s >> 30) & 3)) & 255);  // gradient + 2 LSB of noise
        px[((size_t) y * w + x) * 4 + c] = c == 3 ? 255 : v;
      }
}

int main(int argc, char **argv) {
  const int iw = 1920, ih = 1080, ow = 2560, oh = 1440;
  const int reps = argc > 1 ? atoi(argv[1]) : 20;
  bt709_ctx *ctx = nullptr;
  if (bt709_ctx_create(0, &ctx) != BT709_OK) {
    fprintf(stderr, "no device\n");
    return 2;
  }
  printf("{\"probe\": \"texture-unit pass 2 vs exact walker\", \"in\": \"%dx%d\", \"out\": \"%dx%d\", \"cases\": [", iw, ih, ow, oh);
  for (int mode = 0; mode < 2; mode++) {
    std::vector<uint8_t> src;
    fill(src, iw, ih, mode, 709u + mode);
    cudaArray_t arr;
    cudaChannelFormatDesc fd = cudaCreateChannelDesc<uchar4>();
    CK(cudaMallocArray(&arr, &fd, iw, ih));
    CK(cudaMemcpy2DToArray(arr, 0, 0, src.data(), (size_t) iw * 4, (size_t) iw * 4, ih, cudaMemcpyHostToDevice));
    cudaResourceDesc rd{};
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = arr;
    cudaTextureDesc td{};
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModeLinear;
    td.readMode = cudaReadModeNormalizedFloat;
    td.sRGB = 1;
    td.normalizedCoords = 1;
    cudaTextureObject_t tex;
    CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
    uint8_t *d_in, *d_exact, *d_tex;
    CK(cudaMalloc(&d_in, src.size()));
    CK(cudaMalloc(&d_exact, (size_t) ow * oh * 4));
    CK(cudaMalloc(&d_tex, (size_t) ow * oh * 4));
    CK(cudaMemcpy(d_in, src.data(), src.size(), cudaMemcpyHostToDevice));
    bt709_plane pin = {d_in, (size_t) iw * 4, 0}, pout = {d_exact, (size_t) ow * 4, 0};
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float ms_exact = 0, ms_tex = 0;
    bt709_rescale_bgra(ctx, &pin, iw, ih, &pout, ow, oh, 1, nullptr, 1);
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; i++) bt709_rescale_bgra(ctx, &pin, iw, ih, &pout, ow, oh, 1, nullptr, 0);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms_exact, e0, e1));
    dim3 blk(32, 8), grd((ow + 31) / 32, (oh + 7) / 8);
    k_tex_rescale<<<grd, blk>>>(tex, d_tex, (size_t) ow * 4, ow, oh);
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; i++) k_tex_rescale<<<grd, blk>>>(tex, d_tex, (size_t) ow * 4, ow, oh);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms_tex, e0, e1));
    std::vector<uint8_t> a((size_t) ow * oh * 4), b(a.size());
    CK(cudaMemcpy(a.data(), d_exact, a.size(), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(b.data(), d_tex, b.size(), cudaMemcpyDeviceToHost));
    long long hist[256] = {0};
    int mx = 0;
    for (size_t i = 0; i < a.size(); i++) {
      if ((i & 3) == 3) continue;  // alpha is 255 on both sides
      int d = abs((int) a[i] - (int) b[i]);
      hist[d]++;
      if (d > mx) mx = d;
    }
    const double n = (double) ow * oh * 3;
    printf("%s{\"content\": \"%s\", \"exact_fraction\": %.6f, \"within_1\": %.6f, \"within_2\": %.6f, \"max_abs_delta\": %d, "
           "\"exact_walker_us_per_frame\": %.1f, \"texture_kernel_us_per_frame\": %.1f}",
           mode ? ", " : "", mode ? "smooth gradient + 2 LSB noise" : "white noise", hist[0] / n, (hist[0] + hist[1]) / n,
           (hist[0] + hist[1] + hist[2]) / n, mx, ms_exact * 1e3 / reps, ms_tex * 1e3 / reps);
    cudaDestroyTextureObject(tex);
    cudaFreeArray(arr);
    cudaFree(d_in);
    cudaFree(d_exact);
    cudaFree(d_tex);
  }
  printf("], \"note\": \"the texture kernel re-encodes with double pow: a probe, not a product path; one frame per launch on "
         "both sides\"}\n");
  bt709_ctx_destroy(ctx);
  return 0;
}
