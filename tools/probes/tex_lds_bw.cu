// Probe: can table lookups issued as texture fetches (TEX data pipe of L1TEX) run alongside
// shared-memory lookups (LSU data pipe) and add lookup throughput?  The encode and rescale kernels
// sit at ~87 % of the LSU data-pipe wavefront rate while the L1 data banks are ~34 % busy.
// Each thread does ITER rounds of NL conflict-free LDS.64 gathers (per-lane replicated 256-entry
// float2 table, the encode kernel's layout) plus NT tex1Dfetch<float2> gathers from a 256-entry
// linear texture, with pseudo-random byte indices.  Prints SM cycles per round per warp.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tex_lds_bw tex_lds_bw.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

template <int NL, int NT>
__global__ void __launch_bounds__(1024, 1) k(cudaTextureObject_t tex, const float2 *tab, float *out, int iters) {
  extern __shared__ __align__(16) float2 s_tab[];  // entry b replica r at b*32 + r
  for (int i = threadIdx.x; i < 256 * 32; i += blockDim.x) s_tab[i] = tab[i >> 5];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const uint32_t base = (uint32_t) __cvta_generic_to_shared(s_tab) + lane * 8;
  uint32_t x = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 12345u;
  float acc0 = 0.f, acc1 = 0.f;
  for (int it = 0; it < iters; it++) {
    x = x * 1664525u + 1013904223u;
    uint32_t r = x;
#pragma unroll
    for (int j = 0; j < NL; j++) {
      const uint32_t b = (r >> ((j & 3) * 8)) & 255u;
      float2 v;
      asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(base + (b << 8)));
      acc0 += v.x;
      acc1 += v.y;
      if ((j & 3) == 3) r = r * 22695477u + 1u;
    }
#pragma unroll
    for (int j = 0; j < NT; j++) {
      const int b = (int) ((r >> ((j & 3) * 8)) & 255u);
      const float2 v = tex1Dfetch<float2>(tex, b);
      acc0 += v.x;
      acc1 += v.y;
      if ((j & 3) == 3) r = r * 22695477u + 1u;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc0 + acc1;
}

template <int NL, int NT>
static void run(cudaTextureObject_t tex, const float2 *tab, float *out, int sms, int clock_khz) {
  const int iters = 2000;
  cudaFuncSetAttribute(k<NL, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<NL, NT><<<sms, 1024, 65536>>>(tex, tab, out, 100);
  cudaEventRecord(e0);
  k<NL, NT><<<sms, 1024, 65536>>>(tex, tab, out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double cyc = ms * 1e-3 * clock_khz * 1e3;            // SM cycles
  const double per_round_warp = cyc / iters / 32.0;           // 32 warps per SM, all sharing one L1TEX
  printf("NL=%2d NT=%2d  %8.3f ms  %7.2f SM-cycles per warp-round  (%.2f per lookup)  err=%d\n", NL, NT, ms,
         per_round_warp, per_round_warp / (NL + NT), (int) cudaGetLastError());
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  std::vector<float2> h(256);
  for (int i = 0; i < 256; i++) h[i] = make_float2(i * 0.5f, i * 0.25f);
  float2 *tab;
  float *out;
  cudaMalloc(&tab, 256 * sizeof(float2));
  cudaMalloc(&out, prop.multiProcessorCount * 1024 * sizeof(float));
  cudaMemcpy(tab, h.data(), 256 * sizeof(float2), cudaMemcpyHostToDevice);
  cudaResourceDesc rd = {};
  rd.resType = cudaResourceTypeLinear;
  rd.res.linear.devPtr = tab;
  rd.res.linear.desc = cudaCreateChannelDesc<float2>();
  rd.res.linear.sizeInBytes = 256 * sizeof(float2);
  cudaTextureDesc td = {};
  td.readMode = cudaReadModeElementType;
  cudaTextureObject_t tex;
  printf("create tex: %d\n", (int) cudaCreateTextureObject(&tex, &rd, &td, nullptr));
  const int sms = prop.multiProcessorCount, khz = prop.clockRate;
  printf("%s, %d SMs, %d kHz\n", prop.name, sms, khz);
  run<12, 0>(tex, tab, out, sms, khz);
  run<8, 0>(tex, tab, out, sms, khz);
  run<0, 4>(tex, tab, out, sms, khz);
  run<0, 12>(tex, tab, out, sms, khz);
  run<8, 4>(tex, tab, out, sms, khz);
  run<8, 2>(tex, tab, out, sms, khz);
  run<10, 2>(tex, tab, out, sms, khz);
  run<6, 6>(tex, tab, out, sms, khz);
  return 0;
}
