// bt709_misc.cu — the reference's test-harness 4:4:4 packed path, the NV12 <-> planar chroma shuffles
// for the .y4m format, and the per-device kernel attribute setup.
#include "bt709_device.cuh"

namespace bt709 {

// ===================================================================== packed 4:4:4 harness
// +convertSoftware: (BGRAToBT709Converter.m:89-144): Apple196_from_sRGB_convertRGBToYCbCr per pixel
__global__ void __launch_bounds__(256) k_convert444(const Packed444Args p) {
  __shared__ float s_enc[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_enc[i] = __ldg(&p.enc256[i]);
  __syncthreads();
  for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (long long) gridDim.x * blockDim.x) {
    const uint32_t px = p.in[i];
    float rn = s_enc[(px >> 16) & 255], gn = s_enc[(px >> 8) & 255], bn = s_enc[px & 255];
    float ey = __fadd_rn(__fadd_rn(__fmul_rn(0.2126f, rn), __fmul_rn(0.7152f, gn)), __fmul_rn(0.0722f, bn));
    float eb = __fdiv_rn(__fsub_rn(bn, ey), 1.8556f);
    float er = __fdiv_rn(__fsub_rn(rn, ey), 1.5748f);
    uint32_t Y = (uint32_t) round_pos(__fadd_rn(__fmul_rn(ey, 219.0f), 16.0f));
    uint32_t Cb = (uint32_t) round_pos(__fadd_rn(__fmul_rn(eb, 224.0f), 128.0f));
    uint32_t Cr = (uint32_t) round_pos(__fadd_rn(__fmul_rn(er, 224.0f), 128.0f));
    p.out[i] = (Cr << 16) | (Cb << 8) | Y;
  }
}

// +unconvertSoftware: (BGRAToBT709Converter.m:146-198): Apple196_to_sRGB per pixel, A = 0
__global__ void __launch_bounds__(256) k_unconvert444(const Packed444Args p) {
  const float fN = (float) (1 << p.cells_log2);
  for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += (long long) gridDim.x * blockDim.x) {
    const uint32_t px = p.in[i];
    float yt = __fmul_rn(__fmul_rn((float) (px & 255) - 16.0f, p.m.inv255), p.m.ys);
    float cbn = __fmul_rn((float) ((px >> 8) & 255) - 128.0f, p.m.inv255);
    float crn = __fmul_rn((float) ((px >> 16) & 255) - 128.0f, p.m.inv255);
    float xr = add_sat(yt, __fmul_rn(crn, p.m.r_cr));
    float xg = add_sat(__fadd_rn(yt, __fmul_rn(cbn, p.m.g_cb)), __fmul_rn(crn, p.m.g_cr));
    float xb = add_sat(yt, __fmul_rn(cbn, p.m.b_cb));
    uint32_t R = lut_code(p.lut, xr, p.clamp_lo, fN), G = lut_code(p.lut, xg, p.clamp_lo, fN),
             B = lut_code(p.lut, xb, p.clamp_lo, fN);
    p.out[i] = (R << 16) | (G << 8) | B;
  }
}

// +copyBT709ToCoreVideo: (BGRAToBT709Converter.m:1042-1099): the chroma of a 2x2 block ends up
// being that of its last row's even column (the odd row overwrites the even row's write).
__global__ void __launch_bounds__(256) k_copy_packed(const CopyPackedArgs p) {
  const long long n = (long long) p.w * p.h;
  for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x) {
    const int col = (int) (i % p.w), row = (int) (i / p.w);
    const uint32_t px = p.in[i];
    p.y[(long long) row * p.py + col] = (uint8_t) (px & 255);
    const bool last_row_of_block = (row & 1) || row == p.h - 1;
    if ((col & 1) == 0 && last_row_of_block) {
      uint8_t *c = p.c + (long long) (row >> 1) * p.pc + col;
      c[0] = (uint8_t) ((px >> 8) & 255);
      c[1] = (uint8_t) ((px >> 16) & 255);
    }
  }
}

// ======================================================================= NV12 <-> planar chroma
// +copyYCBCr: (BGRAToBT709Converter.m:688-790) splits CbCr pairs into U and V planes for the .y4m
// writer; the inverse feeds .y4m frames to the decoder.  Pure byte shuffles, 2 B/sample each way.
template <bool SPLIT>
__global__ void __launch_bounds__(256) k_cbcr_planar(const PlanarArgs p) {
  const long long n = (long long) p.nframes * p.ch * p.cw;
  for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x) {
    const int x = (int) (i % p.cw);
    const long long r = i / p.cw;
    const int y = (int) (r % p.ch);
    const long long f = r / p.ch;
    uint8_t *c = p.c + f * p.sc + (long long) y * p.pc + 2 * x;
    uint8_t *u = p.u + f * p.su + (long long) y * p.pu + x;
    uint8_t *v = p.v + f * p.sv + (long long) y * p.pv + x;
    if (SPLIT) { *u = c[0]; *v = c[1]; } else { c[0] = *u; c[1] = *v; }
  }
}

int launch_cbcr_planar(const PlanarArgs &a, bool split, void *stream) {
  const long long n = (long long) a.nframes * a.ch * a.cw;
  long long g = (n + 255) / 256;
  g = g > 148 * 16 ? 148 * 16 : (g < 1 ? 1 : g);
  if (split)
    k_cbcr_planar<true><<<(unsigned) g, 256, 0, (cudaStream_t) stream>>>(a), note_launch();
  else
    k_cbcr_planar<false><<<(unsigned) g, 256, 0, (cudaStream_t) stream>>>(a), note_launch();
  return (int) cudaGetLastError();
}

static unsigned grid_1d(long long n) {
  long long g = (n + 255) / 256;
  return (unsigned) (g > 148 * 16 ? 148 * 16 : (g < 1 ? 1 : g));
}

int launch_convert444(const Packed444Args &a, void *stream) {
  k_convert444<<<grid_1d(a.n), 256, 0, (cudaStream_t) stream>>>(a), note_launch();
  return (int) cudaGetLastError();
}
int launch_unconvert444(const Packed444Args &a, void *stream) {
  k_unconvert444<<<grid_1d(a.n), 256, 0, (cudaStream_t) stream>>>(a), note_launch();
  return (int) cudaGetLastError();
}
int launch_copy_packed(const CopyPackedArgs &a, void *stream) {
  k_copy_packed<<<grid_1d((long long) a.w * a.h), 256, 0, (cudaStream_t) stream>>>(a), note_launch();
  return (int) cudaGetLastError();
}

int decode_kernels_init();
int rescale_kernels_init();
int encode_kernels_init();

// per device (called from bt709_ctx_create with the device current)
int kernels_init() {
  int e;
  if ((e = decode_kernels_init())) return e;
  if ((e = rescale_kernels_init())) return e;
  if ((e = fused_kernels_init())) return e;  // after the rescale init: it probes where dynamic shared memory starts
  return encode_kernels_init();
}

}  // namespace bt709
