// bt709_kernels.cuh — launch-argument structs shared by the kernel translation units (bt709_decode/rescale/encode/misc.cu)
// and the C ABI (bt709_api.cpp).  Plain PODs, passed by value.
#pragma once
#include <cstdint>

#include "bt709_tables.h"

namespace bt709 {

struct DecodeArgs {
  const uint8_t *y, *c, *a;  // a == nullptr: opaque
  uint8_t *out;
  long long py, pc, pa, po;  // row pitches, bytes
  long long sy, sc, sa, so;  // frame strides, bytes
  int w, h, nframes;
  int rows2;            // ceil(h / 2) row pairs per frame
  int tiles_per_row;    // 128-pixel tiles per row pair (fast kernels)
  int rps;              // row pairs a CTA covers per step = 32 / tiles_per_row (>= 1)
  int tile_stride;      // min(tiles_per_row, 32): warps loop over tiles when a row has more than 32
  // fast kernel addressing: 32-bit offsets in load/store units (4 B for Y/CbCr/alpha, 16 B for BGRA)
  unsigned step_rp, d_rp;               // row pairs a CTA band advances per step; its remainder mod rows2
  unsigned py4, pc4, pa4, po16;         // pitches in units
  unsigned sy4, sc4, sa4, so16;         // frame strides in units
  unsigned dy, dc, da, dout;            // offset increment per step
  unsigned wy, wc, wa, wout;            // extra increment when the row pair wraps into the next frame
  // ring kernel, linear tile order (launcher sets it for contiguous batches whose rows have fewer than 32 tiles
  // and leave warps idle in the band layout): 0 = band layout, else tiles per row.  The 32 warps of a CTA then
  // take 32 consecutive tiles of the row-major tile sequence; the per-step increments step the tile COLUMN and
  // the wrap increments fire when the column passes the end of the row (same code as the frame wrap).
  int linear_tiles;
  // k_decode_lin (contiguous full-tile batches): the batch is one tall image whose 128-pixel x 2-row tiles are
  // taken in row-major order by `lin_warps` warps (a multiple of the tiles per row, so a warp keeps its tile
  // column for the whole kernel and steps `lin_drp` row pairs down per step: no wrap logic, constant strides)
  int lin_warps, lin_drp, lin_wpc;      // active warps in the grid; row pairs per step; warps per CTA
  // bulk-staged kernel: shared-memory ring geometry
  int nstages, stage_bytes, nrows;      // ring depth; bytes per stage; rows (= bulk copies) per stage
  const uint32_t *lut;  // cell LUT of the selected gamma, N + 1 words, device memory
  float clamp_lo;
  int cells_log2;
  // optional log-spaced cell LUT of the same step function (bt709_tables.h: LogLut); nullptr = none
  const uint32_t *log_lut;
  float log_clamp_lo;
  int log_cells;
  unsigned log_first;
  // optional biased log-cell LUT of the same step function (bt709_tables.h: BiasedLogLut); nullptr = none.
  // k_decode_lin<-1, ...> reads it through lut / clamp_lo / log_cells / log_first with log_bias set.
  const uint32_t *blog_lut;
  float blog_clamp_lo, blog_bias;
  int blog_cells;
  unsigned blog_first;
  float log_bias;
  MatrixConsts m;
  float one;  // 1.0f, deliberately a runtime value (see decode_row)
};

struct RescaleArgs {
  const uint8_t *in;
  uint8_t *out;
  long long pi, po, si, so;
  int iw, ih, ow, oh, nframes;
  float scale_x, scale_y;   // (float)iw / (float)ow, (float)ih / (float)oh
  const float *lin256;      // sRGB byte -> linear float
  const uint32_t *enc_lut;  // linear float -> sRGB byte (cell LUT)
  float enc_clamp_lo;
  int enc_cells_log2;
  // optional log-cell form of the same encode step function (bt709_tables.h: LogLut); nullptr = none
  const uint32_t *enc_log_lut;
  float enc_log_clamp_lo;
  int enc_log_cells;
  unsigned enc_log_first;
  // optional biased log-cell form (bt709_tables.h: BiasedLogLut), one replica per lane in k_rescale_pair
  const uint32_t *enc_blog_lut;
  float enc_blog_clamp_lo, enc_blog_bias;
  int enc_blog_cells;
  unsigned enc_blog_first;
  float one;  // 1.0f, deliberately a runtime value (blocks FFMA2 contraction, see k_rescale_cols)
  int known_opaque;  // the caller guarantees A = 255 everywhere (pass 1 without an alpha stream wrote the source)
  int pair_th;       // k_rescale_pair: output rows per warp tile (launcher: kPairTH, less for calls of a few frames)
};

struct EncodeArgs {
  const uint8_t *in;
  uint8_t *y, *c;
  long long pi, py, pc, si, sy, sc;
  int w, h, nframes;
  const float *lin256;     // input-gamma byte -> linear float          (BT709_tolinearNorm)
  const float *compn256;   // byte -> unorm8(from_linear(to_linear(b)))  (per-pixel luma path)
  const uint32_t *out_lut;  // linear float -> output-gamma byte (cell LUT, BT709_from_linear)
  float out_clamp_lo;
  int out_cells_log2;
  // optional biased log-cell form of the same step function (bt709_tables.h: BiasedLogLut): ~645 cells for the Apple and
  // sRGB curves instead of 4096 uniform ones, so one replica per lane fits shared memory (k_encode_fast: OMODE 2)
  const uint32_t *out_blog_lut;
  float out_blog_clamp_lo, out_blog_bias;
  int out_blog_cells;
  unsigned out_blog_first;
  // fast kernel: band layout and 32-bit unit offsets (16 B for BGRA, 4 B for Y / CbCr), see k_encode_fast
  int tiles_per_row, rps, tile_stride;
  // linear tile order (contiguous batches whose rows have fewer than 32 tiles and would leave warps idle in the band
  // layout, e.g. 15 or 30 tiles for 1920 / 3840-wide rows): 0 = band layout, else the number of active warps of the grid,
  // a multiple of tiles_per_row - warp g takes tile column g % tiles_per_row of row pairs g / tiles_per_row + k * step_rp
  int lin_warps;
  unsigned step_rp, d_rp, pi16, py4, pc4, si16, sy4, sc4, di, dy, dc, wi, wy, wc;
  float one;  // 1.0f, deliberately a runtime value
  unsigned zero;  // 0, deliberately a runtime value (scheduling fence by data dependence, see k_encode_fast)
};

struct Packed444Args {
  const uint32_t *in;
  uint32_t *out;
  long long n;           // pixels
  const float *enc256;   // convert: byte -> apple_oetf(srgb_eotf(b/255))
  const uint32_t *lut;   // unconvert: Apple decode cell LUT
  float clamp_lo;
  int cells_log2;
  MatrixConsts m;
};

struct CopyPackedArgs {
  const uint32_t *in;
  uint8_t *y, *c;
  long long py, pc;
  int w, h;
};

struct PlanarArgs {
  uint8_t *c, *u, *v;
  long long pc, pu, pv, sc, su, sv;
  int cw, ch, nframes;
};

// Fused decode -> rescale (bt709_fused.cu): one persistent CTA per SM takes (frame, band of 64 output rows) units;
// it decodes the band's source rows into its own slice of `scratch` (a few hundred KB that it overwrites band after
// band, so the lines live in L2 and the intermediate image never reaches HBM) and then runs the paired-column
// walker of pass 2 over them.
struct FusedArgs {
  DecodeArgs d;       // NV12 source (opaque: no alpha plane), gamma tables
  RescaleArgs r;      // iw/ih = frame size, ow/oh/out = destination, encode tables (in/pi/si unused)
  uint8_t *scratch;   // device memory, gridDim.x * cta_stride bytes
  long long cta_stride;
  unsigned spitch;    // bytes per scratch row (>= w * 4, multiple of 16)
  int bands;          // ceil(oh / band_rows)
  int band_rows;      // output rows per band: kWalkTH, less for calls with fewer bands than SMs
};

// number of kernels the calling thread has launched through the launchers below (bt709_launch_count accounting)
extern thread_local unsigned long long tl_kernel_launches;
inline void note_launch() { ++tl_kernel_launches; }

// launchers (one per kernel family).  Each returns the cudaError_t of the launch.
int launch_decode(const DecodeArgs &a, int sm_count, void *stream);
int launch_rescale(const RescaleArgs &a, int sm_count, void *stream);
// fused decode -> rescale; scratch_bytes: size of the caller's scratch area.  Returns -1 when the fused kernel does not
// apply (geometry, alignment, alpha stream, scratch too small): the caller then runs the two passes.
int launch_fused(const DecodeArgs &d, const RescaleArgs &r, uint8_t *scratch, size_t scratch_bytes, int sm_count, void *stream);
int fused_kernels_init();
int launch_encode(const EncodeArgs &a, int sm_count, void *stream);
int launch_convert444(const Packed444Args &a, void *stream);
int launch_unconvert444(const Packed444Args &a, void *stream);
int launch_copy_packed(const CopyPackedArgs &a, void *stream);
int launch_cbcr_planar(const PlanarArgs &a, bool split, void *stream);
// one-time per-device kernel attribute setup (dynamic shared memory opt-in)
int kernels_init();

}  // namespace bt709
