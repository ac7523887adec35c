// bt709_fused.cu — decode -> rescale in ONE kernel (AAPLRenderer.m:940-977 is the
// two-pass path it replaces, MetalBT709Decoder.m:612-618 the author's note that a single pass would halve the traffic).
//
// The same-size sRGB intermediate of the two-pass path (the reference's _resizeTexture) is 8.3 MB per 1080p frame
// written to and read back from HBM.  Here a persistent CTA takes one (frame, band of 64 output rows) unit at a time:
//   phase A  all warps decode the source rows the band needs (plus the row above and below, rounded to chroma row
//            pairs) with the pass-1 arithmetic (decode_row: same float32 steps, same cell LUT, same 8-bit sRGB
//            quantisation) into the CTA's private slice of a scratch area;
//   phase B  after a CTA barrier the warps run the paired-column walker of pass 2 (pair_tile, opaque code: pass 1
//            without an alpha stream writes A = 255) over those rows and stream the destination rows out.
// A slice is ~52 rows (400 KB for 1080p) and is overwritten band after band, 148 slices = 59 MB, so the lines stay
// in the 126 MB L2: HBM sees 1.5 B per source pixel in and 4 B per destination pixel out (17.856 MB per 1080p ->
// 1440p frame instead of 34.4 MB) and the result is bit-identical to decode-then-rescale by construction.
//
// Both phases are instruction-bound, not HBM-bound (profiles/), so the fused kernel is not faster than the two
// passes by much; what it saves is the 16.6 MB per frame of HBM traffic (and its power).  Shared memory holds the
// tables of BOTH passes: the encode LUT (biased log cells, one replica per lane, 82 KB), the sRGB byte -> linear table
// (64 KB) and the decode cell LUT with 8 replicas (33 KB; 2-3 way bank conflicts on random content - the per-lane
// replicated 128 KB version of the pass-1 kernels does not fit next to the other two).  The Linear mode needs no
// decode LUT of its own: its step function IS the sRGB encode, so it reads the encode LUT.
#include "bt709_decode_row.cuh"
#include "bt709_pair.cuh"

namespace bt709 {

size_t pair_layout_bytes(unsigned base, int cells, unsigned extra_bytes, int warps, int row_words);  // bt709_rescale.cu
extern int g_dyn_smem_base;

// CL2/RL2: decode LUT geometry (10, 3), or (-1, 5) = the encode LUT itself (Linear mode)
template <int CL2, int RL2>
__global__ void __launch_bounds__(kFusedWarps * 32, 1) k_fused(const FusedArgs a) {
  extern __shared__ __align__(16) uint32_t s_mem[];
  constexpr int R = 1 << RL2, N = 1 << (CL2 > 0 ? CL2 : 0);
  constexpr uint32_t kDecBytes = CL2 > 0 ? (((uint32_t) (N + 1) * R * 4u + 127u) & ~127u) : 0u;
  const DecodeArgs &d = a.d;
  const RescaleArgs &p = a.r;
  const PairCtx c = pair_tables(p, s_mem, kDecBytes);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  uint32_t sbase;
  if (CL2 > 0) {
    const uint32_t smem0 = (uint32_t) __cvta_generic_to_shared(s_mem);
    uint32_t *s_dec = s_mem + ((c.extra - smem0) >> 2);
    fill_replicated<RL2>(s_dec, R, d.lut, N + 1);
    __syncthreads();
    sbase = c.extra + ((lane & (R - 1)) << 2) - (0x4B000000u << (2 + RL2));
  } else {
    sbase = c.enc_base;  // cvta(encode LUT) - (first << 7) + (lane << 2): exactly the <-1, 5> lookup base
  }
  asm volatile("" : "+r"(sbase));  // one register, not a base + constant that ptxas re-adds at every lookup
  const float clamp_lo = d.clamp_lo;
  DecodeK2 k;
  k.kY = pk(-(8388608.0f + 16.0f), -(8388608.0f + 16.0f));
  k.kC = pk(-(8388608.0f + 128.0f), -(8388608.0f + 128.0f));
  k.inv255 = pk(d.m.inv255, d.m.inv255);
  k.ys = pk(d.m.ys, d.m.ys);
  k.g_cb = pk(d.m.g_cb, d.m.g_cb);
  k.b_cb = pk(d.m.b_cb, d.m.b_cb);
  k.r_cr = pk(d.m.r_cr, d.m.r_cr);
  k.g_cr = pk(d.m.g_cr, d.m.g_cr);
  k.fN = pk((float) N, (float) N);
  k.magic = pk(8388608.0f, 8388608.0f);
  k.one = pk(d.one, d.one);
  k.bias = pk(d.log_bias, d.log_bias);

  unsigned long long pol_last;  // L2 cache policy of the slice: keep (evict last)
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_last));
  const unsigned T = (unsigned) d.tiles_per_row;
  constexpr int TW = 64 * kPairNP;
  const int tiles_x = (p.ow + TW - 1) / TW;
  const unsigned units = (unsigned) d.nframes * (unsigned) a.bands;
  uint8_t *scr = a.scratch + (long long) blockIdx.x * a.cta_stride;
  const unsigned spitch = a.spitch;
  for (unsigned u = blockIdx.x; u < units; u += gridDim.x) {
    const unsigned f = u / (unsigned) a.bands, band = u - f * (unsigned) a.bands;
    const int oy0 = (int) band * a.band_rows;
    const int oyl = min(oy0 + a.band_rows, p.oh) - 1;
    // source rows of the band: top tap of its first row .. bottom tap of its last row, same float32 steps as the walker
    const float sy0 = __fadd_rn(__fmul_rn(__fadd_rn((float) oy0, 0.5f), p.scale_y), -0.5f);
    const float syl = __fadd_rn(__fmul_rn(__fadd_rn((float) oyl, 0.5f), p.scale_y), -0.5f);
    const int r0 = max((int) floorf(sy0), 0), r1 = min(max((int) floorf(syl) + 1, 0), p.ih - 1);
    const unsigned rp0 = (unsigned) r0 >> 1, nrp = ((unsigned) r1 >> 1) - rp0 + 1;

    // ---- phase A: decode row pairs rp0 .. rp0 + nrp - 1 into the slice (128-pixel x 2-row tiles, one per warp step).
    // The words of the next two tiles are in flight during a tile's arithmetic; (row pair, tile column) advance
    // incrementally; the slice is written with the L2 evict_last policy so that the streaming output of phase B
    // does not push it out to HBM before the next band overwrites it.
    const uint8_t *yf = d.y + (long long) f * d.sy, *cf = d.c + (long long) f * d.sc;
    const unsigned ntiles = nrp * T;
    const unsigned step_rp = (unsigned) nwarps / T, step_tile = (unsigned) nwarps - step_rp * T;
    struct Words { uint32_t y0, y1, c; };
    unsigned lrp = (unsigned) warp / T, ltile = (unsigned) warp - lrp * T;  // (row pair, tile) of the next load
    auto load = [&](Words &wd) {
      const unsigned rp = rp0 + lrp, col4 = ltile * 32 + lane;
      wd.y0 = ld_stream_u32(yf + (long long) (2 * rp) * d.py + (size_t) col4 * 4);
      wd.y1 = ld_stream_u32(yf + (long long) (2 * rp + 1) * d.py + (size_t) col4 * 4);
      wd.c = ld_stream_u32(cf + (long long) rp * d.pc + (size_t) col4 * 4);
      lrp += step_rp;
      ltile += step_tile;
      if (ltile >= T) { ltile -= T; lrp++; }
    };
    Words q0{0, 0, 0}, q1{0, 0, 0};
    if ((unsigned) warp < ntiles) load(q0);
    if ((unsigned) warp + nwarps < ntiles) load(q1);
    unsigned srp = (unsigned) warp / T, stile = (unsigned) warp - srp * T;  // (row pair, tile) being decoded
    for (unsigned t = warp; t < ntiles; t += nwarps) {
      const Words cur = q0;
      q0 = q1;
      if (t + 2 * nwarps < ntiles) load(q1);
      const unsigned col4 = stile * 32 + lane;
      f32x2 cbn = mul2(add2(pku(magic_byte<0>(cur.c), magic_byte<2>(cur.c)), k.kC), k.inv255);
      f32x2 crn = mul2(add2(pku(magic_byte<1>(cur.c), magic_byte<3>(cur.c)), k.kC), k.inv255);
      f32x2 cb_g = mul2(cbn, k.g_cb), cb_b = mul2(cbn, k.b_cb);
      f32x2 cr_r = mul2(crn, k.r_cr), cr_g = mul2(crn, k.g_cr);
      uint8_t *o = scr + (size_t) (2 * srp) * spitch + (size_t) col4 * 16;
      st_l2hint_u128(o, decode_row<CL2, RL2, false>(cur.y0, 0u, cb_g, cb_b, cr_r, cr_g, k, clamp_lo, sbase), pol_last);
      st_l2hint_u128(o + spitch, decode_row<CL2, RL2, false>(cur.y1, 0u, cb_g, cb_b, cr_r, cr_g, k, clamp_lo, sbase), pol_last);
      srp += step_rp;
      stile += step_tile;
      if (stile >= T) { stile -= T; srp++; }
    }
    __syncthreads();  // the slice is complete and visible to the whole CTA

    // ---- phase B: pass 2 over the slice; row y of the frame is slice row y - 2 * rp0
    const uint8_t *src = scr - (long long) (2 * rp0) * spitch;
    uint8_t *dst = p.out + (long long) f * p.so;
    for (int tx = warp; tx < tiles_x; tx += nwarps) pair_tile<false, false, kPairNP, false, kFusedPD>(p, c, src, spitch, dst, tx * TW, oy0, lane, pol_last, a.band_rows);
    __syncthreads();  // every warp is done reading before the next band overwrites the slice
  }
}

int launch_fused(const DecodeArgs &d0, const RescaleArgs &r0, uint8_t *scratch, size_t scratch_bytes, int sm_count, void *stream) {
  static const bool off = getenv("BT709_NO_FUSED") != nullptr;
  if (off || !scratch || d0.a != nullptr) return -1;
  DecodeArgs d = d0;
  RescaleArgs r = r0;
  d.rows2 = (d.h + 1) / 2;
  d.tiles_per_row = (d.w + 127) / 128;
  // geometry of the two halves: full 128-pixel tiles, even height; horizontal scale <= 1 and at most 3 source rows per
  // output row (the walker's range); aligned planes
  if (d.w % 128 != 0 || d.h % 2 != 0 || r.iw != d.w || r.ih != d.h || r.iw > r.ow || (long long) r.ih > 3ll * r.oh) return -1;
  if (!(aligned(d.y, 4) && aligned(d.py, 4) && aligned(d.sy, 4) && aligned(d.c, 4) && aligned(d.pc, 4) && aligned(d.sc, 4) &&
        aligned(r.out, 8) && aligned(r.po, 8) && aligned(r.so, 8) && r.ow % 2 == 0 && aligned(scratch, 16)))
    return -1;
  if (!r.enc_blog_lut || g_dyn_smem_base < 0 || (long long) r.oh * r.po >= (1ll << 32) || r.po < 0) return -1;
  const bool linear = d.blog_lut && d.blog_cells > 0;  // Linear mode: the decode step function is the encode LUT
  if (!linear && d.cells_log2 != 10) return -1;
  const unsigned dec_bytes = linear ? 0u : ((1025u * 8u * 4u + 127u) & ~127u);
  const size_t smem = pair_layout_bytes((unsigned) g_dyn_smem_base, r.enc_blog_cells, dec_bytes, kFusedWarps, kWalkRowInfo);
  if (smem > kMaxDynSmem) return -1;
  FusedArgs a{};
  // band height: kWalkTH rows, halved while the call has fewer (frame, band) units than SMs (a CTA works on one unit)
  a.band_rows = kWalkTH;
  while (a.band_rows > 16 && (long long) d.nframes * ((r.oh + a.band_rows - 1) / a.band_rows) < (long long) sm_count) a.band_rows >>= 1;
  a.bands = (r.oh + a.band_rows - 1) / a.band_rows;
  const long long units = (long long) d.nframes * a.bands;
  if (units >= (1ll << 31)) return -1;
  // rows a band can need: its source span, both taps, rounding to chroma row pairs, and the walker's prefetch overrun
  const int rows_max = (int) ((double) a.band_rows * (double) r.ih / (double) r.oh) + 8 + kFusedPD + 2;
  a.spitch = (unsigned) (((size_t) d.w * 4 + 255) & ~(size_t) 255);
  a.cta_stride = (long long) rows_max * a.spitch;
  // one CTA per SM, fewer when there are fewer units or the scratch area has room for fewer slices
  long long g = units < (long long) sm_count ? units : (long long) sm_count;
  const long long room = (long long) (scratch_bytes / (unsigned long long) a.cta_stride);
  if (room < g) g = room;
  if (g < 1) return -1;
  const unsigned grid = (unsigned) g;
  if (linear) {
    d.lut = d.blog_lut;
    d.clamp_lo = d.blog_clamp_lo;
    d.log_cells = d.blog_cells;
    d.log_first = d.blog_first;
    d.log_bias = d.blog_bias;
  }
  a.d = d;
  a.r = r;
  a.scratch = scratch;
  cudaStream_t s = (cudaStream_t) stream;
  if (linear)
    k_fused<-1, 5><<<grid, kFusedWarps * 32, smem, s>>>(a), note_launch();
  else
    k_fused<10, 3><<<grid, kFusedWarps * 32, smem, s>>>(a), note_launch();
  return (int) cudaGetLastError();
}

int fused_kernels_init() {
  int e = (int) cudaFuncSetAttribute(k_fused<10, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kMaxDynSmem);
  if (e) return e;
  return (int) cudaFuncSetAttribute(k_fused<-1, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kMaxDynSmem);
}

}  // namespace bt709
