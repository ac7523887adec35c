// bt709_decode.cu — pass 1: video-range BT.709 NV12 -> sRGB BGRA8 (sm_100a kernels + launcher).
//
// Replaces the decode kernels/fragments of AAPLShaders.metal:180-438 and the dispatch of
// MetalBT709Decoder.m:497-610.  All arithmetic that feeds a quantiser is done with explicitly
// rounded float32 operations (never contracted to FMA) in the order the reference's CPU code uses
// (BT709.h:389-397, 424-426, 494-500), and every transfer curve is an exact table lookup
// (bt709_tables.h), so results are bit-identical to the reference's CPU path, not merely +-1 LSB.
//
// No tensor cores: the path is a per-pixel map with 5.5 B/pixel of HBM traffic and no contraction.
// What matters is (a) 16-byte coalesced stores and wide enough load requests with enough bytes in
// flight per SM, (b) keeping the per-pixel instruction count under the issue budget (about 37
// warp-instructions per 32 pixels at 80 % of the HBM roofline), (c) keeping the three curve lookups
// per pixel off the MUFU pipe and free of shared-memory bank conflicts (per-lane replicated LUT).
//
// Kernels (all share decode_row / tile_compute_store):
//   k_decode_ring     default: per-warp cp.async shared-memory ring for the inputs; band layout, or linear
//                     tile order when the tiles per row do not divide the CTA's 32 warps (1920, 3840 wide)
//   k_decode_fast     fallback: per-thread LDG with register prefetch (any 4-byte aligned planes, ragged and
//                     2-mod-4 widths, rows wider than 32 tiles)
//   k_decode_bulk     opt-in (BT709_BULK=1): cp.async.bulk row copies + mbarriers; measured slower
//   k_decode_generic  any size, pitch and alignment; one thread per 2x2 block
//
// Tile shape of the fast kernels: a warp owns 128 pixels x 2 rows (one chroma row); lane l handles
// pixels [4l, 4l+4) of both rows, so every warp-level access is contiguous:
//   Y     2 rows x 128 B,   CbCr  1 x 128 B (read once, used for both rows: nearest 4:2:0 upsample),
//   BGRA  2 rows x STG.128 = 512 B per instruction.
// A pair of packed floats always holds the same pixel position of chroma pair A (pixels 0,1) and
// chroma pair B (pixels 2,3), so chroma products pair up naturally with the luma terms.
#include "bt709_decode_row.cuh"

namespace bt709 {

// Work layout.  A warp owns a 128-pixel x 2-row tile (lane l: pixels [4l, 4l+4) of both rows).
// The warps of the CTA sit side by side: they cover `rps` complete row pairs per step
// (rps = warps / tiles per row), so a CTA touches 2*rps whole contiguous rows of Y, rps rows of
// CbCr and 2*rps rows of BGRA at a time, and consecutive CTAs take consecutive bands.  That keeps
// DRAM pages hot (a warp streaming its own row did not: 4736 independent row streams measured
// 51 % long-scoreboard stalls at 4.2 TB/s).  Each warp's (tile, row-in-band) is fixed for the whole
// kernel; per step only (frame, row pair) advance, by a constant with a carry, and every address
// is two IMAD.WIDE away.  The loads of the next step are issued before the current step is
// computed; the loop is unrolled twice so the two load buffers ping-pong without moves.
// FULL = (w % 128 == 0 && h % 2 == 0): no per-lane or per-row predicates anywhere.
template <bool ALPHA>
struct TileLoad {
  uint32_t y0, y1, c, a0, a1;
};

// Addressing.  Every plane is addressed as base + 32-bit offset in load/store units (4 bytes
// for Y / CbCr / alpha, 16 bytes for BGRA), so an address is a single IMAD.WIDE.U32 and stepping
// to the next (frame, row pair) is one 32-bit add per plane plus a predicated add on a frame
// wrap.  The launcher splits batches whose offsets would not fit in 32 bits.
template <bool ALPHA, bool FULL>
__device__ __forceinline__ void tile_load(const DecodeArgs &p, unsigned oy, unsigned oc, unsigned oa, unsigned rp,
                                          bool valid, TileLoad<ALPHA> &t) {
  const bool has_row1 = FULL || rp * 2 + 1 < (unsigned) p.h;
  if (!FULL) t.y0 = t.y1 = t.c = t.a0 = t.a1 = 0;
  if (FULL || valid) {
    t.y0 = ld_stream_u32(ptr_mad(p.y, oy, 4));
    t.c = ld_stream_u32(ptr_mad(p.c, oc, 4));
    if (has_row1) t.y1 = ld_stream_u32(ptr_mad(p.y, oy + p.py4, 4));
    if (ALPHA) {
      t.a0 = ld_stream_u32(ptr_mad(p.a, oa, 4));
      if (has_row1) t.a1 = ld_stream_u32(ptr_mad(p.a, oa + p.pa4, 4));
    }
  }
}

template <int CL2, int RL2, bool ALPHA, bool FULL>
__device__ __forceinline__ void tile_compute_store(const DecodeArgs &p, unsigned oo, unsigned rp, bool valid, bool half,
                                                   const TileLoad<ALPHA> &t, const DecodeK2 &k, float clamp_lo,
                                                   uint32_t sbase) {
  if (!FULL && !valid) return;
  const bool has_row1 = FULL || rp * 2 + 1 < (unsigned) p.h;
  // chroma pairs A (bytes 0,1) and B (bytes 2,3): BT709.h:499-500 then the matrix columns
  f32x2 cbn = mul2(add2(pku(magic_byte<0>(t.c), magic_byte<2>(t.c)), k.kC), k.inv255);
  f32x2 crn = mul2(add2(pku(magic_byte<1>(t.c), magic_byte<3>(t.c)), k.kC), k.inv255);
  f32x2 cb_g = mul2(cbn, k.g_cb), cb_b = mul2(cbn, k.b_cb);
  f32x2 cr_r = mul2(crn, k.r_cr), cr_g = mul2(crn, k.g_cr);
  // half (w % 4 == 2 only): the last lane of a row owns two pixels and stores 8 bytes; its 4-byte loads stay
  // inside the row pitch, which is a multiple of 4 on this path
  auto store = [&](unsigned off16, uint4 v) {
    uint8_t *dst = (uint8_t *) ptr_mad(p.out, off16, 16);
    if (!FULL && half)
      st_stream_u64(dst, make_uint2(v.x, v.y));
    else
      st_stream_u128(dst, v);
  };
  store(oo, decode_row<CL2, RL2, ALPHA>(t.y0, t.a0, cb_g, cb_b, cr_r, cr_g, k, clamp_lo, sbase));
  if (has_row1) store(oo + p.po16, decode_row<CL2, RL2, ALPHA>(t.y1, t.a1, cb_g, cb_b, cr_r, cr_g, k, clamp_lo, sbase));
}

// WRAP = false: the batch is one contiguous tall image (frame stride == rows * pitch for every
// plane), so stepping never crosses a frame seam and the carry logic disappears.
template <int CL2, int RL2, bool ALPHA, bool FULL, bool WRAP, int PD, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) k_decode_fast(const DecodeArgs p) {
  extern __shared__ __align__(128) uint32_t s_lut[];
  constexpr int N = 1 << CL2, R = 1 << RL2;
  const int entries = CL2 ? N + 1 : p.log_cells;  // CL2 == 0: log-spaced cell LUT (bt709_tables.h: LogLut)
  fill_replicated<RL2>(s_lut, R, p.lut, entries);
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const unsigned warp = threadIdx.x >> 5;
  // the index register holds 2^23 + cell (uniform cells) or first + cell (log cells): fold the bias in
  const uint32_t sbase = (uint32_t) __cvta_generic_to_shared(s_lut) + ((lane & (R - 1)) << 2) -
                         ((CL2 ? 0x4B000000u : p.log_first) << (2 + RL2));
  const float clamp_lo = p.clamp_lo;
  DecodeK2 k;
  k.kY = pk(-(8388608.0f + 16.0f), -(8388608.0f + 16.0f));
  k.kC = pk(-(8388608.0f + 128.0f), -(8388608.0f + 128.0f));
  k.inv255 = pk(p.m.inv255, p.m.inv255);
  k.ys = pk(p.m.ys, p.m.ys);
  k.g_cb = pk(p.m.g_cb, p.m.g_cb);
  k.b_cb = pk(p.m.b_cb, p.m.b_cb);
  k.r_cr = pk(p.m.r_cr, p.m.r_cr);
  k.g_cr = pk(p.m.g_cr, p.m.g_cr);
  k.fN = pk((float) N, (float) N);
  k.magic = pk(8388608.0f, 8388608.0f);
  k.one = pk(p.one, p.one);
  k.bias = pk(0.0f, 0.0f);

  // this warp's place in the CTA's band: fixed for the whole kernel
  const unsigned r_local = warp / p.tiles_per_row;
  if (r_local >= (unsigned) p.rps) return;  // spare warps when the warp count is not a multiple of the tiles per row
  const unsigned tile0 = warp - r_local * p.tiles_per_row;
  const unsigned rows2 = p.rows2;
  const unsigned total_rp = (unsigned) p.nframes * rows2;
  const unsigned rpg0 = blockIdx.x * p.rps + r_local;
  if (rpg0 >= total_rp) return;
  const unsigned f0 = rpg0 / rows2, rp0 = rpg0 - f0 * rows2;
  const unsigned nsteps = (total_rp - rpg0 + p.step_rp - 1) / p.step_rp;  // >= 1

  for (unsigned tile = tile0; tile < (unsigned) p.tiles_per_row; tile += p.tile_stride) {
    const unsigned col4 = tile * 32 + lane;  // column in 4-pixel units
    const bool valid = FULL || (int) (col4 * 4) < p.w;
    const bool half = !FULL && (int) (col4 * 4 + 2) == p.w;  // w % 4 == 2: this lane owns the last two pixels
    // Two offset streams: the load stream runs PD steps ahead of the store stream.  Measured
    // (tools/probes/gran_bw.cu): with 1024 threads per SM the HBM pipe needs >= 24 B of loads in
    // flight per thread to reach 6 TB/s at this byte mix; 12 B (one step ahead) gave 4.4 TB/s.
    unsigned oy = f0 * p.sy4 + rp0 * 2 * p.py4 + col4;
    unsigned oc = f0 * p.sc4 + rp0 * p.pc4 + col4;
    unsigned oa = ALPHA ? f0 * p.sa4 + rp0 * 2 * p.pa4 + col4 : 0;
    unsigned oo = f0 * p.so16 + rp0 * 2 * p.po16 + col4;
    unsigned rpl = rp0, rps_ = rp0;  // row pair of the next load / of the next store
    constexpr int NB = PD + 1;
    TileLoad<ALPHA> buf[NB];
    auto advance_load = [&]() {
      rpl += p.d_rp;
      oy += p.dy; oc += p.dc;
      if (ALPHA) oa += p.da;
      if (WRAP && rpl >= rows2) { rpl -= rows2; oy += p.wy; oc += p.wc; if (ALPHA) oa += p.wa; }
    };
#pragma unroll
    for (int j = 0; j < PD; j++) {
      if ((unsigned) j < nsteps) {
        tile_load<ALPHA, FULL>(p, oy, oc, oa, rpl, valid, buf[j]);
        advance_load();
      }
    }
    unsigned kk = 0;
    while (true) {
#pragma unroll
      for (int j = 0; j < NB; j++) {
        if (kk + PD < nsteps) {
          tile_load<ALPHA, FULL>(p, oy, oc, oa, rpl, valid, buf[(j + PD) % NB]);
          advance_load();
        }
        tile_compute_store<CL2, RL2, ALPHA, FULL>(p, oo, rps_, valid, half, buf[j], k, clamp_lo, sbase);
        rps_ += p.d_rp;
        oo += p.dout;
        if (WRAP && rps_ >= rows2) { rps_ -= rows2; oo += p.wout; }
        if (++kk >= nsteps) goto tile_done;
      }
    }
  tile_done:;
  }
}

__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

// ================================================================= decode, cp.async warp ring
// Same arithmetic and CTA-band layout as k_decode_fast, but each warp stages its own inputs through
// a private shared-memory ring: per step ONE cp.async instruction (16 bytes per lane; lanes 0-7
// fetch the 128 B of Y row 0, lanes 8-15 Y row 1, lanes 16-23 the CbCr row) replaces three
// LDG.32 and their address arithmetic, and the bytes in flight live in shared memory instead of
// registers (D - 1 steps ahead).  Needs FULL geometry and 16-byte aligned Y / CbCr planes.
template <int CL2, int RL2, bool ALPHA, bool WRAP, int D, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) k_decode_ring(const DecodeArgs p) {
  extern __shared__ __align__(128) uint32_t s_lut[];
  constexpr int N = 1 << CL2, R = 1 << RL2;
  const int entries = CL2 ? N + 1 : p.log_cells;  // CL2 == 0: log-spaced cell LUT
  const uint32_t kLutBytes = (uint32_t) entries * R * 4;
  constexpr uint32_t kStage = ALPHA ? 640 : 384;  // Y0 | Y1 | CbCr (| A0 | A1), 128 B each
  fill_replicated<RL2>(s_lut, R, p.lut, entries);
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const unsigned warp = threadIdx.x >> 5;
  const uint32_t smem0 = (uint32_t) __cvta_generic_to_shared(s_lut);
  const uint32_t sbase = smem0 + ((lane & (R - 1)) << 2) - ((CL2 ? 0x4B000000u : p.log_first) << (2 + RL2));
  const uint32_t ring = smem0 + kLutBytes + warp * (D * kStage);
  const float clamp_lo = p.clamp_lo;
  DecodeK2 k;
  k.kY = pk(-(8388608.0f + 16.0f), -(8388608.0f + 16.0f));
  k.kC = pk(-(8388608.0f + 128.0f), -(8388608.0f + 128.0f));
  k.inv255 = pk(p.m.inv255, p.m.inv255);
  k.ys = pk(p.m.ys, p.m.ys);
  k.g_cb = pk(p.m.g_cb, p.m.g_cb);
  k.b_cb = pk(p.m.b_cb, p.m.b_cb);
  k.r_cr = pk(p.m.r_cr, p.m.r_cr);
  k.g_cr = pk(p.m.g_cr, p.m.g_cr);
  k.fN = pk((float) N, (float) N);
  k.magic = pk(8388608.0f, 8388608.0f);
  k.one = pk(p.one, p.one);
  k.bias = pk(0.0f, 0.0f);

  // this warp's first tile: band layout (fixed tile column, the row pair steps) or linear tile order (the
  // tile column steps and wraps into the next row pair; WRAP only, one tall image: see DecodeArgs.linear_tiles)
  unsigned tile, rows2, f0, rp0, nsteps, rpl;
  if (WRAP && p.linear_tiles) {
    const unsigned T = p.linear_tiles, total = (unsigned) p.rows2 * T, S = gridDim.x * (THREADS / 32);
    const unsigned t0 = blockIdx.x * (THREADS / 32) + warp;
    if (t0 >= total) return;
    rp0 = t0 / T;
    tile = t0 - rp0 * T;
    f0 = 0;
    rows2 = T;    // the wrap test below compares the stepped tile column with the tiles per row
    rpl = tile;
    nsteps = (total - t0 + S - 1) / S;
  } else {
    const unsigned r_local = warp / p.tiles_per_row;
    if (r_local >= (unsigned) p.rps) return;
    tile = warp - r_local * p.tiles_per_row;
    rows2 = p.rows2;
    const unsigned total_rp = (unsigned) p.nframes * rows2;
    const unsigned rpg0 = blockIdx.x * p.rps + r_local;
    if (rpg0 >= total_rp) return;
    f0 = rpg0 / rows2;
    rp0 = rpg0 - f0 * rows2;
    rpl = rp0;
    nsteps = (total_rp - rpg0 + p.step_rp - 1) / p.step_rp;
  }

  // ---- copy role of this lane: which plane row it fetches, as base + 32-bit offset in 4-byte units
  const int role = lane >> 3, sub = lane & 7;  // 0: Y row 0, 1: Y row 1, 2: CbCr, 3: idle (alpha: second pass)
  const uint8_t *cbase = role < 2 ? p.y : p.c;
  unsigned coff = role < 2 ? f0 * p.sy4 + (rp0 * 2 + role) * p.py4 : f0 * p.sc4 + rp0 * p.pc4;
  coff += tile * 32 + sub * 4;
  const unsigned cinc = role < 2 ? p.dy : p.dc, cwrap = role < 2 ? p.wy : p.wc;
  const uint32_t cdst = role * 128 + sub * 16;
  // alpha rows (ALPHA only): lanes 0-7 row 0, lanes 8-15 row 1, via a second cp.async
  unsigned aoff = ALPHA ? f0 * p.sa4 + (rp0 * 2 + (role & 1)) * p.pa4 + tile * 32 + sub * 4 : 0;
  const uint32_t adst = 384 + (role & 1) * 128 + sub * 16;

  auto issue = [&](uint32_t slot) {
    if (role < 3)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ring + slot * kStage + cdst),
                   "l"(ptr_mad(cbase, coff, 4))
                   : "memory");
    if (ALPHA && role < 2)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ring + slot * kStage + adst),
                   "l"(ptr_mad(p.a, aoff, 4))
                   : "memory");
    // step this lane's source to the warp's row pair in the CTA's next band
    rpl += p.d_rp;
    coff += cinc;
    if (ALPHA) aoff += p.da;
    if (WRAP && rpl >= rows2) { rpl -= rows2; coff += cwrap; if (ALPHA) aoff += p.wa; }
  };

  const unsigned col4 = tile * 32 + lane;
  unsigned oo = f0 * p.so16 + rp0 * 2 * p.po16 + col4;
  unsigned rps_ = rpl;  // store-side copy of the wrap counter (row pair, or tile column in linear order)
#pragma unroll
  for (int j = 0; j < D - 1; j++) {
    if ((unsigned) j < nsteps) issue(j);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  uint32_t slot = 0, fill = D - 1;
  for (unsigned kk = 0; kk < nsteps; kk++) {
    if (kk + D - 1 < nsteps) issue(fill);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group %0;" ::"n"(D - 1) : "memory");
    __syncwarp();
    const uint32_t st = ring + slot * kStage + lane * 4;
    TileLoad<ALPHA> t;
    t.y0 = lds_u32(st);
    t.y1 = lds_u32(st + 128);
    t.c = lds_u32(st + 256);
    if (ALPHA) {
      t.a0 = lds_u32(st + 384);
      t.a1 = lds_u32(st + 512);
    }
    __syncwarp();  // every lane has its words before this slot is refilled
    tile_compute_store<CL2, RL2, ALPHA, true>(p, oo, rps_, true, false, t, k, clamp_lo, sbase);
    rps_ += p.d_rp;
    oo += p.dout;
    if (WRAP && rps_ >= rows2) { rps_ -= rows2; oo += p.wout; }
    slot = slot + 1 == D ? 0 : slot + 1;
    fill = fill + 1 == D ? 0 : fill + 1;
  }
}

// ================================================================= decode, linear tile stream
// The kernel for contiguous batches of full tiles (w % 128 == 0, even h, frames back to back): the batch is one
// tall image, its tiles are numbered row-major and warp g of the grid takes tiles g, g + S, g + 2S, ...  S (the
// number of participating warps) is rounded DOWN to a multiple of the tiles per row T, so a warp keeps its tile
// column for the whole kernel and moves S / T row pairs down per step: every pointer advances by a per-lane
// constant and there is no wrap test, no frame carry and no 32-bit offset limit (at most T - 1 < 32 of the 4736
// warps idle, against 2 of every 32 in the band layout of 1920- or 3840-wide rows).
// Same per-warp cp.async ring as k_decode_ring, with the loop overhead cut to the bone (the ring kernel spends
// ~40 of its ~246 instructions per step on addressing, wrap logic and ring bookkeeping):
//   * the loop is unrolled over the D ring slots, so slot addresses are immediates;
//   * ONE __syncwarp per step: the refill of a slot is issued after the barrier of the NEXT step, by which time
//     every lane has read it;
//   * all 32 lanes issue the copy; the 8 lanes without a row use a source size of 0 (zero fill into a pad area
//     of the stage), so the instruction needs no predicate;
//   * running 64-bit pointers (two adds per step each) instead of offset + wrap + LEA pairs.
// COMPACT: no pad area in the stage (the 209 KB log-cell LUT of the Linear mode leaves room for 30 compact
// ring slices); the idle lanes are then predicated off instead.
// RAGGED: the width is even but not a multiple of 128 (720, 960, 1440 ...): the last tile column is partial.  A warp keeps
// its tile column, so what changes is per-lane and loop-invariant: the source size of the lane's 16-byte copy (0..16
// bytes, the rest is zero-filled) and whether the lane stores 16, 8 (w % 4 == 2) or no bytes per row.
// STRIDED: the frames of the batch are not back to back (e.g. one allocation per frame with the CbCr plane right behind
// the Y plane, what a hardware decoder hands over): the tile sequence still runs over (frame, row pair, tile column) in
// row-major order, a step is `lin_drp` row pairs = dq whole frames + dr row pairs further, and a pointer gets one extra
// per-lane constant whenever its row-pair-in-frame counter wraps (one compare and three predicated adds per step on
// the load side and on the store side).
template <int CL2, int RL2, bool ALPHA, int D, bool COMPACT, bool EARLY, bool RAGGED = false, bool STRIDED = false>
__global__ void __launch_bounds__(1024, 1) k_decode_lin(const DecodeArgs p) {
  extern __shared__ __align__(128) uint32_t s_lut[];
  constexpr int N = CL2 > 0 ? 1 << CL2 : 1, R = 1 << RL2;
  const int entries = CL2 > 0 ? N + 1 : p.log_cells;  // CL2 == 0: log cells, CL2 < 0: biased log cells
  const uint32_t kLutBytes = ((uint32_t) entries * R * 4 + 127u) & ~127u;
  // Y0 | Y1 | CbCr | pad (| A0 | A1 | pad | pad), 128 B each; COMPACT: without the pads
  constexpr uint32_t kStage = COMPACT ? (ALPHA ? 640 : 384) : (ALPHA ? 1024 : 512), kAlphaOff = COMPACT ? 384 : 512;
  // Programmatic dependent launch (small calls, launch_decode_lin1): let the next launch of the stream be scheduled
  // now, and wait here until the previous kernel of the stream has completed and its writes are visible.  Both are
  // no-ops for a launch without the attribute.  Nothing above this line touches global memory.
  pdl_launch_dependents();
  pdl_wait();

  const int lane = threadIdx.x & 31;
  const unsigned warp = threadIdx.x >> 5;
  const uint32_t smem0 = (uint32_t) __cvta_generic_to_shared(s_lut);
  const uint32_t sbase = smem0 + ((lane & (R - 1)) << 2) - ((CL2 > 0 ? 0x4B000000u : p.log_first) << (2 + RL2));
  const uint32_t ring = smem0 + kLutBytes + warp * (D * kStage);
  const float clamp_lo = p.clamp_lo;
  DecodeK2 k;
  k.kY = pk(-(8388608.0f + 16.0f), -(8388608.0f + 16.0f));
  k.kC = pk(-(8388608.0f + 128.0f), -(8388608.0f + 128.0f));
  k.inv255 = pk(p.m.inv255, p.m.inv255);
  k.ys = pk(p.m.ys, p.m.ys);
  k.g_cb = pk(p.m.g_cb, p.m.g_cb);
  k.b_cb = pk(p.m.b_cb, p.m.b_cb);
  k.r_cr = pk(p.m.r_cr, p.m.r_cr);
  k.g_cr = pk(p.m.g_cr, p.m.g_cr);
  k.fN = pk((float) N, (float) N);
  k.magic = pk(8388608.0f, 8388608.0f);
  k.one = pk(p.one, p.one);
  k.bias = pk(p.log_bias, p.log_bias);

  const unsigned g = blockIdx.x * (unsigned) p.lin_wpc + warp;
  const unsigned T = (unsigned) p.tiles_per_row, rows2 = (unsigned) p.rows2, drp = (unsigned) p.lin_drp;
  const unsigned rp0 = g / T, tile = g - rp0 * T;
  const unsigned total_rp = STRIDED ? (unsigned) p.nframes * rows2 : rows2;  // (contiguous: the launcher folded the frames into rows2)
  // (a warp without work still fills its share of the LUT and meets the CTA barrier below)
  const unsigned nsteps = g < (unsigned) p.lin_warps && rp0 < total_rp ? (total_rp - rp0 + drp - 1) / drp : 0u;
  // STRIDED: (frame, row pair in the frame) of the first tile; a step = dq frames + dr row pairs
  const unsigned f0 = STRIDED ? rp0 / rows2 : 0u, r0 = STRIDED ? rp0 - f0 * rows2 : rp0;
  const unsigned dq = STRIDED ? drp / rows2 : 0u, dr = STRIDED ? drp - dq * rows2 : drp;

  // ---- copy role of this lane: lanes 0-7 Y row 0, 8-15 Y row 1, 16-23 the CbCr row, 24-31 zero fill
  const int role = lane >> 3, sub = lane & 7;
  const uint8_t *src = role < 2 ? p.y + (long long) f0 * p.sy + ((long long) r0 * 2 + role) * p.py
                                : p.c + (long long) f0 * p.sc + (long long) r0 * p.pc;
  src += tile * 128 + sub * 16;
  const long long sinc = role < 2 ? (long long) dq * p.sy + (long long) dr * 2 * p.py : (long long) dq * p.sc + (long long) dr * p.pc;
  // extra step when the row pair wraps into the next frame
  const long long swrap = role < 2 ? p.sy - (long long) rows2 * 2 * p.py : p.sc - (long long) rows2 * p.pc;
  unsigned rl = r0;  // row pair in the frame of the next stage to be issued
  // bytes of this lane's 16-byte chunk that lie inside the row (Y and CbCr rows are both w bytes for an even w)
  const int chunk_left = RAGGED ? min(max(p.w - (int) (tile * 128 + sub * 16), 0), 16) : 16;
  const uint32_t ssize = role < 3 ? (uint32_t) chunk_left : 0u;
  const uint32_t cdst = ring + role * 128 + sub * 16;
  // alpha stream (ALPHA only): lanes 0-7 row 0, 8-15 row 1, the rest zero fill
  const uint8_t *asrc = nullptr;
  long long ainc = 0;
  uint32_t asize = 0;
  long long awrap = 0;
  if (ALPHA) {
    asrc = p.a + (long long) f0 * p.sa + ((long long) r0 * 2 + (role & 1)) * p.pa + tile * 128 + sub * 16;
    ainc = (long long) dq * p.sa + (long long) dr * 2 * p.pa;
    awrap = p.sa - (long long) rows2 * 2 * p.pa;
    asize = role < 2 ? (uint32_t) chunk_left : 0u;
  }
  auto issue = [&](int slot) {
    if (COMPACT) {
      if (role < 3) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(cdst + slot * kStage), "l"(src) : "memory");
    } else {
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(cdst + slot * kStage), "l"(src), "r"(ssize) : "memory");
    }
    src += sinc;
    if (ALPHA) {
      if (COMPACT) {
        if (role < 2)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(cdst + kAlphaOff + slot * kStage), "l"(asrc) : "memory");
      } else {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(cdst + kAlphaOff + slot * kStage), "l"(asrc),
                     "r"(asize)
                     : "memory");
      }
      asrc += ainc;
    }
    if (STRIDED) {
      rl += dr;
      if (rl >= rows2) {
        rl -= rows2;
        src += swrap;
        if (ALPHA) asrc += awrap;
      }
    }
  };

  uint8_t *out0 = p.out + (long long) f0 * p.so + (long long) r0 * 2 * p.po + (size_t) (tile * 32 + lane) * 16;
  const long long po = p.po, oinc = (long long) dq * p.so + (long long) dr * 2 * p.po;
  const long long owrap = p.so - (long long) rows2 * 2 * p.po;
  unsigned ro = r0;  // row pair in the frame of the tile being stored
  const int px_left = RAGGED ? p.w - (int) (tile * 32 + lane) * 4 : 4;  // pixels of this lane inside the row (even)
#pragma unroll
  for (int j = 0; j < D - 1; j++) {
    if ((unsigned) j < nsteps) issue(j);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  // the LUT fill comes AFTER the first stages are on their way: their DRAM latency hides behind it (a call of one
  // 1080p frame is two steps per warp, so the fill and the first load are a third of its run time)
  fill_replicated<RL2>(s_lut, R, p.lut, entries);
  __syncthreads();
  const uint32_t st0 = ring + lane * 4;
  for (unsigned kk = 0; kk < nsteps; kk += D) {
#pragma unroll
    for (int s = 0; s < D; s++) {
      if (kk + s >= nsteps) break;  // warp-uniform
      if (EARLY) {
        // the refill of the slot read one step ago is issued BEFORE waiting for this step's stage: two stages are in
        // flight while the warp waits, one while it computes (measured: the access front of depth 3 with late issue
        // spreads too far at full clock, 978 against 1017 Gpix/s in a 20-step burst)
        if (kk + s + D - 1 < nsteps) issue((s + D - 1) % D);
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group %0;" ::"n"(D - 1) : "memory");
        __syncwarp();
      } else {
        asm volatile("cp.async.wait_group %0;" ::"n"(D - 2) : "memory");
        __syncwarp();  // the step's stage is complete for every lane, and every lane is done reading the previous one
        if (kk + s + D - 1 < nsteps) issue((s + D - 1) % D);
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
      const uint32_t st = st0 + s * kStage;
      TileLoad<ALPHA> t;
      t.y0 = lds_u32(st);
      t.y1 = lds_u32(st + 128);
      t.c = lds_u32(st + 256);
      if (ALPHA) {
        t.a0 = lds_u32(st + kAlphaOff);
        t.a1 = lds_u32(st + kAlphaOff + 128);
      }
      if (EARLY) __syncwarp();  // every lane has its words before the next step refills this slot
      f32x2 cbn = mul2(add2(pku(magic_byte<0>(t.c), magic_byte<2>(t.c)), k.kC), k.inv255);
      f32x2 crn = mul2(add2(pku(magic_byte<1>(t.c), magic_byte<3>(t.c)), k.kC), k.inv255);
      f32x2 cb_g = mul2(cbn, k.g_cb), cb_b = mul2(cbn, k.b_cb);
      f32x2 cr_r = mul2(crn, k.r_cr), cr_g = mul2(crn, k.g_cr);
      const uint4 v0 = decode_row<CL2, RL2, ALPHA>(t.y0, t.a0, cb_g, cb_b, cr_r, cr_g, k, clamp_lo, sbase);
      const uint4 v1 = decode_row<CL2, RL2, ALPHA>(t.y1, t.a1, cb_g, cb_b, cr_r, cr_g, k, clamp_lo, sbase);
      if (!RAGGED || px_left >= 4) {
        st_stream_u128(out0, v0);
        st_stream_u128(out0 + po, v1);
      } else if (px_left == 2) {
        st_stream_u64(out0, make_uint2(v0.x, v0.y));
        st_stream_u64(out0 + po, make_uint2(v1.x, v1.y));
      }
      out0 += oinc;
      if (STRIDED) {
        ro += dr;
        if (ro >= rows2) {
          ro -= rows2;
          out0 += owrap;
        }
      }
    }
  }
}

// ===================================================================== decode, bulk-staged
// Same arithmetic and the same CTA-band work layout as k_decode_fast, but the input planes reach
// the SM through the bulk async-copy engine (cp.async.bulk -> UBLKCP, the 1-D TMA path) into a
// shared-memory ring instead of through per-thread LDG into registers.
//
// Why: with per-thread loads the bytes in flight per SM are bounded by registers
// (32 warps x 3 x 128 B = 12 KB), and the measured throughput was the same 4.46 TB/s with the
// arithmetic compiled out - a latency (Little's law) bound, not an issue bound.  A ring of up to 8
// stages x 11.5 KB keeps ~90 KB of reads in flight per SM, costs no registers, and removes every
// load address computation from the consumer warps.
//
//   stage  = one CTA band: 2*rps rows of Y and rps rows of CbCr (and 2*rps rows of alpha),
//            each row one bulk copy of w bytes, issued by lane 0 of warp j for row j
//   full[s]   mbarrier, count = rows per stage, completes when all bytes of the stage have landed
//   empty[s]  mbarrier, count = consumer warps, completes when every warp has read its words
// Requires w % 128 == 0, h even, 16-byte aligned planes and pitches (else k_decode_fast).
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

template <int CL2, int RL2, bool ALPHA, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) k_decode_bulk(const DecodeArgs p) {
  extern __shared__ __align__(128) uint32_t s_lut[];
  constexpr int N = 1 << CL2, R = 1 << RL2;
  constexpr uint32_t kLutBytes = (uint32_t) (N + 1) * R * 4;
  fill_replicated<RL2>(s_lut, R, p.lut, N + 1);

  const uint32_t smem0 = (uint32_t) __cvta_generic_to_shared(s_lut);
  const uint32_t ring0 = smem0 + kLutBytes;
  const uint32_t S = p.nstages;
  const uint32_t bar_full = ring0 + S * p.stage_bytes, bar_empty = bar_full + 8 * S;
  const unsigned nwarps_active = p.rps * p.tiles_per_row;
  if (threadIdx.x == 0) {
    for (uint32_t i = 0; i < S; i++) {
      mbar_init(bar_full + 8 * i, p.nrows);
      mbar_init(bar_empty + 8 * i, nwarps_active);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const unsigned warp = threadIdx.x >> 5;
  const uint32_t sbase = smem0 + ((lane & (R - 1)) << 2) - (0x4B000000u << (2 + RL2));
  const float clamp_lo = p.clamp_lo;
  DecodeK2 k;
  k.kY = pk(-(8388608.0f + 16.0f), -(8388608.0f + 16.0f));
  k.kC = pk(-(8388608.0f + 128.0f), -(8388608.0f + 128.0f));
  k.inv255 = pk(p.m.inv255, p.m.inv255);
  k.ys = pk(p.m.ys, p.m.ys);
  k.g_cb = pk(p.m.g_cb, p.m.g_cb);
  k.b_cb = pk(p.m.b_cb, p.m.b_cb);
  k.r_cr = pk(p.m.r_cr, p.m.r_cr);
  k.g_cr = pk(p.m.g_cr, p.m.g_cr);
  k.fN = pk((float) N, (float) N);
  k.magic = pk(8388608.0f, 8388608.0f);
  k.one = pk(p.one, p.one);
  k.bias = pk(0.0f, 0.0f);

  if (warp >= nwarps_active) return;  // spare warps
  const unsigned rows2 = p.rows2, rps = p.rps, w = p.w;
  const unsigned total_rp = (unsigned) p.nframes * rows2;
  const unsigned nbands = (total_rp + rps - 1) / rps;
  if (blockIdx.x >= nbands) return;
  const unsigned nsteps = (nbands - blockIdx.x + gridDim.x - 1) / gridDim.x;
  const unsigned r_local = warp / p.tiles_per_row;
  const unsigned tile = warp - r_local * p.tiles_per_row;

  // ---- producer role: lane 0 of warp j feeds row j of every stage
  const bool producer = lane == 0 && warp < (unsigned) p.nrows;
  unsigned prow_rr = 0, prow_dst = 0;          // row pair within the band; byte offset within a stage
  const uint8_t *pbase = nullptr;              // plane base
  unsigned pstride = 0, ppitch = 0, prow_mul = 0, prow_add = 0;  // frame stride; pitch; row = rp*mul + add
  if (producer) {
    const unsigned j = warp;
    if (j < 2 * rps) {
      prow_rr = j >> 1; pbase = p.y; pstride = (unsigned) p.sy; ppitch = (unsigned) p.py; prow_mul = 2; prow_add = j & 1;
    } else if (j < 3 * rps) {
      prow_rr = j - 2 * rps; pbase = p.c; pstride = (unsigned) p.sc; ppitch = (unsigned) p.pc; prow_mul = 1; prow_add = 0;
    } else {
      const unsigned q = j - 3 * rps;
      prow_rr = q >> 1; pbase = p.a; pstride = (unsigned) p.sa; ppitch = (unsigned) p.pa; prow_mul = 2; prow_add = q & 1;
    }
    prow_dst = j * w;
  }
  auto produce = [&](unsigned kp) {
    if (!producer || kp >= nsteps) return;
    const uint32_t slot = kp % S;
    if (kp >= S) mbar_wait(bar_empty + 8 * slot, ((kp / S) - 1) & 1);
    const unsigned rpg = (blockIdx.x + kp * gridDim.x) * rps + prow_rr;
    const uint32_t bar = bar_full + 8 * slot;
    if (rpg < total_rp) {
      const unsigned f = rpg / rows2, rp = rpg - f * rows2;
      const uint8_t *src = ptr_mad(ptr_mad(pbase, f, pstride), rp * prow_mul + prow_add, ppitch);
      mbar_expect_tx(bar, w);
      bulk_g2s(ring0 + slot * p.stage_bytes + prow_dst, src, w, bar);
    } else {
      mbar_arrive(bar);  // band runs past the last row pair: nothing to copy for this row
    }
  };

  // ---- consumer role
  const unsigned col4 = tile * 32 + lane;
  const uint32_t my_y = 2 * r_local * w + col4 * 4;              // byte offsets within a stage
  const uint32_t my_c = (2 * rps + r_local) * w + col4 * 4;
  const uint32_t my_a = (3 * rps + 2 * r_local) * w + col4 * 4;
  unsigned rpg = blockIdx.x * rps + r_local;
  unsigned f0 = rpg / rows2, rp = rpg - f0 * rows2;
  unsigned oo = f0 * p.so16 + rp * 2 * p.po16 + col4;

  for (unsigned kp = 0; kp + 1 < S; kp++) produce(kp);
  for (unsigned kk = 0; kk < nsteps; kk++) {
    produce(kk + S - 1);
    const uint32_t slot = kk % S;
    mbar_wait(bar_full + 8 * slot, (kk / S) & 1);
    const uint32_t st = ring0 + slot * p.stage_bytes;
    TileLoad<ALPHA> t;
    t.y0 = lds_u32(st + my_y);
    t.y1 = lds_u32(st + my_y + w);
    t.c = lds_u32(st + my_c);
    if (ALPHA) {
      t.a0 = lds_u32(st + my_a);
      t.a1 = lds_u32(st + my_a + w);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_empty + 8 * slot);
    if (rpg < total_rp) tile_compute_store<CL2, RL2, ALPHA, true>(p, oo, rp, true, false, t, k, clamp_lo, sbase);
    // advance to this warp's row pair in the CTA's next band
    rpg += p.step_rp;
    rp += p.d_rp;
    oo += p.dout;
    if (rp >= rows2) { rp -= rows2; oo += p.wout; }
  }
}

// ========================================================================= decode, generic
// Any width/height/pitch/alignment (odd sizes, unaligned planes): one thread per 2x2 block,
// byte accesses, LUT read through the read-only path.  Same arithmetic, same results.
template <bool ALPHA>
__global__ void __launch_bounds__(256) k_decode_generic(const DecodeArgs p) {
  const int cw = (p.w + 1) >> 1;
  const long long blocks = (long long) p.nframes * p.rows2 * cw;
  const float fN = (float) (1 << p.cells_log2);
  const float kY = 16.0f, kC = 128.0f;
  // one 32-bit store per pixel whenever the output plane allows it (it almost always does)
  const bool word_stores = (((unsigned long long) p.out | (unsigned long long) p.po | (unsigned long long) p.so) & 3ull) == 0;
  for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < blocks;
       i += (long long) gridDim.x * blockDim.x) {
    const int bx = (int) (i % cw);
    const long long t = i / cw;
    const int rp = (int) (t % p.rows2);
    const long long f = t / p.rows2;
    const uint8_t *cp = p.c + f * p.sc + (long long) rp * p.pc + bx * 2;
    float cbn = __fmul_rn((float) cp[0] - kC, p.m.inv255);
    float crn = __fmul_rn((float) cp[1] - kC, p.m.inv255);
    float cb_g = __fmul_rn(cbn, p.m.g_cb), cb_b = __fmul_rn(cbn, p.m.b_cb);
    float cr_r = __fmul_rn(crn, p.m.r_cr), cr_g = __fmul_rn(crn, p.m.g_cr);
    for (int dy = 0; dy < 2; dy++) {
      const int row = rp * 2 + dy;
      if (row >= p.h) break;
      for (int dx = 0; dx < 2; dx++) {
        const int col = bx * 2 + dx;
        if (col >= p.w) break;
        float yf = (float) p.y[f * p.sy + (long long) row * p.py + col] - kY;
        float yt = __fmul_rn(__fmul_rn(yf, p.m.inv255), p.m.ys);
        float xr = add_sat(yt, cr_r);
        float xg = add_sat(__fadd_rn(yt, cb_g), cr_g);
        float xb = add_sat(yt, cb_b);
        uint32_t R = lut_code(p.lut, xr, p.clamp_lo, fN);
        uint32_t G = lut_code(p.lut, xg, p.clamp_lo, fN);
        uint32_t B = lut_code(p.lut, xb, p.clamp_lo, fN);
        uint32_t A = 255;
        if (ALPHA) {
          float af = (float) p.a[f * p.sa + (long long) row * p.pa + col] - kY;
          float an = add_sat(__fmul_rn(__fmul_rn(af, p.m.inv255), p.m.ys), 0.0f);
          A = (uint32_t) round_pos(__fmul_rn(an, 255.0f));
        }
        uint8_t *o = p.out + f * p.so + (long long) row * p.po + (long long) col * 4;
        if (word_stores) {
          *reinterpret_cast<uint32_t *>(o) = B | (G << 8) | (R << 16) | (A << 24);
        } else {
          o[0] = (uint8_t) B; o[1] = (uint8_t) G; o[2] = (uint8_t) R; o[3] = (uint8_t) A;
        }
      }
    }
  }
}

#ifndef BT709_DECODE_THREADS
#define BT709_DECODE_THREADS 1024
#endif
constexpr int kDecodeThreads = BT709_DECODE_THREADS;
#ifndef BT709_PREFETCH_DEPTH
#define BT709_PREFETCH_DEPTH 3
#endif
constexpr int kPrefetchDepth = BT709_PREFETCH_DEPTH;  // steps the loads run ahead of the math
#ifndef BT709_RING_DEPTH
#define BT709_RING_DEPTH 2
#endif
constexpr int kRingDepth = BT709_RING_DEPTH;  // stages of the per-warp cp.async ring

template <int CL2, int RL2, bool ALPHA, bool FULL, bool WRAP>
static int launch_decode_fast2(const DecodeArgs &a0, int sm_count, cudaStream_t s) {
  const size_t smem = (size_t) (CL2 ? (1 << CL2) + 1 : a0.log_cells) * (1 << RL2) * 4;
  auto k = k_decode_fast<CL2, RL2, ALPHA, FULL, WRAP, kPrefetchDepth, kDecodeThreads>;
  auto kring = k_decode_ring<CL2, RL2, ALPHA, WRAP, kRingDepth, kDecodeThreads>;
  // ring slices exist for the warps that have a tile only.  The fully replicated log-cell LUT (209 KB)
  // leaves room for 30 of them (1920 or 3840 wide, no alpha); otherwise that mode prefetches in registers
  const int wpc = kDecodeThreads / 32;
  const bool linear = WRAP && a0.linear_tiles != 0;  // linear tile order: every warp has a tile
  const int active_warps = (linear || a0.tiles_per_row >= wpc) ? wpc : a0.tiles_per_row * a0.rps;
  const size_t smem_ring = smem + (size_t) active_warps * kRingDepth * (ALPHA ? 640 : 384);
  static const bool no_ring = getenv("BT709_NO_RING") != nullptr;
  // (the ring kernel gives every warp one fixed tile column: rows wider than 32 tiles take the register path)
  const bool use_ring = smem_ring <= kMaxDynSmem && a0.tiles_per_row <= wpc && FULL && !no_ring && aligned(a0.y, 16) && aligned(a0.py, 16) && aligned(a0.sy, 16) &&
                        aligned(a0.c, 16) && aligned(a0.pc, 16) && aligned(a0.sc, 16) &&
                        (!ALPHA || (aligned(a0.a, 16) && aligned(a0.pa, 16) && aligned(a0.sa, 16)));
  // frames per launch such that every 32-bit unit offset stays below 2^32
  const unsigned long long lim = 0xFFFF0000ull;
  unsigned long long per = lim;
  auto fit = [&](long long stride, unsigned unit, long long frame_bytes) {
    if (stride <= 0) return;
    unsigned long long n = (lim * unit - (unsigned long long) frame_bytes) / (unsigned long long) stride + 1;
    if (n < per) per = n;
  };
  fit(a0.sy, 4, (long long) a0.h * a0.py);
  fit(a0.sc, 4, (long long) a0.rows2 * a0.pc);
  fit(a0.so, 16, (long long) a0.h * a0.po);
  if (ALPHA) fit(a0.sa, 4, (long long) a0.h * a0.pa);
  if (per < 1) per = 1;
  if (a0.nframes == 1) {
    // one (possibly very tall) image: split by rows instead, in chunks whose offsets fit in 32 bits
    const unsigned long long out_bytes = (unsigned long long) a0.h * a0.po;
    if (out_bytes > lim * 16 || (unsigned long long) a0.h * a0.py > lim * 4) {
      long long rows_per = (long long) ((lim * 16) / (unsigned long long) a0.po);
      if ((unsigned long long) rows_per * a0.py > lim * 4) rows_per = (long long) ((lim * 4) / (unsigned long long) a0.py);
      rows_per &= ~1ll;
      for (long long r0 = 0; r0 < a0.h; r0 += rows_per) {
        DecodeArgs a = a0;
        a.h = (int) (a0.h - r0 < rows_per ? a0.h - r0 : rows_per);
        a.rows2 = (a.h + 1) / 2;
        a.y = a0.y + r0 * a0.py;
        a.c = a0.c + (r0 / 2) * a0.pc;
        a.out = a0.out + r0 * a0.po;
        if (ALPHA) a.a = a0.a + r0 * a0.pa;
        int e = launch_decode_fast2<CL2, RL2, ALPHA, FULL, WRAP>(a, sm_count, s);
        if (e) return e;
      }
      return 0;
    }
  }
  for (int f0 = 0; f0 < a0.nframes; f0 += (int) per) {
    DecodeArgs a = a0;
    a.nframes = (int) ((unsigned long long) (a0.nframes - f0) < per ? (unsigned long long) (a0.nframes - f0) : per);
    a.y = a0.y + (long long) f0 * a0.sy;
    a.c = a0.c + (long long) f0 * a0.sc;
    a.out = a0.out + (long long) f0 * a0.so;
    if (ALPHA) a.a = a0.a + (long long) f0 * a0.sa;
    const unsigned total_rp = (unsigned) a.nframes * a.rows2;
    unsigned grid = (total_rp + a.rps - 1) / a.rps;
    if (grid > (unsigned) sm_count) grid = sm_count;
    // per-step increments of the unit offsets, and the extra increment on a frame wrap (mod 2^32)
    a.step_rp = grid * a.rps;
    const unsigned d_f = a.step_rp / a.rows2;
    a.d_rp = a.step_rp - d_f * a.rows2;
    a.py4 = (unsigned) (a.py / 4); a.pc4 = (unsigned) (a.pc / 4); a.pa4 = (unsigned) (a.pa / 4); a.po16 = (unsigned) (a.po / 16);
    a.sy4 = (unsigned) (a.sy / 4); a.sc4 = (unsigned) (a.sc / 4); a.sa4 = (unsigned) (a.sa / 4); a.so16 = (unsigned) (a.so / 16);
    a.dy = d_f * a.sy4 + a.d_rp * 2 * a.py4;
    a.dc = d_f * a.sc4 + a.d_rp * a.pc4;
    a.da = d_f * a.sa4 + a.d_rp * 2 * a.pa4;
    a.dout = d_f * a.so16 + a.d_rp * 2 * a.po16;
    a.wy = a.sy4 - a.rows2 * 2 * a.py4;
    a.wc = a.sc4 - a.rows2 * a.pc4;
    a.wa = a.sa4 - a.rows2 * 2 * a.pa4;
    a.wout = a.so16 - a.rows2 * 2 * a.po16;
    if (linear && use_ring) {
      // one tall image, tiles taken in row-major order, 32 per CTA and step: the "wrap" is the end of a row
      const unsigned T = (unsigned) a.tiles_per_row, total_tiles = (unsigned) a.rows2 * T;
      grid = (total_tiles + wpc - 1) / wpc;
      if (grid > (unsigned) sm_count) grid = sm_count;
      const unsigned S = grid * wpc, drp = S / T, dcol = S - drp * T;
      a.d_rp = dcol;
      a.dy = drp * 2 * a.py4 + dcol * 32; a.dc = drp * a.pc4 + dcol * 32;
      a.da = drp * 2 * a.pa4 + dcol * 32; a.dout = drp * 2 * a.po16 + dcol * 32;
      a.wy = 2 * a.py4 - T * 32; a.wc = a.pc4 - T * 32; a.wa = 2 * a.pa4 - T * 32; a.wout = 2 * a.po16 - T * 32;
    } else {
      a.linear_tiles = 0;
    }
    if (use_ring)
      kring<<<grid, kDecodeThreads, smem_ring, s>>>(a), note_launch();
    else
      k<<<grid, kDecodeThreads, smem, s>>>(a), note_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int) e;
  }
  return 0;
}


template <int CL2, int RL2, bool ALPHA>
static bool bulk_ok(const DecodeArgs &a, int *nstages, int *stage_bytes, int *nrows) {
  constexpr size_t lut = (size_t) ((1 << CL2) + 1) * (1 << RL2) * 4;
  const int wpc = kDecodeThreads / 32;
  if (!(a.w % 128 == 0 && a.h % 2 == 0 && a.tiles_per_row <= wpc)) return false;
  if (!(aligned(a.y, 16) && aligned(a.py, 16) && aligned(a.sy, 16) && aligned(a.c, 16) && aligned(a.pc, 16) &&
        aligned(a.sc, 16)))
    return false;
  if (ALPHA && !(aligned(a.a, 16) && aligned(a.pa, 16) && aligned(a.sa, 16))) return false;
  *nrows = (ALPHA ? 5 : 3) * a.rps;
  if (*nrows > a.rps * a.tiles_per_row) return false;  // one producer lane per row
  *stage_bytes = *nrows * a.w;
  long long room = (long long) kMaxDynSmem - (long long) lut - 16 * 8 * 2;
  int n = (int) (room / *stage_bytes);
  if (n > 8) n = 8;
  if (n < 2) return false;
  *nstages = n;
  return true;
}

// ================================================================= decode, linear tile stream, register loads
// Same tile order and constant strides as k_decode_lin, but the three words of a tile come straight into registers
// (LDG.32 x 3) with the loads of the next PD tiles in flight, instead of through the cp.async ring.  The r2 probe
// tools/probes/row_bw.cu measured this pattern (no arithmetic) at 1050 Gpix/s with two units in flight against 1009
// for the ring.  No shared memory beyond the LUT, no __syncwarp.  Opaque only (the alpha variant keeps the ring).
template <int CL2, int RL2, int PD>
__global__ void __launch_bounds__(1024, 1) k_decode_reg(const DecodeArgs p) {
  extern __shared__ __align__(128) uint32_t s_lut[];
  constexpr int N = CL2 > 0 ? 1 << CL2 : 1, R = 1 << RL2;
  const int entries = CL2 > 0 ? N + 1 : p.log_cells;
  fill_replicated<RL2>(s_lut, R, p.lut, entries);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const unsigned warp = threadIdx.x >> 5;
  const uint32_t smem0 = (uint32_t) __cvta_generic_to_shared(s_lut);
  const uint32_t sbase = smem0 + ((lane & (R - 1)) << 2) - ((CL2 > 0 ? 0x4B000000u : p.log_first) << (2 + RL2));
  const float clamp_lo = p.clamp_lo;
  DecodeK2 k;
  k.kY = pk(-(8388608.0f + 16.0f), -(8388608.0f + 16.0f));
  k.kC = pk(-(8388608.0f + 128.0f), -(8388608.0f + 128.0f));
  k.inv255 = pk(p.m.inv255, p.m.inv255);
  k.ys = pk(p.m.ys, p.m.ys);
  k.g_cb = pk(p.m.g_cb, p.m.g_cb);
  k.b_cb = pk(p.m.b_cb, p.m.b_cb);
  k.r_cr = pk(p.m.r_cr, p.m.r_cr);
  k.g_cr = pk(p.m.g_cr, p.m.g_cr);
  k.fN = pk((float) N, (float) N);
  k.magic = pk(8388608.0f, 8388608.0f);
  k.one = pk(p.one, p.one);
  k.bias = pk(p.log_bias, p.log_bias);

  const unsigned g = blockIdx.x * (unsigned) p.lin_wpc + warp;
  const unsigned T = (unsigned) p.tiles_per_row, rows2 = (unsigned) p.rows2, drp = (unsigned) p.lin_drp;
  const unsigned rp0 = g / T, tile = g - rp0 * T;
  // (a warp without work still fills its share of the LUT and meets the CTA barrier below)
  const unsigned nsteps = g < (unsigned) p.lin_warps && rp0 < rows2 ? (rows2 - rp0 + drp - 1) / drp : 0u;
  const uint8_t *yp = p.y + (long long) rp0 * 2 * p.py + (size_t) (tile * 32 + lane) * 4;
  const uint8_t *cp = p.c + (long long) rp0 * p.pc + (size_t) (tile * 32 + lane) * 4;
  const long long py = p.py, yinc = (long long) drp * 2 * p.py, cinc = (long long) drp * p.pc;
  uint8_t *out0 = p.out + (long long) rp0 * 2 * p.po + (size_t) (tile * 32 + lane) * 16;
  const long long po = p.po, oinc = (long long) drp * 2 * p.po;
  constexpr int NB = PD + 1;
  TileLoad<false> buf[NB];
  auto load = [&](TileLoad<false> &t) {
    t.y0 = ld_stream_u32(yp);
    t.y1 = ld_stream_u32(yp + py);
    t.c = ld_stream_u32(cp);
    yp += yinc;
    cp += cinc;
  };
#pragma unroll
  for (int j = 0; j < PD; j++)
    if ((unsigned) j < nsteps) load(buf[j]);
  for (unsigned kk = 0; kk < nsteps; kk += NB) {
#pragma unroll
    for (int s = 0; s < NB; s++) {
      if (kk + s >= nsteps) break;  // warp-uniform
      if (kk + s + PD < nsteps) load(buf[(s + PD) % NB]);
      const TileLoad<false> &t = buf[s];
      f32x2 cbn = mul2(add2(pku(magic_byte<0>(t.c), magic_byte<2>(t.c)), k.kC), k.inv255);
      f32x2 crn = mul2(add2(pku(magic_byte<1>(t.c), magic_byte<3>(t.c)), k.kC), k.inv255);
      f32x2 cb_g = mul2(cbn, k.g_cb), cb_b = mul2(cbn, k.b_cb);
      f32x2 cr_r = mul2(crn, k.r_cr), cr_g = mul2(crn, k.g_cr);
      st_stream_u128(out0, decode_row<CL2, RL2, false>(t.y0, 0u, cb_g, cb_b, cr_r, cr_g, k, clamp_lo, sbase));
      st_stream_u128(out0 + po, decode_row<CL2, RL2, false>(t.y1, 0u, cb_g, cb_b, cr_r, cr_g, k, clamp_lo, sbase));
      out0 += oinc;
    }
  }
}

// Ring of k_decode_lin: depth 2 with the refill issued BEFORE the wait ("2e": two stages in flight while the warp
// waits, one while it computes) or depth 3 with the refill issued after it ("3l": two stages in flight throughout, one
// __syncwarp less).  Measured on one box, 600 x 1080p, BT.709 curve, 20 / 100 / 1000 steps (Gpix/s):
//   2e 1016 / 953 / 899     3l 994 / 973 / 918     round-1 ring kernel 1017 / 953 / 900
// i.e. the deeper late-issue ring is 2 % slower in a burst at full clock (the access front spreads) and 2 % faster once
// the power cap has pulled the clocks down.  The Linear mode and the alpha variant are faster with 3l either way
// (Linear 1012 vs 966 in a burst).  BT709_RING=2e|3l overrides the choice.
// k_decode_lin for a contiguous batch of full tiles (decoded as one tall image).  Returns -1 when it does not apply.
template <int CL2, int RL2, bool ALPHA, bool COMPACT, int D, bool EARLY, bool RAGGED = false, bool STRIDED = false>
static int launch_decode_lin1(const DecodeArgs &a0, int sm_count, cudaStream_t s, bool strided = false) {
  if constexpr (!RAGGED && !COMPACT) {  // partial last tile column: the same kernel with per-lane copy sizes and store widths
    if (a0.w % 128 != 0) return launch_decode_lin1<CL2, RL2, ALPHA, COMPACT, D, EARLY, true, STRIDED>(a0, sm_count, s, strided);
  }
  if (!RAGGED && a0.w % 128 != 0) return -1;  // (the compact stage has no zero-fill form)
  if constexpr (!STRIDED && !COMPACT && RL2 == 5) {  // frames that are not back to back: the wrap-aware instantiation (batches only)
    if (strided) return launch_decode_lin1<CL2, RL2, ALPHA, COMPACT, D, EARLY, RAGGED, true>(a0, sm_count, s, true);
  }
  if (!STRIDED && strided) return -1;
  DecodeArgs a = a0;
  if (!STRIDED) {
    a.h *= a.nframes;
    a.rows2 *= a.nframes;
    a.nframes = 1;
  }
  const size_t lut_bytes = (((size_t) (CL2 > 0 ? (1 << CL2) + 1 : a.log_cells) * (1 << RL2) * 4) + 127) & ~(size_t) 127;
  if (lut_bytes >= kMaxDynSmem) return -1;
  const size_t slice = (size_t) D * (COMPACT ? (ALPHA ? 640 : 384) : (ALPHA ? 1024 : 512));
  int wpc = (int) ((kMaxDynSmem - lut_bytes) / slice);
  if (wpc > 32) wpc = 32;
  if (wpc < 8) return -1;
  const unsigned T = (unsigned) a.tiles_per_row;
  const unsigned long long total_tiles = (unsigned long long) a.nframes * a.rows2 * T;
  unsigned long long grid = (total_tiles + wpc - 1) / wpc;
  if (grid > (unsigned long long) sm_count) grid = sm_count;
  const unsigned W = (unsigned) grid * wpc;
  if (W < T) return -1;
  a.lin_wpc = wpc;
  a.lin_warps = (int) (W / T * T);
  a.lin_drp = a.lin_warps / (int) T;
  // Small calls (the few-replica LUT instantiations, RL2 < 5) are launched with programmatic stream serialization:
  // back-to-back calls on one stream (the reference's one-frame-per-call pattern) overlap this launch's scheduling
  // with the previous kernel's tail; the kernel waits (griddepcontrol.wait) before it touches global memory.
  // BT709_PDL=0 disables it, BT709_PDL=all applies it to every launch that supports it (pdl_enabled).
  if constexpr (RAGGED || STRIDED) {  // opt-in to the dynamic shared memory here: the plain instantiations are set up at init
    static const int attr_rc = (int) cudaFuncSetAttribute(k_decode_lin<CL2, RL2, ALPHA, D, COMPACT, EARLY, RAGGED, STRIDED>,
                                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kMaxDynSmem);
    if (attr_rc) return attr_rc;
  }
  const cudaError_t le = launch_pdl(k_decode_lin<CL2, RL2, ALPHA, D, COMPACT, EARLY, RAGGED, STRIDED>, (unsigned) grid, (unsigned) wpc * 32,
                                    lut_bytes + (size_t) wpc * slice, s, pdl_enabled(RL2 < 5), a);
  note_launch();
  if (le != cudaSuccess) return (int) le;
  return (int) cudaGetLastError();
}

// ring choice: BT709_RING=2e|3l, else `late3_default` (see the table above)
template <int CL2, int RL2, bool ALPHA, bool COMPACT>
static int launch_decode_lin(const DecodeArgs &a, int sm_count, cudaStream_t s, bool late3_default, bool strided = false) {
  static const char *env = getenv("BT709_RING");
  static const char *reg_env = getenv("BT709_REG");  // experiment: register loads, 1-3 tiles ahead (opaque, replicated LUT)
  if constexpr (!ALPHA && !COMPACT && RL2 == 5) {
    if (reg_env && a.w % 128 == 0 && !strided) {
      DecodeArgs b = a;
      b.h *= b.nframes; b.rows2 *= b.nframes; b.nframes = 1;
      const size_t lut_bytes = (((size_t) (CL2 > 0 ? (1 << (CL2 > 0 ? CL2 : 0)) + 1 : b.log_cells) * 32 * 4) + 127) & ~(size_t) 127;
      const unsigned T = (unsigned) b.tiles_per_row;
      unsigned long long grid = ((unsigned long long) b.rows2 * T + 31) / 32;
      if (grid > (unsigned long long) sm_count) grid = sm_count;
      const unsigned W = (unsigned) grid * 32;
      if (W >= T) {
        b.lin_wpc = 32; b.lin_warps = (int) (W / T * T); b.lin_drp = b.lin_warps / (int) T;
        const int pd = atoi(reg_env);
        if (pd == 1) k_decode_reg<CL2, RL2, 1><<<(unsigned) grid, 1024, lut_bytes, s>>>(b), note_launch();
        else if (pd == 3) k_decode_reg<CL2, RL2, 3><<<(unsigned) grid, 1024, lut_bytes, s>>>(b), note_launch();
        else k_decode_reg<CL2, RL2, 2><<<(unsigned) grid, 1024, lut_bytes, s>>>(b), note_launch();
        return (int) cudaGetLastError();
      }
    }
  }
  const bool late3 = env ? (env[0] == '3') : late3_default;
  if (late3) return launch_decode_lin1<CL2, RL2, ALPHA, COMPACT, 3, false>(a, sm_count, s, strided);
  return launch_decode_lin1<CL2, RL2, ALPHA, COMPACT, 2, true>(a, sm_count, s, strided);
}

// Picks the k_decode_lin instantiation: per-lane replicated LUT for batches, a few replicas (a much shorter
// table fill per CTA, some bank conflicts) when the whole call is only a few steps per warp.
template <bool ALPHA>
static int try_decode_lin(const DecodeArgs &a, int sm_count, cudaStream_t s) {
  static const bool off = getenv("BT709_NO_LIN2") != nullptr;
  static const char *small_env = getenv("BT709_SMALL_STEPS");
  static const int small_steps = small_env ? atoi(small_env) : 4;
  if (off) return -1;
  static const bool no_ragged = getenv("BT709_NO_RAGGED_LIN") != nullptr;  // reproduction: ragged widths on k_decode_fast
  if (!((a.w % 2 == 0) && (a.h % 2 == 0)) || (a.w % 128 != 0 && no_ragged)) return -1;
  const bool contiguous = a.nframes == 1 || (a.sy == (long long) a.h * a.py && a.sc == (long long) a.rows2 * a.pc &&
                                             a.so == (long long) a.h * a.po && (!ALPHA || a.sa == (long long) a.h * a.pa));
  // Strided batches: measured on the Y|CbCr-per-frame layout (session 36), the wrap-aware instantiation is level with the
  // round-1 ring kernel on full-width rows (945 vs 940 Gpix/s for 1080p, 944 vs 982 for 2048-wide sRGB: the layout, not
  // the kernel, costs the 8 % against a contiguous batch) and 23 % faster on ragged widths, which the ring kernel cannot
  // take (828 vs 674).  So: ragged widths only.  BT709_STRIDED_LIN=all / 0: always / never.
  static const char *strided_env = getenv("BT709_STRIDED_LIN");
  const bool strided = !contiguous;
  const bool no_strided = strided_env ? strided_env[0] == '0' || (strided_env[0] != 'a' && a.w % 128 == 0) : a.w % 128 == 0;
  // (strided: non-negative strides that keep the frames of a plane apart, so that the wrap constants are what they say)
  if (strided && (no_strided || a.sy < (long long) a.h * a.py || a.sc < (long long) a.rows2 * a.pc || a.so < (long long) a.h * a.po ||
                  (ALPHA && a.sa < (long long) a.h * a.pa) || !aligned(a.sy, 16) || !aligned(a.sc, 16) || !aligned(a.so, 16) ||
                  (ALPHA && !aligned(a.sa, 16))))
    return -1;
  if ((long long) a.nframes * a.h >= (1ll << 30)) return -1;
  if (!(aligned(a.y, 16) && aligned(a.py, 16) && aligned(a.c, 16) && aligned(a.pc, 16) && aligned(a.out, 16) && aligned(a.po, 16) &&
        (!ALPHA || (aligned(a.a, 16) && aligned(a.pa, 16)))))
    return -1;
  // steps per warp with a full grid of 32-warp CTAs
  const unsigned long long tiles = (unsigned long long) a.nframes * a.rows2 * a.tiles_per_row;
  const bool small = tiles <= (unsigned long long) small_steps * sm_count * 32;
  static const bool no_blog = getenv("BT709_NO_BLOG") != nullptr;
  if (a.blog_lut && a.blog_cells > 0 && !no_blog) {
    // Linear mode, biased log cells: 645 cells, so a replica per lane leaves room for all 32 padded ring slices
    DecodeArgs b = a;
    b.lut = a.blog_lut;
    b.clamp_lo = a.blog_clamp_lo;
    b.log_cells = a.blog_cells;
    b.log_first = a.blog_first;
    b.log_bias = a.blog_bias;
    if (small) return launch_decode_lin<-1, 2, ALPHA, false>(b, sm_count, s, true, strided);
    return launch_decode_lin<-1, 5, ALPHA, false>(b, sm_count, s, true, strided);
  }
  if (a.log_lut && a.log_cells > 0) {
    DecodeArgs b = a;
    b.lut = a.log_lut;
    b.clamp_lo = a.log_clamp_lo;
    if (small) return launch_decode_lin<0, 2, ALPHA, false>(b, sm_count, s, true, strided);
    return launch_decode_lin<0, 5, ALPHA, true>(b, sm_count, s, true, strided);
  }
  if (a.cells_log2 == 10) {
    if (small) return launch_decode_lin<10, 3, ALPHA, false>(a, sm_count, s, ALPHA, strided);
    return launch_decode_lin<10, 5, ALPHA, false>(a, sm_count, s, ALPHA, strided);
  }
  return -1;
}

template <int CL2, int RL2, bool ALPHA>
static int launch_decode_fast(const DecodeArgs &a0, int sm_count, cudaStream_t s) {
  DecodeArgs a = a0;
  int nstages = 0, stage_bytes = 0, nrows = 0;
  // The bulk-staged variant is correct (tests/test_gpu_decode.py runs it) but measured slower than
  // per-thread loads (774 vs 811 Gpix/s on 600 x 1080p): the path is bound by HBM *write*
  // bandwidth (profiles/r01_membw_probe.json), not by read latency.  Opt-in for reproduction.
  static const bool use_bulk = getenv("BT709_BULK") != nullptr;
  if constexpr (CL2 != 0)
  if (use_bulk && bulk_ok<CL2, RL2, ALPHA>(a, &nstages, &stage_bytes, &nrows)) {
    constexpr size_t lut = (size_t) ((1 << CL2) + 1) * (1 << RL2) * 4;
    a.nstages = nstages; a.stage_bytes = stage_bytes; a.nrows = nrows;
    const size_t smem = lut + (size_t) nstages * stage_bytes + (size_t) nstages * 16;
    // the output offsets are 32-bit in 16-byte units: split batches that would overflow
    const unsigned long long lim = 0xFFFF0000ull;
    unsigned long long per = a.so > 0 ? (lim * 16 - (unsigned long long) a.h * a.po) / (unsigned long long) a.so + 1 : lim;
    if (per < 1) per = 1;
    for (int f0 = 0; f0 < a0.nframes; f0 += (int) per) {
      a.nframes = (int) ((unsigned long long) (a0.nframes - f0) < per ? (unsigned long long) (a0.nframes - f0) : per);
      a.y = a0.y + (long long) f0 * a0.sy;
      a.c = a0.c + (long long) f0 * a0.sc;
      a.out = a0.out + (long long) f0 * a0.so;
      if (ALPHA) a.a = a0.a + (long long) f0 * a0.sa;
      const unsigned total_rp = (unsigned) a.nframes * a.rows2;
      unsigned grid = (total_rp + a.rps - 1) / a.rps;
      if (grid > (unsigned) sm_count) grid = sm_count;
      a.step_rp = grid * a.rps;
      const unsigned d_f = a.step_rp / a.rows2;
      a.d_rp = a.step_rp - d_f * a.rows2;
      a.po16 = (unsigned) (a.po / 16); a.so16 = (unsigned) (a.so / 16);
      a.dout = d_f * a.so16 + a.d_rp * 2 * a.po16;
      a.wout = a.so16 - a.rows2 * 2 * a.po16;
      k_decode_bulk<CL2, RL2, ALPHA, kDecodeThreads><<<grid, kDecodeThreads, smem, s>>>(a), note_launch();
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) return (int) e;
    }
    return 0;
  }
  const bool full = (a.w % 128 == 0) && (a.h % 2 == 0);
  if (!full) return launch_decode_fast2<CL2, RL2, ALPHA, false, true>(a, sm_count, s);
  // contiguous frames: decode the batch as one tall image (no frame seams to step across)
  const bool contiguous = a.nframes == 1 || (a.sy == (long long) a.h * a.py && a.sc == (long long) a.rows2 * a.pc &&
                                             a.so == (long long) a.h * a.po && (!ALPHA || a.sa == (long long) a.h * a.pa));
  if (contiguous && (long long) a.nframes * a.h < (1ll << 30)) {
    a.h *= a.nframes;
    a.rows2 *= a.nframes;
    a.nframes = 1;
    // rows of 15 or 30 tiles (1920, 3840 wide) leave 2 of the 32 warps idle in the band layout: take the tiles
    // in linear order instead (ring kernel only; falls back to the band layout when the ring is not usable)
    static const bool no_linear = getenv("BT709_NO_LINEAR") != nullptr;
    const int wpc = kDecodeThreads / 32;
    // (only if the ring still fits with all 32 warps active: the 209 KB log-cell LUT leaves room for 30)
    const size_t lut_bytes = (size_t) (CL2 ? (1 << CL2) + 1 : a.log_cells) * (1 << RL2) * 4;
    const bool ring_fits = lut_bytes + (size_t) wpc * kRingDepth * (ALPHA ? 640 : 384) <= kMaxDynSmem;
    if (!no_linear && ring_fits && a.tiles_per_row < wpc && wpc % a.tiles_per_row != 0) {
      a.linear_tiles = a.tiles_per_row;
      return launch_decode_fast2<CL2, RL2, ALPHA, true, true>(a, sm_count, s);
    }
    return launch_decode_fast2<CL2, RL2, ALPHA, true, false>(a, sm_count, s);
  }
  return launch_decode_fast2<CL2, RL2, ALPHA, true, true>(a, sm_count, s);
}

int launch_decode(const DecodeArgs &a0, int sm_count, void *stream) {
  cudaStream_t s = (cudaStream_t) stream;
  DecodeArgs a = a0;
  a.rows2 = (a.h + 1) / 2;
  // 128-pixel tiles; the CTA's warps cover rps whole row pairs per step (see k_decode_fast)
  a.tiles_per_row = (a.w + 127) / 128;
  const int wpc = kDecodeThreads / 32;
  a.rps = a.tiles_per_row >= wpc ? 1 : wpc / a.tiles_per_row;
  a.tile_stride = a.tiles_per_row >= wpc ? wpc : a.tiles_per_row;  // > 4096 px wide: warps loop over tiles
  const bool alpha = a.a != nullptr;
  const unsigned long long units = (unsigned long long) a.nframes * a.rows2 + 2ull * sm_count * 32;
  bool fast = (a.w % 2 == 0) && aligned(a.y, 4) && aligned(a.py, 4) && aligned(a.sy, 4) && aligned(a.c, 4) &&
              aligned(a.pc, 4) && aligned(a.sc, 4) && aligned(a.out, 16) && aligned(a.po, 16) &&
              aligned(a.so, 16) && units < 0x7FFFFFFFull && a.sy < (1ll << 32) && a.sc < (1ll << 32) &&
              a.so < (1ll << 34) && a.sa < (1ll << 32) && a.py < (1ll << 28) && a.pc < (1ll << 28) && a.po < (1ll << 30) &&
              a.pa < (1ll << 28) && a.sy >= 0 && a.sc >= 0 && a.so >= 0 && a.sa >= 0 &&
              (a.cells_log2 == 10 || a.cells_log2 == 12);
  if (alpha) fast = fast && aligned(a.a, 4) && aligned(a.pa, 4) && aligned(a.sa, 4);
  if (fast) {
    {
      const int e = alpha ? try_decode_lin<true>(a, sm_count, s) : try_decode_lin<false>(a, sm_count, s);
      if (e != -1) return e;  // -1: not applicable, use the older kernels below
    }
    // Linear mode: the steep pure-sRGB-encode step function needs 4096 uniform cells (8 replicas, 2-3 way
    // bank conflicts) but only ~1650 log-spaced ones, which fit shared memory with one replica per lane
    if (a.log_lut && a.log_cells > 0 && (size_t) a.log_cells * 128 <= kMaxDynSmem) {
      a.lut = a.log_lut;
      a.clamp_lo = a.log_clamp_lo;
      return alpha ? launch_decode_fast<0, 5, true>(a, sm_count, s) : launch_decode_fast<0, 5, false>(a, sm_count, s);
    }
    if (a.cells_log2 == 10)
      return alpha ? launch_decode_fast<10, 5, true>(a, sm_count, s) : launch_decode_fast<10, 5, false>(a, sm_count, s);
    return alpha ? launch_decode_fast<12, 3, true>(a, sm_count, s) : launch_decode_fast<12, 3, false>(a, sm_count, s);
  }
  const long long blocks = (long long) a.nframes * a.rows2 * ((a.w + 1) / 2);
  long long grid = (blocks + 255) / 256;
  if (grid > (long long) sm_count * 16) grid = (long long) sm_count * 16;
  if (alpha)
    k_decode_generic<true><<<(unsigned) grid, 256, 0, s>>>(a), note_launch();
  else
    k_decode_generic<false><<<(unsigned) grid, 256, 0, s>>>(a), note_launch();
  return (int) cudaGetLastError();
}

template <bool ALPHA>
static int optin_decode_smem_log() {
  int e = (int) cudaFuncSetAttribute(k_decode_fast<0, 5, ALPHA, true, true, kPrefetchDepth, kDecodeThreads>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kMaxDynSmem);
  if (e) return e;
  e = (int) cudaFuncSetAttribute(k_decode_fast<0, 5, ALPHA, true, false, kPrefetchDepth, kDecodeThreads>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kMaxDynSmem);
  if (e) return e;
  e = (int) cudaFuncSetAttribute(k_decode_fast<0, 5, ALPHA, false, true, kPrefetchDepth, kDecodeThreads>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kMaxDynSmem);
  if (e) return e;
  e = (int) cudaFuncSetAttribute(k_decode_ring<0, 5, ALPHA, true, kRingDepth, kDecodeThreads>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kMaxDynSmem);
  if (e) return e;
  e = (int) cudaFuncSetAttribute(k_decode_lin<0, 5, ALPHA, 2, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int) kMaxDynSmem);
  if (e) return e;
  e = (int) cudaFuncSetAttribute(k_decode_lin<0, 5, ALPHA, 3, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int) kMaxDynSmem);
  if (e) return e;
  e = (int) cudaFuncSetAttribute(k_decode_lin<0, 2, ALPHA, 2, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int) kMaxDynSmem);
  if (e) return e;
  e = (int) cudaFuncSetAttribute(k_decode_lin<0, 2, ALPHA, 3, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int) kMaxDynSmem);
  if (e) return e;
  e = (int) cudaFuncSetAttribute(k_decode_lin<-1, 5, ALPHA, 2, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int) kMaxDynSmem);
  if (e) return e;
  e = (int) cudaFuncSetAttribute(k_decode_lin<-1, 5, ALPHA, 3, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int) kMaxDynSmem);
  if (e) return e;
  e = (int) cudaFuncSetAttribute(k_decode_lin<-1, 2, ALPHA, 2, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int) kMaxDynSmem);
  if (e) return e;
  e = (int) cudaFuncSetAttribute(k_decode_lin<-1, 2, ALPHA, 3, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int) kMaxDynSmem);
  if (e) return e;
  e = (int) cudaFuncSetAttribute(k_decode_lin<10, 5, ALPHA, 2, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int) kMaxDynSmem);
  if (e) return e;
  e = (int) cudaFuncSetAttribute(k_decode_lin<10, 5, ALPHA, 3, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int) kMaxDynSmem);
  if (e) return e;
  e = (int) cudaFuncSetAttribute(k_decode_lin<10, 3, ALPHA, 2, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int) kMaxDynSmem);
  if (e) return e;
  e = (int) cudaFuncSetAttribute(k_decode_lin<10, 3, ALPHA, 3, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int) kMaxDynSmem);
  if (e) return e;
  return (int) cudaFuncSetAttribute(k_decode_ring<0, 5, ALPHA, false, kRingDepth, kDecodeThreads>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kMaxDynSmem);
}

template <int CL2, int RL2, bool ALPHA>
static int optin_decode_smem() {
  constexpr size_t smem = (size_t) ((1 << CL2) + 1) * (1 << RL2) * 4;
  int e = (int) cudaFuncSetAttribute(k_decode_fast<CL2, RL2, ALPHA, true, true, kPrefetchDepth, kDecodeThreads>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
  if (e) return e;
  e = (int) cudaFuncSetAttribute(k_decode_fast<CL2, RL2, ALPHA, true, false, kPrefetchDepth, kDecodeThreads>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
  if (e) return e;
  e = (int) cudaFuncSetAttribute(k_decode_ring<CL2, RL2, ALPHA, true, kRingDepth, kDecodeThreads>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kMaxDynSmem);
  if (e) return e;
  e = (int) cudaFuncSetAttribute(k_decode_ring<CL2, RL2, ALPHA, false, kRingDepth, kDecodeThreads>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kMaxDynSmem);
  if (e) return e;
  e = (int) cudaFuncSetAttribute(k_decode_bulk<CL2, RL2, ALPHA, kDecodeThreads>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kMaxDynSmem);
  if (e) return e;
  return (int) cudaFuncSetAttribute(k_decode_fast<CL2, RL2, ALPHA, false, true, kPrefetchDepth, kDecodeThreads>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
}

template <int CL2, int RL2>
static int optin_decode_reg() {
  int e = (int) cudaFuncSetAttribute(k_decode_reg<CL2, RL2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kMaxDynSmem);
  if (e) return e;
  e = (int) cudaFuncSetAttribute(k_decode_reg<CL2, RL2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kMaxDynSmem);
  if (e) return e;
  return (int) cudaFuncSetAttribute(k_decode_reg<CL2, RL2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kMaxDynSmem);
}

// per device (called from bt709_ctx_create with the device current): opt in to > 48 KB of dynamic
// shared memory for the replicated-LUT decode kernels
int decode_kernels_init() {
  int e;
  if ((e = optin_decode_reg<10, 5>())) return e;
  if ((e = optin_decode_reg<-1, 5>())) return e;
  if ((e = optin_decode_smem<10, 5, false>())) return e;
  if ((e = optin_decode_smem<10, 5, true>())) return e;
  if ((e = optin_decode_smem<12, 3, false>())) return e;
  if ((e = optin_decode_smem<12, 3, true>())) return e;
  if ((e = optin_decode_smem_log<false>())) return e;
  if ((e = optin_decode_smem_log<true>())) return e;
  return 0;
}

}  // namespace bt709
