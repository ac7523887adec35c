// bt709_rescale.cu — pass 2: bilinear resample of an sRGB BGRA8 image in linear light
// (samplingShader, AAPLShaders.metal:73-85 + MetalScaleRenderContext.m:55-105).
#include "bt709_pair.cuh"

namespace bt709 {

// ================================================================================= rescale
// Pass 2 (samplingShader, AAPLShaders.metal:73-85): one thread per output pixel.  Source texels
// come through the read-only path (each is reused by ~1.8 output pixels for 1080p -> 1440p and
// the four taps of neighbouring threads overlap, so L1/L2 absorb the reuse and HBM sees each
// source row about once).  sRGB->linear is a 256-float shared table (exact: the input is 8-bit);
// linear->sRGB is the cell LUT in shared memory.  float32 steps mirror oracle_rescale_bgra.
__global__ void __launch_bounds__(256) k_rescale(const RescaleArgs p) {
  extern __shared__ __align__(16) uint32_t s_mem[];
  float *s_lin = reinterpret_cast<float *>(s_mem);
  uint32_t *s_enc = s_mem + 256;
  const int nenc = (1 << p.enc_cells_log2) + 1;
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lin[i] = __ldg(&p.lin256[i]);
  for (int i = threadIdx.x; i < nenc; i += blockDim.x) s_enc[i] = __ldg(&p.enc_lut[i]);
  __syncthreads();
  const float fN = (float) (1 << p.enc_cells_log2);

  const int tiles_x = (p.ow + 31) / 32, tiles_y = (p.oh + 7) / 8;
  const long long tiles = (long long) p.nframes * tiles_x * tiles_y;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
    const int bx = (int) (t % tiles_x);
    const long long r = t / tiles_x;
    const int by = (int) (r % tiles_y);
    const long long f = r / tiles_y;
    const int ox = bx * 32 + tx, oy = by * 8 + ty;
    if (ox >= p.ow || oy >= p.oh) continue;
    float sx = __fadd_rn(__fmul_rn(__fadd_rn((float) ox, 0.5f), p.scale_x), -0.5f);
    float sy = __fadd_rn(__fmul_rn(__fadd_rn((float) oy, 0.5f), p.scale_y), -0.5f);
    float fx0 = floorf(sx), fy0 = floorf(sy);
    float fx = __fsub_rn(sx, fx0), fy = __fsub_rn(sy, fy0);
    int x0 = (int) fx0, y0 = (int) fy0;
    int x1 = min(max(x0 + 1, 0), p.iw - 1), y1 = min(max(y0 + 1, 0), p.ih - 1);
    x0 = min(max(x0, 0), p.iw - 1);
    y0 = min(max(y0, 0), p.ih - 1);
    const uint8_t *src = p.in + f * p.si;
    const uint32_t p00 = __ldg(reinterpret_cast<const uint32_t *>(src + (long long) y0 * p.pi) + x0);
    const uint32_t p10 = __ldg(reinterpret_cast<const uint32_t *>(src + (long long) y0 * p.pi) + x1);
    const uint32_t p01 = __ldg(reinterpret_cast<const uint32_t *>(src + (long long) y1 * p.pi) + x0);
    const uint32_t p11 = __ldg(reinterpret_cast<const uint32_t *>(src + (long long) y1 * p.pi) + x1);
    uint32_t outw = 0;
#pragma unroll
    for (int ch = 0; ch < 4; ch++) {
      const int sh = ch * 8;
      float t00, t10, t01, t11;
      if (ch < 3) {
        t00 = s_lin[(p00 >> sh) & 255]; t10 = s_lin[(p10 >> sh) & 255];
        t01 = s_lin[(p01 >> sh) & 255]; t11 = s_lin[(p11 >> sh) & 255];
      } else {
        const float k = 1.0f / 255.0f;
        t00 = __fmul_rn((float) (p00 >> 24), k); t10 = __fmul_rn((float) (p10 >> 24), k);
        t01 = __fmul_rn((float) (p01 >> 24), k); t11 = __fmul_rn((float) (p11 >> 24), k);
      }
      float top = __fadd_rn(t00, __fmul_rn(fx, __fsub_rn(t10, t00)));
      float bot = __fadd_rn(t01, __fmul_rn(fx, __fsub_rn(t11, t01)));
      float v = add_sat(top, __fmul_rn(fy, __fsub_rn(bot, top)));
      uint32_t code = ch < 3 ? lut_code(s_enc, v, p.enc_clamp_lo, fN) : (uint32_t) round_pos(__fmul_rn(v, 255.0f));
      outw |= code << sh;
    }
    *reinterpret_cast<uint32_t *>(p.out + f * p.so + (long long) oy * p.po + (long long) ox * 4) = outw;
  }
}

// -------------------------------------------------------------------- rescale, column walker
// Fast path for upscales, same size and moderate downscales (launch_rescale decides; when the horizontal
// scale is <= 1 it prefers k_rescale_cols2 below, which shares this design with two columns per thread).
// A warp owns a tile of 32 output columns x kWalkTH output rows; a thread owns one output column
// and walks down the source rows the tile needs.  For each source row it converts its two
// horizontal taps to linear light (256-entry tables replicated per lane: conflict-free) and blends
// them (the oracle's `top`/`bot`, same float32 steps), keeps the two most recent blended rows in
// registers and emits every output row that lies between them, so a blended row is shared by the
// ~1/scale output rows that touch it.  A warp reads ~96-128 contiguous source bytes and writes 128
// contiguous output bytes per row; nothing but the tables lives in shared memory.
//
// Everything that depends on the output row only (fy, the top source row, the row index)
// is warp-uniform: the warp computes it once per tile, one row per lane, into its own slice of
// shared memory, and the row loop reads it back with one LDS.64 per output row.  The walker runs
// over UNCLAMPED source rows (top row -1 at the upper edge) and clamps only the row it fetches, so
// the clamped-edge cases of the oracle (both taps on the same row) need no special code.
//
// The oracle's clamp01 after the vertical blend is the identity for these operands (for a, b in
// [0,1] and f in [0,1]: d = RN(b-a), p = RN(f*d), |p| <= |d|, and RN(a+p) stays inside [0,1] because
// RN(1-a) <= (1-a) + 2^-25 and RN(a-b) <= a), so the blends run on packed FFMA2/FMUL2 without .sat.
struct Lin4 {
  f32x2 rg, ba;  // (R, G), (B, A) in linear light (A is linear already)
};

struct WalkTables {
  uint32_t lin_lane;   // shared address of the {linear, alpha} byte table | lane << 2 (the PRMT operand)
  uint32_t rows_base;  // shared address of this warp's row-info slice
  uint32_t enc_base;   // encode-LUT base with the index bias and the lane's replica folded in
};

// Fills the CTA's tables (all threads, ends with a __syncthreads) and returns the per-thread addresses.
// Layout (dynamic shared memory):
//   [lin table, 64 KB, placed at the first 64 KB-ALIGNED shared address]
//     word b*64 + r      : sRGB byte b -> linear float, replica r (one per lane)
//     word b*64 + 32 + r : alpha byte b -> b * (1/255), replica r
//     entry stride 256 B and 64 KB alignment: the lookup address is ONE PRMT of the texel word and a
//     per-lane register holding (table address | lane << 2), no add
//   [row info, one slice per warp]  in the alignment gap in front of the table when it fits, else after it
//   [encode cell LUT, R interleaved replicas: cell i of replica r at word i*R + r]  after the table
// ELOG: the encode LUT is the log-cell form (index = upper half of the float) with 2^ERL2 replicas.  It is too big
// for the space behind the aligned table, so it goes FIRST and the table follows at the next 64 KB boundary:
//   [encode log-cell LUT][pad][lin table @ 64 KB boundary][row info]
template <int ERL2, int WARPS, bool ELOG = false>
__device__ __forceinline__ WalkTables walk_tables(const RescaleArgs &p, uint32_t *s_mem) {
  const uint32_t smem0 = (uint32_t) __cvta_generic_to_shared(s_mem);
  constexpr int kRowWords = WARPS * kWalkRowInfo;
  const int nenc = (ELOG ? p.enc_log_cells : (1 << p.enc_cells_log2) + 1) << ERL2;
  const uint32_t head = ELOG ? (uint32_t) (nenc + 4) * 4u : 0u;  // bytes in front of the table's alignment pad
  const uint32_t pad = head + ((0x10000u - ((smem0 + head) & 0xFFFFu)) & 0xFFFFu);
  float *s_lin = reinterpret_cast<float *>(s_mem + (pad >> 2));
  const bool rows_in_gap = !ELOG && pad >= kRowWords * 4;
  uint32_t *s_after = s_mem + (pad >> 2) + 256 * 64;
  uint32_t *s_rows = rows_in_gap ? s_mem : s_after;
  uint32_t *s_enc = ELOG ? s_mem + 4 : s_after + (rows_in_gap ? 0 : kRowWords) + 4;
  if (ELOG && pad + 65536u + kRowWords * 4u > (uint32_t) kMaxDynSmem) __trap();  // dynamic shared memory starts 1 KB into the window
  for (int i = threadIdx.x; i < 256 * 32; i += blockDim.x) {
    s_lin[(i >> 5) * 64 + (i & 31)] = __ldg(&p.lin256[i >> 5]);
    s_lin[(i >> 5) * 64 + 32 + (i & 31)] = __fmul_rn((float) (i >> 5), 1.0f / 255.0f);
  }
  const uint32_t *glut = ELOG ? p.enc_log_lut : p.enc_lut;
  for (int i = threadIdx.x; i < nenc; i += blockDim.x) s_enc[i] = __ldg(&glut[i >> ERL2]);
  // the encode-LUT address is (index register << 2) + enc_base, the index register holding 2^23 + cell (uniform
  // cells) or first + cell (log cells): the bias is folded into enc_base, which goes through shared memory so
  // that ptxas sees an opaque register and emits a single LEA per lookup
  if (threadIdx.x == 0)
    s_enc[-1] = (uint32_t) __cvta_generic_to_shared(s_enc) - ((ELOG ? p.enc_log_first : 0x4B000000u) << (2 + ERL2));
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  WalkTables t;
  t.lin_lane = (smem0 + pad) | (lane << 2);
  t.rows_base = (uint32_t) __cvta_generic_to_shared(s_rows + warp * kWalkRowInfo);
  t.enc_base = *reinterpret_cast<volatile uint32_t *>(s_enc - 1) + ((lane & ((1 << ERL2) - 1)) << 2);
  return t;
}

// ERL2: log2 of the encode-LUT replication (bank-conflict relief: lanes use replica lane % R); ELOG: log-cell LUT
template <int ERL2, bool ELOG>
__global__ void __launch_bounds__(kWalkWarps * 32, 1) k_rescale_cols(const RescaleArgs p) {
  extern __shared__ __align__(16) uint32_t s_mem[];
  const WalkTables wt = walk_tables<ERL2, kWalkWarps, ELOG>(p, s_mem);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t lin_lane = wt.lin_lane, rows_base = wt.rows_base, enc_base = wt.enc_base;
  const float fN = (float) (1 << p.enc_cells_log2);
  const float clamp_lo = ELOG ? p.enc_log_clamp_lo : p.enc_clamp_lo;
  const f32x2 one = pk(p.one, p.one), neg1 = pk(-1.0f, -1.0f);
  const f32x2 fN2 = pk(fN, fN), magic2 = pk(8388608.0f, 8388608.0f);

  auto lin_of = [&](uint32_t px, uint32_t sel, int extra) -> float {
    const uint32_t addr = __byte_perm(px, lin_lane, sel);  // table | byte << 8 | lane << 2
    float v;
    if (extra)
      asm("ld.shared.f32 %0, [%1+128];" : "=f"(v) : "r"(addr));
    else
      asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
  };
  auto texel = [&](uint32_t px) -> Lin4 {  // BGRA8 sRGB texel -> linear (R,G), (B, A/255)
    Lin4 t;
    t.rg = pk(lin_of(px, 0x7624, 0), lin_of(px, 0x7614, 0));
    t.ba = pk(lin_of(px, 0x7604, 0), lin_of(px, 0x7634, 1));
    return t;
  };
  auto enc_idx = [&](float t) -> uint32_t {  // cell-LUT entry for the index held in the float 2^23 + idx
    uint32_t e;
    asm("ld.shared.u32 %0, [%1];" : "=r"(e) : "r"((__float_as_uint(t) << (2 + ERL2)) + enc_base));
    return e;
  };

  const int tiles_x = (p.ow + 31) >> 5, tiles_y = (p.oh + kWalkTH - 1) / kWalkTH;
  const unsigned tiles_per_frame = tiles_x * tiles_y;
  const unsigned tiles = (unsigned) p.nframes * tiles_per_frame;
  const int ylast = p.ih - 1;
  const unsigned pi = (unsigned) p.pi, po = (unsigned) p.po;
  for (unsigned t = blockIdx.x * kWalkWarps + warp; t < tiles; t += gridDim.x * kWalkWarps) {
    const unsigned f = t / tiles_per_frame, tr = t - f * tiles_per_frame;
    const int ty = tr / tiles_x, tx = tr - ty * tiles_x;
    const int oy0 = ty * kWalkTH;
    // ---- row info of this tile: {fy, top source row (unclamped), output row}
    __syncwarp();
#pragma unroll
    for (int k = 0; k < kWalkTH / 32; k++) {
      const int r = k * 32 + lane, oy = oy0 + r;
      const float sy = __fadd_rn(__fmul_rn(__fadd_rn((float) oy, 0.5f), p.scale_y), -0.5f);
      const float fy0 = floorf(sy);
      uint4 e;
      e.x = __float_as_uint(__fsub_rn(sy, fy0));
      e.y = oy < p.oh ? (uint32_t) (int) fy0 : 0x7fffffffu;
      e.z = (unsigned) oy;
      e.w = 0;
      asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(rows_base + r * 16), "r"(e.x), "r"(e.y), "r"(e.z), "r"(e.w) : "memory");
    }
    if (lane == 0)
      asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(rows_base + kWalkTH * 16), "r"(0x7fffffffu) : "memory");
    __syncwarp();

    // ---- horizontal sample position of this column: same float32 steps as oracle_rescale_bgra
    const int ox = tx * 32 + lane;
    const bool live = ox < p.ow;
    const float sx = __fadd_rn(__fmul_rn(__fadd_rn((float) ox, 0.5f), p.scale_x), -0.5f);
    const float fx0 = floorf(sx);
    const float fx = __fsub_rn(sx, fx0);
    const int xi = (int) fx0;
    const int x0 = min(max(xi, 0), p.iw - 1), x1 = min(max(xi + 1, 0), p.iw - 1);
    const f32x2 fx2 = pk(fx, fx);
    const uint8_t *src0 = p.in + (long long) f * p.si + (size_t) x0 * 4;
    // output row pointer: stepped BEFORE each row is computed (stepping after the store made the compiler keep
    // a copy of the old pointer alive), so it starts one row above the tile
    uint8_t *dp = p.out + (long long) f * p.so + (size_t) min(ox, p.ow - 1) * 4 + (long long) (oy0 - 1) * (long long) po;

    auto blend = [&](uint32_t w0, uint32_t w1) -> Lin4 {  // top/bot of the oracle: t0 + fx * (t1 - t0)
      const Lin4 t0 = texel(w0), t1 = texel(w1);
      Lin4 h;
      h.rg = fma2_rn(mul2(fx2, fma2_rn(t0.rg, neg1, t1.rg)), one, t0.rg);
      h.ba = fma2_rn(mul2(fx2, fma2_rn(t0.ba, neg1, t1.ba)), one, t0.ba);
      return h;
    };
    // taps of the next source row; the fetch pointer stops advancing at the last row (clamp to edge)
    // (two pointers: x1 - x0 is 0 at the clamped image edges and 1 elsewhere, a per-thread quantity)
    const uint8_t *fp0 = nullptr, *fp1 = nullptr;
    // loads the taps of source row clamp(row, 0, ih - 1), which is where the pointers stand, and steps
    // them down unless the next row clamps to the same one
    auto fetch = [&](int row, uint32_t &w0, uint32_t &w1) {
      w0 = __ldg(reinterpret_cast<const uint32_t *>(fp0));
      w1 = __ldg(reinterpret_cast<const uint32_t *>(fp1));
      const unsigned inc = (unsigned) row < (unsigned) ylast ? pi : 0u;  // also 0 for row == -1
      fp0 += inc;
      fp1 += inc;
    };
    auto emit = [&](float fy, const Lin4 &A, const Lin4 &B) {
      // v = top + fy * (bot - top) per channel; clamp01 is the identity here (see above)
      const f32x2 fy2 = pk(fy, fy);
      const f32x2 vrg = fma2_rn(mul2(fy2, fma2_rn(A.rg, neg1, B.rg)), one, A.rg);
      const f32x2 vba = fma2_rn(mul2(fy2, fma2_rn(A.ba, neg1, B.ba)), one, A.ba);
      float vr, vg, vb, va, tr_, tg_, tb_;
      upk(vrg, vr, vg);
      upk(vba, vb, va);
      vr = fmaxf(vr, clamp_lo); vg = fmaxf(vg, clamp_lo); vb = fmaxf(vb, clamp_lo);
      if (ELOG) {  // log cells: the index is the upper half of the float
        tr_ = __uint_as_float(__byte_perm(__float_as_uint(vr), 0, 0x4432));
        tg_ = __uint_as_float(__byte_perm(__float_as_uint(vg), 0, 0x4432));
        tb_ = __uint_as_float(__byte_perm(__float_as_uint(vb), 0, 0x4432));
      } else {
        upk(fma2_rz(pk(vr, vg), fN2, magic2), tr_, tg_);
        asm("fma.rz.f32 %0, %1, %2, 0f4B000000;" : "=f"(tb_) : "f"(vb), "f"(fN));
      }
      const uint32_t sr = enc_idx(tr_) + __float_as_uint(vr);  // byte 3 = sRGB code
      const uint32_t sg = enc_idx(tg_) + __float_as_uint(vg);
      const uint32_t sb = enc_idx(tb_) + __float_as_uint(vb);
      // alpha: round(va * 255), half away from zero: the integer part of RZ(x + 32768.5) sits in byte 1
      float ta;
      asm("add.rz.f32 %0, %1, 0f47000080;" : "=f"(ta) : "f"(__fmul_rn(va, 255.0f)));
      const uint32_t bg = __byte_perm(sb, sg, 0x0073);
      const uint32_t w = __byte_perm(bg, __byte_perm(sr, __float_as_uint(ta), 0x0053), 0x5410);
      if (live) *reinterpret_cast<uint32_t *>(dp) = w;
    };

    // ---- walk.  H[j & 1] = blended row y, H[(j + 1) & 1] = blended row y + 1; the raw taps of the
    // next kWalkPD rows sit in registers, fetched that many rows ahead of use.
    uint32_t ip = rows_base;  // row-info cursor
    uint4 info;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(info.x), "=r"(info.y), "=r"(info.z), "=r"(info.w) : "r"(ip));
    int y = (int) info.y;     // first top row, >= -1
    uint32_t q0[kWalkPD], q1[kWalkPD], a0w, a1w;
    fp0 = ptr_mad(src0, (unsigned) max(y, 0), pi);
    fp1 = fp0 + (x1 - x0) * 4;
    fetch(y, a0w, a1w);
#pragma unroll
    for (int j = 0; j < kWalkPD; j++) fetch(y + 1 + j, q0[j], q1[j]);
    Lin4 H[2];
    H[0] = blend(a0w, a1w);
    bool done = false;
    while (!done) {  // every condition below is warp-uniform
#pragma unroll
      for (int j = 0; j < 2 * kWalkPD; j++) {
        H[(j + 1) & 1] = blend(q0[j % kWalkPD], q1[j % kWalkPD]);                 // row min(y + 1, ih - 1)
        fetch(y + 1 + kWalkPD, q0[j % kWalkPD], q1[j % kWalkPD]);                 // refill the slot
        while ((int) info.y <= y) {  // emit the output rows whose top tap is y (<=: can never spin)
          dp += po;
          emit(__uint_as_float(info.x), H[j & 1], H[(j + 1) & 1]);
          ip += 16;
          asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(info.x), "=r"(info.y), "=r"(info.z), "=r"(info.w) : "r"(ip));
        }
        y++;
        if (info.y == 0x7fffffffu) { done = true; break; }
      }
    }
  }
}

// -------------------------------------------------------------- rescale, two columns per thread
// Same walk, but a thread owns TWO adjacent output columns (a warp: 64 columns x kWalkTH rows).  For horizontal scales <= 1 the taps of the two columns lie in
// three consecutive (clamped) source texels p0 <= p1 <= p2: column a uses (T0, T1), column b uses (T0, T1) or
// (T1, T2).  Per source row that is 3 loads and 3 texel conversions instead of 4 + 4, one pointer update, one
// loop; per output row one row-info read, one pointer step and one 8-byte store for both pixels.
// 24 warps x 80 registers measured best (16 warps: latency-bound, 32 warps x 64 registers: same as 24).
constexpr int kWalk2Warps = 24;

template <int ERL2, bool ELOG>
__global__ void __launch_bounds__(kWalk2Warps * 32, 1) k_rescale_cols2(const RescaleArgs p) {
  extern __shared__ __align__(16) uint32_t s_mem[];
  const WalkTables wt = walk_tables<ERL2, kWalk2Warps, ELOG>(p, s_mem);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t lin_lane = wt.lin_lane, rows_base = wt.rows_base, enc_base = wt.enc_base;
  const float fN = (float) (1 << p.enc_cells_log2);
  const float clamp_lo = ELOG ? p.enc_log_clamp_lo : p.enc_clamp_lo;
  const f32x2 one = pk(p.one, p.one), neg1 = pk(-1.0f, -1.0f);
  const f32x2 fN2 = pk(fN, fN), magic2 = pk(8388608.0f, 8388608.0f);

  auto lin_of = [&](uint32_t px, uint32_t sel, int extra) -> float {
    const uint32_t addr = __byte_perm(px, lin_lane, sel);
    float v;
    if (extra)
      asm("ld.shared.f32 %0, [%1+128];" : "=f"(v) : "r"(addr));
    else
      asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
  };
  auto texel = [&](uint32_t px) -> Lin4 {
    Lin4 t;
    t.rg = pk(lin_of(px, 0x7624, 0), lin_of(px, 0x7614, 0));
    t.ba = pk(lin_of(px, 0x7604, 0), lin_of(px, 0x7634, 1));
    return t;
  };
  auto enc_idx = [&](float t) -> uint32_t {
    uint32_t e;
    asm("ld.shared.u32 %0, [%1];" : "=r"(e) : "r"((__float_as_uint(t) << (2 + ERL2)) + enc_base));
    return e;
  };
  auto lerp = [&](const Lin4 &t0, const Lin4 &t1, f32x2 f2) -> Lin4 {  // t0 + f * (t1 - t0), the oracle's steps
    Lin4 h;
    h.rg = fma2_rn(mul2(f2, fma2_rn(t0.rg, neg1, t1.rg)), one, t0.rg);
    h.ba = fma2_rn(mul2(f2, fma2_rn(t0.ba, neg1, t1.ba)), one, t0.ba);
    return h;
  };
  auto pixel = [&](const Lin4 &v) -> uint32_t {  // linear (R,G),(B,A) -> BGRA8 word (sRGB encode, alpha linear)
    float vr, vg, vb, va, tr_, tg_, tb_;
    upk(v.rg, vr, vg);
    upk(v.ba, vb, va);
    vr = fmaxf(vr, clamp_lo); vg = fmaxf(vg, clamp_lo); vb = fmaxf(vb, clamp_lo);
    if (ELOG) {  // log cells: the index is the upper half of the float (PRMT, see lut2 in bt709_decode.cu)
      tr_ = __uint_as_float(__byte_perm(__float_as_uint(vr), 0, 0x4432));
      tg_ = __uint_as_float(__byte_perm(__float_as_uint(vg), 0, 0x4432));
      tb_ = __uint_as_float(__byte_perm(__float_as_uint(vb), 0, 0x4432));
    } else {
      upk(fma2_rz(pk(vr, vg), fN2, magic2), tr_, tg_);
      asm("fma.rz.f32 %0, %1, %2, 0f4B000000;" : "=f"(tb_) : "f"(vb), "f"(fN));
    }
    const uint32_t sr = enc_idx(tr_) + __float_as_uint(vr);
    const uint32_t sg = enc_idx(tg_) + __float_as_uint(vg);
    const uint32_t sb = enc_idx(tb_) + __float_as_uint(vb);
    float ta;
    asm("add.rz.f32 %0, %1, 0f47000080;" : "=f"(ta) : "f"(__fmul_rn(va, 255.0f)));
    const uint32_t bg = __byte_perm(sb, sg, 0x0073);
    return __byte_perm(bg, __byte_perm(sr, __float_as_uint(ta), 0x0053), 0x5410);
  };

  const int tiles_x = (p.ow + 63) >> 6, tiles_y = (p.oh + kWalkTH - 1) / kWalkTH;
  const unsigned tiles_per_frame = tiles_x * tiles_y;
  const unsigned tiles = (unsigned) p.nframes * tiles_per_frame;
  const int ylast = p.ih - 1, xlast = p.iw - 1;
  const unsigned pi = (unsigned) p.pi, po = (unsigned) p.po;
  for (unsigned t = blockIdx.x * kWalk2Warps + warp; t < tiles; t += gridDim.x * kWalk2Warps) {
    const unsigned f = t / tiles_per_frame, tr = t - f * tiles_per_frame;
    const int ty = tr / tiles_x, tx = tr - ty * tiles_x;
    const int oy0 = ty * kWalkTH;
    __syncwarp();
#pragma unroll
    for (int k = 0; k < kWalkTH / 32; k++) {
      const int r = k * 32 + lane, oy = oy0 + r;
      const float sy = __fadd_rn(__fmul_rn(__fadd_rn((float) oy, 0.5f), p.scale_y), -0.5f);
      const float fy0 = floorf(sy);
      asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(rows_base + r * 16), "r"(__float_as_uint(__fsub_rn(sy, fy0))),
                   "r"(oy < p.oh ? (uint32_t) (int) fy0 : 0x7fffffffu) : "memory");
    }
    if (lane == 0) asm volatile("st.shared.v2.u32 [%0], {%1,%1};" ::"r"(rows_base + kWalkTH * 16), "r"(0x7fffffffu) : "memory");
    __syncwarp();

    // ---- the two columns of this thread
    const int oxa = tx * 64 + lane * 2;
    const bool live_a = oxa < p.ow, live_b = oxa + 1 < p.ow;
    const float sxa = __fadd_rn(__fmul_rn(__fadd_rn((float) oxa, 0.5f), p.scale_x), -0.5f);
    const float sxb = __fadd_rn(__fmul_rn(__fadd_rn((float) (oxa + 1), 0.5f), p.scale_x), -0.5f);
    const float fxa0 = floorf(sxa), fxb0 = floorf(sxb);
    const int xia = (int) fxa0, xib = (int) fxb0;      // xib - xia is 0 or 1 for scales <= 1
    const bool bsel = xib != xia;                      // column b uses (T1, T2) instead of (T0, T1)
    const f32x2 fxa2 = pk(__fsub_rn(sxa, fxa0), __fsub_rn(sxa, fxa0));
    const f32x2 fxb2 = pk(__fsub_rn(sxb, fxb0), __fsub_rn(sxb, fxb0));
    const int p0 = min(max(xia, 0), xlast), p1 = min(max(xia + 1, 0), xlast), p2 = min(max(xia + 2, 0), xlast);
    const unsigned d1 = (unsigned) (p1 - p0) * 4u, d2 = (unsigned) (p2 - p0) * 4u;  // 0/4 and 0/4/8 (clamped edges)
    const uint8_t *src0 = p.in + (long long) f * p.si + (size_t) p0 * 4;
    uint8_t *dp = p.out + (long long) f * p.so + (size_t) min(oxa, p.ow - 1) * 4 + (long long) (oy0 - 1) * (long long) po;
    const bool wide_store = live_b && ((reinterpret_cast<uintptr_t>(dp) | po) & 7u) == 0;

    const uint8_t *fp = nullptr;
    auto fetch = [&](int row, uint32_t &w0, uint32_t &w1, uint32_t &w2) {
      w0 = __ldg(reinterpret_cast<const uint32_t *>(fp));
      w1 = __ldg(reinterpret_cast<const uint32_t *>(fp + d1));
      w2 = __ldg(reinterpret_cast<const uint32_t *>(fp + d2));
      fp += (unsigned) row < (unsigned) ylast ? pi : 0u;
    };
    auto blend2 = [&](uint32_t w0, uint32_t w1, uint32_t w2, Lin4 &ha, Lin4 &hb) {
      const Lin4 t0 = texel(w0), t1 = texel(w1), t2 = texel(w2);
      ha = lerp(t0, t1, fxa2);
      Lin4 u0, u1;
      u0.rg = bsel ? t1.rg : t0.rg; u0.ba = bsel ? t1.ba : t0.ba;
      u1.rg = bsel ? t2.rg : t1.rg; u1.ba = bsel ? t2.ba : t1.ba;
      hb = lerp(u0, u1, fxb2);
    };

    uint32_t ip = rows_base;
    uint32_t info_fy, info_y;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(info_fy), "=r"(info_y) : "r"(ip));
    int y = (int) info_y;
    uint32_t q0[kWalkPD], q1[kWalkPD], q2[kWalkPD], a0w, a1w, a2w;
    fp = ptr_mad(src0, (unsigned) max(y, 0), pi);
    fetch(y, a0w, a1w, a2w);
#pragma unroll
    for (int j = 0; j < kWalkPD; j++) fetch(y + 1 + j, q0[j], q1[j], q2[j]);
    Lin4 Ha[2], Hb[2];
    blend2(a0w, a1w, a2w, Ha[0], Hb[0]);
    bool done = false;
    while (!done) {
#pragma unroll
      for (int j = 0; j < 2 * kWalkPD; j++) {
        blend2(q0[j % kWalkPD], q1[j % kWalkPD], q2[j % kWalkPD], Ha[(j + 1) & 1], Hb[(j + 1) & 1]);
        fetch(y + 1 + kWalkPD, q0[j % kWalkPD], q1[j % kWalkPD], q2[j % kWalkPD]);
        while ((int) info_y <= y) {
          dp += po;
          const float fy = __uint_as_float(info_fy);
          const f32x2 fy2 = pk(fy, fy);
          const uint32_t wa = pixel(lerp(Ha[j & 1], Ha[(j + 1) & 1], fy2));
          const uint32_t wb = pixel(lerp(Hb[j & 1], Hb[(j + 1) & 1], fy2));
          if (wide_store) {
            *reinterpret_cast<uint2 *>(dp) = make_uint2(wa, wb);
          } else {
            if (live_a) *reinterpret_cast<uint32_t *>(dp) = wa;
            if (live_b) *reinterpret_cast<uint32_t *>(dp + 4) = wb;
          }
          ip += 16;
          asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(info_fy), "=r"(info_y) : "r"(ip));
        }
        y++;
        if (info_y == 0x7fffffffu) { done = true; break; }
      }
    }
  }
}

__global__ void __launch_bounds__(kPairWarps * 32, 1) k_rescale_pair(const RescaleArgs p) {
  extern __shared__ __align__(16) uint32_t s_mem[];
  pdl_launch_dependents();
  const PairCtx c = pair_tables(p, s_mem, 0, kPairRowInfo);
  pdl_wait();  // (calls of a few frames are launched with programmatic stream serialization; the tables are constant)
  const int th = p.pair_th;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int TW = 64 * kPairNP;  // tile width
  const int tiles_x = (p.ow + TW - 1) / TW, tiles_y = (p.oh + th - 1) / th;
  const unsigned tiles_per_frame = tiles_x * tiles_y;
  const unsigned tiles = (unsigned) p.nframes * tiles_per_frame;
  bool with_alpha = false;  // warp-uniform: the code this warp tries first on its next tile
  // Tile order: every tile that is not in the first or last tile column first, those two columns at the end.  The
  // walker has separate code for tiles without clamped taps (pair_tile: INNER), each variant a ~20 KB loop; with edge
  // tiles sprinkled through the order, one or two warps of every CTA sit in the other variant at any time and the
  // instruction cache thrashes (measured: 288 against 316 Gpix/s).  (The variant itself is still chosen by a vote on
  // the real tap offsets: the order only decides WHEN a warp meets an edge tile.)
  const unsigned inner_x = tiles_x > 2 ? (unsigned) tiles_x - 2u : 0u, edge_x = (unsigned) tiles_x - inner_x;
  const unsigned n_inner = (unsigned) p.nframes * (unsigned) tiles_y * inner_x;
  // (tile t goes to CTA t % grid: a call with fewer tiles than warps still uses every SM)
  for (unsigned t = blockIdx.x + gridDim.x * warp; t < tiles; t += gridDim.x * kPairWarps) {
    unsigned f;
    int ty, tx;
    if (t < n_inner) {
      const unsigned per = (unsigned) tiles_y * inner_x;
      f = t / per;
      const unsigned tr = t - f * per;
      ty = (int) (tr / inner_x);
      tx = (int) (tr - (unsigned) ty * inner_x) + 1;
    } else {
      const unsigned e = t - n_inner, per = (unsigned) tiles_y * edge_x;
      f = e / per;
      const unsigned tr = e - f * per;
      ty = (int) (tr / edge_x);
      const unsigned k = tr - (unsigned) ty * edge_x;
      tx = inner_x ? (k ? tiles_x - 1 : 0) : (int) k;
    }
    const uint8_t *src = p.in + (long long) f * p.si;
    uint8_t *dst = p.out + (long long) f * p.so;
    const unsigned pi = (unsigned) p.pi;
    const int ox0 = tx * TW, oy0 = ty * th;
    // the alpha code runs one pair per thread (a 64-column half tile at a time) to stay within the register budget
    auto alpha_tile = [&]() -> bool {
      bool opaque = true;
#pragma unroll 1
      for (int q = 0; q < kPairNP; q++)
        if (ox0 + q * 64 < p.ow) opaque = pair_tile<true, true, 1, true, kPairPD>(p, c, src, pi, dst, ox0 + q * 64, oy0, lane, 0ull, th) && opaque;
      return opaque;
    };
    if (p.known_opaque) {  // the source was written by pass 1 without an alpha stream: nothing to watch for
      pair_tile<false, true, kPairNP, false, kPairPD>(p, c, src, pi, dst, ox0, oy0, lane, 0ull, th);
    } else if (!with_alpha) {
      if (!pair_tile<false, true, kPairNP, true, kPairPD>(p, c, src, pi, dst, ox0, oy0, lane, 0ull, th)) {  // a tap with A != 255: redo with alpha
        alpha_tile();
        with_alpha = true;
      }
    } else {
      if (alpha_tile()) with_alpha = false;  // opaque again
    }
  }
}

// shared-memory address (low 16 bits matter) at which a kernel's dynamic shared memory starts; -1 = unknown
int g_dyn_smem_base = -1;  // shared with bt709_fused.cu
__global__ void k_probe_smem_base(unsigned *out) {
  extern __shared__ __align__(16) uint32_t s_mem[];
  *out = (unsigned) __cvta_generic_to_shared(s_mem);
}
// dynamic bytes the ELOG layout of walk_tables needs: [LUT, 16 replicas][pad to 64 KB][table 64 KB][row info]
static size_t elog_layout_bytes(unsigned base, int cells, int warps) {
  const unsigned head = ((unsigned) cells * 16u + 4u) * 4u;
  const unsigned pad = head + ((0x10000u - ((base + head) & 0xFFFFu)) & 0xFFFFu);
  return (size_t) pad + 65536u + (size_t) warps * kWalkRowInfo * 4u;
}

// dynamic bytes k_rescale_pair needs: [LUT, 32 replicas][pad to 64 KB][table 64 KB][row info]
size_t pair_layout_bytes(unsigned base, int cells, unsigned extra_bytes, int warps, int row_words) {
  const unsigned head = ((((unsigned) cells * 32u + 4u) * 4u + 127u) & ~127u) + extra_bytes;
  const unsigned pad = head + ((0x10000u - ((base + head) & 0xFFFFu)) & 0xFFFFu);
  return (size_t) pad + 65536u + (size_t) warps * row_words * 4u;
}

static size_t walk_smem_bytes(int enc_cells_log2, int rep_log2) {  // 64 KB of alignment slack in front of the table
  return 65536 + (256 * 64 + (size_t) kWalkWarps * kWalkRowInfo + 4 + ((((size_t) 1 << enc_cells_log2) + 1) << rep_log2)) * 4;
}

int launch_rescale(const RescaleArgs &a, int sm_count, void *stream) {
  // Column walker: any horizontal scale (two taps per column either way) and vertical scales up to 3 source
  // rows per output row.  It blends every source row of the strip, so a vertical downscale by s costs ~s blends
  // per output row; beyond 3 the per-pixel kernel (4 taps, no shared rows) is as good.  Offsets must fit 32 bits.
  if ((long long) a.ih <= 3ll * a.oh && a.enc_cells_log2 <= 13 && (long long) a.oh * a.po < (1ll << 32) &&
      (long long) a.ih * a.pi < (1ll << 32) && a.pi >= 0 && a.po >= 0) {
    // the kernel counts warp tiles in 32 bits: split batches with more than 2^31 of them across launches
    const long long tiles_per_frame = (long long) ((a.ow + 31) / 32) * ((a.oh + kWalkTH - 1) / kWalkTH);
    long long frames_per_launch = (1ll << 31) / tiles_per_frame;
    if (frames_per_launch < 1) frames_per_launch = 1;
    for (long long f0 = 0; f0 < a.nframes; f0 += frames_per_launch) {
      RescaleArgs b = a;
      b.nframes = (int) (a.nframes - f0 < frames_per_launch ? a.nframes - f0 : frames_per_launch);
      b.in = a.in + f0 * a.si;
      b.out = a.out + f0 * a.so;
      const long long ctas = (b.nframes * tiles_per_frame + kWalkWarps - 1) / kWalkWarps;
      const unsigned grid2 = (unsigned) (ctas < (long long) sm_count ? ctas : (long long) sm_count);
      static const bool no_cols2 = getenv("BT709_NO_COLS2") != nullptr;
      static const bool no_elog = getenv("BT709_NO_ELOG") != nullptr;
      const size_t smem_u = walk_smem_bytes(a.enc_cells_log2, 2);  // uniform-cell LUT, 4 replicas
      // log-cell LUT with 16 replicas in front of the 64 KB-aligned table: usable when that layout fits the
      // dynamic shared memory, which depends on where the window starts (probed once, rescale_kernels_init)
      const bool elog = b.enc_log_lut && !no_elog && g_dyn_smem_base >= 0 &&
                        elog_layout_bytes((unsigned) g_dyn_smem_base, b.enc_log_cells,
                                          a.iw <= a.ow && !no_cols2 ? kWalk2Warps : kWalkWarps) <= kMaxDynSmem;
      cudaStream_t st = (cudaStream_t) stream;
      static const bool no_pair = getenv("BT709_NO_PAIR") != nullptr;
      // (even width and 8-byte aligned rows: every column pair is stored with one 8-byte store)
      const bool pair_ok = (a.ow % 2 == 0) && aligned(a.out, 8) && aligned(a.po, 8) && aligned(a.so, 8);
      if (a.iw <= a.ow && pair_ok && !no_pair && !no_cols2 && b.enc_blog_lut && g_dyn_smem_base >= 0 &&
          pair_layout_bytes((unsigned) g_dyn_smem_base, b.enc_blog_cells, 0, kPairWarps, kPairRowInfo) <= kMaxDynSmem) {
        // horizontal scale <= 1: paired columns, opaque speculation, conflict-free encode LUT (k_rescale_pair)
        // tile height: kPairTH rows, halved while the call has fewer tiles than the grid has warps (one 1080p -> 1440p
        // frame: 460 tiles of 64 rows for 2368 warps; 1800 of 16 rows) - shorter tiles cost a few source rows of halo
        const long long tiles_x = (a.ow + 64 * kPairNP - 1) / (64 * kPairNP);
        int th = kPairTH;
        while (th > 16 && tiles_x * ((a.oh + th - 1) / th) * b.nframes < (long long) sm_count * kPairWarps) th >>= 1;
        b.pair_th = th;
        const long long tiles64 = tiles_x * ((a.oh + th - 1) / th) * b.nframes;
        const unsigned g2 = (unsigned) (tiles64 < (long long) sm_count ? tiles64 : (long long) sm_count);
        const bool small = tiles64 < 4ll * sm_count * kPairWarps;
        const cudaError_t le = launch_pdl(k_rescale_pair, g2, kPairWarps * 32,
                                          pair_layout_bytes((unsigned) g_dyn_smem_base, b.enc_blog_cells, 0, kPairWarps, kPairRowInfo), st,
                                          pdl_enabled(small), b);
        note_launch();
        if (le != cudaSuccess) return (int) le;
      } else if (a.iw <= a.ow && !no_cols2 && (elog || smem_u <= kMaxDynSmem)) {
        // horizontal scale <= 1: two adjacent columns per thread share their middle texel
        const long long tiles64 = (long long) ((a.ow + 63) / 64) * ((a.oh + kWalkTH - 1) / kWalkTH) * b.nframes;
        const long long ctas2 = (tiles64 + kWalk2Warps - 1) / kWalk2Warps;
        const unsigned g2 = (unsigned) (ctas2 < (long long) sm_count ? ctas2 : (long long) sm_count);
        if (elog) k_rescale_cols2<4, true><<<g2, kWalk2Warps * 32, kMaxDynSmem, st>>>(b), note_launch();
        else k_rescale_cols2<2, false><<<g2, kWalk2Warps * 32, smem_u, st>>>(b), note_launch();
      } else if (elog) {
        k_rescale_cols<4, true><<<grid2, kWalkWarps * 32, kMaxDynSmem, st>>>(b), note_launch();
      } else if (smem_u <= kMaxDynSmem) {
        k_rescale_cols<2, false><<<grid2, kWalkWarps * 32, smem_u, st>>>(b), note_launch();
      } else {
        k_rescale_cols<0, false><<<grid2, kWalkWarps * 32, walk_smem_bytes(a.enc_cells_log2, 0), st>>>(b), note_launch();
      }
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) return (int) e;
    }
    return 0;
  }
  const size_t smem = (256 + (size_t) (1 << a.enc_cells_log2) + 1) * 4;
  const long long tiles = (long long) a.nframes * ((a.ow + 31) / 32) * ((a.oh + 7) / 8);
  long long grid = tiles < (long long) sm_count * 8 ? tiles : (long long) sm_count * 8;
  k_rescale<<<(unsigned) grid, 256, smem, (cudaStream_t) stream>>>(a), note_launch();
  return (int) cudaGetLastError();
}

int rescale_kernels_init() {
  const int m = (int) kMaxDynSmem;
  int e;
  if (g_dyn_smem_base < 0) {  // where does dynamic shared memory start?  (1 KB into the window on sm_100)
    unsigned *d = nullptr, h = 0;
    if (cudaMalloc(&d, sizeof(unsigned)) == cudaSuccess) {
      k_probe_smem_base<<<1, 1, 16>>>(d);
      if (cudaMemcpy(&h, d, sizeof(unsigned), cudaMemcpyDeviceToHost) == cudaSuccess) g_dyn_smem_base = (int) (h & 0xFFFFu);
      cudaFree(d);
    }
    (void) cudaGetLastError();
  }
  if ((e = (int) cudaFuncSetAttribute(k_rescale_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, m))) return e;
  if ((e = (int) cudaFuncSetAttribute(k_rescale_cols<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, m))) return e;
  if ((e = (int) cudaFuncSetAttribute(k_rescale_cols<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, m))) return e;
  if ((e = (int) cudaFuncSetAttribute(k_rescale_cols<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, m))) return e;
  if ((e = (int) cudaFuncSetAttribute(k_rescale_cols2<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, m))) return e;
  return (int) cudaFuncSetAttribute(k_rescale_cols2<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, m);
}

}  // namespace bt709
