// bt709_tables.h — host-side construction of the constant tables a context uploads once.
//
// Every transfer-curve evaluation on the device is an exact table lookup:
//   * an 8-bit input (encode, rescale taps) indexes a 256-entry table;
//   * a float32 input x in [0,1] (matrix output, 2x2 linear mean, bilinear blend) is
//     mapped to its 8-bit code by a *step function*  code(x) = #{k : thr[k] <= x},
//     where thr[k] is the smallest float32 whose reference result is >= k.  The
//     reference evaluates the curve in double and rounds (BT709.h:711-713), which is a
//     monotone map float32 -> {0..255}, so 255 thresholds describe it completely and
//     the device result is bit-identical to the reference for every input.
//   * the step function is stored as a uniform-cell table ("cell LUT"): the cell index
//     is trunc(x * N), each cell holds at most one threshold, and one 32-bit entry
//     encodes {code at the cell start, mantissa of the threshold}; see cell_lut_eval().
#pragma once
#include <cstdint>
#include <vector>

namespace bt709 {

// kBt709LinearRgb: the reference's "previous API" pair BT709_convertRGBToYCbCr / BT709_convertYCbCrToRGB with
// applyGammaMap = 1 (linear-light RGB bytes <-> BT.709, BT709.h:320-345, 560-607); packed 4:4:4 entries only
enum Gamma { kApple = 0, kSrgb = 1, kLinear = 2, kBt709 = 3, kBt709LinearRgb = 4 };

// ---- scalar curve math with the reference's float32/double semantics ----------------
float srgb_eotf(float v);   // sRGB.h:43-57
float srgb_oetf(float v);   // sRGB.h:62-74
float bt709_eotf(float v);  // BT709.h:68-81
float bt709_oetf(float v);  // BT709.h:90-103
float apple_eotf(float v);  // BT709.h:125-137
float apple_oetf(float v);  // BT709.h:139-151
int quant255(float v);      // BT709.h:711-713

// code(x) of the decode path: matrix output -> sRGB byte, per gamma mode
int decode_code(int gamma, float x);
// code(v) of BT709_from_linear (BT709.h:1150-1167): linear float -> output-gamma byte
int from_linear_code(int gamma, float v);
// BT709_tolinearNorm (BT709.h:1098-1148) for one byte
float to_linear(int gamma, int byte);

// ---- step functions ------------------------------------------------------------------
struct StepFn {
  // thr[k], k = 1..255: smallest float32 x in [0,1] with code(x) >= k (2.0f if never).
  // thr[0] = 0.
  float thr[256];
  int eval(float x) const;  // reference evaluation by counting (host side, for checks)
};

typedef int (*CodeFn)(int mode, float x);
// Bisection over float32 bit patterns; returns false if code() is found non-monotone
// on the verification sample.
bool build_step_fn(CodeFn fn, int mode, StepFn *out);

// ---- cell LUT ------------------------------------------------------------------------
struct CellLut {
  int cells_log2 = 0;             // N = 1 << cells_log2 uniform cells over [0,1], plus cell N = {1.0}
  float clamp_lo = 0.f;           // inputs are clamped up to this power of two first (code is 0 below it)
  std::vector<uint32_t> entries;  // N + 1 words
};
// Fails (returns false) if some cell would hold two thresholds or thr[1] < clamp_lo.
bool build_cell_lut(const StepFn &fn, int cells_log2, CellLut *out);
// Smallest cells_log2 in [lo, hi] for which build_cell_lut succeeds; -1 if none.
int build_cell_lut_auto(const StepFn &fn, int lo, int hi, CellLut *out);
// Host model of the device lookup (bit-for-bit what the kernels do).
int cell_lut_eval(const CellLut &lut, float x);

// ---- log-cell LUT ----------------------------------------------------------------------
// Same step function, cells spaced like the float32 format itself: cell = bits(x) >> 16 (sign,
// exponent and 7 mantissa bits), first cell = the one clamp_lo starts.  A steep curve such as the
// pure sRGB encode of the decoder's Linear mode needs 4096 uniform cells but only ~1650 of these.
//     xc   = max(x, clamp_lo)
//     code = (entries[(bits(xc) >> 16) - first] + bits(xc)) >> 24
//     entries[i] = (base << 24) + 2^24 - bits(threshold inside the cell, or the next cell's start)
struct LogLut {
  float clamp_lo = 0.f;           // bits(clamp_lo) has its low 16 bits clear and clamp_lo <= thr[1]
  uint32_t first = 0;             // bits(clamp_lo) >> 16
  std::vector<uint32_t> entries;  // (bits(1.0f) >> 16) - first + 1 words
};
// Fails (returns false) if some cell would hold two thresholds.
bool build_log_lut(const StepFn &fn, LogLut *out);
int log_lut_eval(const LogLut &lut, float x);

// ---- biased log-cell LUT -----------------------------------------------------------------
// The log-cell idea with the cell taken from u = x + bias (one extra float add) instead of x itself:
//     xc   = max(x, clamp_lo)
//     u    = xc + bias                              (float32, round to nearest)
//     code = (entries[(bits(u) >> 16) - first] + bits(xc)) >> 24
// Adding a constant is monotone, so every cell is still an interval of x; below the bias the cells become
// uniform (width bias * 2^-7) instead of shrinking with x, which is all the sRGB encode needs there (its
// thresholds are 1/3295 apart on the linear toe).  With bias = 2^-5 the pure sRGB encode takes 5 binades of u
// = 644 cells instead of 1634, so a replica per lane (bank-conflict free for any content) fits shared memory
// next to the other tables: 82 KB instead of 209 KB.
struct BiasedLogLut {
  float clamp_lo = 0.f;
  float bias = 0.f;
  uint32_t first = 0;             // bits(clamp_lo + bias) >> 16
  std::vector<uint32_t> entries;  // (bits(1.0f + bias) >> 16) - first + 1 words
};
// Fails (returns false) if some cell would hold two thresholds or span more than the add trick allows.
bool build_biased_log_lut(const StepFn &fn, float bias, BiasedLogLut *out);
int biased_log_lut_eval(const BiasedLogLut &lut, float x);

// ---- matrix constants, computed in float32 exactly as BT709.h:389-397 ------------------
struct MatrixConsts {
  float inv255;  // 1.0f/255.0f
  float ys;      // 255/219
  float r_cr, g_cb, g_cr, b_cb;
};
MatrixConsts matrix_consts();

// host-only exhaustive self-check (bt709_selftest_host in the C ABI)
int selftest_host();

}  // namespace bt709
