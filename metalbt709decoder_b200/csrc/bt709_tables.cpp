// bt709_tables.cpp — see bt709_tables.h.  Host-only; runs once per context.
#include "bt709_tables.h"

#include <cmath>
#include <cstring>

namespace bt709 {

static inline uint32_t f2u(float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  return u;
}
static inline float u2f(uint32_t u) {
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}

// ------------------------------------------------------------------ curves
// Stored values are float32; pow is evaluated in double and the whole right-hand side
// is rounded to float once, as in the reference headers.

float srgb_eotf(float v) {
  if (v <= 0.04045f) return v * (1.0f / 12.92f);
  float s = (v + 0.055f) * (1.0f / (1.0f + 0.055f));
  return (float) std::pow((double) s, (double) 2.4f);
}

float srgb_oetf(float v) {
  if (v <= 0.0031308f) return v * 12.92f;
  const float a = 0.055f;
  const float g = 1.0f / 2.4f;
  return (float) ((double) (1.0f + a) * std::pow((double) v, (double) g) - (double) a);
}

float bt709_eotf(float v) {
  if (v < 0.081f) return v * (1.0f / 4.5f);
  const float a = 0.099f;
  const float g = 1.0f / 0.45f;
  float s = (v + a) * (1.0f / (1.0f + a));
  return (float) std::pow((double) s, (double) g);
}

// BT709.h:90-103 (the whole right-hand side is a double expression rounded to float once)
float bt709_oetf(float v) {
  if (v < 0.018f) return v * 4.5f;
  const float a = 0.099f;
  const float g = 0.45f;
  return (float) ((double) (1.0f + a) * std::pow((double) v, (double) g) - (double) a);
}

static const float kAppleGamma = 1.960938f;

float apple_eotf(float v) {
  if (v < 0.05583828f) return v * (1.0f / 16.0f);
  return (float) std::pow((double) v, (double) kAppleGamma);
}

float apple_oetf(float v) {
  if (v < 0.00349f) return v * 16.0f;
  const float g = 1.0f / kAppleGamma;
  return (float) std::pow((double) v, (double) g);
}

int quant255(float v) { return (int) std::round((double) (v * 255.0f)); }

int decode_code(int gamma, float x) {
  switch (gamma) {
    case kApple: return quant255(srgb_oetf(apple_eotf(x)));
    case kBt709: return quant255(srgb_oetf(bt709_eotf(x)));
    case kLinear: return quant255(srgb_oetf(x));
    case kBt709LinearRgb: return quant255(bt709_eotf(x));  // BT709_convertYCbCrToRGB(applyGammaMap = 1): linear RGB out
    default: return quant255(x);
  }
}

int from_linear_code(int gamma, float v) {
  switch (gamma) {
    case kSrgb: return quant255(srgb_oetf(v));
    case kApple: return quant255(apple_oetf(v));
    default: return quant255(v);
  }
}

float to_linear(int gamma, int byte) {
  float n = byte * (1.0f / 255.0f);
  switch (gamma) {
    case kSrgb: return srgb_eotf(n);
    case kApple: return apple_eotf(n);
    default: return n;
  }
}

// ------------------------------------------------------------- step functions
int StepFn::eval(float x) const {
  int lo = 0, hi = 255;  // largest k with thr[k] <= x
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (thr[mid] <= x)
      lo = mid;
    else
      hi = mid - 1;
  }
  return lo;
}

bool build_step_fn(CodeFn fn, int mode, StepFn *out) {
  const uint32_t kOne = 0x3F800000u;  // bits of 1.0f; positive floats order like their bits
  out->thr[0] = 0.0f;
  const int top = fn(mode, 1.0f);
  if (fn(mode, 0.0f) != 0 || top < 0 || top > 255) return false;
  for (int k = 1; k <= 255; k++) {
    if (k > top) {
      out->thr[k] = 2.0f;
      continue;
    }
    uint32_t lo = 0, hi = kOne;  // invariant: code(lo) < k <= code(hi)
    while (hi - lo > 1) {
      uint32_t mid = lo + (hi - lo) / 2;
      if (fn(mode, u2f(mid)) >= k)
        hi = mid;
      else
        lo = mid;
    }
    out->thr[k] = u2f(hi);
    if (k > 1 && !(out->thr[k] >= out->thr[k - 1])) return false;
  }
  // verification sample: the counted value must equal the direct evaluation
  uint32_t s = 0x9E3779B9u;
  for (int i = 0; i < 200000; i++) {
    s = s * 1664525u + 1013904223u;
    float x = u2f((uint32_t) (((uint64_t) s * (kOne + 1ull)) >> 32));
    if (fn(mode, x) != out->eval(x)) return false;
  }
  for (int k = 1; k <= top; k++) {
    uint32_t b = f2u(out->thr[k]);
    for (int d = -2; d <= 2; d++) {
      int64_t bb = (int64_t) b + d;
      if (bb < 0 || bb > (int64_t) kOne) continue;
      float x = u2f((uint32_t) bb);
      if (fn(mode, x) != out->eval(x)) return false;
    }
  }
  return true;
}

// ------------------------------------------------------------------ cell LUT
// Device lookup (kernels and cell_lut_eval below):
//     xc   = max(x, clamp_lo)                       clamp_lo = 2^-(n+1)
//     idx  = trunc(xc * N)                          fma.rz(xc, N, 2^23), low mantissa bits
//     code = (entries[idx] + bits(xc)) >> 24
// All floats of one cell share sign and exponent, so bits(xc) = (exp << 23) + mant and
//     entries[idx] = (base << 24) + (2^24 - field) - (exp << 23)      (mod 2^32)
// makes the sum (base << 24) + 2^24 - field + mant, whose top byte is base + (mant >= field).
// field = mantissa of the one threshold inside the cell, or 2^23 (never reached) if none.
bool build_cell_lut(const StepFn &fn, int cells_log2, CellLut *out) {
  const int n = cells_log2;
  const uint32_t N = 1u << n;
  const float clamp_lo = std::ldexp(1.0f, -(n + 1));
  if (fn.thr[1] < clamp_lo) return false;  // clamping would change a result
  out->cells_log2 = n;
  out->clamp_lo = clamp_lo;
  out->entries.assign(N + 1, 0u);
  for (uint32_t i = 0; i <= N; i++) {
    float lower = (i == 0) ? clamp_lo : (float) i / (float) N;
    uint32_t lb = f2u(lower);
    uint32_t base = (uint32_t) fn.eval(lower);
    uint32_t field = 1u << 23;
    if (i < N) {
      float upper = (float) (i + 1) / (float) N;
      float last = u2f(f2u(upper) - 1);  // largest float below the next cell
      if ((f2u(last) >> 23) != (lb >> 23)) return false;
      int cnt = fn.eval(last) - (int) base;
      if (cnt > 1) return false;
      if (cnt == 1) field = f2u(fn.thr[base + 1]) & 0x7FFFFFu;
    }
    uint32_t exp = lb >> 23;
    out->entries[i] = (base << 24) + ((1u << 24) - field) - (exp << 23);
  }
  return true;
}

bool build_log_lut(const StepFn &fn, LogLut *out) {
  const uint32_t kOneHi = 0x3F800000u >> 16;
  if (!(fn.thr[1] > 0.0f) || fn.thr[1] > 1.0f) return false;
  const uint32_t first = f2u(fn.thr[1]) >> 16;  // the cell that holds (or starts at) the first threshold
  out->first = first;
  out->clamp_lo = u2f(first << 16);
  out->entries.assign(kOneHi - first + 1, 0u);
  for (uint32_t c = first; c <= kOneHi; c++) {
    const uint32_t lo = c << 16, hi = lo + 0xFFFFu;  // bit patterns of the cell's first and last float
    const uint32_t base = (uint32_t) fn.eval(u2f(lo));
    const int cnt = fn.eval(u2f(hi)) - (int) base;
    if (cnt > 1) return false;
    // bits of the threshold inside the cell; if there is none, one past the cell (never reached)
    const uint32_t tb = cnt == 1 ? f2u(fn.thr[base + 1]) : lo + 0x10000u;
    out->entries[c - first] = (base << 24) + (1u << 24) - tb;
  }
  return true;
}

int log_lut_eval(const LogLut &lut, float x) {
  float xc = x > lut.clamp_lo ? x : lut.clamp_lo;
  return (int) ((lut.entries[(f2u(xc) >> 16) - lut.first] + f2u(xc)) >> 24);
}

bool build_biased_log_lut(const StepFn &fn, float bias, BiasedLogLut *out) {
  if (!(fn.thr[1] > 0.0f) || fn.thr[1] > 1.0f || !(bias > 0.0f)) return false;
  // clamp: the largest power of two not above the first threshold (code 0 below it either way)
  int e;
  std::frexp(fn.thr[1], &e);  // thr[1] = m * 2^e, m in [0.5, 1)
  const float clamp_lo = std::ldexp(1.0f, e - 1);
  auto cell_of = [&](float x) { return f2u(x + bias) >> 16; };  // float32 add, round to nearest (no contraction: plain +)
  const uint32_t first = cell_of(clamp_lo), last = cell_of(1.0f);
  out->clamp_lo = clamp_lo;
  out->bias = bias;
  out->first = first;
  out->entries.assign(last - first + 1, 0u);
  const uint32_t blo = f2u(clamp_lo), bhi = f2u(1.0f);
  uint32_t start = blo;  // bit pattern of the first x of the current cell
  for (uint32_t c = first; c <= last; c++) {
    if (cell_of(u2f(start)) != c) return false;  // an empty cell: the bias is too small for this resolution
    // one past the last x of cell c: bisection over the (monotone) cell index
    uint32_t end;
    if (c == last) {
      end = bhi + 1;
    } else {
      uint32_t lo = start, hi = bhi;  // cell_of(lo) == c < cell_of(hi)
      while (hi - lo > 1) {
        const uint32_t mid = lo + (hi - lo) / 2;
        if (cell_of(u2f(mid)) > c) hi = mid; else lo = mid;
      }
      end = hi;
    }
    const uint32_t base = (uint32_t) fn.eval(u2f(start));
    const int cnt = fn.eval(u2f(end - 1)) - (int) base;
    if (cnt > 1) return false;
    const uint32_t tb = cnt == 1 ? f2u(fn.thr[base + 1]) : end;
    // the add trick: (2^24 - tb + bits(x)) must stay inside [0, 2^25) for every x of the cell
    if (tb - start > (1u << 24) || (cnt == 1 && (end - 1) - tb >= (1u << 24))) return false;
    out->entries[c - first] = (base << 24) + (1u << 24) - tb;
    start = end;
  }
  return true;
}

int biased_log_lut_eval(const BiasedLogLut &lut, float x) {
  const float xc = x > lut.clamp_lo ? x : lut.clamp_lo;
  const float u = xc + lut.bias;
  return (int) ((lut.entries[(f2u(u) >> 16) - lut.first] + f2u(xc)) >> 24);
}

int build_cell_lut_auto(const StepFn &fn, int lo, int hi, CellLut *out) {
  for (int n = lo; n <= hi; n++)
    if (build_cell_lut(fn, n, out)) return n;
  return -1;
}

int cell_lut_eval(const CellLut &lut, float x) {
  float xc = x > lut.clamp_lo ? x : lut.clamp_lo;
  uint32_t idx = (uint32_t) (xc * (float) (1u << lut.cells_log2));  // exact product, truncation
  return (int) ((lut.entries[idx] + f2u(xc)) >> 24);
}

// ------------------------------------------------------------------- matrix
MatrixConsts matrix_consts() {
  // BT709.h:40-53, 389-397 — float32 arithmetic, left to right
  const float Kr = 0.2126f, Kg = 0.7152f, Kb = 0.0722f;
  const float crSpan = 1.5748f, cbSpan = 1.8556f;
  const float kr_kg = Kr / Kg, kb_kg = Kb / Kg;
  const float ys = 255.0f / (float) (235 - 16);
  const float cs = 255.0f / (float) (240 - 16);
  MatrixConsts m;
  m.inv255 = 1.0f / 255.0f;
  m.ys = ys;
  m.r_cr = cs * crSpan;
  m.g_cb = -1.0f * cs * cbSpan * kb_kg;
  m.g_cr = -1.0f * cs * crSpan * kr_kg;
  m.b_cb = cs * cbSpan;
  return m;
}

}  // namespace bt709

// ---------------------------------------------------------------- host self-test
#include <atomic>
#include <thread>
#include <vector>

namespace bt709 {

int selftest_host() {
  const MatrixConsts m = matrix_consts();
  // 1. cell LUT == direct evaluation on every possible G-channel matrix output, all four gammas
  for (int g = 0; g < 4; g++) {
    StepFn fn;
    CellLut lut;
    if (!build_step_fn(decode_code, g, &fn)) return 1;
    if (!build_cell_lut(fn, g == kLinear ? 12 : 10, &lut)) return 2;
    LogLut llut;  // the Linear mode's fast path uses the log-cell form of the same step function
    const bool has_log = g == kLinear;
    if (has_log && !build_log_lut(fn, &llut)) return 9;
    std::atomic<long> bad(0);
    std::vector<std::thread> th;
    for (int t = 0; t < 8; t++)
      th.emplace_back([&, t] {
        long b = 0;
        for (int Y = t * 32; Y < (t + 1) * 32; Y++) {
          const float yt = ((Y - 16) * m.inv255) * m.ys;
          for (int Cb = 0; Cb < 256; Cb++) {
            const float cbg = ((Cb - 128) * m.inv255) * m.g_cb;
            for (int Cr = 0; Cr < 256; Cr++) {
              float x = (yt + cbg) + ((Cr - 128) * m.inv255) * m.g_cr;
              x = x < 0.f ? 0.f : (x > 1.f ? 1.f : x);
              const int want = decode_code(g, x);
              if (cell_lut_eval(lut, x) != want) b++;
              if (has_log && log_lut_eval(llut, x) != want) b++;
            }
          }
        }
        bad += b;
      });
    for (auto &x : th) x.join();
    if (bad.load()) return 3;
  }
  // 1b. biased log-cell LUT of the sRGB encode (pass 2, fused kernel, Linear-mode decode) == the step function:
  //     every float within 64 ulp of a threshold or of a cell boundary, and 4 M floats spread over [0, 1]
  //     (the exhaustive sweep over all 1.07e9 floats of [0, 1] was run once at development time: 0 mismatches)
  //     - and the same for the biased log-cell LUTs of BT709_from_linear (Apple and sRGB curves), which the 4:2:0
  //     encode's chroma lookup uses with one replica per lane (k_encode_fast, OMODE 2)
  struct BlogCase { CodeFn fn; int mode; };
  const BlogCase blog_cases[] = {{decode_code, kLinear}, {from_linear_code, kApple}, {from_linear_code, kSrgb}};
  for (const BlogCase &bc : blog_cases) {
    StepFn fn;
    BiasedLogLut bl;
    if (!build_step_fn(bc.fn, bc.mode, &fn)) return 1;
    if (!build_biased_log_lut(fn, 1.0f / 32.0f, &bl)) return 10;
    long bad = 0;
    auto check = [&](uint32_t bits) {
      if (bits > 0x3F800000u) return;
      const float x = u2f(bits);
      if (biased_log_lut_eval(bl, x) != fn.eval(x)) bad++;
    };
    for (int k = 1; k <= 255; k++)
      if (fn.thr[k] <= 1.0f)
        for (int d = -64; d <= 64; d++) check(f2u(fn.thr[k]) + d);
    // cell boundaries: the first x whose biased value lands in each cell
    for (uint32_t c = bl.first; c < bl.first + (uint32_t) bl.entries.size(); c++) {
      uint32_t lo = 0, hi = 0x3F800000u;
      while (hi - lo > 1) {
        const uint32_t mid = lo + (hi - lo) / 2;
        if ((f2u(u2f(mid) + bl.bias) >> 16) >= c) hi = mid; else lo = mid;
      }
      for (int d = -64; d <= 64; d++) check(hi + d);
    }
    for (uint32_t b = 0; b <= 0x3F800000u; b += 263) check(b);
    if (bad) return 11;
  }
  // 2. x / c == fma(fma(-q0, c, x), rc, q0), q0 = x * rc, for every value the 4:2:0 encode divides
  {
    const float Kr = 0.2126f, Kg = 0.7152f, Kb = 0.0722f, cb = 1.8556f, cr = 1.5748f;
    const float rcb = 1.0f / cb, rcr = 1.0f / cr;
    std::atomic<long> bad(0);
    std::vector<std::thread> th;
    for (int t = 0; t < 8; t++)
      th.emplace_back([&, t] {
        long b = 0;
        for (int R = t * 32; R < (t + 1) * 32; R++)
          for (int G = 0; G < 256; G++)
            for (int B = 0; B < 256; B++) {
              const float rn = R * (1.0f / 255.0f), gn = G * (1.0f / 255.0f), bn = B * (1.0f / 255.0f);
              const float ey = (Kr * rn) + (Kg * gn) + (Kb * bn);
              const float xb = bn - ey, xr = rn - ey;
              const float q0 = xb * rcb, p0 = xr * rcr;
              if (std::fmaf(std::fmaf(-q0, cb, xb), rcb, q0) != xb / cb) b++;
              if (std::fmaf(std::fmaf(-p0, cr, xr), rcr, p0) != xr / cr) b++;
            }
        bad += b;
      });
    for (auto &x : th) x.join();
    if (bad.load()) return 4;
  }
  return 0;
}

}  // namespace bt709
