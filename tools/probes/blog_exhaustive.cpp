// blog_exhaustive.cpp — the biased log-cell LUTs against their step functions on EVERY float32 of [0, 1] (1.07e9 values
// each): the sRGB encode of pass 2 / the fused kernel / the Linear-mode decode, and BT709_from_linear for the Apple and
// sRGB curves (the 4:2:0 encode's chroma lookup).  Host only; ~2 s on 16 threads.
// build: g++ -O2 -std=c++17 -I../../metalbt709decoder_b200/csrc -o blog_exhaustive blog_exhaustive.cpp \
//            ../../metalbt709decoder_b200/csrc/bt709_tables.cpp -lpthread
#include <atomic>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#include "bt709_tables.h"
using namespace bt709;

static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

int main() {
  struct Case { const char *name; CodeFn fn; int mode; };
  const Case cases[] = {{"decode_code(Linear) = sRGB encode", decode_code, kLinear},
                        {"from_linear_code(Apple)", from_linear_code, kApple},
                        {"from_linear_code(sRGB)", from_linear_code, kSrgb}};
  long total_bad = 0;
  for (const Case &c : cases) {
    StepFn fn;
    BiasedLogLut bl;
    if (!build_step_fn(c.fn, c.mode, &fn) || !build_biased_log_lut(fn, 1.0f / 32.0f, &bl)) {
      printf("%s: build failed\n", c.name);
      return 2;
    }
    std::atomic<long> bad(0);
    std::vector<std::thread> th;
    const uint64_t n = 0x3F800000ull + 1;  // bit patterns of [0, 1]
    const int T = 16;
    for (int t = 0; t < T; t++)
      th.emplace_back([&, t] {
        const uint32_t lo = (uint32_t) (n * t / T), hi = (uint32_t) (n * (t + 1) / T);
        long b = 0;
        int code = fn.eval(u2f(lo));  // the step function only changes at its thresholds: track it while walking up
        for (uint32_t bits = lo; bits < hi; bits++) {
          const float x = u2f(bits);
          while (code < 255 && x >= fn.thr[code + 1]) code++;
          if (biased_log_lut_eval(bl, x) != code) b++;
        }
        bad += b;
      });
    for (auto &x : th) x.join();
    printf("%s: %ld mismatches over all %llu floats of [0, 1] (%zu cells, bias %g)\n", c.name, bad.load(), (unsigned long long) n,
           bl.entries.size(), bl.bias);
    total_bad += bad.load();
  }
  return total_bad ? 1 : 0;
}
