// bt709_device.cuh — device-side helpers shared by the kernel translation units: streaming
// loads/stores, exact rounding helpers, the packed f32x2 arithmetic and the cell-LUT lookups.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>

#include "bt709_kernels.cuh"

namespace bt709 {

// ------------------------------------------------------------------ small device helpers
__device__ __forceinline__ float add_sat(float a, float b) {
  float r;  // clamps to [0,1] in the adder: saturatef of BT709.h:444-446 for free
  asm("add.rn.sat.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}

// byte k of `word` as an exact float minus `bias` (bias = 16 for Y/alpha, 128 for chroma):
// PRMT builds the float 2^23 + byte, one FADD removes 2^23 + bias.  No I2F (XU pipe).
template <int K>
__device__ __forceinline__ float byte_minus(uint32_t word, float bias_plus_2p23) {
  uint32_t bits = __byte_perm(word, 0x4B000000u, 0x7540 + K);
  return __fadd_rn(__uint_as_float(bits), -bias_plus_2p23);
}

__device__ __forceinline__ uint32_t ld_stream_u32(const void *p) {
  uint32_t v;
  asm volatile("ld.global.cs.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ uint4 ld_stream_u128(const void *p) {
  uint4 v;
  asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_stream_u128(void *p, uint4 v) {
  asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_stream_u64(void *p, uint2 v) {
  asm volatile("st.global.cs.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void st_stream_u32(void *p, uint32_t v) {
  asm volatile("st.global.cs.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// 16-byte store with an L2 cache policy (createpolicy handle), e.g. evict_last for data that is read back soon
__device__ __forceinline__ void st_l2hint_u128(void *p, uint4 v, unsigned long long policy) {
  asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(policy)
               : "memory");
}

// base + a * b with 32-bit unsigned a, b: one IMAD.WIDE.U32 (the compiler's own 64-bit address
// arithmetic spent ~7 instructions per pixel on sign extensions and carry chains)
__device__ __forceinline__ const uint8_t *ptr_mad(const uint8_t *base, unsigned a, unsigned b) {
  unsigned long long r;
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"((unsigned long long) base));
  return (const uint8_t *) r;
}

// round-half-away-from-zero of a non-negative float, exactly: the +0.5 is done with
// round-toward-zero so it can never be rounded up across an integer.
__device__ __forceinline__ int round_pos(float r) {
  float t;
  asm("add.rz.f32 %0, %1, 0f3F000000;" : "=f"(t) : "f"(r));
  return __float2int_rz(t);
}

// ---- cell LUT lookups (bt709_tables.cpp: build_cell_lut) --------------------------------
// shared, per-lane replicated: entry i of replica r lives at word (i << RL2) + r
template <int CL2, int RL2>
__device__ __forceinline__ uint32_t lut_sum_smem(float x, float clamp_lo, uint32_t sbase) {
  float xc = fmaxf(x, clamp_lo);
  float t;
  asm("fma.rz.f32 %0, %1, %2, 0f4B000000;" : "=f"(t) : "f"(xc), "f"((float) (1 << CL2)));
  uint32_t addr = (__float_as_uint(t) << (2 + RL2)) + sbase;
  uint32_t e;
  asm("ld.shared.u32 %0, [%1];" : "=r"(e) : "r"(addr));
  return e + __float_as_uint(xc);  // byte 3 is the 8-bit code
}

// generic: table pointer (global through the read-only path, or shared), runtime size
__device__ __forceinline__ uint32_t lut_code(const uint32_t *lut, float x, float clamp_lo, float fN) {
  float xc = fmaxf(x, clamp_lo);
  float t;
  asm("fma.rz.f32 %0, %1, %2, 0f4B000000;" : "=f"(t) : "f"(xc), "f"(fN));
  uint32_t idx = __float_as_uint(t) - 0x4B000000u;
  return (lut[idx] + __float_as_uint(xc)) >> 24;
}

// ------------------------------------------------------------------ replicated table fill
// Entry e of src (n entries) -> R = 1 << RL2 consecutive words at dst + e * stride_words (stride_words >= R, dst and
// the stride 16-byte aligned), R >= 4.  A warp loads 32 entries with ONE coalesced read and hands them round by shuffle;
// R / 4 lanes write one entry's replicas with 16-byte stores, so a store instruction covers 512 contiguous bytes
// (stride_words == R) without bank conflicts.  The old loop `dst[i] = __ldg(&src[i >> RL2])` read every entry R times:
// R * n / threads dependent L2 round trips per thread, 2-4 us per launch - a third of a one-frame call.
template <int RL2>
__device__ __forceinline__ void fill_replicated(uint32_t *dst, int stride_words, const uint32_t *src, int n) {
  static_assert(RL2 >= 2 && RL2 <= 5, "4 to 32 replicas");
  constexpr int LPE = 1 << (RL2 - 2);   // lanes per entry (16 bytes each)
  constexpr int EPI = 32 / LPE;         // entries per store instruction
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int base = warp * 32; base < n; base += nwarps * 32) {
    const uint32_t mine = base + lane < n ? __ldg(src + base + lane) : 0u;
#pragma unroll
    for (int j = 0; j < LPE; j++) {
      const int sub = j * EPI + lane / LPE;  // entry of the block this lane stores in round j
      const uint32_t v = __shfl_sync(0xffffffffu, mine, sub);
      if (base + sub < n)
        *reinterpret_cast<uint4 *>(dst + (size_t) (base + sub) * stride_words + (lane % LPE) * 4) = make_uint4(v, v, v, v);
    }
  }
}

// Programmatic dependent launch (kernels launched with cudaLaunchAttributeProgrammaticStreamSerialization, see
// launch_pdl): let the next launch of the stream be scheduled / wait until the previous kernel of the stream has
// completed and its writes are visible.  Both are no-ops in a launch without the attribute.  A kernel must not touch
// global memory other than its constant tables before pdl_wait().
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ------------------------------------------------------------------ packed float32 pairs
// Blackwell packed instructions (add/mul/fma .f32x2 -> FADD2/FMUL2/FFMA2): identical IEEE results per
// lane, half the issue slots.  NOTE: ptxas contracts a packed mul.rn.f32x2 feeding an add.rn.f32x2
// into one FFMA2 (one rounding instead of two).  Wherever an add follows a multiply and bit-exactness
// matters, write it as fma2_rn(product, one, addend) with `one` = 1.0f loaded from the kernel
// arguments, or do the add in scalar form.
typedef unsigned long long f32x2;

__device__ __forceinline__ f32x2 pk(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ f32x2 pku(uint32_t lo, uint32_t hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ void upk(f32x2 v, float &lo, float &hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ void upku(f32x2 v, uint32_t &lo, uint32_t &hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 fma2_rz(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rz.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

__device__ __forceinline__ f32x2 fma2_rn(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

// x + 32768.5 rounded toward zero, packed: the integer part of each (non-negative, < 32768) lane lands in byte 1
__device__ __forceinline__ f32x2 add2_rz_pair(f32x2 a) {
  f32x2 r, k;
  asm("mov.b64 %0, {%1, %1};" : "=l"(k) : "r"(0x47000080u));
  asm("add.rz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(k));
  return r;
}

template <int K>
__device__ __forceinline__ uint32_t magic_byte(uint32_t word) {  // bits of the float 2^23 + byte K
  return __byte_perm(word, 0x4B000000u, 0x7540 + K);
}

// ------------------------------------------------------------------ host-side helpers
// <<<grid, block, smem, stream>>> with or without programmatic stream serialization (small calls: the launch overlaps
// the tail of the previous kernel of the stream; the kernel itself calls pdl_wait() before it touches its inputs)
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t s, bool pdl,
                                     Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, args...);
  if (pdl && e == cudaErrorNotSupported) {  // a driver / partition without programmatic launch: plain stream order
    (void) cudaGetLastError();
    cfg.numAttrs = 0;
    e = cudaLaunchKernelEx(&cfg, kernel, args...);
  }
  return e;
}
// BT709_PDL=0: never; BT709_PDL=all: every launch that supports it; default: small calls only
static inline bool pdl_enabled(bool small_call) {
  static const char *env = getenv("BT709_PDL");
  if (!env) return small_call;
  if (env[0] == 'a') return true;
  return env[0] != '0' && small_call;
}

static inline bool aligned(const void *p, size_t a) { return ((uintptr_t) p & (a - 1)) == 0; }
static inline bool aligned(long long v, size_t a) { return ((unsigned long long) v & (a - 1)) == 0; }

constexpr size_t kMaxDynSmem = 232448;  // 227 KB of dynamic shared memory per CTA on sm_100

}  // namespace bt709
