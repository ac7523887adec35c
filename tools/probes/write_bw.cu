// Probe: how fast can one B200 write HBM, by store flavour?  (context for the decode roofline)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("err %s line %d\n", cudaGetErrorString(e), __LINE__); return 1;}}while(0)

template<int MODE> __global__ void __launch_bounds__(1024) k_write(uint4* out, size_t n16) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  uint4 v = make_uint4(i, i+1, i+2, i+3);
  for (; i < n16; i += stride) {
    if (MODE == 0) out[i] = v;
    if (MODE == 1) asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(out+i), "r"(v.x),"r"(v.y),"r"(v.z),"r"(v.w) : "memory");
    if (MODE == 2) asm volatile("st.global.wt.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(out+i), "r"(v.x),"r"(v.y),"r"(v.z),"r"(v.w) : "memory");
    if (MODE == 3) asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(out+i), "r"(v.x),"r"(v.y),"r"(v.z),"r"(v.w) : "memory");
  }
}
// 256-bit stores
__global__ void __launch_bounds__(1024) k_write256(uint4* out, size_t n32) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n32; i += stride) {
    unsigned a=i;
    asm volatile("st.global.v8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" :: "l"(out+2*i), "r"(a) : "memory");
  }
}
// bulk S2G: each CTA repeatedly writes 16 KB chunks from smem with cp.async.bulk
__global__ void __launch_bounds__(128) k_bulk(uint8_t* out, size_t nchunks, int chunk) {
  extern __shared__ __align__(128) uint8_t sm[];
  for (int i = threadIdx.x; i < chunk/4; i += blockDim.x) ((uint32_t*)sm)[i] = i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(sm);
    int inflight = 0;
    for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(out + c*(size_t)chunk), "r"(s), "r"(chunk) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      if (++inflight >= 8) { asm volatile("cp.async.bulk.wait_group.read 7;" ::: "memory"); }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}
// read-only
__global__ void __launch_bounds__(1024) k_read(const uint4* in, size_t n16, uint32_t* sink) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  uint32_t acc = 0;
  for (; i + 3*stride < n16; i += 4*stride) {
    uint4 a = __ldcs(in+i), b = __ldcs(in+i+stride), c = __ldcs(in+i+2*stride), d = __ldcs(in+i+3*stride);
    acc += a.x ^ b.y ^ c.z ^ d.w;
  }
  if (acc == 0x12345678) *sink = acc;
}
template<class F> float timeit(F f, int n=10){ for(int i=0;i<2;i++) f(); cudaDeviceSynchronize(); cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a); for(int i=0;i<n;i++) f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms,a,b); return ms/n/1e3f; }
int main(){
  size_t bytes = (size_t)4<<30; uint8_t* buf; CK(cudaMalloc(&buf, bytes)); uint32_t* sink; CK(cudaMalloc(&sink,4));
  size_t n16 = bytes/16;
  for (int grid : {148, 148*2, 148*16}) {
    printf("grid %d: ", grid);
    printf("st %.0f  ", bytes/timeit([&]{k_write<0><<<grid,1024>>>((uint4*)buf,n16);})/1e9);
    printf("st.cs %.0f  ", bytes/timeit([&]{k_write<1><<<grid,1024>>>((uint4*)buf,n16);})/1e9);
    printf("st.wt %.0f  ", bytes/timeit([&]{k_write<2><<<grid,1024>>>((uint4*)buf,n16);})/1e9);
    printf("st.noalloc %.0f  ", bytes/timeit([&]{k_write<3><<<grid,1024>>>((uint4*)buf,n16);})/1e9);
    printf("st256 %.0f  ", bytes/timeit([&]{k_write256<<<grid,1024>>>((uint4*)buf,n16/2);})/1e9);
    printf("read %.0f\n", bytes/timeit([&]{k_read<<<grid,1024>>>((const uint4*)buf,n16,sink);})/1e9);
  }
  for (int chunk : {4096, 16384, 65536}) {
    cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, chunk);
    printf("bulk S2G chunk %d: %.0f GB/s\n", chunk, bytes/timeit([&]{k_bulk<<<148*2,128,chunk>>>(buf, bytes/chunk, chunk);})/1e9);
  }
  CK(cudaGetLastError());
  // small buffer resident in L2 (64 MB): write bandwidth into L2
  size_t small = (size_t)64<<20;
  printf("L2-resident 64MB: st %.0f GB/s\n", small*16/timeit([&]{ for(int r=0;r<16;r++) k_write<0><<<148*2,1024>>>((uint4*)buf, small/16);})/1e9);
  return 0;
}
