// Probe: achievable HBM GB/s for a linear stream at the decode byte mix (1.5 B read : 4 B written)
// as a function of how many read bytes each thread keeps in flight.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint4 ld16(const void*p){uint4 v; asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3},[%4];":"=r"(v.x),"=r"(v.y),"=r"(v.z),"=r"(v.w):"l"(p)); return v;}
__device__ __forceinline__ uint2 ld8(const void*p){uint2 v; asm volatile("ld.global.cs.v2.u32 {%0,%1},[%2];":"=r"(v.x),"=r"(v.y):"l"(p)); return v;}
__device__ __forceinline__ void st16(void*p, uint4 v){asm volatile("st.global.cs.v4.u32 [%0],{%1,%2,%3,%4};"::"l"(p),"r"(v.x),"r"(v.y),"r"(v.z),"r"(v.w):"memory");}
// one "unit" per thread-iteration = 16 px: 16 B of Y, 8 B of C, 64 B of BGRA.  UNROLL units in flight.
template<int UNROLL> __global__ void __launch_bounds__(1024) k(const uint8_t* y, const uint8_t* c, uint8_t* out, size_t units) {
  const size_t nthreads = (size_t)gridDim.x*blockDim.x;
  const int lane = threadIdx.x & 31;
  for (size_t u0 = (size_t)blockIdx.x*blockDim.x + threadIdx.x; u0 < units; u0 += nthreads*UNROLL) {
    uint4 a[UNROLL]; uint2 b[UNROLL];
    #pragma unroll
    for (int j=0;j<UNROLL;j++){ size_t u=u0+j*nthreads; if(u<units){ a[j]=ld16(y+u*16); b[j]=ld8(c+u*8);} }
    #pragma unroll
    for (int j=0;j<UNROLL;j++){ size_t u=u0+j*nthreads; if(u<units){
        size_t wbase = (u - lane) * 64;   // the warp's 32 units -> 2048 B; 4 lane-contiguous stores
        #pragma unroll
        for (int s=0;s<4;s++) st16(out + wbase + (size_t)(s*32+lane)*16, make_uint4(a[j].x+s, a[j].y, b[j].x, b[j].y));
    } }
  }
}
template<class F> float timeit(F f, int n=10){ for(int i=0;i<2;i++) f(); cudaDeviceSynchronize(); cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a); for(int i=0;i<n;i++) f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms,a,b); return ms/n/1e3f; }
int main(){
  size_t px=(size_t)600*1920*1080; size_t units=px/16;
  uint8_t *y,*c,*out; cudaMalloc(&y,px); cudaMalloc(&c,px/2); cudaMalloc(&out,px*4); cudaMemset(y,1,px); cudaMemset(c,2,px/2);
  for (int grid : {148, 296}) {
    float t1=timeit([&]{k<1><<<grid,1024>>>(y,c,out,units);});
    float t2=timeit([&]{k<2><<<grid,1024>>>(y,c,out,units);});
    float t4=timeit([&]{k<4><<<grid,1024>>>(y,c,out,units);});
    printf("grid %d  unroll1 (24 B/thr in flight): %.0f Gpix/s %.0f GB/s | unroll2: %.0f Gpix/s %.0f GB/s | unroll4: %.0f Gpix/s %.0f GB/s\n", grid,
      px/t1/1e9, px*5.5/t1/1e9, px/t2/1e9, px*5.5/t2/1e9, px/t4/1e9, px*5.5/t4/1e9);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
