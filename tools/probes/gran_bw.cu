// Probe: decode byte mix (1.5 B read : 4 B written per pixel), linear streams; vary the load width per
// lane (VEC pixels per thread-unit: 4 -> LDG.32, 8 -> LDG.64, 16 -> LDG.128) and the units in flight.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
template<int B> struct V; template<> struct V<4>{typedef uint32_t T;}; template<> struct V<8>{typedef uint2 T;}; template<> struct V<16>{typedef uint4 T;};
template<int B> __device__ __forceinline__ typename V<B>::T ld(const void*p){ return __ldcs((const typename V<B>::T*)p); }
__device__ __forceinline__ void st16(void*p, uint4 v){asm volatile("st.global.cs.v4.u32 [%0],{%1,%2,%3,%4};"::"l"(p),"r"(v.x),"r"(v.y),"r"(v.z),"r"(v.w):"memory");}
__device__ __forceinline__ uint32_t first(uint32_t v){return v;} __device__ __forceinline__ uint32_t first(uint2 v){return v.x;} __device__ __forceinline__ uint32_t first(uint4 v){return v.x;}
// unit = VEC px: VEC bytes of Y, VEC/2 bytes of C, VEC*4 bytes out (VEC/4 lane-contiguous STG.128 per warp)
template<int VEC, int UNROLL> __global__ void __launch_bounds__(1024) k(const uint8_t* y, const uint8_t* c, uint8_t* out, size_t units) {
  const size_t nthreads = (size_t)gridDim.x*blockDim.x;
  const int lane = threadIdx.x & 31;
  for (size_t u0 = (size_t)blockIdx.x*blockDim.x + threadIdx.x; u0 < units; u0 += nthreads*UNROLL) {
    typename V<VEC>::T a[UNROLL]; typename V<(VEC>=8?VEC/2:4)>::T b[UNROLL];
    #pragma unroll
    for (int j=0;j<UNROLL;j++){ size_t u=u0+j*nthreads; if(u<units){ a[j]=ld<VEC>(y+u*VEC);
        if constexpr (VEC>=8) b[j]=ld<VEC/2>(c+u*(VEC/2)); else { if ((u&1)==0) b[j]=ld<4>(c+(u>>1)*4); } } }
    #pragma unroll
    for (int j=0;j<UNROLL;j++){ size_t u=u0+j*nthreads; if(u<units){
        size_t wbase = (u - lane) * (VEC*4);
        #pragma unroll
        for (int s=0;s<VEC/4;s++) st16(out + wbase + (size_t)(s*32+lane)*16, make_uint4(first(a[j])+s, first(b[j]), 3, 4));
    } }
  }
}
template<class F> float timeit(F f, int n=10){ for(int i=0;i<2;i++) f(); cudaDeviceSynchronize(); cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a); for(int i=0;i<n;i++) f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms,a,b); return ms/n/1e3f; }
template<int VEC,int UN> void run(const uint8_t*y,const uint8_t*c,uint8_t*out,size_t px){
  float t=timeit([&]{k<VEC,UN><<<148,1024>>>(y,c,out,px/VEC);});
  printf("px/unit %2d  units in flight %d : %5.0f Gpix/s %5.0f GB/s\n", VEC, UN, px/t/1e9, px*5.5/t/1e9);
}
int main(){
  size_t px=(size_t)600*1920*1080;
  uint8_t *y,*c,*out; cudaMalloc(&y,px); cudaMalloc(&c,px/2); cudaMalloc(&out,px*4); cudaMemset(y,1,px); cudaMemset(c,2,px/2);
  run<4,1>(y,c,out,px); run<4,2>(y,c,out,px); run<4,4>(y,c,out,px); run<4,8>(y,c,out,px);
  run<8,1>(y,c,out,px); run<8,2>(y,c,out,px); run<8,4>(y,c,out,px);
  run<16,1>(y,c,out,px); run<16,2>(y,c,out,px);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
