// Probe (round 2): can a ONE-ROW unit layout with real NV12 addressing beat the 2-row-tile layouts?  Round 1 measured
// (profiles/r01_membw_probe.json): 2-row tiles + cp.async ring, no arithmetic: 1006 Gpix/s; a linear one-row stream
// with 4 units of register loads in flight: 1123 Gpix/s - but that probe read the chroma plane linearly, which NV12
// does not allow.  Here a unit is 4 pixels of one row: LDG.32 of Y, LDG.32 of the CbCr word of its row PAIR (so the
// chroma row is read twice, the second time from L2), one STG.128.  Units in flight per thread: 1, 2, 4, 6, 8.
// Second family: 2-row units (Y row 0, Y row 1, CbCr word: 3 loads, two STG.128) with the same register scheme.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ void st16(void*p, uint4 v){asm volatile("st.global.cs.v4.u32 [%0],{%1,%2,%3,%4};"::"l"(p),"r"(v.x),"r"(v.y),"r"(v.z),"r"(v.w):"memory");}
__device__ __forceinline__ uint32_t ldcs(const void*p){uint32_t v; asm volatile("ld.global.cs.u32 %0,[%1];":"=r"(v):"l"(p)); return v;}
template<int UN> __global__ void __launch_bounds__(1024) k1(const uint8_t* y, const uint8_t* c, uint8_t* out, unsigned w4, size_t units) {
  const size_t nthreads = (size_t)gridDim.x*blockDim.x;
  for (size_t u0 = (size_t)blockIdx.x*blockDim.x + threadIdx.x; u0 < units; u0 += nthreads*UN) {
    uint32_t a[UN], b[UN];
    #pragma unroll
    for (int j=0;j<UN;j++){ size_t u=u0+j*nthreads; if(u<units){ size_t row=u/w4; unsigned col=u-row*w4; a[j]=ldcs(y+u*4); b[j]=ldcs(c+((row>>1)*w4+col)*4); } }
    #pragma unroll
    for (int j=0;j<UN;j++){ size_t u=u0+j*nthreads; if(u<units) st16(out+u*16, make_uint4(a[j], b[j], a[j]^b[j], 4)); }
  }
}
// 2-row units: u indexes (row pair, col4)
template<int UN> __global__ void __launch_bounds__(1024) k2(const uint8_t* y, const uint8_t* c, uint8_t* out, unsigned w4, size_t units) {
  const size_t nthreads = (size_t)gridDim.x*blockDim.x;
  for (size_t u0 = (size_t)blockIdx.x*blockDim.x + threadIdx.x; u0 < units; u0 += nthreads*UN) {
    uint32_t a[UN], a2[UN], b[UN];
    #pragma unroll
    for (int j=0;j<UN;j++){ size_t u=u0+j*nthreads; if(u<units){ size_t rp=u/w4; unsigned col=u-rp*w4; a[j]=ldcs(y+((2*rp)*w4+col)*4); a2[j]=ldcs(y+((2*rp+1)*w4+col)*4); b[j]=ldcs(c+u*4); } }
    #pragma unroll
    for (int j=0;j<UN;j++){ size_t u=u0+j*nthreads; if(u<units){ size_t rp=u/w4; unsigned col=u-rp*w4;
        st16(out+((2*rp)*w4+col)*16, make_uint4(a[j], b[j], 3, 4)); st16(out+((2*rp+1)*w4+col)*16, make_uint4(a2[j], b[j], 3, 4)); } }
  }
}
template<class F> float timeit(F f, int n=10){ for(int i=0;i<2;i++) f(); cudaDeviceSynchronize(); cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a); for(int i=0;i<n;i++) f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms,a,b); return ms/n/1e3f; }
template<int UN> void run1(const uint8_t*y,const uint8_t*c,uint8_t*out,unsigned w,size_t px,int grid,int block){
  float t=timeit([&]{k1<UN><<<grid,block>>>(y,c,out,w/4,px/4);});
  printf("1-row units, %d in flight, grid %4d x %4d : %5.0f Gpix/s %5.0f GB/s\n", UN, grid, block, px/t/1e9, px*5.5/t/1e9);
}
template<int UN> void run2(const uint8_t*y,const uint8_t*c,uint8_t*out,unsigned w,size_t px,int grid,int block){
  float t=timeit([&]{k2<UN><<<grid,block>>>(y,c,out,w/4,px/8);});
  printf("2-row units, %d in flight, grid %4d x %4d : %5.0f Gpix/s %5.0f GB/s\n", UN, grid, block, px/t/1e9, px*5.5/t/1e9);
}
int main(){
  const unsigned w=1920; size_t px=(size_t)600*1920*1080;
  uint8_t *y,*c,*out; cudaMalloc(&y,px); cudaMalloc(&c,px/2); cudaMalloc(&out,px*4); cudaMemset(y,1,px); cudaMemset(c,2,px/2);
  run1<1>(y,c,out,w,px,148,1024); run1<2>(y,c,out,w,px,148,1024); run1<4>(y,c,out,w,px,148,1024); run1<6>(y,c,out,w,px,148,1024); run1<8>(y,c,out,w,px,148,1024);
  run1<2>(y,c,out,w,px,296,1024); run1<4>(y,c,out,w,px,296,1024); run1<4>(y,c,out,w,px,296,512); run1<4>(y,c,out,w,px,592,512);
  run2<1>(y,c,out,w,px,148,1024); run2<2>(y,c,out,w,px,148,1024); run2<3>(y,c,out,w,px,148,1024); run2<4>(y,c,out,w,px,148,1024);
  run2<2>(y,c,out,w,px,296,1024); run2<2>(y,c,out,w,px,296,512);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
