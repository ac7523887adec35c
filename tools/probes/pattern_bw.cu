// Probe: which part of the decode kernel's memory pattern costs bandwidth?
// All variants move the decode byte mix for F frames of WxH: 1.5 B read + 4 B written per pixel.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("err %s line %d\n", cudaGetErrorString(e), __LINE__); return 1;}}while(0)
constexpr int W=1920,H=1080;
__device__ __forceinline__ uint32_t ldcs(const void*p){uint32_t v; asm volatile("ld.global.cs.u32 %0,[%1];":"=r"(v):"l"(p)); return v;}
__device__ __forceinline__ void stcs(void*p, uint4 v){asm volatile("st.global.cs.v4.u32 [%0],{%1,%2,%3,%4};"::"l"(p),"r"(v.x),"r"(v.y),"r"(v.z),"r"(v.w):"memory");}

// MODE 0: linear mix: thread handles 4 px: reads 4 B Y + (every other) 4 B C, writes 16 B.  grid-stride over all pixels/4.
// MODE 1: band layout like k_decode_fast (128px x 2 rows per warp, CTA band of 2 row pairs), no prefetch, loads then stores
// MODE 2: band layout, writes only
// MODE 3: band layout, reads only (results xor-ed into a sink)
// MODE 4: band layout with one-step prefetch (like the real kernel)
template<int MODE> __global__ void __launch_bounds__(1024,1) k(const uint8_t* y, const uint8_t* c, uint8_t* out, int frames, uint32_t* sink) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (MODE == 0) {
    size_t n4 = (size_t)frames*W*H/4; size_t i = (size_t)blockIdx.x*blockDim.x+threadIdx.x, st=(size_t)gridDim.x*blockDim.x;
    for (; i<n4; i+=st) {
      uint32_t a = ldcs(y+i*4); uint32_t b = (i&1)? 0u : ldcs(c+(i>>1)*4);
      stcs(out+i*16, make_uint4(a,b,a+1,b+1));
    }
    return;
  }
  const int tiles = W/128, rps = 32/tiles;  // 15, 2
  const int r_local = warp / tiles; if (r_local >= rps) return; const int tile = warp - r_local*tiles;
  const int rows2 = H/2; const unsigned total = (unsigned)frames*rows2;
  const int col = tile*128 + lane*4;
  uint32_t acc = 0;
  uint32_t py0=0, py1=0, pc=0; bool have=false;
  for (unsigned rpg = blockIdx.x*rps + r_local; rpg < total; rpg += gridDim.x*rps) {
    unsigned f = rpg / rows2, rp = rpg - f*rows2;
    const uint8_t* yp = y + (size_t)f*W*H + (size_t)(2*rp)*W + col;
    const uint8_t* cp = c + (size_t)f*W*H/2 + (size_t)rp*W + col;
    uint8_t* op = out + (size_t)f*W*H*4 + (size_t)(2*rp)*W*4 + (size_t)col*4;
    uint32_t y0=0,y1=0,cc=0;
    if (MODE==1 || MODE==3) { y0=ldcs(yp); y1=ldcs(yp+W); cc=ldcs(cp); }
    if (MODE==4) {
      if (!have) { py0=ldcs(yp); py1=ldcs(yp+W); pc=ldcs(cp); }
      y0=py0; y1=py1; cc=pc;
      unsigned nr = rpg + gridDim.x*rps;
      if (nr < total) { unsigned nf=nr/rows2, nrp=nr-nf*rows2; const uint8_t* nyp = y+(size_t)nf*W*H+(size_t)(2*nrp)*W+col; const uint8_t* ncp=c+(size_t)nf*W*H/2+(size_t)nrp*W+col; py0=ldcs(nyp); py1=ldcs(nyp+W); pc=ldcs(ncp); have=true; }
    }
    if (MODE==3) { acc ^= y0^y1^cc; }
    else { stcs(op, make_uint4(y0,cc,y0+1,cc+1)); stcs(op+(size_t)W*4, make_uint4(y1,cc,y1+1,cc+2)); }
  }
  if (MODE==3 && acc==0x1234567) *sink=acc;
}
template<class F> float timeit(F f, int n=10){ for(int i=0;i<2;i++) f(); cudaDeviceSynchronize(); cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a); for(int i=0;i<n;i++) f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms,a,b); return ms/n/1e3f; }
int main(){
  int frames=600; size_t px=(size_t)frames*W*H;
  uint8_t *y,*c,*out; uint32_t* sink; CK(cudaMalloc(&y,px)); CK(cudaMalloc(&c,px/2)); CK(cudaMalloc(&out,px*4)); CK(cudaMalloc(&sink,4));
  cudaMemset(y,1,px); cudaMemset(c,2,px/2);
  for (int grid : {148, 296, 148*8}) {
    float t0=timeit([&]{k<0><<<grid,1024>>>(y,c,out,frames,sink);});
    printf("linear mix grid %d: %.0f Gpix/s  %.0f GB/s\n", grid, px/t0/1e9, px*5.5/t0/1e9);
  }
  float t1=timeit([&]{k<1><<<148,1024>>>(y,c,out,frames,sink);});
  printf("band no-prefetch: %.0f Gpix/s %.0f GB/s\n", px/t1/1e9, px*5.5/t1/1e9);
  float t2=timeit([&]{k<2><<<148,1024>>>(y,c,out,frames,sink);});
  printf("band writes only: %.0f Gpix/s %.0f GB/s (write)\n", px/t2/1e9, px*4.0/t2/1e9);
  float t3=timeit([&]{k<3><<<148,1024>>>(y,c,out,frames,sink);});
  printf("band reads only: %.0f Gpix/s %.0f GB/s (read)\n", px/t3/1e9, px*1.5/t3/1e9);
  float t4=timeit([&]{k<4><<<148,1024>>>(y,c,out,frames,sink);});
  printf("band prefetch: %.0f Gpix/s %.0f GB/s\n", px/t4/1e9, px*5.5/t4/1e9);
  CK(cudaGetLastError());
  return 0;
}
