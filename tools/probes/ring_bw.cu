// Probe: decode memory pattern (CTA bands, 128 px x 2 rows per warp-step) with the inputs staged
// through a per-warp shared-memory ring filled by one 16-byte cp.async per lane (24 lanes) per
// step, D steps deep.  No arithmetic: this measures what the memory pattern alone can sustain.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
constexpr int W=1920,H=1080;
__device__ __forceinline__ void stcs(void*p, uint4 v){asm volatile("st.global.cs.v4.u32 [%0],{%1,%2,%3,%4};"::"l"(p),"r"(v.x),"r"(v.y),"r"(v.z),"r"(v.w):"memory");}
template<int D, int ROWPAIRS, int BLOCKED> __global__ void __launch_bounds__(1024,1) k(const uint8_t* y, const uint8_t* c, uint8_t* out, int frames) {
  extern __shared__ __align__(128) uint8_t ring[];   // [warp][stage][384*ROWPAIRS]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tiles = W/128, rps = 32/tiles;  // 15, 2
  const int r_local = warp / tiles; if (r_local >= rps) return; const int tile = warp - r_local*tiles;
  const int rows2 = H/2; const unsigned total = (unsigned)frames*rows2 / ROWPAIRS;   // units of ROWPAIRS row pairs
  constexpr int STAGE = 384*ROWPAIRS;
  const uint32_t myring = (uint32_t)__cvta_generic_to_shared(ring) + warp*D*STAGE;
  // lane role for the copy: lanes 0-7 row0, 8-15 row1, 16-23 chroma (per row pair)
  const int role = lane >> 3, sub = lane & 7;
  const unsigned step = gridDim.x*rps;
  auto src_of = [&](unsigned u, int q) -> const uint8_t* {   // unit u, row pair q within unit
    unsigned rpg = u*ROWPAIRS + q; unsigned f = rpg / rows2, rp = rpg - f*rows2;
    if (role < 2) return y + (size_t)f*W*H + (size_t)(2*rp+role)*W + tile*128 + sub*16;
    return c + (size_t)f*W*H/2 + (size_t)rp*W + tile*128 + sub*16;
  };
  auto issue = [&](unsigned u, int slot) {
    if (role < 3) {
      #pragma unroll
      for (int q=0;q<ROWPAIRS;q++) {
        uint32_t dst = myring + slot*STAGE + q*384 + role*128 + sub*16;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src_of(u,q)) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  unsigned u0 = blockIdx.x*rps + r_local;
  unsigned nsteps = u0 < total ? (total - u0 + step - 1)/step : 0;
  unsigned ustep = step;
  if (BLOCKED) {   // each CTA owns one contiguous run of bands
    unsigned bands = (total + rps - 1)/rps, per = (bands + gridDim.x - 1)/gridDim.x;
    unsigned b0 = blockIdx.x*per, b1 = min(b0+per, bands);
    u0 = b0*rps + r_local; ustep = rps; nsteps = b1 > b0 ? b1 - b0 : 0;
    if (nsteps && u0 + (nsteps-1)*ustep >= total) nsteps--;
  }
  for (int j=0;j<D-1;j++) { if ((unsigned)j<nsteps) issue(u0 + j*ustep, j); else asm volatile("cp.async.commit_group;" ::: "memory"); }
  for (unsigned k=0;k<nsteps;k++) {
    if (k + D - 1 < nsteps) issue(u0 + (k+D-1)*ustep, (k+D-1)%D); else asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group %0;" :: "n"(D-1) : "memory");
    __syncwarp();
    const int slot = k % D;
    #pragma unroll
    for (int q=0;q<ROWPAIRS;q++) {
      uint32_t base = myring + slot*STAGE + q*384 + lane*4;
      uint32_t y0,y1,cc;
      asm volatile("ld.shared.u32 %0,[%1];":"=r"(y0):"r"(base)); asm volatile("ld.shared.u32 %0,[%1];":"=r"(y1):"r"(base+128)); asm volatile("ld.shared.u32 %0,[%1];":"=r"(cc):"r"(base+256));
      unsigned rpg = (u0 + k*ustep)*ROWPAIRS + q; unsigned f = rpg/rows2, rp = rpg - f*rows2;
      uint8_t* op = out + (size_t)f*W*H*4 + (size_t)(2*rp)*W*4 + (size_t)(tile*128+lane*4)*4;
      stcs(op, make_uint4(y0,cc,y0+1,cc+1)); stcs(op+(size_t)W*4, make_uint4(y1,cc,y1+1,cc+2));
    }
    __syncwarp();
  }
}
template<class F> float timeit(F f, int n=10){ for(int i=0;i<2;i++) f(); cudaDeviceSynchronize(); cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a); for(int i=0;i<n;i++) f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms,a,b); return ms/n/1e3f; }
template<int D,int RP,int BL> void run(const uint8_t*y,const uint8_t*c,uint8_t*out,int frames,size_t px){
  size_t smem = (size_t)32*D*384*RP;
  cudaFuncSetAttribute(k<D,RP,BL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  float t=timeit([&]{k<D,RP,BL><<<148,1024,smem>>>(y,c,out,frames);});
  printf("blocked=%d ring depth %d, row pairs per step %d (smem %zu KB): %5.0f Gpix/s %5.0f GB/s  %s\n", BL, D, RP, smem/1024, px/t/1e9, px*5.5/t/1e9, cudaGetErrorString(cudaGetLastError()));
}
int main(){
  int frames=600; size_t px=(size_t)frames*W*H;
  uint8_t *y,*c,*out; cudaMalloc(&y,px); cudaMalloc(&c,px/2); cudaMalloc(&out,px*4); cudaMemset(y,1,px); cudaMemset(c,2,px/2);
  run<2,1,0>(y,c,out,frames,px); run<3,1,0>(y,c,out,frames,px); run<4,1,0>(y,c,out,frames,px);
  run<2,1,1>(y,c,out,frames,px); run<3,1,1>(y,c,out,frames,px); run<4,1,1>(y,c,out,frames,px); run<8,1,1>(y,c,out,frames,px);
}
