// nvdec_caps.cpp — is NVDEC usable on this box?  dlopen(libnvcuvid.so.1) + cuvidGetDecoderCaps for H.264 4:2:0 8-bit.
// The Video Codec SDK headers are not in the image: the struct below restates CUVIDDECODECAPS of nvcuvid.h / cuviddec.h
// (NVIDIA Video Codec SDK 12.x).  build: g++ -O2 -o nvdec_caps nvdec_caps.cpp -I/usr/local/cuda/include -ldl
#include <cuda.h>
#include <dlfcn.h>
#include <cstdio>
#include <cstring>
#include <initializer_list>

struct Caps {               // CUVIDDECODECAPS
  int eCodecType;           // cudaVideoCodec: H264 = 4
  int eChromaFormat;        // cudaVideoChromaFormat: 420 = 1
  unsigned int nBitDepthMinus8;
  unsigned int reserved1[3];
  unsigned char bIsSupported;
  unsigned char nNumNVDECs;
  unsigned short nOutputFormatMask;
  unsigned int nMaxWidth, nMaxHeight, nMaxMBCount;
  unsigned short nMinWidth, nMinHeight;
  unsigned char bIsHistogramSupported, nCounterBitDepth;
  unsigned short nMaxHistogramBins;
  unsigned int reserved3[10];
};

int main() {
  void *cu = dlopen("libcuda.so.1", RTLD_NOW);
  if (!cu) { printf("no libcuda: %s\n", dlerror()); return 1; }
  auto pInit = (CUresult (*)(unsigned)) dlsym(cu, "cuInit");
  auto pDevGet = (CUresult (*)(CUdevice *, int)) dlsym(cu, "cuDeviceGet");
  auto pRetain = (CUresult (*)(CUcontext *, CUdevice)) dlsym(cu, "cuDevicePrimaryCtxRetain");
  auto pSet = (CUresult (*)(CUcontext)) dlsym(cu, "cuCtxSetCurrent");
  CUdevice dev; CUcontext ctx;
  printf("cuInit %d\n", pInit(0));
  printf("cuDeviceGet %d\n", pDevGet(&dev, 0));
  printf("retain %d\n", pRetain(&ctx, dev));
  printf("set %d\n", pSet(ctx));
  const char *names[] = {"libnvcuvid.so.1", "/usr/lib/libnvcuvid.so.1", "/usr/local/nvidia/lib/libnvcuvid.so"};
  void *nv = nullptr;
  for (const char *n : names) { nv = dlopen(n, RTLD_NOW); printf("dlopen %s: %s\n", n, nv ? "ok" : dlerror()); if (nv) break; }
  if (!nv) return 2;
  auto pCaps = (CUresult (*)(Caps *)) dlsym(nv, "cuvidGetDecoderCaps");
  printf("cuvidGetDecoderCaps sym %p, cuvidCreateVideoParser %p, cuvidCreateDecoder %p, cuvidMapVideoFrame64 %p\n", (void *) pCaps,
         dlsym(nv, "cuvidCreateVideoParser"), dlsym(nv, "cuvidCreateDecoder"), dlsym(nv, "cuvidMapVideoFrame64"));
  if (!pCaps) return 3;
  for (int codec : {4 /*H264*/, 8 /*HEVC*/}) {
    Caps c; memset(&c, 0, sizeof c);
    c.eCodecType = codec; c.eChromaFormat = 1; c.nBitDepthMinus8 = 0;
    CUresult r = pCaps(&c);
    printf("codec %d: rc %d supported %d nvdecs %d outmask 0x%x max %ux%u mb %u min %ux%u\n", codec, (int) r, c.bIsSupported, c.nNumNVDECs,
           c.nOutputFormatMask, c.nMaxWidth, c.nMaxHeight, c.nMaxMBCount, c.nMinWidth, c.nMinHeight);
  }
  return 0;
}
