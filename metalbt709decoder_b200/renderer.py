"""Host-buffer mirror of the reference's operator objects (include/metal_bt709_decoder.h).

Class and method names follow the reference so that tests read like its XCTest files:
  MetalRenderContext, MetalBT709Decoder.decodeBT709(...), MetalScaleRenderContext.renderScaled(...),
  BGRAToBT709Converter.convert/unconvert/copyBT709ToCoreVideo/createYCbCrFromBGRA,
  cvpbu_ycbcr_subsample, AAPLRenderer.drawFrame.
Pixel storage is host memory (pinned when a device is present), exposed as numpy views; all
arithmetic runs in the CUDA library.
"""
import ctypes

import numpy as np

from . import _lib

MetalBT709GammaApple, MetalBT709GammaSRGB, MetalBT709GammaLinear = 0, 1, 2
MetalBT709GammaBT709Spec = 3  # extension: the BT.709 spec curve of the reference's CPU path (BT709.h:668-738)
BT709GammaSrgb, BT709GammaApple, BT709GammaLinear = 1, 2, 3
BGRAToBT709ConverterSoftware, BGRAToBT709ConverterVImage, BGRAToBT709ConverterMetal = 0, 1, 2
kMBDYCbCrMatrix_Untagged, kMBDYCbCrMatrix_ITU_R_709_2, kMBDYCbCrMatrix_ITU_R_601_4 = 0, 1, 2
(kMBDTransferFunction_Untagged, kMBDTransferFunction_ITU_R_709_2, kMBDTransferFunction_sRGB,
 kMBDTransferFunction_Linear) = 0, 1, 2, 3


def _view(ptr, nbytes):
    buf = (ctypes.c_uint8 * nbytes).from_address(ptr)
    return np.frombuffer(buf, dtype=np.uint8)


class CVPixelBuffer:
    """CVPixelBufferRef stand-in (420YpCbCr8BiPlanarVideoRange) backed by MBDPixelBuffer."""

    def __init__(self, width=None, height=None, _ptr=None):
        self._L = _lib.lib()
        self._p = _ptr if _ptr is not None else self._L.BGRAToBT709Converter_createCoreVideoYCbCrBuffer(width, height)
        if not self._p:
            raise MemoryError("createCoreVideoYCbCrBuffer failed")

    @property
    def struct(self):
        return self._p.contents

    @property
    def width(self):
        return self.struct.width

    @property
    def height(self):
        return self.struct.height

    @property
    def y(self):
        s = self.struct
        return _view(s.y, s.yBytesPerRow * s.height).reshape(s.height, s.yBytesPerRow)[:, :s.width]

    @property
    def cbcr(self):
        s = self.struct
        ch, cw2 = (s.height + 1) // 2, ((s.width + 1) // 2) * 2
        return _view(s.cbcr, s.cbcrBytesPerRow * ch).reshape(ch, s.cbcrBytesPerRow)[:, :cw2]

    def setAttachments(self, matrix, transfer):
        self.struct.yCbCrMatrix = matrix
        self.struct.transferFunction = transfer

    def release(self):
        if self._p:
            self._L.MBDPixelBufferRelease(self._p)
            self._p = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


class BGRATexture:
    """id<MTLTexture> stand-in (BGRA8Unorm / BGRA8Unorm_sRGB) backed by MBDTexture."""

    def __init__(self, ptr):
        self._L = _lib.lib()
        self._p = ptr
        if not ptr:
            raise MemoryError("makeBGRATexture failed")

    @property
    def struct(self):
        return self._p.contents

    @property
    def width(self):
        return self.struct.width

    @property
    def height(self):
        return self.struct.height

    @property
    def pixels(self):
        """(h, w, 4) uint8 view, bytes B,G,R,A"""
        s = self.struct
        return _view(s.pixels, s.bytesPerRow * s.height).reshape(s.height, s.bytesPerRow)[:, :s.width * 4].reshape(
            s.height, s.width, 4)

    @property
    def words(self):
        """(h, w) uint32 view: A<<24 | R<<16 | G<<8 | B"""
        return self.pixels.view(np.uint32).reshape(self.height, self.width)

    def release(self):
        if self._p:
            self._L.MBDTextureRelease(self._p)
            self._p = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


class MetalRenderContext:
    def __init__(self, device=0):
        self._L = _lib.lib()
        self._h = ctypes.c_void_p(self._L.MetalRenderContext_new(int(device)))
        self.device = int(device)

    def setupMetal(self):
        return bool(self._L.MetalRenderContext_setupMetal(self._h))

    def makeBGRATexture(self, width, height, pixels=None, isSRGB=True):
        p = None
        if pixels is not None:
            pixels = np.ascontiguousarray(pixels, dtype=np.uint32)
            assert pixels.size == width * height
            p = pixels.ctypes.data_as(ctypes.c_void_p)
        return BGRATexture(self._L.MetalRenderContext_makeBGRATexture(self._h, width, height, p, int(isSRGB)))

    def getBGRATexturePixels(self, tex):
        """-getBGRATexturePixels: (MetalRenderContext.m:140-203): tightly packed uint32 copy"""
        return tex.words.copy()

    def waitUntilCompleted(self):
        return bool(self._L.MetalRenderContext_waitUntilCompleted(self._h))

    @property
    def launchCount(self):
        return int(self._L.MetalRenderContext_launchCount(self._h))

    def release(self):
        if self._h:
            self._L.MetalRenderContext_release(self._h)
            self._h = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


class MetalBT709Decoder:
    """MetalBT709Decoder.h:21-74"""

    def __init__(self):
        self._L = _lib.lib()
        self._h = ctypes.c_void_p(self._L.MetalBT709Decoder_new())
        self._mrc = None

    @property
    def metalRenderContext(self):
        return self._mrc

    @metalRenderContext.setter
    def metalRenderContext(self, mrc):
        self._mrc = mrc
        self._L.MetalBT709Decoder_setMetalRenderContext(self._h, mrc._h)

    @property
    def gamma(self):
        return self._L.MetalBT709Decoder_gamma(self._h)

    @gamma.setter
    def gamma(self, g):
        self._L.MetalBT709Decoder_setGamma(self._h, int(g))

    def setUseComputeRenderer(self, flag):
        self._L.MetalBT709Decoder_setUseComputeRenderer(self._h, int(flag))

    useComputeRenderer = property(None, setUseComputeRenderer)

    def setHasAlphaChannel(self, flag):
        self._L.MetalBT709Decoder_setHasAlphaChannel(self._h, int(flag))

    hasAlphaChannel = property(None, setHasAlphaChannel)

    def setupMetal(self):
        return bool(self._L.MetalBT709Decoder_setupMetal(self._h))

    def decodeBT709(self, yCbCrPixelBuffer, alphaPixelBuffer, bgraSRGBTexture, commandBuffer, renderWidth,
                    renderHeight, waitUntilCompleted):
        return bool(self._L.MetalBT709Decoder_decodeBT709(
            self._h, yCbCrPixelBuffer._p if yCbCrPixelBuffer is not None else None,
            alphaPixelBuffer._p if alphaPixelBuffer is not None else None,
            bgraSRGBTexture._p if bgraSRGBTexture is not None else None,
            ctypes.c_void_p(commandBuffer) if commandBuffer else None, renderWidth, renderHeight,
            int(waitUntilCompleted)))

    def release(self):
        if self._h:
            self._L.MetalBT709Decoder_release(self._h)
            self._h = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


class MetalScaleRenderContext:
    """MetalScaleRenderContext.h:20-42"""

    def __init__(self):
        self._L = _lib.lib()
        self._h = ctypes.c_void_p(self._L.MetalScaleRenderContext_new())

    def renderScaled(self, mrc, drawable, renderWidth, renderHeight, commandBuffer, bgraTexture,
                     waitUntilCompleted=True):
        return bool(self._L.MetalScaleRenderContext_renderScaled(
            self._h, mrc._h, drawable._p, renderWidth, renderHeight,
            ctypes.c_void_p(commandBuffer) if commandBuffer else None, bgraTexture._p, int(waitUntilCompleted)))

    def release(self):
        if self._h:
            self._L.MetalScaleRenderContext_release(self._h)
            self._h = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


def _u32(a):
    a = np.ascontiguousarray(a, dtype=np.uint32)
    return a, a.ctypes.data_as(ctypes.c_void_p)


class BGRAToBT709Converter:
    """Class-method facade of BGRAToBT709Converter.h:30-133 (every method takes the render context)."""

    @staticmethod
    def createCoreVideoYCbCrBuffer(width, height):
        return CVPixelBuffer(width, height)

    @staticmethod
    def setBT709Attributes(pb):
        return bool(_lib.lib().BGRAToBT709Converter_setBT709Attributes(pb._p))

    @staticmethod
    def convert(mrc, inBGRAPixels, width, height, type):
        """returns (worked, outBT709Pixels)"""
        src, sp = _u32(inBGRAPixels)
        out = np.zeros(width * height, np.uint32)
        ok = _lib.lib().BGRAToBT709Converter_convert(mrc._h, sp, out.ctypes.data_as(ctypes.c_void_p), width, height,
                                                     int(type))
        return bool(ok), out

    @staticmethod
    def unconvert(mrc, inBT709Pixels, width, height, type):
        src, sp = _u32(inBT709Pixels)
        out = np.zeros(width * height, np.uint32)
        ok = _lib.lib().BGRAToBT709Converter_unconvert(mrc._h, sp, out.ctypes.data_as(ctypes.c_void_p), width, height,
                                                       int(type))
        return bool(ok), out

    @staticmethod
    def copyBT709ToCoreVideo(mrc, inBT709Pixels, cvPixelBuffer):
        src, sp = _u32(inBT709Pixels)
        return bool(_lib.lib().BGRAToBT709Converter_copyBT709ToCoreVideo(mrc._h, sp, cvPixelBuffer._p))

    @staticmethod
    def createYCbCrFromBGRA(mrc, bgraPixels, width, height, isLinear=False, asSRGBGamma=False):
        src, sp = _u32(bgraPixels)
        p = _lib.lib().BGRAToBT709Converter_createYCbCrFromBGRA(mrc._h, sp, width, height, int(isLinear),
                                                                int(asSRGBGamma))
        return CVPixelBuffer(_ptr=p) if p else None


def cvpbu_ycbcr_subsample(mrc, inPixels, width, height, dst, inputGamma, outputGamma):
    src, sp = _u32(inPixels)
    return bool(_lib.lib().cvpbu_ycbcr_subsample(mrc._h, sp, width, height, dst._p, int(inputGamma), int(outputGamma)))


class AAPLRenderer:
    """The per-frame 1-pass / 2-pass policy of -[AAPLRenderer drawInMTKView:] (AAPLRenderer.m:874-987)."""

    def __init__(self, mrc, decodeGamma=MetalBT709GammaSRGB, hasAlphaChannel=False):
        self._L = _lib.lib()
        self._mrc = mrc
        self._h = ctypes.c_void_p(self._L.AAPLRenderer_new(mrc._h, int(decodeGamma), int(hasAlphaChannel)))
        if not self._h:
            raise RuntimeError("AAPLRenderer_new failed (no CUDA device?)")

    def drawFrame(self, yCbCr, alpha, drawable, waitUntilCompleted=True):
        return bool(self._L.AAPLRenderer_drawFrame(self._h, yCbCr._p, alpha._p if alpha is not None else None,
                                                   drawable._p, int(waitUntilCompleted)))

    def release(self):
        if self._h:
            self._L.AAPLRenderer_release(self._h)
            self._h = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass
