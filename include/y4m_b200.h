/*
 * y4m_b200.h — YUV4MPEG2 (.y4m) writer and reader for 8-bit 4:2:0 planar frames.
 *
 * Writer: same structs, enum and function names as the reference's header-only
 * Renderer/y4m_writer.h (which needs NSString and cannot be compiled off-Apple), and a
 * byte-identical stream:
 *   "YUV4MPEG2 W<w> H<h> F<num>:<den> Ip A1:1 C420jpeg\n" "XYSCSS=420JPEG\n"     (y4m_writer.h:61-192)
 *   per frame  "FRAME\n" + Y (w*h) + U (w/2*h/2) + V (w/2*h/2)                    (y4m_writer.h:194-241)
 * (the XYSCSS comment really is written as a second line by the reference; readers skip it).
 * Reader: the inverse, for round trips and for feeding .y4m files to the decode path.
 */
#ifndef Y4M_B200_H
#define Y4M_B200_H

#include <stdint.h>
#include <stdio.h>

#include "bt709_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef enum { /* y4m_writer.h:17-25 */
  Y4MHeaderFPS_1,
  Y4MHeaderFPS_15,
  Y4MHeaderFPS_24,
  Y4MHeaderFPS_25,
  Y4MHeaderFPS_29_97,
  Y4MHeaderFPS_30,
  Y4MHeaderFPS_60
} Y4MHeaderFPS;

typedef struct { /* y4m_writer.h:27-31 */
  int width;
  int height;
  Y4MHeaderFPS fps;
} Y4MHeaderStruct;

typedef struct { /* y4m_writer.h:35-44 */
  uint8_t *yPtr;
  int yLen;
  uint8_t *uPtr;
  int uLen;
  uint8_t *vPtr;
  int vLen;
} Y4MFrameStruct;

/* y4m_writer.h:48-57, 61-192, 194-241.  Return codes as in the reference: 0 ok, 2 short write, 3 bad fps. */
BT709_API FILE *y4m_open_file(const char *outFilePath);
BT709_API int y4m_write_header(FILE *outFile, Y4MHeaderStruct *hsPtr);
BT709_API int y4m_write_frame(FILE *outFile, Y4MFrameStruct *fsPtr);

/* Reader.  y4m_read_header parses the stream header (and the reference's XYSCSS line if present) and
 * fills width/height/fps; returns 0, 1 on a malformed header / unsupported chroma format, or 3 when the
 * frame rate is not one of the enum's (the writer's own bad-fps code).  y4m_read_frame reads one FRAME into caller buffers of
 * w*h, (w/2)*(h/2), (w/2)*(h/2) bytes; returns 0, 1 at end of file, 2 on a truncated frame. */
BT709_API int y4m_read_header(FILE *inFile, Y4MHeaderStruct *hsPtr);
BT709_API int y4m_read_frame(FILE *inFile, const Y4MHeaderStruct *hsPtr, uint8_t *y, uint8_t *u, uint8_t *v);

/* NV12 <-> planar on the device (the +copyYCBCr: step of BGRAToBT709Converter.m:688-790, which the
 * reference does in a CPU loop): split the interleaved CbCr plane into U and V planes, or merge them.
 * cw, ch = chroma width and height in samples (w/2, h/2). */
BT709_API int bt709_cbcr_to_planar(const bt709_ctx *ctx, const bt709_plane *cbcr, const bt709_plane *u,
                                   const bt709_plane *v, int cw, int ch, int nframes, void *stream, int wait);
BT709_API int bt709_planar_to_cbcr(const bt709_ctx *ctx, const bt709_plane *u, const bt709_plane *v,
                                   const bt709_plane *cbcr, int cw, int ch, int nframes, void *stream, int wait);

#ifdef __cplusplus
}
#endif
#endif
