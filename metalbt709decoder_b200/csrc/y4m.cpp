// y4m.cpp — YUV4MPEG2 writer (byte-identical to the reference's Renderer/y4m_writer.h) and reader.
#include "../../include/y4m_b200.h"

#include <cstdlib>
#include <cstring>
#include <string>

extern "C" {

FILE *y4m_open_file(const char *outFilePath) {
  FILE *f = fopen(outFilePath, "wb");
  if (f == nullptr) fprintf(stderr, "could not open output Y4M file \"%s\"\n", outFilePath);
  return f;
}

static const char *fps_segment(int fps) {
  switch (fps) {  // y4m_writer.h:102-140
    case Y4MHeaderFPS_1: return "F1:1 ";
    case Y4MHeaderFPS_15: return "F15:1 ";
    case Y4MHeaderFPS_24: return "F24:1 ";
    case Y4MHeaderFPS_25: return "F25:1 ";
    case Y4MHeaderFPS_29_97: return "F30000:1001 ";
    case Y4MHeaderFPS_30: return "F30:1 ";
    case Y4MHeaderFPS_60: return "F60:1 ";
    default: return nullptr;
  }
}

int y4m_write_header(FILE *outFile, Y4MHeaderStruct *hs) {
  const char *fps = fps_segment((int) hs->fps);
  if (!fps) return 3;
  std::string h = "YUV4MPEG2 ";
  h += "W" + std::to_string(hs->width) + " ";
  h += "H" + std::to_string(hs->height) + " ";
  h += fps;
  h += "Ip ";            // progressive
  h += "A1:1 ";          // square pixels
  h += "C420jpeg\n";     // 4:2:0, centre-sited chroma
  h += "XYSCSS=420JPEG\n";  // the reference emits this comment as a line of its own
  return fwrite(h.data(), h.size(), 1, outFile) == 1 ? 0 : 2;
}

int y4m_write_frame(FILE *outFile, Y4MFrameStruct *fs) {
  if (fwrite("FRAME\n", 6, 1, outFile) != 1) return 2;
  if (fs->yLen > 0 && fwrite(fs->yPtr, (size_t) fs->yLen, 1, outFile) != 1) return 2;
  if (fs->uLen > 0 && fwrite(fs->uPtr, (size_t) fs->uLen, 1, outFile) != 1) return 2;
  if (fs->vLen > 0 && fwrite(fs->vPtr, (size_t) fs->vLen, 1, outFile) != 1) return 2;
  return 0;
}

static bool read_line(FILE *f, std::string *out, size_t limit = 4096) {
  out->clear();
  int c;
  while ((c = fgetc(f)) != EOF) {
    if (c == '\n') return true;
    out->push_back((char) c);
    if (out->size() > limit) return false;
  }
  return !out->empty();
}

int y4m_read_header(FILE *in, Y4MHeaderStruct *hs) {
  std::string line;
  if (!read_line(in, &line) || line.compare(0, 10, "YUV4MPEG2 ") != 0) return 1;
  int w = 0, h = 0;
  long num = 0, den = 0;
  std::string chroma = "420jpeg";
  size_t pos = 10;
  while (pos < line.size()) {
    size_t end = line.find(' ', pos);
    if (end == std::string::npos) end = line.size();
    const std::string tok = line.substr(pos, end - pos);
    pos = end + 1;
    if (tok.empty()) continue;
    const char *v = tok.c_str() + 1;
    switch (tok[0]) {
      case 'W': w = atoi(v); break;
      case 'H': h = atoi(v); break;
      case 'F': { char *e; num = strtol(v, &e, 10); den = (*e == ':') ? strtol(e + 1, nullptr, 10) : 1; break; }
      case 'C': chroma = v; break;
      default: break;  // I, A, X...
    }
  }
  if (w <= 0 || h <= 0 || (w & 1) || (h & 1) || chroma.compare(0, 3, "420") != 0) return 1;
  hs->width = w;
  hs->height = h;
  int fps = -1;
  const struct { long n, d; int e; } tab[] = {{1, 1, Y4MHeaderFPS_1}, {15, 1, Y4MHeaderFPS_15}, {24, 1, Y4MHeaderFPS_24},
                                             {25, 1, Y4MHeaderFPS_25}, {30000, 1001, Y4MHeaderFPS_29_97},
                                             {30, 1, Y4MHeaderFPS_30}, {60, 1, Y4MHeaderFPS_60}};
  for (const auto &t : tab) if (t.n == num && t.d == den) fps = t.e;
  if (fps < 0) return 3;  // a frame rate the writer's enum cannot express (same code as y4m_write_header's bad fps)
  hs->fps = (Y4MHeaderFPS) fps;
  // the reference's stray comment line, if this file came from its writer
  long at = ftell(in);
  if (at >= 0) {
    char buf[15];
    if (fread(buf, 1, 15, in) == 15 && memcmp(buf, "XYSCSS=420JPEG\n", 15) == 0) return 0;
    fseek(in, at, SEEK_SET);
  }
  return 0;
}

int y4m_read_frame(FILE *in, const Y4MHeaderStruct *hs, uint8_t *y, uint8_t *u, uint8_t *v) {
  std::string line;
  if (!read_line(in, &line)) return 1;
  if (line.compare(0, 5, "FRAME") != 0) return 2;
  const size_t ylen = (size_t) hs->width * hs->height, clen = (size_t) (hs->width / 2) * (hs->height / 2);
  if (fread(y, 1, ylen, in) != ylen) return 2;
  if (fread(u, 1, clen, in) != clen) return 2;
  if (fread(v, 1, clen, in) != clen) return 2;
  return 0;
}

}  // extern "C"
