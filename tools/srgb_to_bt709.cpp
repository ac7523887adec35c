// srgb_to_bt709 — command-line twin of the reference's srgb_to_bt709/srgb_to_bt709.m:
//   srgb_to_bt709 ?OPTIONS? OUTPUT.y4m
//   -frame F.png | -frames F0001.png | -gamma apple|srgb|linear | -alpha 0|1 | -fps 1|15|24|25|2997|30|60
// Same options, defaults, messages and exit codes (srgb_to_bt709.m:74-82, 1217-1323); the BT.709
// 4:2:0 encode runs on the GPU through the host layer (include/metal_bt709_decoder.h) and the frames are
// written with the byte-identical .y4m writer (include/y4m_b200.h).
//
// Out of the hot path, restated only as far as needed: image input is 8-bit non-interlaced PNG
// (gray / RGB / RGBA / gray+alpha, decoded here with zlib) or binary PPM (P6); it is taken to be sRGB
// (the reference's "input is sRGB colorspace" case; ColorSync conversion of other profiles is not
// reproduced).  Pixels with alpha are premultiplied, as CoreGraphics renders them into the reference's
// framebuffer (srgb_to_bt709.m:805-810), with round-to-nearest.
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../include/metal_bt709_decoder.h"
#include "../include/y4m_b200.h"

static void usage() {  // srgb_to_bt709.m:74-82
  printf("srgb_to_bt709 ?OPTIONS? OUTPUT.y4m\n");
  printf("OPTIONS:\n");
  printf("-frame F.png (input is a single frame)\n");
  printf("-frames F0001.png (first frame of N input frames)\n");
  printf("-gamma apple|srgb|linear (default is apple)\n");
  printf("-fps 1|15|24|25|2997|30|60 (default to 30 with -frames)\n");
  fflush(stdout);
}

struct Image {
  int w = 0, h = 0;
  bool has_alpha = false;
  std::vector<uint32_t> px;  // A<<24 | R<<16 | G<<8 | B, straight (not premultiplied) alpha
};

static bool read_file(const std::string &path, std::vector<uint8_t> *out) {
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) return false;
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  out->resize(n > 0 ? (size_t) n : 0);
  bool ok = n >= 0 && fread(out->data(), 1, out->size(), f) == out->size();
  fclose(f);
  return ok;
}

static uint32_t be32(const uint8_t *p) { return (uint32_t) p[0] << 24 | (uint32_t) p[1] << 16 | (uint32_t) p[2] << 8 | p[3]; }

static bool decode_png(const std::vector<uint8_t> &d, Image *img) {
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
  if (d.size() < 33 || memcmp(d.data(), sig, 8) != 0) return false;
  size_t pos = 8;
  int w = 0, h = 0, depth = 0, ctype = -1, interlace = 0;
  std::vector<uint8_t> idat;
  while (pos + 12 <= d.size()) {
    const uint32_t len = be32(&d[pos]);
    const char *type = (const char *) &d[pos + 4];
    if (pos + 12 + len > d.size()) return false;
    const uint8_t *body = &d[pos + 8];
    if (!memcmp(type, "IHDR", 4)) {
      if (len != 13) return false;  // a short IHDR would be read past its end
      w = (int) be32(body); h = (int) be32(body + 4); depth = body[8]; ctype = body[9]; interlace = body[12];
    } else if (!memcmp(type, "IDAT", 4)) {
      idat.insert(idat.end(), body, body + len);
    } else if (!memcmp(type, "IEND", 4)) {
      break;
    }
    pos += 12 + len;
  }
  int ch = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
  if (w <= 0 || h <= 0 || depth != 8 || ch == 0 || interlace != 0) {
    fprintf(stderr, "unsupported PNG (need 8-bit gray/RGB/RGBA, non-interlaced, not indexed)\n");
    return false;
  }
  // sanity bound before allocating from header fields (16384 x 16384 is far beyond any frame this tool converts)
  if (w > 16384 || h > 16384) {
    fprintf(stderr, "unsupported PNG (dimensions %d x %d too large)\n", w, h);
    return false;
  }
  const size_t stride = (size_t) w * ch;
  std::vector<uint8_t> raw((stride + 1) * h);
  uLongf rawlen = raw.size();
  if (uncompress(raw.data(), &rawlen, idat.data(), idat.size()) != Z_OK || rawlen != raw.size()) return false;
  std::vector<uint8_t> prev(stride, 0), cur(stride);
  img->w = w; img->h = h; img->has_alpha = (ch == 2 || ch == 4);
  img->px.resize((size_t) w * h);
  for (int y = 0; y < h; y++) {
    const uint8_t *line = &raw[(stride + 1) * y];
    const int ft = line[0];
    for (size_t i = 0; i < stride; i++) {
      const int a = i >= (size_t) ch ? cur[i - ch] : 0, b = prev[i], c = i >= (size_t) ch ? prev[i - ch] : 0;
      int pred = 0;
      switch (ft) {
        case 0: pred = 0; break;
        case 1: pred = a; break;
        case 2: pred = b; break;
        case 3: pred = (a + b) >> 1; break;
        case 4: { int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
                  pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); break; }
        default: return false;
      }
      cur[i] = (uint8_t) (line[1 + i] + pred);
    }
    for (int x = 0; x < w; x++) {
      const uint8_t *p = &cur[(size_t) x * ch];
      uint32_t r, g, b, a = 255;
      if (ch == 1) { r = g = b = p[0]; }
      else if (ch == 2) { r = g = b = p[0]; a = p[1]; }
      else { r = p[0]; g = p[1]; b = p[2]; if (ch == 4) a = p[3]; }
      img->px[(size_t) y * w + x] = a << 24 | r << 16 | g << 8 | b;
    }
    prev.swap(cur);
  }
  return true;
}

static bool decode_ppm(const std::vector<uint8_t> &d, Image *img) {
  if (d.size() < 11 || d[0] != 'P' || d[1] != '6') return false;
  size_t pos = 2;
  int vals[3], n = 0;
  while (n < 3 && pos < d.size()) {
    while (pos < d.size() && (d[pos] == ' ' || d[pos] == '\n' || d[pos] == '\r' || d[pos] == '\t')) pos++;
    if (pos < d.size() && d[pos] == '#') { while (pos < d.size() && d[pos] != '\n') pos++; continue; }
    int v = 0; bool any = false;
    while (pos < d.size() && d[pos] >= '0' && d[pos] <= '9') { v = v * 10 + (d[pos++] - '0'); any = true; }
    if (!any) return false;
    vals[n++] = v;
  }
  pos++;  // single whitespace after maxval
  if (n != 3 || vals[2] != 255 || d.size() < pos + (size_t) vals[0] * vals[1] * 3) return false;
  img->w = vals[0]; img->h = vals[1]; img->has_alpha = false;
  img->px.resize((size_t) img->w * img->h);
  for (size_t i = 0; i < img->px.size(); i++) {
    const uint8_t *p = &d[pos + i * 3];
    img->px[i] = 0xFFu << 24 | (uint32_t) p[0] << 16 | (uint32_t) p[1] << 8 | p[2];
  }
  return true;
}

static bool load_image(const std::string &path, Image *img) {
  std::vector<uint8_t> d;
  if (!read_file(path, &d)) {
    fprintf(stderr, "can't read image data from file \"%s\"\n", path.c_str());  // srgb_to_bt709.m:92-95
    exit(1);
  }
  return decode_png(d, img) || decode_ppm(d, img);
}

static bool file_exists(const std::string &p) {
  FILE *f = fopen(p.c_str(), "rb");
  if (f) fclose(f);
  return f != nullptr;
}

// srgb_to_bt709.m:555-663: F0001.png, F0002.png ... until the first missing file
static std::vector<std::string> frames_from_first(const std::string &first) {
  if (!file_exists(first)) {
    fprintf(stderr, "error: first filename \"%s\" does not exist\n", first.c_str());
    exit(1);
  }
  const size_t slash = first.find_last_of('/');
  const std::string dir = slash == std::string::npos ? "" : first.substr(0, slash + 1);
  std::string tail = slash == std::string::npos ? first : first.substr(slash + 1);
  const size_t dot = tail.find_last_of('.');
  const std::string ext = dot == std::string::npos ? "" : tail.substr(dot);
  const std::string stem = dot == std::string::npos ? tail : tail.substr(0, dot);
  size_t start = stem.size();
  while (start > 1 && stem[start - 1] >= '0' && stem[start - 1] <= '9') start--;
  if (start == stem.size() || start == 0) {
    fprintf(stderr, "error: could not find frame number in first filename \"%s\"\n", first.c_str());
    exit(1);
  }
  const std::string name = stem.substr(0, start), num = stem.substr(start);
  const int width = (int) num.size();
  std::vector<std::string> out;
  for (int i = atoi(num.c_str()); i < 9999999; i++) {
    char buf[32];
    snprintf(buf, sizeof buf, "%0*d", width, i);
    const std::string p = dir + name + buf + ext;
    if (!file_exists(p)) break;
    out.push_back(p);
  }
  return out;
}

int main(int argc, const char **argv) {
  if (argc < 4) {  // srgb_to_bt709.m:1196-1199
    usage();
    exit(1);
  }
  const char *in = nullptr, *out = nullptr;
  bool frames_pattern = false, is_alpha = false;
  std::string gamma = "apple";
  int fps = Y4MHeaderFPS_30;
  for (int i = 1; i < argc;) {
    const char *arg = argv[i];
    if (arg[0] == '-') {
      if (i + 1 >= argc) { usage(); exit(1); }
      const char *val = argv[i + 1];
      if (!strcmp(arg, "-alpha")) {
        if (!strcmp(val, "1")) is_alpha = true;
        else if (!strcmp(val, "0")) is_alpha = false;
        else { printf("option -alpha unknown value \"%s\"", val); exit(3); }
      } else if (!strcmp(arg, "-gamma")) {
        if (!strcmp(val, "apple") || !strcmp(val, "srgb") || !strcmp(val, "linear")) gamma = val;
        else { printf("option -gamma unknown value \"%s\"", val); exit(3); }
      } else if (!strcmp(arg, "-fps")) {
        if (!strcmp(val, "1")) fps = Y4MHeaderFPS_1;
        else if (!strcmp(val, "15")) fps = Y4MHeaderFPS_15;
        else if (!strcmp(val, "24")) fps = Y4MHeaderFPS_24;
        else if (!strcmp(val, "25")) fps = Y4MHeaderFPS_25;
        else if (!strcmp(val, "2997")) fps = Y4MHeaderFPS_29_97;
        else if (!strcmp(val, "30")) fps = Y4MHeaderFPS_30;
        else if (!strcmp(val, "60")) fps = Y4MHeaderFPS_60;
        else { printf("option -fps unknown value \"%s\"", val); exit(3); }
      } else if (!strcmp(arg, "-frame") || !strcmp(arg, "-frames")) {
        if (in) { printf("%s filename \"%s\" must appear just once\n", arg, val); exit(3); }
        in = val;
        if (!strcmp(arg, "-frame")) fps = Y4MHeaderFPS_1;  // one frame shown for one second
        else frames_pattern = true;
      } else {
        printf("unknown option \"%s\"\n", arg);
        exit(3);
      }
      i += 2;
    } else {
      if (out) { printf("output filename \"%s\" must appear just once\n", arg); exit(3); }
      out = arg;
      i++;
    }
  }
  if (!in) {
    printf("int filename not found, either -frame or -frames must be used to indicate input image(s)\n");
    exit(3);
  }
  if (!out) { usage(); exit(3); }

  std::vector<std::string> files = frames_pattern ? frames_from_first(in) : std::vector<std::string>{in};
  const bool is_linear = gamma == "linear", is_srgb = gamma == "srgb";

  MetalRenderContext *mrc = MetalRenderContext_new(0);
  if (!MetalRenderContext_setupMetal(mrc)) {
    fprintf(stderr, "no CUDA device: srgb_to_bt709 has no CPU path\n");
    return 1;
  }
  const std::string outs = out;
  const size_t dot = outs.find_last_of('.');
  const std::string alpha_out = (dot == std::string::npos ? outs : outs.substr(0, dot)) + "_alpha.y4m";

  // pass 0: colour (premultiplied RGB); pass 1 (only with -alpha 1): the alpha channel as linear gray
  for (int pass = 0; pass < (is_alpha ? 2 : 1); pass++) {
    const std::string &path = pass == 0 ? outs : alpha_out;
    FILE *f = y4m_open_file(path.c_str());
    if (!f) return 1;
    bool wrote_header = false;
    std::vector<uint8_t> Y, Cb, Cr;
    for (const std::string &file : files) {
      Image img;
      if (!load_image(file, &img)) { fprintf(stderr, "can't decode image \"%s\"\n", file.c_str()); return 1; }
      if ((img.w % 2) || (img.h % 2)) {  // BGRAToBT709Converter.m:540-541
        fprintf(stderr, "width and height must be even, got %d x %d\n", img.w, img.h);
        return 2;
      }
      std::vector<uint32_t> bgra(img.px.size());
      for (size_t i = 0; i < bgra.size(); i++) {
        const uint32_t p = img.px[i], a = p >> 24;
        if (pass == 1) {  // srgb_to_bt709.m:870-936: A copied over R, G and B, alpha forced to 0xFF... as linear data
          bgra[i] = 0xFFu << 24 | a << 16 | a << 8 | a;
        } else {
          auto pm = [a](uint32_t c) { return (c * a + 127u) / 255u; };
          bgra[i] = 0xFFu << 24 | pm((p >> 16) & 255) << 16 | pm((p >> 8) & 255) << 8 | pm(p & 255);
        }
      }
      // alpha planes are always encoded as linear (srgb_to_bt709.m:952-954)
      const bool lin = pass == 1 ? true : is_linear, srgb = pass == 1 ? false : is_srgb;
      MBDPixelBuffer *pb = BGRAToBT709Converter_createYCbCrFromBGRA(mrc, bgra.data(), img.w, img.h, lin, srgb);
      if (!pb) return 1;
      Y.resize((size_t) img.w * img.h);
      Cb.resize((size_t) (img.w / 2) * (img.h / 2));
      Cr.resize(Cb.size());
      if (!BGRAToBT709Converter_copyYCBCr(mrc, pb, Y.data(), Cb.data(), Cr.data())) return 1;
      if (!wrote_header) {
        Y4MHeaderStruct hs = {img.w, img.h, (Y4MHeaderFPS) fps};
        const int r = y4m_write_header(f, &hs);
        if (r != 0) return r;
        wrote_header = true;
      }
      Y4MFrameStruct fs = {Y.data(), (int) Y.size(), Cb.data(), (int) Cb.size(), Cr.data(), (int) Cr.size()};
      const int r = y4m_write_frame(f, &fs);
      if (r != 0) return r;
      MBDPixelBufferRelease(pb);
    }
    fclose(f);
    fprintf(stdout, "wrote %s\n", path.c_str());
  }
  MetalRenderContext_release(mrc);
  return 0;
}
