// TEST INFRASTRUCTURE -- CPU restatement of the sample reconstruction of a baseline JPEG as Pillow's bundled
// libjpeg-turbo performs it with default settings (what the reference's frame read, eval.py:324-327 ->
// detectron2 read_image -> PIL.Image.open(...).convert("RGB"), executes).  libjpeg-turbo / IJG libjpeg are third-party
// dependencies that are not vendored under /root/reference; the algorithms restated here are the published ones of IJG
// release 6b, which libjpeg-turbo reproduces bit for bit in C and SIMD:
//   * jidctint.c  jpeg_idct_islow: 13-bit fixed-point Loeffler-Ligtenberg-Moschytz inverse DCT, two passes, PASS1_BITS 2,
//                 output through the range-limit table with the +128 level shift folded in (wrap-around included);
//   * jdsample.c  h2v1_fancy_upsample / h2v2_fancy_upsample: triangle filter (3/4, 1/4), the image edge replicated;
//   * jdcolor.c   ycc_rgb_convert: 16-bit fixed-point tables (1.40200, 1.77200, 0.71414, 0.34414).
// PINNED on Pillow itself: tests/test_jpeg_decode.py compares this oracle with PIL.Image.open pixel for pixel over
// subsamplings, qualities and odd sizes; the CUDA decoder (gomatching_b200/csrc/jpeg_decode.cu) is then compared with
// both.  The entropy decoder is shared with the product (gomatching_b200/csrc/jpeg_entropy.h: host code, no arithmetic
// beyond Huffman decoding); Pillow is what checks it.  Nothing in the product links or calls this file.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../gomatching_b200/csrc/jpeg_entropy.h"

namespace {

inline uint8_t clamp8(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

// the post-IDCT range-limit table addressed with (x & 1023): 128..255, 255 x 384, 0 x 384, 0..127
inline uint8_t idct_limit(int x) {
  x &= 1023;
  if (x < 128) return (uint8_t)(x + 128);
  if (x < 512) return 255;
  if (x < 896) return 0;
  return (uint8_t)(x - 896);
}

constexpr int CONST_BITS = 13, PASS1_BITS = 2;
constexpr int F_0_298631336 = 2446, F_0_390180644 = 3196, F_0_541196100 = 4433, F_0_765366865 = 6270, F_0_899976223 = 7373,
              F_1_175875602 = 9633, F_1_501321110 = 12299, F_1_847759065 = 15137, F_1_961570560 = 16069,
              F_2_053119869 = 16819, F_2_562915447 = 20995, F_3_072711026 = 25172;

inline int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

void idct_islow(const int16_t* coef, const uint16_t* q, uint8_t* out, int stride) {
  int ws[64];
  for (int c = 0; c < 8; ++c) {
    int in[8];
    for (int r = 0; r < 8; ++r) in[r] = (int)coef[8 * r + c] * (int)q[8 * r + c];
    int z2 = in[2], z3 = in[6];
    int z1 = (z2 + z3) * F_0_541196100;
    int tmp2 = z1 + z3 * (-F_1_847759065);
    int tmp3 = z1 + z2 * F_0_765366865;
    z2 = in[0]; z3 = in[4];
    int tmp0 = (z2 + z3) * (1 << CONST_BITS), tmp1 = (z2 - z3) * (1 << CONST_BITS);
    const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    tmp0 = in[7]; tmp1 = in[5]; tmp2 = in[3]; tmp3 = in[1];
    z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
    int z4 = tmp1 + tmp3;
    const int z5 = (z3 + z4) * F_1_175875602;
    tmp0 *= F_0_298631336; tmp1 *= F_2_053119869; tmp2 *= F_3_072711026; tmp3 *= F_1_501321110;
    z1 *= -F_0_899976223; z2 *= -F_2_562915447; z3 *= -F_1_961570560; z4 *= -F_0_390180644;
    z3 += z5; z4 += z5;
    tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
    ws[8 * 0 + c] = descale(tmp10 + tmp3, CONST_BITS - PASS1_BITS);
    ws[8 * 7 + c] = descale(tmp10 - tmp3, CONST_BITS - PASS1_BITS);
    ws[8 * 1 + c] = descale(tmp11 + tmp2, CONST_BITS - PASS1_BITS);
    ws[8 * 6 + c] = descale(tmp11 - tmp2, CONST_BITS - PASS1_BITS);
    ws[8 * 2 + c] = descale(tmp12 + tmp1, CONST_BITS - PASS1_BITS);
    ws[8 * 5 + c] = descale(tmp12 - tmp1, CONST_BITS - PASS1_BITS);
    ws[8 * 3 + c] = descale(tmp13 + tmp0, CONST_BITS - PASS1_BITS);
    ws[8 * 4 + c] = descale(tmp13 - tmp0, CONST_BITS - PASS1_BITS);
  }
  for (int r = 0; r < 8; ++r) {
    const int* w = ws + 8 * r;
    int z2 = w[2], z3 = w[6];
    int z1 = (z2 + z3) * F_0_541196100;
    int tmp2 = z1 + z3 * (-F_1_847759065);
    int tmp3 = z1 + z2 * F_0_765366865;
    int tmp0 = (w[0] + w[4]) * (1 << CONST_BITS), tmp1 = (w[0] - w[4]) * (1 << CONST_BITS);
    const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    tmp0 = w[7]; tmp1 = w[5]; tmp2 = w[3]; tmp3 = w[1];
    z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
    int z4 = tmp1 + tmp3;
    const int z5 = (z3 + z4) * F_1_175875602;
    tmp0 *= F_0_298631336; tmp1 *= F_2_053119869; tmp2 *= F_3_072711026; tmp3 *= F_1_501321110;
    z1 *= -F_0_899976223; z2 *= -F_2_562915447; z3 *= -F_1_961570560; z4 *= -F_0_390180644;
    z3 += z5; z4 += z5;
    tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
    const int sh = CONST_BITS + PASS1_BITS + 3;
    uint8_t* o = out + (size_t)r * stride;
    o[0] = idct_limit(descale(tmp10 + tmp3, sh)); o[7] = idct_limit(descale(tmp10 - tmp3, sh));
    o[1] = idct_limit(descale(tmp11 + tmp2, sh)); o[6] = idct_limit(descale(tmp11 - tmp2, sh));
    o[2] = idct_limit(descale(tmp12 + tmp1, sh)); o[5] = idct_limit(descale(tmp12 - tmp1, sh));
    o[3] = idct_limit(descale(tmp13 + tmp0, sh)); o[4] = idct_limit(descale(tmp13 - tmp0, sh));
  }
}

// full-resolution plane (width x height) of one component from its downsampled plane `p` (row pitch `pitch`; `cw` x `ch`
// real samples)
void upsample(const uint8_t* p, int pitch, int cw, int ch, int hs, int vs, int width, int height, std::vector<uint8_t>& out) {
  out.assign((size_t)width * height, 0);
  if (hs == 1 && vs == 1) {
    for (int y = 0; y < height; ++y) memcpy(&out[(size_t)y * width], p + (size_t)y * pitch, width);
    return;
  }
  if (cw <= 2) {                                             // jinit_upsampler: fancy only if downsampled_width > 2
    for (int y = 0; y < height; ++y)
      for (int x = 0; x < width; ++x) out[(size_t)y * width + x] = p[(size_t)(vs == 2 ? y >> 1 : y) * pitch + (hs == 2 ? x >> 1 : x)];
    return;
  }
  std::vector<int> colsum(cw);
  for (int y = 0; y < height; ++y) {
    const int r = vs == 2 ? y >> 1 : y;
    const uint8_t* row0 = p + (size_t)r * pitch;
    if (vs == 2) {
      int rn = (y & 1) ? r + 1 : r - 1;                      // the nearer neighbour row; the image edge is replicated
      if (rn < 0) rn = 0;
      if (rn > ch - 1) rn = ch - 1;
      const uint8_t* row1 = p + (size_t)rn * pitch;
      for (int x = 0; x < cw; ++x) colsum[x] = 3 * row0[x] + row1[x];
    } else {
      for (int x = 0; x < cw; ++x) colsum[x] = row0[x];
    }
    uint8_t* o = &out[(size_t)y * width];
    for (int x = 0; x < width; ++x) {
      int v;
      if (hs == 2) {
        const int i = x >> 1;
        const int cur = colsum[i];
        if (vs == 2) {
          if (!(x & 1)) v = i == 0 ? (cur * 4 + 8) >> 4 : (cur * 3 + colsum[i - 1] + 8) >> 4;
          else v = i == cw - 1 ? (cur * 4 + 7) >> 4 : (cur * 3 + colsum[i + 1] + 7) >> 4;
        } else {
          if (!(x & 1)) v = i == 0 ? cur : (cur * 3 + colsum[i - 1] + 1) >> 2;
          else v = i == cw - 1 ? cur : (cur * 3 + colsum[i + 1] + 2) >> 2;
        }
      } else {                                               // hs == 1, vs == 2 is rejected by the entropy stage
        v = colsum[x];
      }
      o[x] = (uint8_t)v;
    }
  }
}

}  // namespace

extern "C" {

// 0 = ok, else msda_jpeg::Error.  Writes width / height; if `rgb` is non-null and `cap` >= 3 * w * h fills it (RGB HWC).
int jpeg_oracle_decode(const uint8_t* data, size_t len, int* width, int* height, uint8_t* rgb, size_t cap) {
  msda_jpeg::Decoded d;
  const int rc = msda_jpeg::entropy_decode(data, len, d);
  if (rc != 0) return rc;
  *width = d.width;
  *height = d.height;
  if (!rgb) return 0;
  if (cap < (size_t)3 * d.width * d.height) return -1;
  std::vector<uint8_t> plane[3], full[3];
  for (int c = 0; c < d.ncomp; ++c) {
    const msda_jpeg::Component& cp = d.comp[c];
    const int pitch = cp.blocks_w * 8;
    plane[c].assign((size_t)pitch * cp.blocks_h * 8, 0);
    for (int by = 0; by < cp.blocks_h; ++by)
      for (int bx = 0; bx < cp.blocks_w; ++bx)
        idct_islow(d.coef.data() + cp.coef_offset + ((size_t)by * cp.blocks_w + bx) * 64, d.quant[cp.tq],
                   &plane[c][(size_t)by * 8 * pitch + bx * 8], pitch);
    upsample(plane[c].data(), pitch, cp.width, cp.height, d.max_h / cp.h, d.max_v / cp.v, d.width, d.height, full[c]);
  }
  const size_t n = (size_t)d.width * d.height;
  if (d.ncomp == 1) {
    for (size_t i = 0; i < n; ++i) rgb[3 * i] = rgb[3 * i + 1] = rgb[3 * i + 2] = full[0][i];
  } else if (!d.ycc) {
    for (size_t i = 0; i < n; ++i) { rgb[3 * i] = full[0][i]; rgb[3 * i + 1] = full[1][i]; rgb[3 * i + 2] = full[2][i]; }
  } else {
    for (size_t i = 0; i < n; ++i) {
      const int y = full[0][i], cb = full[1][i] - 128, cr = full[2][i] - 128;
      rgb[3 * i] = clamp8(y + ((91881 * cr + 32768) >> 16));
      rgb[3 * i + 1] = clamp8(y + ((-22554 * cb + 32768 - 46802 * cr) >> 16));
      rgb[3 * i + 2] = clamp8(y + ((116130 * cb + 32768) >> 16));
    }
  }
  return 0;
}

}  // extern "C"
