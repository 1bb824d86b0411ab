// Host side of the JPEG frame decoder: marker parsing and Huffman (entropy) decoding of baseline / extended-sequential
// 8-bit JPEG streams into quantised DCT coefficient blocks.  The sequential bit-stream walk stays on the host -- it is
// what nvJPEG's hybrid back-end keeps there too -- and everything that is per-block or per-pixel arithmetic
// (dequantisation, inverse DCT, chroma upsampling, colour conversion) runs on the device (jpeg_decode.cu).
//
// Replaces, together with jpeg_decode.cu, the frame read of the reference's video loop:
//     img = read_image(path, format="BGR")                        eval.py:324-327
// = detectron2.data.detection_utils.read_image -> PIL.Image.open(...).convert("RGB") -> numpy -> [:, :, ::-1], i.e.
// Pillow's bundled libjpeg-turbo with its default settings (JDCT_ISLOW, fancy upsampling).  detectron2 and Pillow are
// third-party dependencies outside /root/reference; the arithmetic restated here is the published
// libjpeg (IJG release 6b) algorithm that libjpeg-turbo reproduces bit for bit, and parity is pinned on Pillow itself
// (tests/test_jpeg_decode.py compares every pixel with PIL.Image.open).
//
// Supported: SOF0 / SOF1 (Huffman, 8-bit), 1 or 3 components, sampling factors 1 or 2 with the first component the
// largest, interleaved or non-interleaved scans, restart intervals, JFIF / Adobe colour-space markers.  Everything else
// (progressive, arithmetic coding, 12-bit, CMYK, other sampling) is reported as unsupported -- there is no silent
// fallback.
#pragma once
#include <stdint.h>
#include <string.h>

#include <vector>

namespace msda_jpeg {

enum Error { kOk = 0, kTruncated = 1, kCorrupt = 2, kUnsupported = 3 };

struct Component {
  int id = 0, h = 1, v = 1, tq = 0;       // sampling factors, quantisation table
  int td = 0, ta = 0;                     // Huffman tables of the current scan
  int blocks_w = 0, blocks_h = 0;         // allocated blocks (padded to whole MCUs)
  int width = 0, height = 0;              // downsampled_width / downsampled_height: the real samples
  int dc_pred = 0;
  size_t coef_offset = 0;                 // first int16 of this component in Decoded::coef
};

struct Decoded {
  int width = 0, height = 0, ncomp = 0;
  int max_h = 1, max_v = 1;
  bool ycc = true;                        // 3 components: YCbCr (true) or RGB stored directly
  Component comp[3];
  uint16_t quant[4][64];                  // natural (row-major) order
  bool quant_set[4] = {false, false, false, false};
  std::vector<int16_t> coef;              // per component: [blocks_h][blocks_w][64], natural order
};

namespace detail {

static const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                    41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                    30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

constexpr int kLook = 12;                 // bits of lookahead in the symbol and fast-AC tables

inline int extend(int v, int s) { return v < (1 << (s - 1)) ? v - (1 << s) + 1 : v; }

struct HuffTable {
  bool set = false;
  uint8_t bits[17];
  uint8_t vals[256];
  int mincode[17], maxcode[18], valptr[17];
  uint8_t look_len[1 << kLook], look_val[1 << kLook];   // symbol of every kLook-bit prefix that holds a whole code
  int32_t fast_ac[1 << kLook];   // AC tables: (value << 8) | (run << 4) | (code + magnitude bits) when both fit the prefix
  int build() {
    int code = 0, k = 0;
    for (int l = 1; l <= 16; ++l) {
      valptr[l] = k;
      mincode[l] = code;
      code += bits[l];
      k += bits[l];
      maxcode[l] = bits[l] ? code - 1 : -1;
      if (code > (1 << l)) return kCorrupt;
      code <<= 1;
    }
    maxcode[17] = 0x7fffffff;
    if (k > 256) return kCorrupt;
    memset(look_len, 0, sizeof(look_len));
    memset(fast_ac, 0, sizeof(fast_ac));
    int c = 0;
    k = 0;
    for (int l = 1; l <= kLook; ++l) {
      for (int i = 0; i < bits[l]; ++i, ++k, ++c) {
        const int first = c << (kLook - l);
        for (int j = 0; j < (1 << (kLook - l)); ++j) {
          look_len[first + j] = (uint8_t)l;
          look_val[first + j] = vals[k];
          const int run = vals[k] >> 4, mag = vals[k] & 15;
          if (mag && l + mag <= kLook) {
            const int v = extend(((first + j) >> (kLook - l - mag)) & ((1 << mag) - 1), mag);
            fast_ac[first + j] = (v * 256) + (run * 16) + (l + mag);
          }
        }
      }
      c <<= 1;
    }
    return kOk;
  }
};

struct BitReader {
  const uint8_t* p;
  const uint8_t* end;
  uint64_t acc = 0;
  int nbits = 0;
  bool hit_marker = false;      // a marker was met: the stream is padded with zero bits from here on
  bool eof = false;             // ... or the data ended without one
  void fill() {                 // tops the accumulator up to more than 56 bits
    if (!hit_marker && end - p >= 8 && nbits <= 32) {
      uint64_t v;
      memcpy(&v, p, 8);
      v = __builtin_bswap64(v);
      const uint64_t inv = ~v;                                                  // a zero byte of inv = an 0xFF byte of v
      if (!((inv - 0x0101010101010101ull) & ~inv & 0x8080808080808080ull)) {    // no marker / stuffing in the next 8 bytes
        const int take = (64 - nbits) >> 3;                                     // whole bytes that fit
        acc |= (v >> (64 - 8 * take)) << (64 - nbits - 8 * take);
        p += take;
        nbits += 8 * take;
        return;
      }
    }
    while (nbits <= 56) {
      uint64_t byte = 0;
      if (!hit_marker && p < end) {
        byte = *p;
        if (byte == 0xff) {
          if (p + 1 < end && p[1] == 0x00) {
            p += 2;
          } else {
            hit_marker = true;    // leave p at the marker
            byte = 0;
          }
        } else {
          ++p;
        }
      } else {
        if (!hit_marker) eof = true;
        hit_marker = true;
      }
      acc |= byte << (56 - nbits);
      nbits += 8;
    }
  }
  int peek(int n) { return (int)(acc >> (64 - n)); }
  void skip(int n) { acc <<= n; nbits -= n; }
  int get(int n) {              // n <= 16; the caller keeps nbits >= 32
    if (n == 0) return 0;
    const int v = peek(n);
    skip(n);
    return v;
  }
  void reset() { acc = 0; nbits = 0; hit_marker = false; }
};

// needs nbits >= 16
inline int decode_symbol(BitReader& br, const HuffTable& t) {
  const int look = br.peek(kLook);
  int l = t.look_len[look];
  if (l) {
    br.skip(l);
    return t.look_val[look];
  }
  l = kLook + 1;
  int code = br.peek(l);
  while (l <= 16 && code > t.maxcode[l]) {
    ++l;
    if (l > 16) break;
    code = br.peek(l);
  }
  if (l > 16) return -1;
  br.skip(l);
  const int idx = t.valptr[l] + code - t.mincode[l];
  if (idx < 0 || idx >= 256) return -1;
  return t.vals[idx];
}

inline int decode_block(BitReader& br, const HuffTable& dc, const HuffTable& ac, int& pred, int16_t* out) {
  if (br.nbits < 32) br.fill();
  int s = decode_symbol(br, dc);
  if (s < 0 || s > 16) return kCorrupt;
  int diff = 0;
  if (s) diff = extend(br.get(s), s);
  pred += diff;
  out[0] = (int16_t)pred;
  for (int k = 1; k < 64;) {
    if (br.nbits < 32) br.fill();
    const int f = ac.fast_ac[br.peek(kLook)];
    if (f) {                     // run, size and the magnitude bits all inside the lookahead
      k += (f >> 4) & 15;
      if (k > 63) return kCorrupt;
      br.skip(f & 15);
      out[kZigzag[k++]] = (int16_t)(f >> 8);
      continue;
    }
    const int rs = decode_symbol(br, ac);
    if (rs < 0) return kCorrupt;
    const int r = rs >> 4, sz = rs & 15;
    if (sz == 0) {
      if (r != 15) break;        // EOB
      k += 16;
      continue;
    }
    k += r;
    if (k > 63) return kCorrupt;
    out[kZigzag[k]] = (int16_t)extend(br.get(sz), sz);
    ++k;
  }
  return kOk;
}

inline int be16(const uint8_t* p) { return (p[0] << 8) | p[1]; }

}  // namespace detail

// Parse `data` and entropy-decode every scan.  Returns kOk or an Error; `out` holds the coefficient blocks.
inline int entropy_decode(const uint8_t* data, size_t len, Decoded& out) {
  using namespace detail;
  if (len < 4 || data[0] != 0xff || data[1] != 0xd8) return kCorrupt;
  std::vector<HuffTable> tables(8);      // 4 DC + 4 AC; ~20 KB each with the lookahead tables: not on the stack
  HuffTable* dc_tab = tables.data();
  HuffTable* ac_tab = tables.data() + 4;
  int restart_interval = 0;
  bool saw_sof = false, saw_jfif = false, saw_adobe = false;
  int adobe_transform = 0;
  int scans_left = 0;
  size_t pos = 2;
  while (true) {
    // next marker
    while (pos < len && data[pos] != 0xff) ++pos;
    while (pos < len && data[pos] == 0xff) ++pos;
    if (pos >= len) return saw_sof && scans_left == 0 ? kOk : kTruncated;
    const int m = data[pos++];
    if (m == 0xd9) break;                                   // EOI
    if (m == 0x00 || m == 0x01 || (m >= 0xd0 && m <= 0xd7)) continue;     // stuffed byte / TEM / stray RSTn
    if (pos + 2 > len) return kTruncated;
    const int seglen = be16(data + pos);
    if (seglen < 2 || pos + seglen > len) return kTruncated;
    const uint8_t* seg = data + pos + 2;
    const int n = seglen - 2;
    if (m == 0xc0 || m == 0xc1) {                            // SOF0 / SOF1
      if (saw_sof || n < 6) return kCorrupt;
      if (seg[0] != 8) return kUnsupported;
      out.height = be16(seg + 1);
      out.width = be16(seg + 3);
      out.ncomp = seg[5];
      if (out.ncomp != 1 && out.ncomp != 3) return kUnsupported;
      if (out.width <= 0 || out.height <= 0 || n < 6 + 3 * out.ncomp) return kCorrupt;
      if ((long long)out.width * out.height > (1LL << 28)) return kUnsupported;        // 268 Mpixel: refuse absurd headers
      out.max_h = out.max_v = 1;
      for (int c = 0; c < out.ncomp; ++c) {
        Component& cp = out.comp[c];
        cp.id = seg[6 + 3 * c];
        cp.h = seg[7 + 3 * c] >> 4;
        cp.v = seg[7 + 3 * c] & 15;
        cp.tq = seg[8 + 3 * c];
        if (cp.tq > 3) return kCorrupt;
        if (cp.h < 1 || cp.h > 2 || cp.v < 1 || cp.v > 2) return kUnsupported;
        if (cp.h > out.max_h) out.max_h = cp.h;
        if (cp.v > out.max_v) out.max_v = cp.v;
      }
      if (out.ncomp == 1) { out.comp[0].h = out.comp[0].v = 1; out.max_h = out.max_v = 1; }   // like libjpeg: one block per MCU
      if (out.ncomp == 3) {
        if (out.comp[0].h != out.max_h || out.comp[0].v != out.max_v) return kUnsupported;
        if (out.comp[1].h != out.comp[2].h || out.comp[1].v != out.comp[2].v) return kUnsupported;
        if (out.max_h / out.comp[1].h == 1 && out.max_v / out.comp[1].v == 2) return kUnsupported;   // 4:4:0
      }
      const int mcus_w = (out.width + 8 * out.max_h - 1) / (8 * out.max_h);
      const int mcus_h = (out.height + 8 * out.max_v - 1) / (8 * out.max_v);
      size_t total = 0;
      for (int c = 0; c < out.ncomp; ++c) {
        Component& cp = out.comp[c];
        cp.blocks_w = mcus_w * cp.h;
        cp.blocks_h = mcus_h * cp.v;
        cp.width = (out.width * cp.h + out.max_h - 1) / out.max_h;
        cp.height = (out.height * cp.v + out.max_v - 1) / out.max_v;
        cp.coef_offset = total;
        total += (size_t)cp.blocks_w * cp.blocks_h * 64;
      }
      out.coef.assign(total, 0);
      saw_sof = true;
      scans_left = out.ncomp;
    } else if (m == 0xc2 || m == 0xc3 || (m >= 0xc5 && m <= 0xcf && m != 0xc4 && m != 0xc8 && m != 0xcc)) {
      return kUnsupported;                                   // progressive, lossless, arithmetic, hierarchical
    } else if (m == 0xcc) {
      return kUnsupported;
    } else if (m == 0xc4) {                                  // DHT
      int o = 0;
      while (o < n) {
        if (o + 17 > n) return kCorrupt;
        const int tc = seg[o] >> 4, th = seg[o] & 15;
        if (tc > 1 || th > 3) return kCorrupt;
        HuffTable& t = tc ? ac_tab[th] : dc_tab[th];
        int count = 0;
        t.bits[0] = 0;
        for (int i = 1; i <= 16; ++i) { t.bits[i] = seg[o + i]; count += t.bits[i]; }
        o += 17;
        if (count > 256 || o + count > n) return kCorrupt;
        memset(t.vals, 0, sizeof(t.vals));
        memcpy(t.vals, seg + o, count);
        o += count;
        if (t.build() != kOk) return kCorrupt;
        t.set = true;
      }
    } else if (m == 0xdb) {                                  // DQT
      int o = 0;
      while (o < n) {
        const int pq = seg[o] >> 4, tq = seg[o] & 15;
        if (tq > 3 || pq > 1) return kCorrupt;
        ++o;
        if (o + 64 * (pq + 1) > n) return kCorrupt;
        for (int i = 0; i < 64; ++i) {
          const int v = pq ? be16(seg + o + 2 * i) : seg[o + i];
          out.quant[tq][kZigzag[i]] = (uint16_t)v;
        }
        out.quant_set[tq] = true;
        o += 64 * (pq + 1);
      }
    } else if (m == 0xdd) {                                  // DRI
      if (n < 2) return kCorrupt;
      restart_interval = be16(seg);
    } else if (m == 0xe0) {
      if (n >= 5 && memcmp(seg, "JFIF", 5) == 0) saw_jfif = true;
    } else if (m == 0xee) {
      if (n >= 12 && memcmp(seg, "Adobe", 5) == 0) { saw_adobe = true; adobe_transform = seg[11]; }
    } else if (m == 0xda) {                                  // SOS
      if (!saw_sof || n < 1) return kCorrupt;
      const int ns = seg[0];
      if (ns < 1 || ns > out.ncomp || n < 1 + 2 * ns + 3) return kCorrupt;
      Component* sc[3];
      for (int i = 0; i < ns; ++i) {
        const int cid = seg[1 + 2 * i];
        sc[i] = nullptr;
        for (int c = 0; c < out.ncomp; ++c)
          if (out.comp[c].id == cid) sc[i] = &out.comp[c];
        if (!sc[i]) return kCorrupt;
        sc[i]->td = seg[2 + 2 * i] >> 4;
        sc[i]->ta = seg[2 + 2 * i] & 15;
        if (sc[i]->td > 3 || sc[i]->ta > 3 || !dc_tab[sc[i]->td].set || !ac_tab[sc[i]->ta].set) return kCorrupt;
        sc[i]->dc_pred = 0;
      }
      const uint8_t* sp = seg + 1 + 2 * ns;
      if (sp[0] != 0 || sp[1] != 63 || sp[2] != 0) return kUnsupported;       // spectral selection / approximation: progressive
      BitReader br;
      br.p = data + pos + seglen;
      br.end = data + len;
      int mcus_w, mcus_h;
      if (ns == 1) {                                         // non-interleaved: MCU = one block, real block counts
        mcus_w = (sc[0]->width + 7) / 8;
        mcus_h = (sc[0]->height + 7) / 8;
      } else {
        if (ns != out.ncomp) return kUnsupported;
        mcus_w = (out.width + 8 * out.max_h - 1) / (8 * out.max_h);
        mcus_h = (out.height + 8 * out.max_v - 1) / (8 * out.max_v);
      }
      int to_restart = restart_interval, next_rst = 0;
      for (int my = 0; my < mcus_h; ++my) {
        for (int mx = 0; mx < mcus_w; ++mx) {
          if (restart_interval && to_restart == 0) {
            // byte-align, expect RSTn
            br.reset();
            const uint8_t* q = br.p;
            while (q < br.end && *q != 0xff) ++q;
            while (q < br.end && *q == 0xff) ++q;
            if (q >= br.end) return kTruncated;
            if (*q != 0xd0 + next_rst) return kCorrupt;
            br.p = q + 1;
            next_rst = (next_rst + 1) & 7;
            to_restart = restart_interval;
            for (int i = 0; i < ns; ++i) sc[i]->dc_pred = 0;
          }
          for (int i = 0; i < ns; ++i) {
            Component& cp = *sc[i];
            const int bh = ns == 1 ? 1 : cp.h, bv = ns == 1 ? 1 : cp.v;
            for (int by = 0; by < bv; ++by)
              for (int bx = 0; bx < bh; ++bx) {
                const int row = my * bv + by, col = mx * bh + bx;
                int16_t* blk = out.coef.data() + cp.coef_offset + ((size_t)row * cp.blocks_w + col) * 64;
                const int rc = decode_block(br, dc_tab[cp.td], ac_tab[cp.ta], cp.dc_pred, blk);
                if (rc != kOk) return rc;
              }
          }
          if (restart_interval) --to_restart;
        }
      }
      if (br.eof) return kTruncated;                         // ran off the end of the data inside the scan
      scans_left -= ns;
      // continue after the entropy-coded segment: the reader stopped at (or before) the next marker
      pos = (size_t)(br.p - data);
      continue;
    }
    pos += seglen;
  }
  if (!saw_sof || scans_left > 0) return kTruncated;
  for (int c = 0; c < out.ncomp; ++c)
    if (!out.quant_set[out.comp[c].tq]) return kCorrupt;
  // colour space of a 3-component image, as libjpeg's default_decompress_parms decides it
  out.ycc = true;
  if (out.ncomp == 3) {
    if (saw_jfif) out.ycc = true;
    else if (saw_adobe) out.ycc = adobe_transform != 0;
    else out.ycc = !(out.comp[0].id == 'R' && out.comp[1].id == 'G' && out.comp[2].id == 'B');
  }
  return kOk;
}

}  // namespace msda_jpeg
