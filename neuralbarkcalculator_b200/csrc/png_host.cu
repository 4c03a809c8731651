// Host-side PNG image-data encoder for the files of the hot path: the processed/ RGB images
// (reference models.py:203, skimage imsave) and the 0/127/255 dual images (models.py:349-356, PIL save).
//
// Once the GPU path runs at > 1 000 scans/s the folder pipeline is bound by zlib: Z_RLE deflate of a 1.9 MB filtered
// image takes ~27 ms per host thread.  This encoder produces the same kind of stream -- PNG filter Sub for RGB, None
// for grey, run-length matches at distance 1, Huffman coding -- about 5x faster, because it does nothing else:
//   * the image is cut into segments of whole rows (<= 65 535 filtered bytes, so a segment can always fall back to ONE
//     stored block); every segment is one deflate block with its own dynamic Huffman code, built from the segment's
//     histogram (two passes over a cache-resident segment), runs never reach back across a segment start -- segments
//     are therefore independent, which is what a GPU version needs (one thread block per segment);
//   * symbols: literal bytes and (length 3..258, distance 1) matches, exactly zlib's Z_RLE vocabulary; runs are found
//     eight bytes at a time (SWAR zero-byte test on f[i..i+8) ^ f[i-1..i+7)), literals are appended two per store;
//   * code lengths limited to 15 bits by rescaling the histogram (rare); the code-length alphabet uses a fixed
//     complete code (4 bits for each of 0..15), so the block header is ~160 bytes and needs no second Huffman build.
// The output is a complete zlib stream (header, blocks, Adler-32) = the payload of ONE IDAT chunk; chunk framing and
// CRC-32 stay in _png.py.  No GPU involved; plain C++ compiled by nvcc's host compiler into libnbc.so.
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace nbc {
namespace {

constexpr int kMaxSegBytes = 65535;
constexpr int kNumLit = 286;      // literal / length alphabet (0..255 literals, 256 end of block, 257..285 lengths)
constexpr int kMaxBits = 15;

struct LenCode {
  uint16_t sym;      // 257..285
  uint8_t extra_bits;
  uint16_t base;
};
// RFC 1951 3.2.5: length codes 257..285
const LenCode kLenCodes[29] = {{257, 0, 3},   {258, 0, 4},   {259, 0, 5},   {260, 0, 6},   {261, 0, 7},   {262, 0, 8},
                               {263, 0, 9},   {264, 0, 10},  {265, 1, 11},  {266, 1, 13},  {267, 1, 15},  {268, 1, 17},
                               {269, 2, 19},  {270, 2, 23},  {271, 2, 27},  {272, 2, 31},  {273, 3, 35},  {274, 3, 43},
                               {275, 3, 51},  {276, 3, 59},  {277, 4, 67},  {278, 4, 83},  {279, 4, 99},  {280, 4, 115},
                               {281, 5, 131}, {282, 5, 163}, {283, 5, 195}, {284, 5, 227}, {285, 0, 258}};

struct LenLut {
  uint16_t sym[259];
  uint8_t extra_bits[259];
  uint16_t extra_val[259];
  LenLut() {
    for (int L = 3; L <= 258; ++L) {
      int k = 28;
      if (L < 258) {
        k = 0;
        while (k + 1 < 28 && kLenCodes[k + 1].base <= L) ++k;
      }
      sym[L] = kLenCodes[k].sym;
      extra_bits[L] = kLenCodes[k].extra_bits;
      extra_val[L] = (uint16_t)(L - kLenCodes[k].base);
    }
  }
};
const LenLut kLenLut;

struct BitWriter {
  uint8_t* p;
  uint8_t* end;      // 8 bytes before the real end of the buffer: put() always stores 8 bytes
  uint64_t acc = 0;
  int n = 0;         // pending bits, < 8 between calls
  bool overflow = false;
  // append the `bits` low bits of v (LSB first, as deflate packs everything except Huffman codes, which arrive here
  // already bit-reversed); bits <= 32.  The whole accumulator is stored, the pointer advances by the complete bytes.
  inline void put(uint32_t v, int bits) {
    acc |= (uint64_t)v << n;
    n += bits;
    if (p > end) {
      overflow = true;
      p = end;
    }
    memcpy(p, &acc, 8);      // little endian hosts only (x86-64 / aarch64)
    p += n >> 3;
    acc >>= n & ~7;
    n &= 7;
  }
  void align_byte() {
    if (n) put(0, 8 - n);
  }
  void bytes(const uint8_t* src, size_t len) {      // byte aligned (after align_byte)
    if ((size_t)(end + 8 - p) < len) {
      overflow = true;
      return;
    }
    memcpy(p, src, len);
    p += len;
  }
};

inline uint32_t reverse_bits(uint32_t code, int len) {
  uint32_t r = 0;
  for (int i = 0; i < len; ++i) r |= ((code >> i) & 1u) << (len - 1 - i);
  return r;
}

// Huffman code lengths (<= kMaxBits) for freq[0..n); symbols with freq 0 get length 0.  Plain two-queue construction
// on the sorted leaves; if the tree is deeper than kMaxBits the frequencies are flattened ((f + 1) / 2) and the tree
// rebuilt -- any prefix code is a valid deflate code, optimality matters little in that rare case.
void huffman_lengths(const uint32_t* freq_in, int n, uint8_t* len_out) {
  uint32_t freq[kNumLit];
  for (int i = 0; i < n; ++i) freq[i] = freq_in[i];
  for (;;) {
    int order[kNumLit], m = 0;
    for (int i = 0; i < n; ++i)
      if (freq[i]) order[m++] = i;
    for (int i = 0; i < n; ++i) len_out[i] = 0;
    if (m == 0) return;
    if (m == 1) {
      len_out[order[0]] = 1;
      return;
    }
    std::sort(order, order + m, [&](int a, int b) { return freq[a] != freq[b] ? freq[a] < freq[b] : a < b; });
    // nodes 0..m-1 leaves (sorted), m..2m-2 internal (created in non-decreasing weight order)
    uint64_t weight[2 * kNumLit];
    int parent[2 * kNumLit];
    for (int i = 0; i < m; ++i) weight[i] = freq[order[i]];
    int leaf = 0, inode = m, next = m;
    auto pop_min = [&]() {
      if (leaf < m && (inode >= next || weight[leaf] <= weight[inode])) return leaf++;
      return inode++;
    };
    while (next < 2 * m - 1) {
      const int a = pop_min(), b = pop_min();
      weight[next] = weight[a] + weight[b];
      parent[a] = parent[b] = next;
      ++next;
    }
    int depth[2 * kNumLit];
    depth[2 * m - 2] = 0;
    int maxd = 0;
    for (int i = 2 * m - 3; i >= 0; --i) {
      depth[i] = depth[parent[i]] + 1;
      if (i < m && depth[i] > maxd) maxd = depth[i];
    }
    if (maxd <= kMaxBits) {
      for (int i = 0; i < m; ++i) len_out[order[i]] = (uint8_t)depth[i];
      return;
    }
    for (int i = 0; i < n; ++i)
      if (freq[i]) freq[i] = (freq[i] + 1) / 2;
  }
}

// canonical codes (RFC 1951 3.2.2), returned bit-reversed for the LSB-first writer
void canonical_codes(const uint8_t* len, int n, uint16_t* code_rev) {
  int bl_count[kMaxBits + 1] = {0};
  for (int i = 0; i < n; ++i) bl_count[len[i]]++;
  bl_count[0] = 0;
  uint32_t next_code[kMaxBits + 2];
  uint32_t code = 0;
  for (int b = 1; b <= kMaxBits; ++b) {
    code = (code + bl_count[b - 1]) << 1;
    next_code[b] = code;
  }
  for (int i = 0; i < n; ++i) code_rev[i] = len[i] ? (uint16_t)reverse_bits(next_code[len[i]]++, len[i]) : 0;
}

uint32_t adler32_update(uint32_t adler, const uint8_t* d, size_t len) {
  uint32_t a = adler & 0xFFFF, b = adler >> 16;
  while (len) {
    size_t k = len < 5552 ? len : 5552;      // largest run for which b cannot overflow 32 bits
    len -= k;
    // 16 bytes at a time: a grows by their sum, b by 16 * a + sum((16 - i) * d[i]) -- two independent dot products
    // the compiler vectorises, instead of 32 dependent additions
    while (k >= 16) {
      uint32_t s1 = 0, s2 = 0;
      for (int i = 0; i < 16; ++i) {
        s1 += d[i];
        s2 += (uint32_t)(16 - i) * d[i];
      }
      b += 16 * a + s2;
      a += s1;
      d += 16, k -= 16;
    }
    while (k--) a += *d++, b += a;
    a %= 65521u, b %= 65521u;
  }
  return (b << 16) | a;
}

struct Run {
  int32_t pos, len;      // f[pos .. pos+len) repeats f[pos-1]: a deflate match of that length at distance 1
};

inline uint64_t load64(const uint8_t* p) {
  uint64_t v;
  memcpy(&v, p, 8);
  return v;
}
// 0x80 in every byte of v that is zero (exact, no borrow between bytes)
inline uint64_t zero_bytes(uint64_t v) {
  const uint64_t k7f = 0x7F7F7F7F7F7F7F7Full;
  return ~(((v & k7f) + k7f) | v | k7f);
}

inline void put_literals(BitWriter& bw, const uint8_t* f, int a, int b, const uint32_t* lit) {
  int i = a;
  for (; i + 1 < b; i += 2) {       // two literals in one append (<= 30 bits)
    const uint32_t e0 = lit[f[i]], e1 = lit[f[i + 1]];
    const int l0 = (int)(e0 >> 16);
    bw.put((e0 & 0xFFFF) | ((e1 & 0xFFFF) << l0), l0 + (int)(e1 >> 16));
  }
  if (i < b) bw.put(lit[f[i]] & 0xFFFF, (int)(lit[f[i]] >> 16));
}

// one deflate block for the filtered bytes f[0..n), n <= kMaxSegBytes
void encode_segment(const uint8_t* f, int n, bool final_block, BitWriter& bw_io, Run* runs) {
  BitWriter bw = bw_io;      // a local copy lives in registers: stores through the byte pointer may alias anything else
  // ---- pass 1: find the runs (>= 3 copies of the previous byte, zlib's Z_RLE vocabulary) and count the symbols.
  // Eight bytes at a time: byte j of d is zero where f[i+j] == f[i+j-1]; three zero bytes in a row start a run.  A
  // texture has almost none, so the common case is six literals per step without a data-dependent branch.
  uint32_t hist[4][256];
  memset(hist, 0, sizeof(hist));
  uint32_t freq[kNumLit];
  memset(freq, 0, sizeof(freq));
  int nr = 0;
  int i = 0;
  if (n > 0) {
    hist[0][f[0]]++;
    i = 1;
  }
  while (i < n) {
    int k;      // literals before the next run start (or before the next look)
    bool run = false;
    if (i + 8 <= n) {
      const uint64_t x = load64(f + i);
      const uint64_t m = zero_bytes(x ^ load64(f + i - 1));
      const uint64_t t = m & (m >> 8) & (m >> 16);
      if (t == 0) {
        hist[0][x & 0xFF]++, hist[1][(x >> 8) & 0xFF]++, hist[2][(x >> 16) & 0xFF]++;
        hist[3][(x >> 24) & 0xFF]++, hist[0][(x >> 32) & 0xFF]++, hist[1][(x >> 40) & 0xFF]++;
        i += 6;
        continue;
      }
      k = __builtin_ctzll(t) >> 3;
      run = true;
    } else {
      const uint8_t prev = f[i - 1];
      run = f[i] == prev && i + 2 < n && f[i + 1] == prev && f[i + 2] == prev;
      k = run ? 0 : 1;
    }
    for (int j = 0; j < k; ++j) hist[j & 3][f[i + j]]++;
    i += k;
    if (run) {
      const uint8_t prev = f[i - 1];
      int r = 3;
      const int lim = std::min(258, n - i);
      while (r < lim && f[i + r] == prev) ++r;
      runs[nr].pos = i, runs[nr].len = r;
      ++nr;
      freq[kLenLut.sym[r]]++;
      i += r;
    }
  }
  for (int s = 0; s < 256; ++s) freq[s] = hist[0][s] + hist[1][s] + hist[2][s] + hist[3][s];
  freq[256] = 1;      // end of block

  uint8_t len[kNumLit];
  huffman_lengths(freq, kNumLit, len);
  int nlit = kNumLit;
  while (nlit > 257 && len[nlit - 1] == 0) --nlit;

  // ---- size of the dynamic block vs one stored block
  uint64_t bits = 3 + 5 + 5 + 4 + 19 * 3 + (uint64_t)(nlit + 1) * 4;
  for (int s = 0; s < 256; ++s) bits += (uint64_t)freq[s] * len[s];
  bits += len[256];
  for (int k = 0; k < 29; ++k) bits += (uint64_t)freq[kLenCodes[k].sym] * (len[kLenCodes[k].sym] + kLenCodes[k].extra_bits + 1);
  const uint64_t stored_bits = 3 + 7 + 32 + (uint64_t)n * 8;
  if (bits >= stored_bits) {
    bw.put(final_block ? 1 : 0, 1);
    bw.put(0, 2);      // BTYPE 00
    bw.align_byte();
    const uint8_t hdr[4] = {(uint8_t)(n & 0xFF), (uint8_t)(n >> 8), (uint8_t)(~n & 0xFF), (uint8_t)((~n >> 8) & 0xFF)};
    bw.bytes(hdr, 4);
    bw.bytes(f, (size_t)n);
    bw_io = bw;
    return;
  }

  // ---- header of a dynamic block
  uint16_t code[kNumLit];
  canonical_codes(len, kNumLit, code);
  bw.put(final_block ? 1 : 0, 1);
  bw.put(2, 2);                       // BTYPE 10
  bw.put((uint32_t)(nlit - 257), 5);  // HLIT
  bw.put(0, 5);                       // HDIST: one distance code
  bw.put(15, 4);                      // HCLEN: all 19 code-length code lengths follow
  static const uint8_t kClOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
  for (int k = 0; k < 19; ++k) bw.put(kClOrder[k] >= 16 ? 0 : 4, 3);      // symbols 0..15: 4 bits each, 16..18 unused
  // with sixteen 4-bit codes the canonical code of symbol s is s itself (MSB first -> reversed here)
  for (int s = 0; s < nlit; ++s) bw.put(reverse_bits(len[s], 4), 4);
  bw.put(reverse_bits(1, 4), 4);      // the single distance code (distance 1): length 1, code '0'

  // ---- pass 2: literals between the runs, the runs as (length, distance 1) matches
  uint32_t lit[256];                  // code | length << 16
  for (int s = 0; s < 256; ++s) lit[s] = code[s] | ((uint32_t)len[s] << 16);
  int pos = 0;
  for (int r = 0; r < nr; ++r) {
    put_literals(bw, f, pos, runs[r].pos, lit);
    const int L = runs[r].len;
    const int sym = kLenLut.sym[L];
    bw.put(code[sym], len[sym]);
    if (kLenLut.extra_bits[L]) bw.put(kLenLut.extra_val[L], kLenLut.extra_bits[L]);
    bw.put(0, 1);                     // distance code 0 = distance 1, no extra bits
    pos = runs[r].pos + L;
  }
  put_literals(bw, f, pos, n, lit);
  bw.put(code[256], len[256]);
  bw_io = bw;
}

}  // namespace
}  // namespace nbc

extern "C" size_t nbc_png_idat_bound(int height, int width, int channels) {
  if (height <= 0 || width <= 0 || (channels != 1 && channels != 3)) return 0;
  const size_t row = (size_t)width * channels + 1;
  const size_t n = row * (size_t)height;
  // every segment costs at most its stored size + 5 bytes of block header (+ one more stored block per 65 535 bytes of
  // a very wide row); 2 bytes zlib header, 4 bytes Adler-32, slack for the 8-byte stores of the bit writer
  return n + 8 * (n / 32768 + (size_t)height + 2) + 64;
}

extern "C" int64_t nbc_png_idat(const uint8_t* pixels, int height, int width, int channels, int64_t row_stride_bytes,
                                const uint8_t* lut256, uint8_t* out, size_t out_cap) {
  using namespace nbc;
  if (pixels == nullptr || out == nullptr || height <= 0 || width <= 0 || (channels != 1 && channels != 3)) {
    set_error("nbc_png_idat: 8-bit grey (1 channel) or RGB (3 channels) image expected");
    return NBC_ERR_INVALID;
  }
  if (lut256 != nullptr && channels != 1) {
    set_error("nbc_png_idat: a lookup table applies to 1-channel images only");
    return NBC_ERR_INVALID;
  }
  const int64_t rb = (int64_t)width * channels;
  if (row_stride_bytes == 0) row_stride_bytes = rb;
  if (row_stride_bytes < rb) {
    set_error("nbc_png_idat: row stride %lld smaller than a row (%lld bytes)", (long long)row_stride_bytes, (long long)rb);
    return NBC_ERR_INVALID;
  }
  if (out_cap < nbc_png_idat_bound(height, width, channels)) {
    set_error("nbc_png_idat: output buffer of %zu bytes, need %zu", out_cap, nbc_png_idat_bound(height, width, channels));
    return NBC_ERR_INVALID;
  }
  const int64_t frow = rb + 1;      // filter byte + row
  BitWriter bw;
  bw.p = out;
  bw.end = out + out_cap - 8;
  bw.put(0x78, 8);                  // zlib header: deflate, 32 KB window
  bw.put(0x01, 8);                  // FLEVEL 0, FCHECK so that 0x7801 % 31 == 0
  uint32_t adler = 1;

  const int rows_per_seg = (int)std::max<int64_t>(1, kMaxSegBytes / frow);
  std::vector<uint8_t> seg((size_t)std::max<int64_t>(frow, (int64_t)rows_per_seg * frow));
  std::vector<Run> runs((size_t)kMaxSegBytes / 3 + 8);
  const bool sub = channels == 3 && width > 1;      // filter 1 (Sub) for RGB, 0 (None) for grey masks
  for (int r0 = 0; r0 < height; r0 += rows_per_seg) {
    const int nr = std::min(rows_per_seg, height - r0);
    for (int r = 0; r < nr; ++r) {
      const uint8_t* src = pixels + (int64_t)(r0 + r) * row_stride_bytes;
      uint8_t* dst = seg.data() + (int64_t)r * frow;
      if (sub) {
        dst[0] = 1;
        dst[1] = src[0], dst[2] = src[1], dst[3] = src[2];
        for (int64_t k = 3; k < rb; ++k) dst[1 + k] = (uint8_t)(src[k] - src[k - 3]);
      } else if (lut256 != nullptr) {
        dst[0] = 0;
        for (int64_t k = 0; k < rb; ++k) dst[1 + k] = lut256[src[k]];
      } else {
        dst[0] = 0;
        memcpy(dst + 1, src, (size_t)rb);
      }
    }
    const int64_t nbytes = (int64_t)nr * frow;
    adler = adler32_update(adler, seg.data(), (size_t)nbytes);
    const bool last_seg = r0 + nr >= height;
    // a row wider than 65 535 bytes is cut into several blocks
    for (int64_t o = 0; o < nbytes; o += kMaxSegBytes) {
      const int n = (int)std::min<int64_t>(kMaxSegBytes, nbytes - o);
      encode_segment(seg.data() + o, n, last_seg && o + n >= nbytes, bw, runs.data());
    }
  }
  bw.align_byte();
  const uint8_t tail[4] = {(uint8_t)(adler >> 24), (uint8_t)(adler >> 16), (uint8_t)(adler >> 8), (uint8_t)adler};
  bw.bytes(tail, 4);
  if (bw.overflow) {
    set_error("nbc_png_idat: output buffer too small (internal bound violated)");
    return NBC_ERR_INVALID;
  }
  return (int64_t)(bw.p - out);
}

// results/combined_images stand-in (reference models.py:280-347, a two-panel matplotlib figure; see pipeline.py): the
// processed image and the class mask in the figure's colours side by side at half resolution (every second pixel of
// every second row), `gap` white columns between them, under a title strip the caller has rendered.
extern "C" int nbc_compose_combined(const uint8_t* proc_rgb, const uint8_t* mask, int height, int width,
                                    const uint8_t* colours, const uint8_t* strip, int strip_h, int gap, uint8_t* canvas) {
  using namespace nbc;
  if (proc_rgb == nullptr || mask == nullptr || colours == nullptr || canvas == nullptr || height <= 0 || width <= 0 ||
      strip_h < 0 || gap < 0 || (strip_h > 0 && strip == nullptr)) {
    set_error("nbc_compose_combined: invalid argument");
    return NBC_ERR_INVALID;
  }
  const int hh = (height + 1) / 2, hw = (width + 1) / 2;
  const int64_t cw = 2 * (int64_t)hw + gap;      // canvas width in pixels
  if (strip_h > 0) memcpy(canvas, strip, (size_t)(strip_h * cw * 3));
  for (int r = 0; r < hh; ++r) {
    const uint8_t* p = proc_rgb + (int64_t)(2 * r) * width * 3;
    const uint8_t* m = mask + (int64_t)(2 * r) * width;
    uint8_t* o = canvas + ((int64_t)strip_h + r) * cw * 3;
    for (int c = 0; c < hw; ++c) {
      o[3 * c + 0] = p[6 * c + 0];
      o[3 * c + 1] = p[6 * c + 1];
      o[3 * c + 2] = p[6 * c + 2];
    }
    memset(o + (int64_t)hw * 3, 255, (size_t)gap * 3);
    uint8_t* q = o + ((int64_t)hw + gap) * 3;
    for (int c = 0; c < hw; ++c) {
      const uint8_t* col = colours + 3 * (m[2 * c] < 3 ? m[2 * c] : 0);
      q[3 * c + 0] = col[0];
      q[3 * c + 1] = col[1];
      q[3 * c + 2] = col[2];
    }
  }
  return 0;
}
