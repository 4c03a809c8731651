// Host-side scan for the dark bands of a raw scan (plain C++ in the library, no GPU work).
//
// A scanner image carries dark bands above and below the bark (that is what trim_black, models.py:157-166, removes).
// Where those bands are EXACTLY zero the 4x resize of K1 maps them to exact zeros whatever the neighbouring rows hold
// (each output row reads its own four source rows only), so they need not cross PCIe at all: the engine copies the rows
// in between and tells K1 the span (nbc_preprocess_4x_span_u8).  This function finds the span: the memory rows before
// the first and after the last row that holds a non-zero byte, widened outwards to whole groups of `group` rows.  It reads
// only the zero rows (plus one), 8 bytes at a time; an image without zero bands costs two cache lines.
#include <cstdint>
#include <cstring>

#include "common.cuh"

namespace {

// true iff the n bytes at p are all zero
inline bool all_zero(const uint8_t* p, int64_t n) {
  int64_t i = 0;
  for (; i < n && (reinterpret_cast<uintptr_t>(p + i) & 7); ++i)
    if (p[i]) return false;
  const uint64_t* q = reinterpret_cast<const uint64_t*>(p + i);
  const int64_t words = (n - i) >> 3;
  int64_t w = 0;
  for (; w + 8 <= words; w += 8) {      // 64 bytes per test: the compiler turns the OR tree into vector ops
    const uint64_t acc = q[w] | q[w + 1] | q[w + 2] | q[w + 3] | q[w + 4] | q[w + 5] | q[w + 6] | q[w + 7];
    if (acc) return false;
  }
  for (; w < words; ++w)
    if (q[w]) return false;
  for (i += words << 3; i < n; ++i)
    if (p[i]) return false;
  return true;
}

}  // namespace

extern "C" int nbc_host_zero_row_span(const uint8_t* raw, int H, int64_t pitch, int64_t row_bytes, int group, int32_t* row0,
                                      int32_t* rows) {
  NBC_REQUIRE(raw && row0 && rows && H > 0 && row_bytes > 0 && pitch >= row_bytes && group > 0,
              "nbc_host_zero_row_span: bad argument");
  int first = 0;
  while (first < H && all_zero(raw + (int64_t)first * pitch, row_bytes)) ++first;
  if (first == H) {      // an all-zero image: nothing to copy
    *row0 = 0, *rows = 0;
    return 0;
  }
  int last = H;          // one past the last non-zero row
  while (last > first + 1 && all_zero(raw + (int64_t)(last - 1) * pitch, row_bytes)) --last;
  first = first / group * group;
  last = (last + group - 1) / group * group;
  if (last > H) last = H;
  *row0 = first, *rows = last - first;
  return 0;
}
