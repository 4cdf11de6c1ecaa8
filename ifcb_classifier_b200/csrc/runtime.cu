// Error string + device helpers shared by every translation unit.
#include "common.cuh"
#include "../../include/ifcb_b200.h"
#include <cstring>

namespace ifcb {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

int sm_count() {
  static int cached = 0;
  if (cached > 0) return cached;
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  cached = n;
  return n;
}
}  // namespace ifcb

extern "C" int ifcb_abi_version(void) { return IFCB_B200_ABI_VERSION; }
extern "C" const char* ifcb_last_error(void) { return ifcb::get_error(); }
extern "C" int ifcb_sm_count(void) { return ifcb::sm_count(); }

// ---------------------------------------------------------------------------------------------
// Host-side .adc parser (the CSV pyifcb reads with pandas upstream: reference neuston_data.py:446-454 via
// bin.images / ifcb.data.adc).  One pass over the file image, no allocation: per row the three columns the
// ROI table needs (ROI_WIDTH, ROI_HEIGHT, START_BYTE); rows with zero area are dropped, target = 1-based row.
// ---------------------------------------------------------------------------------------------
#include <cstdlib>
extern "C" int64_t ifcb_parse_adc(const char* buf, int64_t len, int col_w, int col_h, int col_b, int64_t max_rows,
                                  int32_t* targets, int64_t* offsets, int32_t* heights, int32_t* widths) {
  if (!buf || len < 0 || col_w < 0 || col_h < 0 || col_b < 0 || !targets || !offsets || !heights || !widths) {
    ifcb::set_error("ifcb_parse_adc: bad argument");
    return -1;
  }
  const int need = col_w > col_h ? (col_w > col_b ? col_w : col_b) : (col_h > col_b ? col_h : col_b);
  int64_t kept = 0, row = 0, i = 0;
  while (i < len) {
    // one line: [i, e)
    int64_t e = i;
    while (e < len && buf[e] != '\n') ++e;
    ++row;
    double v[3] = {0, 0, 0};
    int col = 0;
    int64_t f = i;
    bool complete = false;
    for (int64_t j = i; j <= e; ++j) {
      if (j == e || buf[j] == ',') {
        if (col == col_w || col == col_h || col == col_b) {
          char tmp[64];
          int64_t n = j - f;
          if (n > 63) n = 63;
          for (int64_t t = 0; t < n; ++t) tmp[t] = buf[f + t];
          tmp[n] = 0;
          const double x = strtod(tmp, nullptr);
          if (col == col_w) v[0] = x;
          if (col == col_h) v[1] = x;
          if (col == col_b) v[2] = x;
        }
        if (col == need) complete = true;
        ++col;
        f = j + 1;
      }
    }
    if (e > i && complete) {
      const long long w = (long long)v[0], h = (long long)v[1];
      if (w * h > 0) {
        if (kept >= max_rows) {
          ifcb::set_error("ifcb_parse_adc: more than %lld rows", (long long)max_rows);
          return -1;
        }
        targets[kept] = (int32_t)row;
        offsets[kept] = (int64_t)v[2];
        heights[kept] = (int32_t)h;
        widths[kept] = (int32_t)w;
        ++kept;
      }
    }
    i = e + 1;
  }
  return kept;
}
