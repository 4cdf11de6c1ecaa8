// Error string + device helpers shared by every translation unit.
#include "common.cuh"
#include "../../include/ifcb_b200.h"
#include <cstring>
#include <cmath>

namespace ifcb {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

// TRAIN deterministic mode (Trainer(deterministic=True) upstream): a caller-owned device workspace; while it is set, the weight
// gradient's split-K partials and the BatchNorm reductions' per-block partials are STORED there and summed in a fixed order
// instead of being combined with floating-point atomics.
static void* g_det_ws = nullptr;
static long long g_det_bytes = 0;
void* det_workspace(long long need) { return (g_det_ws && need <= g_det_bytes) ? g_det_ws : nullptr; }
bool det_enabled() { return g_det_ws != nullptr; }

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
extern "C" int ifcb_train_deterministic(void* d_workspace, int64_t bytes) {
  IFCB_ARG_CHECK((d_workspace == nullptr) == (bytes == 0) && bytes >= 0, "ifcb_train_deterministic: pass a workspace and its size, or NULL and 0");
  IFCB_ARG_CHECK((reinterpret_cast<uintptr_t>(d_workspace) & 255) == 0, "ifcb_train_deterministic: workspace must be 256-byte aligned");
  ifcb::g_det_ws = d_workspace;
  ifcb::g_det_bytes = bytes;
  return 0;
}

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

// ---------------------------------------------------------------------------------------------
// JSON text of a score matrix, byte-identical to Python's json.dumps(scores.tolist()) as the reference writes it
// (neuston_callbacks.py:213-230): every float32 is widened to float64 and printed with the shortest digits that
// round-trip (repr): fixed notation for 1e-4 <= |x| < 1e16, else d.ddde-XX, always with a '.0' or an exponent.
// Python's own encoder spends ~0.8 us per number holding the GIL (160 ms per 2048 x 100 bin); this is ~30x faster
// and the ctypes call releases the GIL, so the result writers of do_run overlap the GPU.
// ---------------------------------------------------------------------------------------------
#include <charconv>
namespace {
inline char* put_repr(double v, char* o) {
  if (v != v) { memcpy(o, "NaN", 3); return o + 3; }
  if (v == 0.0) { if (std::signbit(v)) *o++ = '-'; memcpy(o, "0.0", 3); return o + 3; }
  if (v < 0) { *o++ = '-'; v = -v; }
  if (v > 1.7976931348623157e308) { memcpy(o, "Infinity", 8); return o + 8; }
  char tmp[40];
  auto r = std::to_chars(tmp, tmp + sizeof(tmp), v, std::chars_format::scientific);      // d[.ddd]e[+-]XX, shortest digits
  char* e = tmp;
  while (*e != 'e') ++e;
  char digits[24];
  int nd = 0;
  for (char* c = tmp; c < e; ++c) if (*c != '.') digits[nd++] = *c;
  int exp10 = 0;
  { const char* c = e + 1; const bool neg = *c == '-'; if (*c == '+' || *c == '-') ++c; while (c < r.ptr) exp10 = exp10 * 10 + (*c++ - '0'); if (neg) exp10 = -exp10; }
  if (exp10 >= -4 && exp10 < 16) {
    if (exp10 < 0) {                                  // 0.000ddd
      *o++ = '0'; *o++ = '.';
      for (int z = 0; z < -exp10 - 1; ++z) *o++ = '0';
      for (int i = 0; i < nd; ++i) *o++ = digits[i];
    } else {
      int i = 0;
      for (; i <= exp10; ++i) *o++ = i < nd ? digits[i] : '0';
      *o++ = '.';
      if (i >= nd) *o++ = '0';
      for (; i < nd; ++i) *o++ = digits[i];
    }
  } else {
    *o++ = digits[0];
    if (nd > 1) { *o++ = '.'; for (int i = 1; i < nd; ++i) *o++ = digits[i]; }
    *o++ = 'e';
    *o++ = exp10 < 0 ? '-' : '+';
    int a = exp10 < 0 ? -exp10 : exp10;
    char eb[8]; int ne = 0;
    while (a) { eb[ne++] = (char)('0' + a % 10); a /= 10; }
    if (ne < 2) *o++ = '0';
    while (ne) *o++ = eb[--ne];
  }
  return o;
}
}  // namespace

extern "C" int64_t ifcb_format_scores_json(const float* scores, int64_t rows, int64_t cols, char* out, int64_t cap) {
  if (!scores || !out || rows < 0 || cols < 0) { ifcb::set_error("ifcb_format_scores_json: bad argument"); return -1; }
  if (cap < rows * (cols * 28 + 4) + 4) { ifcb::set_error("ifcb_format_scores_json: output buffer too small"); return -1; }
  char* o = out;
  *o++ = '[';
  for (int64_t r = 0; r < rows; ++r) {
    if (r) { *o++ = ','; *o++ = ' '; }
    *o++ = '[';
    for (int64_t c = 0; c < cols; ++c) {
      if (c) { *o++ = ','; *o++ = ' '; }
      o = put_repr((double)scores[r * cols + c], o);
    }
    *o++ = ']';
  }
  *o++ = ']';
  return o - out;
}
