// Host-side .npy reader / writer for the re-encode CLI (reencode_spectrograms.py:49-62 np.load + pad,
// :69-81 trim + np.save).  Plain C++ (no device code); it lives in the library because ctypes releases the GIL
// around a foreign call, so the CLI's I/O threads really run in parallel - np.load / np.save spend most of their
// time per small file in Python under the GIL (profiles/cli_bench_r01.jsonl: 0.8 ms per file, not scaling with
// threads).
//
// Supported on read: format 1.0 / 2.0 / 3.0, C-order, 2-D, little-endian float32 / float64 / float16 - what
// convert_spectrograms.py writes and any "float dtype" mel the reference accepts.  Anything else returns 4 and the
// caller falls back to numpy.  Values are converted to float32 (as `batch_tensor = torch.tensor(..., float32)` does).
// The writer produces exactly the bytes np.save writes for a C-contiguous float32 2-D array (format 1.0, header
// padded with spaces to a multiple of 64 bytes, newline-terminated).
#include <cuda_fp16.h>
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/mqgan_b200.h"
#include "common.cuh"

namespace mq {

struct NpyInfo {
  long long rows, cols, data_offset;
  int dtype;            // 0 = f4, 1 = f8, 2 = f2
};

// 0 ok, 4 unsupported layout / dtype, 5 I/O error
static int npy_parse(FILE* f, NpyInfo* info) {
  unsigned char pre[12];
  if (fread(pre, 1, 10, f) != 10) return 5;
  if (memcmp(pre, "\x93NUMPY", 6) != 0) return 4;
  const int major = pre[6];
  size_t hlen, hoff;
  if (major == 1) {
    hlen = pre[8] | (static_cast<size_t>(pre[9]) << 8);
    hoff = 10;
  } else if (major == 2 || major == 3) {
    if (fread(pre + 10, 1, 2, f) != 2) return 5;
    hlen = pre[8] | (static_cast<size_t>(pre[9]) << 8) | (static_cast<size_t>(pre[10]) << 16) | (static_cast<size_t>(pre[11]) << 24);
    hoff = 12;
  } else {
    return 4;
  }
  if (hlen == 0 || hlen > (1u << 20)) return 4;
  std::string h(hlen, '\0');
  if (fread(&h[0], 1, hlen, f) != hlen) return 5;
  // 'descr': '<f4'
  size_t p = h.find("'descr'");
  if (p == std::string::npos) return 4;
  p = h.find(':', p);
  if (p == std::string::npos) return 4;
  size_t q0 = h.find('\'', p);
  if (q0 == std::string::npos) return 4;
  size_t q1 = h.find('\'', q0 + 1);
  if (q1 == std::string::npos) return 4;
  const std::string descr = h.substr(q0 + 1, q1 - q0 - 1);
  if (descr == "<f4" || descr == "=f4") info->dtype = 0;          // this library only runs on little-endian hosts
  else if (descr == "<f8" || descr == "=f8") info->dtype = 1;
  else if (descr == "<f2" || descr == "=f2") info->dtype = 2;
  else return 4;
  // 'fortran_order': False
  p = h.find("'fortran_order'");
  if (p == std::string::npos) return 4;
  p = h.find(':', p);
  if (p == std::string::npos) return 4;
  size_t v = h.find_first_not_of(" ", p + 1);
  if (v == std::string::npos || h.compare(v, 5, "False") != 0) return 4;
  // 'shape': (T, M)
  p = h.find("'shape'");
  if (p == std::string::npos) return 4;
  size_t lp = h.find('(', p), rp = h.find(')', p);
  if (lp == std::string::npos || rp == std::string::npos || rp < lp) return 4;
  std::vector<long long> dims;
  const char* s = h.c_str() + lp + 1;
  const char* end = h.c_str() + rp;
  while (s < end) {
    while (s < end && (*s == ' ' || *s == ',')) ++s;
    if (s >= end) break;
    char* next = nullptr;
    const long long d = strtoll(s, &next, 10);
    if (next == s || d < 0) return 4;
    dims.push_back(d);
    s = next;
    if (s < end && *s == 'L') ++s;                 // Python 2 longs
  }
  if (dims.size() != 2) return 4;
  info->rows = dims[0];
  info->cols = dims[1];
  info->data_offset = static_cast<long long>(hoff + hlen);
  return 0;
}

}  // namespace mq

using namespace mq;

extern "C" int mq_npy_probe(const char* path, int64_t* rows, int64_t* cols, int* dtype, int64_t* data_offset) {
  MQ_REQUIRE(path && rows && cols, "mq_npy_probe: null argument");
  FILE* f = fopen(path, "rb");
  if (f == nullptr) {
    set_last_error("mq_npy_probe: cannot open %s: %s", path, strerror(errno));
    return 5;
  }
  NpyInfo info;
  const int rc = npy_parse(f, &info);
  fclose(f);
  if (rc != 0) {
    set_last_error("mq_npy_probe: %s: %s", path, rc == 4 ? "not a C-order 2-D little-endian float .npy" : "read error");
    return rc;
  }
  *rows = info.rows;
  *cols = info.cols;
  if (dtype) *dtype = info.dtype;
  if (data_offset) *data_offset = info.data_offset;
  return 0;
}

extern "C" int mq_npy_read_f32(const char* path, float* dst, int64_t dst_rows, int64_t cols, int64_t* rows_out) {
  MQ_REQUIRE(path && dst && dst_rows >= 0 && cols > 0, "mq_npy_read_f32: bad argument");
  FILE* f = fopen(path, "rb");
  if (f == nullptr) {
    set_last_error("mq_npy_read_f32: cannot open %s: %s", path, strerror(errno));
    return 5;
  }
  NpyInfo info;
  int rc = npy_parse(f, &info);
  if (rc == 0 && info.cols != cols) {
    set_last_error("mq_npy_read_f32: %s has %lld columns, expected %lld", path, info.cols, (long long)cols);
    fclose(f);
    return 1;
  }
  if (rc != 0) {
    set_last_error("mq_npy_read_f32: %s: %s", path, rc == 4 ? "not a C-order 2-D little-endian float .npy" : "read error");
    fclose(f);
    return rc;
  }
  const long long take = info.rows < dst_rows ? info.rows : dst_rows;
  const size_t n = static_cast<size_t>(take) * static_cast<size_t>(cols);
  bool ok = true;
  if (info.dtype == 0) {
    ok = fread(dst, sizeof(float), n, f) == n;
  } else {
    const size_t esz = info.dtype == 1 ? 8 : 2;
    const size_t chunk = 1 << 16;
    std::vector<unsigned char> buf(chunk * esz);
    size_t done = 0;
    while (ok && done < n) {
      const size_t m = n - done < chunk ? n - done : chunk;
      ok = fread(buf.data(), esz, m, f) == m;
      if (!ok) break;
      if (info.dtype == 1) {
        const double* s = reinterpret_cast<const double*>(buf.data());
        for (size_t i = 0; i < m; ++i) dst[done + i] = static_cast<float>(s[i]);
      } else {
        const __half* s = reinterpret_cast<const __half*>(buf.data());
        for (size_t i = 0; i < m; ++i) dst[done + i] = __half2float(s[i]);
      }
      done += m;
    }
  }
  fclose(f);
  if (!ok) {
    set_last_error("mq_npy_read_f32: %s: file shorter than its header says", path);
    return 5;
  }
  if (take < dst_rows)                                      // zero padding up to the batch's longest utterance
    memset(dst + n, 0, static_cast<size_t>(dst_rows - take) * static_cast<size_t>(cols) * sizeof(float));
  if (rows_out) *rows_out = info.rows;
  return 0;
}

extern "C" int mq_npy_write_f32(const char* path, const float* src, int64_t rows, int64_t cols) {
  MQ_REQUIRE(path && (src || rows == 0) && rows >= 0 && cols >= 0, "mq_npy_write_f32: bad argument");
  char dict[160];
  const int dl = snprintf(dict, sizeof(dict), "{'descr': '<f4', 'fortran_order': False, 'shape': (%lld, %lld), }",
                          (long long)rows, (long long)cols);
  // magic(6) + version(2) + header length(2) + header, padded with spaces so the data starts 64-byte aligned
  const size_t unpadded = 10 + static_cast<size_t>(dl) + 1;
  const size_t total = (unpadded + 63) / 64 * 64;
  const size_t hlen = total - 10;
  std::string out;
  out.reserve(total);
  out.append("\x93NUMPY", 6);
  out.push_back('\x01');
  out.push_back('\x00');
  out.push_back(static_cast<char>(hlen & 0xff));
  out.push_back(static_cast<char>((hlen >> 8) & 0xff));
  out.append(dict, static_cast<size_t>(dl));
  out.append(total - unpadded, ' ');
  out.push_back('\n');
  FILE* f = fopen(path, "wb");
  if (f == nullptr) {
    set_last_error("mq_npy_write_f32: cannot create %s: %s", path, strerror(errno));
    return 5;
  }
  const size_t n = static_cast<size_t>(rows) * static_cast<size_t>(cols);
  bool ok = fwrite(out.data(), 1, out.size(), f) == out.size();
  if (ok && n > 0) ok = fwrite(src, sizeof(float), n, f) == n;
  ok = (fclose(f) == 0) && ok;
  if (!ok) {
    set_last_error("mq_npy_write_f32: short write to %s: %s", path, strerror(errno));
    return 5;
  }
  return 0;
}
