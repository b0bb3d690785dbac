// g2048_csv.cpp — the reference's transition CSV (training_data.export_csv / import_csv,
// /root/reference/training_data.py:188-248) for HOST buffers.  Byte-for-byte the file numpy.savetxt
// writes there: fmt '%d,'*17 + '%f,' + '%d,'*16 + '%i' [+ ',%f'], header row without comment prefix,
// '\n' line ends.  Boards are exponents in memory and tile values (2^e) in the file.
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/g2048.h"
#include "g2048_internal.h"

namespace {

using g2048::fail;

char* put_u64(char* p, uint64_t v) {
  char tmp[24];
  int k = 0;
  do { tmp[k++] = (char)('0' + v % 10); v /= 10; } while (v);
  while (k) *p++ = tmp[--k];
  return p;
}

// '%f' of a double.  Rewards are almost always small exact integers: print those directly.
char* put_f(char* p, double v) {
  if (v >= 0.0 && v < 1e15 && v == (double)(uint64_t)v) {
    p = put_u64(p, (uint64_t)v);
    std::memcpy(p, ".000000", 7);
    return p + 7;
  }
  return p + std::snprintf(p, 400, "%f", v);
}

char* put_board(char* p, const uint8_t* b) {
  for (int c = 0; c < 16; ++c) {
    p = put_u64(p, b[c] ? (1ull << (b[c] & 63)) : 0ull);
    *p++ = ',';
  }
  return p;
}

std::string header(bool with_returns) {
  std::string h;
  char buf[32];
  for (int m = 1; m <= 4; ++m)
    for (int n = 1; n <= 4; ++n) { std::snprintf(buf, sizeof buf, "%d-%d,", m, n); h += buf; }
  h += "action,reward,";
  for (int m = 1; m <= 4; ++m)
    for (int n = 1; n <= 4; ++n) { std::snprintf(buf, sizeof buf, "next %d-%d,", m, n); h += buf; }
  h += "done";
  if (with_returns) h += ",return";
  h += "\n";
  return h;
}

// tile value -> exponent; -1 when the value is neither 0 nor a power of two in 2..2^31
int exp_of(long long v) {
  if (v == 0) return 0;
  if (v < 2 || v > (1ll << 31) || (v & (v - 1))) return -1;
  return __builtin_ctzll((unsigned long long)v);
}

struct File {
  FILE* f;
  explicit File(FILE* f_) : f(f_) {}
  ~File() { if (f) std::fclose(f); }
};

bool read_all(const char* path, std::vector<char>& out) {
  File fp(std::fopen(path, "rb"));
  if (!fp.f) return false;
  if (std::fseek(fp.f, 0, SEEK_END) == 0) {                       // regular file: one read of the known size
    const long size = std::ftell(fp.f);
    if (size >= 0 && std::fseek(fp.f, 0, SEEK_SET) == 0) {
      out.resize((size_t)size + 1);
      const size_t k = std::fread(out.data(), 1, (size_t)size, fp.f);
      out.resize(k + 1);
      out[k] = '\0';
      return true;
    }
  }
  char buf[1 << 16];
  size_t k;
  while ((k = std::fread(buf, 1, sizeof buf, fp.f)) > 0) out.insert(out.end(), buf, buf + k);
  out.push_back('\0');
  return true;
}

// [begin, end) of the data lines (header skipped, blank lines dropped)
void data_lines(std::vector<char>& txt, std::vector<char*>& lines, int* header_cols) {
  char* p = txt.data();
  char* eol = std::strchr(p, '\n');
  int cols = 1;
  for (char* q = p; *q && q != eol; ++q) cols += (*q == ',');
  *header_cols = cols;
  if (!eol) return;
  p = eol + 1;
  while (*p) {
    char* e = std::strchr(p, '\n');
    if (e) *e = '\0';
    char* q = p;
    while (*q == ' ' || *q == '\r' || *q == '\t') ++q;
    if (*q) lines.push_back(p);
    if (!e) break;
    p = e + 1;
  }
}

// Integer column: optional sign, decimal digits (what "%d" / "%i" print).  *end == p on failure.
inline long long parse_int(char* p, char** end) {
  char* q = p;
  bool neg = false;
  if (*q == '-' || *q == '+') neg = *q++ == '-';
  if (*q < '0' || *q > '9') { *end = p; return 0; }
  unsigned long long v = 0;
  while (*q >= '0' && *q <= '9') v = v * 10u + (unsigned)(*q++ - '0');
  *end = q;
  return neg ? -(long long)v : (long long)v;
}
// Float column.  "%f" output of an integral value below 2^53 — digits '.' zeros, every reward the step
// produces — is converted exactly without strtod; anything else goes through strtod.
inline double parse_float(char* p, char** end) {
  char* q = p;
  bool neg = false;
  if (*q == '-') { neg = true; ++q; }
  unsigned long long v = 0;
  int digits = 0;
  while (*q >= '0' && *q <= '9' && digits < 15) { v = v * 10u + (unsigned)(*q++ - '0'); ++digits; }
  if (digits > 0 && *q == '.') {
    char* z = q + 1;
    while (*z == '0') ++z;
    if (*z == ',' || *z == '\0' || *z == '\r' || *z == '\n') { *end = z; return neg ? -(double)v : (double)v; }
  }
  return std::strtod(p, end);
}

}  // namespace

extern "C" {

int g2048_csv_export(const char* path, const uint8_t* boards, const uint8_t* actions, const double* rewards,
                     const uint8_t* next_boards, const uint8_t* dones, const double* returns, uint64_t n,
                     int append) {
  if (!path) return fail(G2048_ERR_INVALID, "g2048_csv_export: path is NULL");
  if (n && (!boards || !actions || !rewards || !next_boards || !dones))
    return fail(G2048_ERR_INVALID, "g2048_csv_export: NULL column");
  File fp(std::fopen(path, append ? "ab" : "wb"));
  if (!fp.f) return fail(G2048_ERR_INVALID, "g2048_csv_export: cannot open %s: %s", path, std::strerror(errno));
  if (!append) {
    const std::string h = header(returns != nullptr);
    if (std::fwrite(h.data(), 1, h.size(), fp.f) != h.size())
      return fail(G2048_ERR_INVALID, "g2048_csv_export: write to %s failed", path);
  }
  std::vector<char> buf((1 << 20) + 2048);
  char* p = buf.data();
  for (uint64_t i = 0; i < n; ++i) {
    p = put_board(p, boards + 16 * i);
    p = put_u64(p, actions[i]);
    *p++ = ',';
    p = put_f(p, rewards[i]);
    *p++ = ',';
    p = put_board(p, next_boards + 16 * i);
    *p++ = dones[i] ? '1' : '0';
    if (returns) { *p++ = ','; p = put_f(p, returns[i]); }
    *p++ = '\n';
    if ((size_t)(p - buf.data()) >= (1u << 20) || i + 1 == n) {
      const size_t len = (size_t)(p - buf.data());
      if (std::fwrite(buf.data(), 1, len, fp.f) != len)
        return fail(G2048_ERR_INVALID, "g2048_csv_export: write to %s failed", path);
      p = buf.data();
    }
  }
  return G2048_OK;
}

int g2048_csv_rows(const char* path, uint64_t* n_rows, int* has_returns) {
  if (!path || !n_rows) return fail(G2048_ERR_INVALID, "g2048_csv_rows: NULL argument");
  std::vector<char> txt;
  if (!read_all(path, txt)) return fail(G2048_ERR_INVALID, "g2048_csv_rows: cannot open %s: %s", path, std::strerror(errno));
  std::vector<char*> lines;
  int cols = 0;
  data_lines(txt, lines, &cols);
  if (cols != 35 && cols != 36) return fail(G2048_ERR_INVALID, "g2048_csv_rows: %s has %d columns, expected 35 or 36", path, cols);
  *n_rows = lines.size();
  if (has_returns) *has_returns = cols == 36;
  return G2048_OK;
}

int g2048_csv_import(const char* path, uint8_t* boards, uint8_t* actions, double* rewards, uint8_t* next_boards,
                     uint8_t* dones, double* returns, uint64_t n_rows) {
  if (!path) return fail(G2048_ERR_INVALID, "g2048_csv_import: path is NULL");
  if (n_rows && (!boards || !actions || !rewards || !next_boards || !dones))
    return fail(G2048_ERR_INVALID, "g2048_csv_import: NULL column");
  std::vector<char> txt;
  if (!read_all(path, txt)) return fail(G2048_ERR_INVALID, "g2048_csv_import: cannot open %s: %s", path, std::strerror(errno));
  std::vector<char*> lines;
  int cols = 0;
  data_lines(txt, lines, &cols);
  if (cols != 35 && cols != 36) return fail(G2048_ERR_INVALID, "g2048_csv_import: %s has %d columns, expected 35 or 36", path, cols);
  if (lines.size() != n_rows)
    return fail(G2048_ERR_INVALID, "g2048_csv_import: %s has %zu rows, caller expects %llu", path, lines.size(),
                (unsigned long long)n_rows);
  if (returns && cols != 36) return fail(G2048_ERR_INVALID, "g2048_csv_import: %s has no return column", path);
  for (uint64_t i = 0; i < n_rows; ++i) {
    char* p = lines[i];
    for (int c = 0; c < cols; ++c) {
      char* e = nullptr;
      const bool is_float = (c == 17 || c == 35);
      double fv = 0.0;
      long long iv = 0;
      if (is_float) fv = parse_float(p, &e);
      else iv = parse_int(p, &e);
      if (e == p || (c + 1 < cols ? *e != ',' : (*e != '\0' && *e != '\r')))
        return fail(G2048_ERR_INVALID, "g2048_csv_import: %s row %llu column %d: cannot parse", path,
                    (unsigned long long)(i + 1), c + 1);
      p = e + 1;
      if (c < 16 || (c >= 18 && c < 34)) {
        const int ex = exp_of(iv);
        if (ex < 0)
          return fail(G2048_ERR_INVALID, "g2048_csv_import: %s row %llu column %d: %lld is not 0 or a power of two",
                      path, (unsigned long long)(i + 1), c + 1, iv);
        (c < 16 ? boards + 16 * i + c : next_boards + 16 * i + (c - 18))[0] = (uint8_t)ex;
      } else if (c == 16) {
        if (iv < 0 || iv > 3)
          return fail(G2048_ERR_INVALID, "g2048_csv_import: %s row %llu: action %lld not in 0..3", path,
                      (unsigned long long)(i + 1), iv);
        actions[i] = (uint8_t)iv;
      } else if (c == 17) {
        rewards[i] = fv;
      } else if (c == 34) {
        dones[i] = iv != 0;
      } else if (returns) {
        returns[i] = fv;
      }
    }
  }
  return G2048_OK;
}

}  // extern "C"
