// Native witness I/O: the JSON files the Python side writes into rust_files/<tag>/ for the Rust
// driver (src/convolution/Server.py:311-417), read the way vPIN_proof_generation/src/load_data.rs:5-62 and
// load_data_add.rs:5-102 read them, plus a binary sidecar of the same content (LeNet layer 5 is 3 x 6000 rows of 32 decimal
// integers as JSON; the sidecar is what a server keeps next to them for the next proof of the same witness).
//   rust_files/<tag>/pointMult/weight.json               ["123", ...]           decimal strings -> u128 (load_data.rs:18-23)
//   rust_files/<tag>/pointMult/point_mult_p{x,y}_byte.json  [[b0, ..., b31], ...]  little-endian bytes of an F_l element
//   rust_files/<tag>/pointAdd/point_add_{px,py,rx,ry}_byte.json                   the same
//   rust_files/<tag>/pointAdd/point_add_rz_byte.json      [0, 1, ...]            1 marks R = infinity
// Like the reference's `value.as_i64()` filter, array entries that are not integers are skipped; a byte row shorter than 32
// entries is zero-padded, each entry is truncated to 8 bits (the Rust builders cast with `as u8`).
// Sidecar rust_files/<tag>/{pointMult,pointAdd}/witness.bin: "VPINWIT1" | u32 kind (1 mult, 2 add) | u32 0 | u64 count | payload
//   kind 1: count x 16 B weights (u128 LE) | count x 32 B px | count x 32 B py
//   kind 2: count x 32 B px | py | rx | ry | count x 1 B rz
// Pure host code: no CUDA here.
#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/vpin_b200.h"

namespace {

struct IoError {
  vpin_status code;
  std::string msg;
};
thread_local std::string g_io_error;

std::string read_file(const std::string &path) {
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) throw IoError{VPIN_ERR_IO, "cannot open " + path + ": " + strerror(errno)};
  std::string s;
  char buf[1 << 16];
  size_t n;
  while ((n = fread(buf, 1, sizeof(buf), f)) > 0) s.append(buf, n);
  fclose(f);
  return s;
}
bool file_exists(const std::string &path) {
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) return false;
  fclose(f);
  return true;
}

// ---- the JSON subset these files use: arrays of (arrays of) numbers or strings ----
struct Json {
  const std::string &s;
  const std::string &path;
  size_t i = 0;
  Json(const std::string &s_, const std::string &p_) : s(s_), path(p_) {}
  [[noreturn]] void fail(const char *what) const { throw IoError{VPIN_ERR_IO, path + ": malformed JSON (" + what + ") at byte " + std::to_string(i)}; }
  void ws() { while (i < s.size() && (s[i] == ' ' || s[i] == '\n' || s[i] == '\t' || s[i] == '\r')) i++; }
  bool eat(char c) { ws(); if (i < s.size() && s[i] == c) { i++; return true; } return false; }
  void expect(char c) { if (!eat(c)) fail("unexpected character"); }
  char peek() { ws(); if (i >= s.size()) fail("unexpected end"); return s[i]; }
  // a number: returns true and the value if it is an integer that fits i64 (serde_json Value::as_i64), false otherwise
  bool number(int64_t *out) {
    ws();
    size_t start = i;
    bool neg = false, integral = true, overflow = false;
    if (i < s.size() && s[i] == '-') { neg = true; i++; }
    if (i >= s.size() || s[i] < '0' || s[i] > '9') fail("number expected");
    uint64_t v = 0;
    while (i < s.size() && s[i] >= '0' && s[i] <= '9') {
      if (v > (UINT64_MAX - (uint64_t)(s[i] - '0')) / 10) overflow = true; else v = v * 10 + (uint64_t)(s[i] - '0');
      i++;
    }
    if (i < s.size() && (s[i] == '.' || s[i] == 'e' || s[i] == 'E')) {
      integral = false;
      while (i < s.size() && (s[i] == '.' || s[i] == 'e' || s[i] == 'E' || s[i] == '+' || s[i] == '-' || (s[i] >= '0' && s[i] <= '9'))) i++;
    }
    (void)start;
    if (!integral || overflow) return false;
    if (neg) { if (v > (uint64_t)INT64_MAX + 1) return false; *out = (int64_t)(0 - v); }
    else { if (v > (uint64_t)INT64_MAX) return false; *out = (int64_t)v; }
    return true;
  }
  std::string string() {
    expect('"');
    std::string out;
    while (i < s.size() && s[i] != '"') {
      if (s[i] == '\\') fail("escape sequences are not expected in these files");
      out.push_back(s[i++]);
    }
    if (i >= s.size()) fail("unterminated string");
    i++;
    return out;
  }
  void skip_value() {  // anything that is not a number: string, literal, nested container
    char c = peek();
    if (c == '"') { string(); return; }
    if (c == '[' || c == '{') {
      char close = c == '[' ? ']' : '}';
      i++;
      if (eat(close)) return;
      do {
        if (close == '}') { string(); expect(':'); }
        skip_value();
      } while (eat(','));
      expect(close);
      return;
    }
    if (c == '-' || (c >= '0' && c <= '9')) { int64_t d; number(&d); return; }
    while (i < s.size() && ((s[i] >= 'a' && s[i] <= 'z'))) i++;  // true / false / null
  }
  void end() { ws(); if (i != s.size()) fail("trailing characters"); }
};

// [[b, ...], ...] -> count x 32 bytes (load_data.rs:32-42: non-integer entries are dropped)
std::vector<uint8_t> byte_rows(const std::string &path, uint64_t *count) {
  std::string text = read_file(path);
  Json j(text, path);
  std::vector<uint8_t> out;
  uint64_t n = 0;
  j.expect('[');
  if (!j.eat(']')) {
    do {
      j.expect('[');
      uint8_t row[32] = {0};
      size_t k = 0;
      if (!j.eat(']')) {
        do {
          char c = j.peek();
          int64_t v;
          if (c == '-' || (c >= '0' && c <= '9')) {
            if (j.number(&v)) {
              if (k < 32) row[k] = (uint8_t)v;
              k++;
            }
          } else {
            j.skip_value();
          }
        } while (j.eat(','));
        j.expect(']');
      }
      if (k > 32) throw IoError{VPIN_ERR_SIZE_MISMATCH, path + ": row " + std::to_string(n) + " has more than 32 bytes"};
      out.insert(out.end(), row, row + 32);
      n++;
    } while (j.eat(','));
    j.expect(']');
  }
  j.end();
  *count = n;
  return out;
}
// ["123", ...] -> u128 as (lo, hi) pairs (load_data.rs:18-23: u128::from_str, a failure is a panic there, an error here)
std::vector<uint64_t> weight_strings(const std::string &path, uint64_t *count) {
  std::string text = read_file(path);
  Json j(text, path);
  std::vector<uint64_t> out;
  j.expect('[');
  if (!j.eat(']')) {
    do {
      std::string w = j.string();
      if (w.empty()) throw IoError{VPIN_ERR_INVALID_SCALAR, path + ": empty weight"};
      unsigned __int128 v = 0;
      size_t k = w[0] == '+' ? 1 : 0;  // u128::from_str accepts a leading '+'
      if (k == w.size()) throw IoError{VPIN_ERR_INVALID_SCALAR, path + ": bad weight '" + w + "'"};
      for (; k < w.size(); k++) {
        if (w[k] < '0' || w[k] > '9') throw IoError{VPIN_ERR_INVALID_SCALAR, path + ": bad weight '" + w + "'"};
        unsigned __int128 nv = v * 10 + (unsigned)(w[k] - '0');
        if (nv / 10 != v) throw IoError{VPIN_ERR_INVALID_SCALAR, path + ": weight does not fit u128: '" + w + "'"};
        v = nv;
      }
      out.push_back((uint64_t)v);
      out.push_back((uint64_t)(v >> 64));
    } while (j.eat(','));
    j.expect(']');
  }
  j.end();
  *count = out.size() / 2;
  return out;
}
std::vector<int64_t> int_list(const std::string &path, uint64_t *count) {
  std::string text = read_file(path);
  Json j(text, path);
  std::vector<int64_t> out;
  j.expect('[');
  if (!j.eat(']')) {
    do {
      char c = j.peek();
      int64_t v;
      if (c == '-' || (c >= '0' && c <= '9')) { if (j.number(&v)) out.push_back(v); }
      else j.skip_value();
    } while (j.eat(','));
    j.expect(']');
  }
  j.end();
  *count = out.size();
  return out;
}

struct MultWitness { uint64_t n = 0; std::vector<uint64_t> weights; std::vector<uint8_t> px, py; };
struct AddWitness { uint64_t n = 0; std::vector<uint8_t> px, py, rx, ry; std::vector<int64_t> rz; };

const char kMagic[8] = {'V', 'P', 'I', 'N', 'W', 'I', 'T', '1'};
struct BinHeader { char magic[8]; uint32_t kind, reserved; uint64_t count; };

std::string dir_of(const char *root, const char *tag, const char *sub) { return std::string(root) + "/rust_files/" + tag + "/" + sub + "/"; }

void read_exact(FILE *f, void *dst, size_t bytes, const std::string &path) {
  if (bytes && fread(dst, 1, bytes, f) != bytes) { fclose(f); throw IoError{VPIN_ERR_IO, path + ": truncated"}; }
}
FILE *open_bin(const std::string &path, uint32_t kind, uint64_t *count) {
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) throw IoError{VPIN_ERR_IO, "cannot open " + path + ": " + strerror(errno)};
  BinHeader h;
  read_exact(f, &h, sizeof(h), path);
  if (memcmp(h.magic, kMagic, 8) != 0 || h.kind != kind) { fclose(f); throw IoError{VPIN_ERR_IO, path + ": not a witness sidecar of this kind"}; }
  *count = h.count;
  return f;
}

MultWitness load_mult_json(const std::string &d) {
  MultWitness w;
  uint64_t nx = 0, ny = 0;
  w.weights = weight_strings(d + "weight.json", &w.n);
  w.px = byte_rows(d + "point_mult_px_byte.json", &nx);
  w.py = byte_rows(d + "point_mult_py_byte.json", &ny);
  if (nx != w.n || ny != w.n) throw IoError{VPIN_ERR_SIZE_MISMATCH, d + ": weight / px / py have different lengths"};
  return w;
}
AddWitness load_add_json(const std::string &d) {
  AddWitness w;
  uint64_t n[5] = {0, 0, 0, 0, 0};
  w.px = byte_rows(d + "point_add_px_byte.json", &n[0]);
  w.py = byte_rows(d + "point_add_py_byte.json", &n[1]);
  w.rx = byte_rows(d + "point_add_rx_byte.json", &n[2]);
  w.ry = byte_rows(d + "point_add_ry_byte.json", &n[3]);
  w.rz = int_list(d + "point_add_rz_byte.json", &n[4]);
  w.n = n[0];  // load_data_add.rs:37: the count is the length of px
  for (int k = 1; k < 5; k++)
    if (n[k] != w.n) throw IoError{VPIN_ERR_SIZE_MISMATCH, d + ": the five point_add files have different lengths"};
  return w;
}
MultWitness load_mult(const char *root, const char *tag) {
  std::string d = dir_of(root, tag, "pointMult"), bin = d + "witness.bin";
  if (!file_exists(bin)) return load_mult_json(d);
  MultWitness w;
  FILE *f = open_bin(bin, 1, &w.n);
  w.weights.resize(2 * w.n); w.px.resize(32 * w.n); w.py.resize(32 * w.n);
  read_exact(f, w.weights.data(), 16 * w.n, bin);
  read_exact(f, w.px.data(), 32 * w.n, bin);
  read_exact(f, w.py.data(), 32 * w.n, bin);
  fclose(f);
  return w;
}
AddWitness load_add(const char *root, const char *tag) {
  std::string d = dir_of(root, tag, "pointAdd"), bin = d + "witness.bin";
  if (!file_exists(bin)) return load_add_json(d);
  AddWitness w;
  FILE *f = open_bin(bin, 2, &w.n);
  for (auto *v : {&w.px, &w.py, &w.rx, &w.ry}) { v->resize(32 * w.n); read_exact(f, v->data(), 32 * w.n, bin); }
  std::vector<uint8_t> z(w.n);
  read_exact(f, z.data(), w.n, bin);
  fclose(f);
  w.rz.assign(z.begin(), z.end());
  return w;
}
void write_all(const std::string &path, const BinHeader &h, std::initializer_list<std::pair<const void *, size_t>> parts) {
  std::string tmp = path + ".tmp";
  FILE *f = fopen(tmp.c_str(), "wb");
  if (!f) throw IoError{VPIN_ERR_IO, "cannot create " + tmp + ": " + strerror(errno)};
  bool ok = fwrite(&h, 1, sizeof(h), f) == sizeof(h);
  for (auto &p : parts) ok = ok && (p.second == 0 || fwrite(p.first, 1, p.second, f) == p.second);
  ok = (fclose(f) == 0) && ok;
  if (!ok || rename(tmp.c_str(), path.c_str()) != 0) { remove(tmp.c_str()); throw IoError{VPIN_ERR_IO, "cannot write " + path}; }
}

template <class F>
vpin_status guarded(F &&f) {
  try {
    f();
    return VPIN_OK;
  } catch (const IoError &e) {
    g_io_error = e.msg;
    return e.code;
  } catch (const std::bad_alloc &) {
    g_io_error = "host allocation failed";
    return VPIN_ERR_OOM;
  }
}

}  // namespace

extern "C" {

const char *vpin_witness_last_error(void) { return g_io_error.c_str(); }

vpin_status vpin_load_point_mult(const char *root, const char *tag, uint64_t cap, uint64_t *count_out, uint64_t *weights_lo_hi, uint8_t *px32,
                                 uint8_t *py32) {
  return guarded([&] {
    if (!root || !tag || !count_out) throw IoError{VPIN_ERR_BAD_ARGUMENT, "null argument"};
    MultWitness w = load_mult(root, tag);
    *count_out = w.n;
    if (!weights_lo_hi && !px32 && !py32) return;  // size query
    if (!weights_lo_hi || !px32 || !py32) throw IoError{VPIN_ERR_BAD_ARGUMENT, "null output buffer"};
    if (cap < w.n) throw IoError{VPIN_ERR_BUFFER_TOO_SMALL, "capacity " + std::to_string(cap) + " < " + std::to_string(w.n) + " multiplications"};
    memcpy(weights_lo_hi, w.weights.data(), 16 * w.n);
    memcpy(px32, w.px.data(), 32 * w.n);
    memcpy(py32, w.py.data(), 32 * w.n);
  });
}

vpin_status vpin_load_point_add(const char *root, const char *tag, uint64_t cap, uint64_t *count_out, uint8_t *px32, uint8_t *py32, uint8_t *rx32,
                                uint8_t *ry32, int64_t *rz) {
  return guarded([&] {
    if (!root || !tag || !count_out) throw IoError{VPIN_ERR_BAD_ARGUMENT, "null argument"};
    AddWitness w = load_add(root, tag);
    *count_out = w.n;
    if (!px32 && !py32 && !rx32 && !ry32 && !rz) return;  // size query
    if (!px32 || !py32 || !rx32 || !ry32 || !rz) throw IoError{VPIN_ERR_BAD_ARGUMENT, "null output buffer"};
    if (cap < w.n) throw IoError{VPIN_ERR_BUFFER_TOO_SMALL, "capacity " + std::to_string(cap) + " < " + std::to_string(w.n) + " additions"};
    memcpy(px32, w.px.data(), 32 * w.n);
    memcpy(py32, w.py.data(), 32 * w.n);
    memcpy(rx32, w.rx.data(), 32 * w.n);
    memcpy(ry32, w.ry.data(), 32 * w.n);
    memcpy(rz, w.rz.data(), 8 * w.n);
  });
}

vpin_status vpin_witness_json_to_bin(const char *root, const char *tag, uint32_t *written_out) {
  return guarded([&] {
    if (!root || !tag) throw IoError{VPIN_ERR_BAD_ARGUMENT, "null argument"};
    uint32_t written = 0;
    std::string dm = dir_of(root, tag, "pointMult"), da = dir_of(root, tag, "pointAdd");
    if (file_exists(dm + "weight.json")) {
      MultWitness w = load_mult_json(dm);
      BinHeader h;
      memcpy(h.magic, kMagic, 8);
      h.kind = 1; h.reserved = 0; h.count = w.n;
      write_all(dm + "witness.bin", h, {{w.weights.data(), 16 * w.n}, {w.px.data(), 32 * w.n}, {w.py.data(), 32 * w.n}});
      written |= 1;
    }
    if (file_exists(da + "point_add_px_byte.json")) {
      AddWitness w = load_add_json(da);
      std::vector<uint8_t> z(w.n);
      for (uint64_t i = 0; i < w.n; i++) {
        if (w.rz[i] != 0 && w.rz[i] != 1) throw IoError{VPIN_ERR_SIZE_MISMATCH, da + ": rz entries must be 0 or 1"};
        z[i] = (uint8_t)w.rz[i];
      }
      BinHeader h;
      memcpy(h.magic, kMagic, 8);
      h.kind = 2; h.reserved = 0; h.count = w.n;
      write_all(da + "witness.bin", h, {{w.px.data(), 32 * w.n}, {w.py.data(), 32 * w.n}, {w.rx.data(), 32 * w.n}, {w.ry.data(), 32 * w.n}, {z.data(), w.n}});
      written |= 2;
    }
    if (!written) throw IoError{VPIN_ERR_IO, std::string("no witness files under ") + root + "/rust_files/" + tag};
    if (written_out) *written_out = written;
  });
}

}  // extern "C"
