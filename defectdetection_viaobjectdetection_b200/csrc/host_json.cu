// Host-side loader of the reference's on-disk volume format:
//   {"<beam key>": {"<scan>_<label>[_<start>-<end>]": [S numbers] | {"signal": [S numbers], ...}, ...}, ...}
// (signals/improved_multisignal/README.md:67-89).  It restates what JsonSignalDataset._load_all_json_files does
// before windowing (json_dataset.py:36-81): beams in file order, scans STABLY sorted by int(key.split('_')[0]),
// label 0 iff key.split('_')[1] == "Health", defect range float(key.split('_')[2].split('-')[0 / 1]) with
// [0.0, 0.0] on any failure.  Pure host code (no CUDA calls): the parsed beams are handed to the device windowing
// (paut_window_gather) by the caller.
#include <algorithm>
#include <atomic>
#include <cctype>
#include <cerrno>
#include <charconv>
#include <chrono>
#include <cmath>
#include <mutex>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "../../include/paut.h"

namespace {

thread_local std::string g_json_error;

struct Scan {
  std::string key;
  int64_t order = 0;           // int(key.split('_')[0])
  int32_t label = 0;
  float d0 = 0.f, d1 = 0.f;
  size_t off = 0, len = 0;     // samples in Beam::data
  bool bad = false;            // an object without a 'signal' list: np.array(dict) raises, the reference skips the scan
};
struct Beam {
  std::string key;
  std::vector<Scan> scans;     // sorted
  std::vector<float> data;
  int64_t S = 0;               // common signal length, -1 if the scans differ
  int status = 0;              // PAUT_JSON_BEAM_* bits: where the reference would raise out of its per-file try block
  std::string status_msg;
};

struct Parser {
  const char* p;
  const char* end;
  void fail(const std::string& m) const { throw std::runtime_error(m + " at byte " + std::to_string(p_off())); }
  size_t p_off() const { return (size_t)(p - begin); }
  const char* begin;
  void ws() { while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p; }
  bool eat(char c) { ws(); if (p < end && *p == c) { ++p; return true; } return false; }
  void expect(char c) { if (!eat(c)) fail(std::string("expected '") + c + "'"); }
  std::string str() {
    ws();
    if (p >= end || *p != '"') fail("expected a string");
    ++p;
    std::string out;
    while (p < end && *p != '"') {
      if (*p == '\\') {
        if (++p >= end) break;
        switch (*p) {
          case 'n': out += '\n'; break; case 't': out += '\t'; break; case 'r': out += '\r'; break;
          case 'b': out += '\b'; break; case 'f': out += '\f'; break;
          case 'u': {
            if (end - p < 5) fail("bad \\u escape");
            unsigned cp = (unsigned)std::strtoul(std::string(p + 1, p + 5).c_str(), nullptr, 16);
            p += 4;
            if (cp < 0x80) out += (char)cp;
            else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
            else { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
            break;
          }
          default: out += *p;
        }
        ++p;
      } else {
        out += *p++;
      }
    }
    if (p >= end) fail("unterminated string");
    ++p;
    return out;
  }
  // one JSON number (or the NaN / Infinity / -Infinity literals Python's json accepts) -> the float32 the reference
  // gets from np.array(list_of_python_floats, dtype=np.float32): correctly rounded double, then one rounding to fp32
  float number() {
    ws();
    if (p >= end) fail("expected a number");
    if (*p == 'N' && end - p >= 3 && !std::strncmp(p, "NaN", 3)) { p += 3; return NAN; }
    if (*p == 'I' && end - p >= 8 && !std::strncmp(p, "Infinity", 8)) { p += 8; return INFINITY; }
    if (*p == '-' && end - p >= 9 && !std::strncmp(p, "-Infinity", 9)) { p += 9; return -INFINITY; }
    if (!(*p == '-' || (*p >= '0' && *p <= '9'))) fail("expected a number");
    {
      const char* d = *p == '-' ? p + 1 : p;
      if (d >= end || *d < '0' || *d > '9' || (*d == '0' && d + 1 < end && (d[1] == 'x' || d[1] == 'X'))) fail("bad number");
    }
    double v = 0.0;                                     // std::from_chars: correctly rounded, locale-free, ~10x strtod
    const std::from_chars_result r = std::from_chars(p, end, v, std::chars_format::general);
    if (r.ec == std::errc::result_out_of_range) {       // Python: 1e999 -> inf, 1e-999 -> 0.0
      char* e = nullptr;
      v = std::strtod(std::string(p, r.ptr).c_str(), &e);
    } else if (r.ec != std::errc() || r.ptr == p) {
      fail("bad number");
    }
    p = r.ptr;
    return (float)v;
  }
  void numbers(std::vector<float>& out) {
    expect('[');
    if (eat(']')) return;
    do { out.push_back(number()); } while (eat(','));
    expect(']');
  }
  void skip_value() {
    ws();
    if (p >= end) fail("unexpected end");
    if (*p == '"') { str(); return; }
    if (*p == '{') { ++p; if (eat('}')) return; do { str(); expect(':'); skip_value(); } while (eat(',')); expect('}'); return; }
    if (*p == '[') { ++p; if (eat(']')) return; do { skip_value(); } while (eat(',')); expect(']'); return; }
    if (!std::strncmp(p, "true", 4)) { p += 4; return; }
    if (!std::strncmp(p, "false", 5)) { p += 5; return; }
    if (!std::strncmp(p, "null", 4)) { p += 4; return; }
    number();
  }
};

std::vector<std::string> split(const std::string& s, char c) {
  std::vector<std::string> out;
  size_t a = 0;
  for (;;) {
    const size_t b = s.find(c, a);
    out.push_back(s.substr(a, b == std::string::npos ? b : b - a));
    if (b == std::string::npos) break;
    a = b + 1;
  }
  return out;
}

std::string strip(const std::string& s) {
  size_t a = 0, b = s.size();
  while (a < b && std::isspace((unsigned char)s[a])) ++a;
  while (b > a && std::isspace((unsigned char)s[b - 1])) --b;
  return s.substr(a, b - a);
}

// Python int(str): optional whitespace, sign, decimal digits (underscores cannot occur: the key was split on '_')
bool py_int(const std::string& raw, int64_t* out) {
  const std::string s = strip(raw);
  size_t i = 0;
  if (i < s.size() && (s[i] == '+' || s[i] == '-')) ++i;
  if (i >= s.size()) return false;
  for (size_t j = i; j < s.size(); ++j)
    if (s[j] < '0' || s[j] > '9') return false;
  errno = 0;
  const long long v = std::strtoll(s.c_str(), nullptr, 10);
  if (errno == ERANGE) return false;
  *out = v;
  return true;
}

// Python float(str): decimal literal with optional exponent, or inf / infinity / nan (any case); no hex floats
bool py_float(const std::string& raw, double* out) {
  const std::string s = strip(raw);
  size_t i = 0;
  if (i < s.size() && (s[i] == '+' || s[i] == '-')) ++i;
  std::string low;
  for (size_t j = i; j < s.size(); ++j) low += (char)std::tolower((unsigned char)s[j]);
  if (low == "inf" || low == "infinity" || low == "nan") {
    *out = low == "nan" ? NAN : (s[0] == '-' ? -INFINITY : INFINITY);
    return true;
  }
  size_t j = i, digits = 0;
  while (j < s.size() && std::isdigit((unsigned char)s[j])) { ++j; ++digits; }
  if (j < s.size() && s[j] == '.') { ++j; while (j < s.size() && std::isdigit((unsigned char)s[j])) { ++j; ++digits; } }
  if (digits == 0) return false;
  if (j < s.size() && (s[j] == 'e' || s[j] == 'E')) {
    ++j;
    if (j < s.size() && (s[j] == '+' || s[j] == '-')) ++j;
    size_t ed = 0;
    while (j < s.size() && std::isdigit((unsigned char)s[j])) { ++j; ++ed; }
    if (ed == 0) return false;
  }
  if (j != s.size()) return false;
  *out = std::strtod(s.c_str(), nullptr);
  return true;
}

// one beam's object "{ "<scan key>": [..] | {"signal": [..]}, ... }" -> sorted scans (json_dataset.py:44-81)
void parse_beam(Parser& ps, Beam& beam) {
  std::unordered_map<std::string, size_t> index;     // scan key -> position (duplicates)
  ps.expect('{');
  if (!ps.eat('}')) {
    do {
      Scan sc;
      sc.key = ps.str();
      ps.expect(':');
      sc.off = beam.data.size();
      ps.ws();
      if (ps.p < ps.end && *ps.p == '{') {            // {"signal": [...], ...}   json_dataset.py:113-114
        ++ps.p;
        bool found = false;
        if (!ps.eat('}')) {
          do {
            const std::string k = ps.str();
            ps.expect(':');
            if (k == "signal") { beam.data.resize(sc.off); ps.numbers(beam.data); found = true; }
            else ps.skip_value();
          } while (ps.eat(','));
          ps.expect('}');
        }
        if (!found) { beam.data.resize(sc.off); sc.bad = true; }   // caught per scan in the reference (json_dataset.py:108-126)
      } else {
        ps.numbers(beam.data);
      }
      sc.len = beam.data.size() - sc.off;
      // key parsing, json_dataset.py:48,69-79
      const std::vector<std::string> parts = split(sc.key, '_');
      // The reference raises out of its per-file try block here -- at the sort for a key without an integer prefix (every
      // beam, json_dataset.py:48), at the label for a key without a second field (only beams with at least seq_length
      // scans, :51-52, :69) -- and keeps the sequences of the EARLIER beams of the file.  Recorded per beam; the caller
      // stops at the first beam whose status applies.
      if (!py_int(parts[0], &sc.order)) {
        if (!(beam.status & PAUT_JSON_BEAM_BAD_ORDER_KEY)) beam.status_msg = "scan key '" + sc.key + "': int(key.split('_')[0]) fails";
        beam.status |= PAUT_JSON_BEAM_BAD_ORDER_KEY;
        sc.order = 0;
      }
      if (parts.size() < 2) {
        if (!beam.status) beam.status_msg = "scan key '" + sc.key + "' has no label field";
        beam.status |= PAUT_JSON_BEAM_NO_LABEL;
        sc.label = 0;
      } else if (parts[1] == "Health") {
        sc.label = 0;
      } else {
        sc.label = 1;
        double a = 0.0, b = 0.0;
        bool ok = parts.size() >= 3;
        if (ok) {
          const std::vector<std::string> r = split(parts[2], '-');
          ok = r.size() >= 2 && py_float(r[0], &a) && py_float(r[1], &b);
        }
        sc.d0 = ok ? (float)a : 0.f;
        sc.d1 = ok ? (float)b : 0.f;
      }
      // a repeated key: Python's json keeps the LAST value at the position of the first occurrence
      auto seen = index.find(sc.key);
      if (seen != index.end()) {
        Scan& old = beam.scans[seen->second];
        old.off = sc.off; old.len = sc.len; old.bad = sc.bad;
      } else {
        index.emplace(sc.key, beam.scans.size());
        beam.scans.push_back(std::move(sc));
      }
    } while (ps.eat(','));
    ps.expect('}');
  }
  ps.ws();
  if (ps.p != ps.end) ps.fail("trailing data in beam object");
  if (!(beam.status & PAUT_JSON_BEAM_BAD_ORDER_KEY))
    std::stable_sort(beam.scans.begin(), beam.scans.end(), [](const Scan& x, const Scan& y) { return x.order < y.order; });
  beam.S = beam.scans.empty() ? 0 : (int64_t)beam.scans[0].len;
  for (const Scan& s : beam.scans)
    if ((int64_t)s.len != beam.S || s.bad) beam.S = -1;          // a skipped scan makes the beam take the per-scan path
}

}  // namespace

struct paut_json_volume {
  std::vector<Beam> beams;
};

extern "C" {

const char* paut_json_last_error(void) { return g_json_error.c_str(); }

int paut_json_load_host(const char* path, paut_json_volume** out) {
  if (!path || !out) return PAUT_ERR_INVALID;
  *out = nullptr;
  // The file is mapped, not copied: the page faults then happen inside the (parallel) parse.  The span scan needs a
  // NUL after the last byte; a file mapping is zero-filled up to the end of its last page, so only a size that is an
  // exact multiple of the page size takes the read-into-a-string path.
  struct Text {
    const char* ptr = nullptr;
    size_t len = 0;
    void* map = nullptr;
    size_t map_len = 0;
    std::string owned;
    ~Text() { if (map) munmap(map, map_len); }
    const char* data() const { return ptr; }
    size_t size() const { return len; }
  } text;
  {
    const int fd = open(path, O_RDONLY);
    if (fd < 0) { g_json_error = std::string("cannot open ") + path; return PAUT_ERR_INVALID; }
    struct stat st;
    if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode)) { close(fd); g_json_error = std::string("not a regular file: ") + path; return PAUT_ERR_INVALID; }
    const size_t n = (size_t)st.st_size, page = (size_t)sysconf(_SC_PAGESIZE);
    const char* mode = std::getenv("PAUT_JSON_IO");                    // "read" | "mmap" | "populate" (experiment switch)
    const bool use_map = !(mode && !std::strcmp(mode, "read"));
    const int map_flags = MAP_PRIVATE | ((mode && !std::strcmp(mode, "populate")) ? MAP_POPULATE : 0);
    if (use_map && n > 0 && n % page != 0) {
      void* m = mmap(nullptr, n, PROT_READ, map_flags, fd, 0);
      if (m != MAP_FAILED) {
        madvise(m, n, MADV_WILLNEED);
        text.map = m; text.map_len = n; text.ptr = static_cast<const char*>(m); text.len = n;
      }
    }
    if (!text.map) {
      text.owned.resize(n);
      size_t got = 0;
      while (got < n) {
        const ssize_t r = read(fd, &text.owned[got], n - got);
        if (r <= 0) break;
        got += (size_t)r;
      }
      if (got != n) { close(fd); g_json_error = std::string("short read of ") + path; return PAUT_ERR_INVALID; }
      text.ptr = text.owned.c_str(); text.len = n;
    }
    close(fd);
  }
  const bool timing = std::getenv("PAUT_JSON_TIMING") != nullptr;
  const auto t_read = std::chrono::steady_clock::now();
  paut_json_volume* vol = new paut_json_volume();
  try {
    // pass 1 (one thread, byte scan): the spans of the beams' objects.  Beams are independent, so pass 2 parses them
    // on several threads; the result is the same as a sequential parse (beams stay in file order).
    struct Span { std::string key; const char* b; const char* e; };
    std::vector<Span> spans;
    Parser ps{text.data(), text.data() + text.size(), text.data()};
    ps.expect('{');
    if (!ps.eat('}')) {
      do {
        Span sp;
        sp.key = ps.str();
        ps.expect(':');
        ps.ws();
        if (ps.p >= ps.end || *ps.p != '{') ps.fail("expected '{'");
        sp.b = ps.p;
        // only quotes and brackets matter here; strcspn (vectorised in glibc) jumps over the number lists, which
        // are ~99 % of the bytes (the text is NUL-terminated: it lives in a std::string)
        int depth = 0;
        const char* q = ps.p;
        for (;;) {
          q += std::strcspn(q, "\"{}[]");
          if (q >= ps.end) break;
          const char ch = *q;
          if (ch == '\0') { ++q; continue; }                        // stray NUL inside the file: not a delimiter
          if (ch == '"') {                                           // skip the string, honouring escapes
            for (++q; q < ps.end && *q != '"'; ++q)
              if (*q == '\\') ++q;
            if (q >= ps.end) break;
            ++q;
            continue;
          }
          if (ch == '{' || ch == '[') ++depth;
          else if (--depth == 0) break;
          ++q;
        }
        if (q >= ps.end) ps.fail("unterminated beam object");
        sp.e = q + 1;
        ps.p = sp.e;
        spans.push_back(std::move(sp));
      } while (ps.eat(','));
      ps.expect('}');
    }
    ps.ws();
    if (ps.p != ps.end) ps.fail("trailing data");

    const auto t_pass1 = std::chrono::steady_clock::now();
    vol->beams.resize(spans.size());
    std::atomic<size_t> next{0};
    std::mutex err_mu;
    std::string first_error;
    size_t first_error_beam = spans.size();
    auto worker = [&]() {
      for (;;) {
        const size_t i = next.fetch_add(1);
        if (i >= spans.size()) return;
        try {
          Parser bp{spans[i].b, spans[i].e, text.data()};
          vol->beams[i].key = spans[i].key;
          parse_beam(bp, vol->beams[i]);
        } catch (const std::exception& e) {
          std::lock_guard<std::mutex> lk(err_mu);
          if (i < first_error_beam) { first_error_beam = i; first_error = e.what(); }   // the error a sequential parse hits first
        }
      }
    };
    unsigned nthreads = std::thread::hardware_concurrency();
    if (const char* env = std::getenv("PAUT_JSON_THREADS")) nthreads = (unsigned)std::max(1, std::atoi(env));
    nthreads = (unsigned)std::min<size_t>({(size_t)std::max(1u, nthreads), spans.size(), (size_t)16});
    if (text.size() < (size_t(1) << 20)) nthreads = 1;                  // small files: not worth the thread start-up
    if (nthreads <= 1) {
      worker();
    } else {
      std::vector<std::thread> pool;
      for (unsigned t = 0; t < nthreads; ++t) pool.emplace_back(worker);
      for (auto& th : pool) th.join();
    }
    if (timing) {
      const auto t_end = std::chrono::steady_clock::now();
      auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
      std::fprintf(stderr, "[paut_json] %zu bytes, %zu beams, %u threads: span scan %.1f ms, parse %.1f ms\n", text.size(),
                   spans.size(), nthreads, ms(t_read, t_pass1), ms(t_pass1, t_end));
    }
    if (!first_error.empty()) throw std::runtime_error(first_error);
  } catch (const std::exception& e) {
    g_json_error = std::string(path) + ": " + e.what();
    delete vol;
    return PAUT_ERR_INVALID;
  }
  *out = vol;
  return PAUT_OK;
}

void paut_json_free(paut_json_volume* v) { delete v; }

int paut_json_num_beams(const paut_json_volume* v) { return v ? (int)v->beams.size() : PAUT_ERR_INVALID; }

int paut_json_beam_info(const paut_json_volume* v, int beam, const char** key, int64_t* n_scans, int64_t* signal_length) {
  if (!v || beam < 0 || beam >= (int)v->beams.size()) return PAUT_ERR_INVALID;
  const Beam& b = v->beams[beam];
  if (key) *key = b.key.c_str();
  if (n_scans) *n_scans = (int64_t)b.scans.size();
  if (signal_length) *signal_length = b.S;
  return PAUT_OK;
}

const char* paut_json_scan_key(const paut_json_volume* v, int beam, int64_t i) {
  if (!v || beam < 0 || beam >= (int)v->beams.size()) return nullptr;
  const Beam& b = v->beams[beam];
  return i >= 0 && i < (int64_t)b.scans.size() ? b.scans[i].key.c_str() : nullptr;
}

int paut_json_beam_status(const paut_json_volume* v, int beam, const char** message) {
  if (!v || beam < 0 || beam >= (int)v->beams.size()) return PAUT_ERR_INVALID;
  const Beam& b = v->beams[beam];
  if (message) *message = b.status_msg.c_str();
  return b.status;
}

int64_t paut_json_scan_copy_host(const paut_json_volume* v, int beam, int64_t i, float* out, int64_t cap) {
  if (!v || beam < 0 || beam >= (int)v->beams.size()) return PAUT_ERR_INVALID;
  const Beam& b = v->beams[beam];
  if (i < 0 || i >= (int64_t)b.scans.size()) return PAUT_ERR_INVALID;
  const Scan& s = b.scans[i];
  if (s.bad) return PAUT_JSON_SCAN_SKIPPED;
  if (out && cap > 0) std::memcpy(out, b.data.data() + s.off, (size_t)std::min<int64_t>(cap, (int64_t)s.len) * sizeof(float));
  return (int64_t)s.len;
}

int paut_json_beam_copy_host(const paut_json_volume* v, int beam, float* signals, int32_t* labels, float* defects,
                             int64_t* scan_order) {
  if (!v || beam < 0 || beam >= (int)v->beams.size()) return PAUT_ERR_INVALID;
  const Beam& b = v->beams[beam];
  if (signals && b.S < 0) { g_json_error = "beam '" + b.key + "': scans of different lengths"; return PAUT_ERR_UNSUPPORTED; }
  for (size_t i = 0; i < b.scans.size(); ++i) {
    const Scan& s = b.scans[i];
    if (signals && s.len) std::memcpy(signals + i * (size_t)b.S, b.data.data() + s.off, s.len * sizeof(float));
    if (labels) labels[i] = s.label;
    if (defects) { defects[2 * i] = s.d0; defects[2 * i + 1] = s.d1; }
    if (scan_order) scan_order[i] = s.order;
  }
  return PAUT_OK;
}

}  // extern "C"
