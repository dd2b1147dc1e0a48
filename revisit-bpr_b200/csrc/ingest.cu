// Host-side JSONL ingest (no device code): the reference's on-disk interaction format
// (bin/datasets/format-repro.sh:56-81) straight into flat integer arrays.
//   <split>.jsonl                  {"user": u, "item": i}            -> rbpr_ingest_pairs
//   <split>-user-seen-items.jsonl  {"user": u, "seen_items": [..]}   -> rbpr_ingest_lists
//   <split>-grouped.jsonl          {"user": u, "item": [..]}         -> rbpr_ingest_lists
// Replaces the per-line json.loads + scipy dok_matrix loop of
// SparseSamplingInMemoryWithCollator._sparse_matrix (experiments/bpr/dataset.py:183-190) and the
// dict-of-lists readers (dataset.py:16-24): the file is mmap'ed and scanned once; values of other
// keys (numbers, strings, arrays, nested objects) are skipped.
#include <errno.h>
#include <fcntl.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <string>
#include <vector>

#include "../../include/rbpr.h"

namespace {

struct Mapped {
  const char* p = nullptr;
  size_t n = 0;
  int fd = -1;
  ~Mapped() {
    if (p && n) munmap(const_cast<char*>(p), n);
    if (fd >= 0) close(fd);
  }
};

thread_local std::string g_ingest_err;

int fail(const std::string& msg) {
  g_ingest_err = msg;
  return RBPR_ERR_DATA;
}

int map_file(const char* path, Mapped& m) {
  m.fd = open(path, O_RDONLY);
  if (m.fd < 0) return fail(std::string("cannot open ") + path + ": " + strerror(errno));
  struct stat st;
  if (fstat(m.fd, &st) != 0) return fail(std::string("cannot stat ") + path);
  m.n = (size_t)st.st_size;
  if (m.n == 0) return 0;
  void* p = mmap(nullptr, m.n, PROT_READ, MAP_PRIVATE, m.fd, 0);
  if (p == MAP_FAILED) {
    m.n = 0;
    return fail(std::string("cannot mmap ") + path);
  }
  madvise(p, m.n, MADV_SEQUENTIAL);
  m.p = (const char*)p;
  return 0;
}

inline const char* skip_ws(const char* c, const char* e) {
  while (c < e && (*c == ' ' || *c == '\t' || *c == '\r')) ++c;
  return c;
}

// parse a JSON string starting at the opening quote; returns pointer past the closing quote
inline const char* skip_string(const char* c, const char* e) {
  ++c;
  while (c < e && *c != '"') c += (*c == '\\' && c + 1 < e) ? 2 : 1;
  return c < e ? c + 1 : nullptr;
}

inline const char* parse_int(const char* c, const char* e, int64_t* out) {
  bool neg = false;
  if (c < e && *c == '-') {
    neg = true;
    ++c;
  }
  if (c >= e || *c < '0' || *c > '9') return nullptr;
  int64_t v = 0;
  while (c < e && *c >= '0' && *c <= '9') v = v * 10 + (*c++ - '0');
  if (c < e && (*c == '.' || *c == 'e' || *c == 'E')) return nullptr;  // ids are integers
  *out = neg ? -v : v;
  return c;
}

// skip any JSON value (number, string, literal, array, object)
const char* skip_value(const char* c, const char* e) {
  c = skip_ws(c, e);
  if (c >= e) return nullptr;
  if (*c == '"') return skip_string(c, e);
  if (*c == '[' || *c == '{') {
    int depth = 0;
    while (c < e) {
      if (*c == '"') {
        c = skip_string(c, e);
        if (!c) return nullptr;
        continue;
      }
      if (*c == '[' || *c == '{') ++depth;
      if (*c == ']' || *c == '}') {
        --depth;
        if (depth == 0) return c + 1;
      }
      if (*c == '\n') return nullptr;
      ++c;
    }
    return nullptr;
  }
  while (c < e && *c != ',' && *c != '}' && *c != '\n') ++c;
  return c;
}

// One line = one object.  For each key: if it equals `ka`, its value must be an int -> a;
// if it equals `kb`: int -> b (pairs mode) or array of ints appended to `list` (lists mode).
int parse_line(const char* c, const char* e, const char* ka, size_t la, const char* kb, size_t lb,
               bool b_is_list, int64_t* a, int64_t* b, std::vector<int64_t>* list, bool* got_a,
               bool* got_b) {
  c = skip_ws(c, e);
  if (c >= e || *c != '{') return -1;
  ++c;
  while (true) {
    c = skip_ws(c, e);
    if (c < e && *c == '}') return 0;
    if (c >= e || *c != '"') return -1;
    const char* k0 = c + 1;
    const char* kend = skip_string(c, e);
    if (!kend) return -1;
    const size_t klen = (size_t)(kend - 1 - k0);
    c = skip_ws(kend, e);
    if (c >= e || *c != ':') return -1;
    c = skip_ws(c + 1, e);
    const bool is_a = klen == la && memcmp(k0, ka, la) == 0;
    const bool is_b = klen == lb && memcmp(k0, kb, lb) == 0;
    if (is_a) {
      c = parse_int(c, e, a);
      if (!c) return -1;
      *got_a = true;
    } else if (is_b && !b_is_list) {
      c = parse_int(c, e, b);
      if (!c) return -1;
      *got_b = true;
    } else if (is_b) {
      if (c >= e || *c != '[') return -1;
      c = skip_ws(c + 1, e);
      if (c < e && *c == ']') {
        ++c;
      } else {
        while (true) {
          int64_t v;
          c = parse_int(skip_ws(c, e), e, &v);
          if (!c) return -1;
          list->push_back(v);
          c = skip_ws(c, e);
          if (c < e && *c == ',') {
            ++c;
            continue;
          }
          if (c < e && *c == ']') {
            ++c;
            break;
          }
          return -1;
        }
      }
      *got_b = true;
    } else {
      c = skip_value(c, e);
      if (!c) return -1;
    }
    c = skip_ws(c, e);
    if (c < e && *c == ',') {
      ++c;
      continue;
    }
    if (c < e && *c == '}') return 0;
    return -1;
  }
}

template <typename T>
T* to_malloc(const std::vector<T>& v) {
  T* p = (T*)malloc((v.size() ? v.size() : 1) * sizeof(T));
  if (p && !v.empty()) memcpy(p, v.data(), v.size() * sizeof(T));
  return p;
}

}  // namespace

extern "C" {

const char* rbpr_ingest_last_error(void) { return g_ingest_err.c_str(); }

void rbpr_ingest_free(void* p) { free(p); }

int rbpr_ingest_pairs(const char* path, const char* key_a, const char* key_b, int64_t** a_out,
                      int64_t** b_out, int64_t* n_out) {
  if (!path || !key_a || !key_b || !a_out || !b_out || !n_out) return fail("ingest_pairs: null argument");
  Mapped m;
  int rc = map_file(path, m);
  if (rc) return rc;
  std::vector<int64_t> va, vb;
  va.reserve(m.n / 24 + 16);
  vb.reserve(m.n / 24 + 16);
  const size_t la = strlen(key_a), lb = strlen(key_b);
  const char* c = m.p;
  const char* end = m.p + m.n;
  int64_t line = 0;
  while (c < end) {
    const char* nl = (const char*)memchr(c, '\n', (size_t)(end - c));
    const char* le = nl ? nl : end;
    ++line;
    if (skip_ws(c, le) < le) {  // non-blank line
      int64_t a = 0, b = 0;
      bool ga = false, gb = false;
      if (parse_line(c, le, key_a, la, key_b, lb, false, &a, &b, nullptr, &ga, &gb) != 0 || !ga || !gb)
        return fail(std::string(path) + ": line " + std::to_string(line) + " is not an object with integer \"" +
                    key_a + "\" and \"" + key_b + "\"");
      va.push_back(a);
      vb.push_back(b);
    }
    c = le + 1;
  }
  *a_out = to_malloc(va);
  *b_out = to_malloc(vb);
  *n_out = (int64_t)va.size();
  if (!*a_out || !*b_out) return fail("ingest_pairs: out of memory");
  return 0;
}

int rbpr_ingest_lists(const char* path, const char* key_a, const char* key_list, int64_t** a_out,
                      int64_t** offsets_out, int64_t** values_out, int64_t* n_rows_out) {
  if (!path || !key_a || !key_list || !a_out || !offsets_out || !values_out || !n_rows_out)
    return fail("ingest_lists: null argument");
  Mapped m;
  int rc = map_file(path, m);
  if (rc) return rc;
  std::vector<int64_t> va, off(1, 0), vals;
  vals.reserve(m.n / 6 + 16);
  const size_t la = strlen(key_a), lb = strlen(key_list);
  const char* c = m.p;
  const char* end = m.p + m.n;
  int64_t line = 0;
  while (c < end) {
    const char* nl = (const char*)memchr(c, '\n', (size_t)(end - c));
    const char* le = nl ? nl : end;
    ++line;
    if (skip_ws(c, le) < le) {
      int64_t a = 0, b = 0;
      bool ga = false, gb = false;
      if (parse_line(c, le, key_a, la, key_list, lb, true, &a, &b, &vals, &ga, &gb) != 0 || !ga || !gb)
        return fail(std::string(path) + ": line " + std::to_string(line) + " is not an object with integer \"" +
                    key_a + "\" and integer array \"" + key_list + "\"");
      va.push_back(a);
      off.push_back((int64_t)vals.size());
    }
    c = le + 1;
  }
  *a_out = to_malloc(va);
  *offsets_out = to_malloc(off);
  *values_out = to_malloc(vals);
  *n_rows_out = (int64_t)va.size();
  if (!*a_out || !*offsets_out || !*values_out) return fail("ingest_lists: out of memory");
  return 0;
}

}  // extern "C"
