// Expression-matrix ingestion (scope row f4): the on-disk formats on the input side of the path, as host
// C++ behind the C ABI, producing the host CSC arrays plaidgpu_score consumes (the dgCMatrix slots
// @p / @i / @x of SURVEY.md §8 a1):
//   * R save() / saveRDS() files holding a dgCMatrix — the reference's own fixture format
//     (inst/extdata/pbmc3k-50cells.rda, written by dev/extdata.R:15): gzip + "RDX2/RDX3" XDR serialisation;
//   * Matrix Market coordinate files and 10x directories (matrix.mtx[.gz] + features/genes.tsv[.gz] +
//     barcodes.tsv[.gz]) — what Seurat::Read10X / Matrix::readMM feed into the reference's callers.
// Pure host code (text / XDR decoding is not GPU work); gzip through zlib.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include <algorithm>
#include <memory>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "../../include/plaidgpu.h"

struct plaidgpu_spmat {
  int32_t P = 0;
  int64_t N = 0;
  std::vector<int32_t> p, i;
  std::vector<double> x;
  std::vector<std::string> rownames, colnames;
};

namespace {

thread_local std::string g_io_err;

int io_fail(const std::string& msg) {
  g_io_err = msg;
  return PLAIDGPU_ERR_ARG;
}

bool read_file(const std::string& path, std::string& buf) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  if (fseek(f, 0, SEEK_END) == 0) {  // regular file: one allocation, one read
    const long sz = ftell(f);
    rewind(f);
    if (sz > 0) {
      buf.resize((size_t)sz);
      const size_t got = fread(&buf[0], 1, (size_t)sz, f);
      buf.resize(got);
    }
  }
  char tmp[1 << 16];
  size_t n;
  while ((n = fread(tmp, 1, sizeof(tmp), f)) > 0) buf.append(tmp, n);  // pipes, or a file that grew
  fclose(f);
  return true;
}

// gzip (RFC 1952) -> plain; input that is not gzip is passed through
bool gunzip(std::string& buf) {
  if (buf.size() < 2 || (unsigned char)buf[0] != 0x1f || (unsigned char)buf[1] != 0x8b) return true;
  z_stream zs;
  memset(&zs, 0, sizeof(zs));
  if (inflateInit2(&zs, 16 + MAX_WBITS) != Z_OK) return false;
  std::string out;
  out.reserve(buf.size() * 4);
  std::vector<unsigned char> tmp(1 << 20);
  zs.next_in = (Bytef*)buf.data();
  size_t left = buf.size();
  int rc = Z_OK;
  for (;;) {
    if (zs.avail_in == 0) {
      if (left == 0) break;  // input exhausted before the end of the stream: truncated
      const size_t take = std::min<size_t>(left, 1u << 30);
      zs.avail_in = (uInt)take;
      left -= take;
    }
    zs.next_out = tmp.data();
    zs.avail_out = (uInt)tmp.size();
    rc = inflate(&zs, Z_NO_FLUSH);
    if (rc != Z_OK && rc != Z_STREAM_END) break;
    out.append((const char*)tmp.data(), tmp.size() - zs.avail_out);
    if (rc == Z_STREAM_END) {
      if (zs.avail_in == 0 && left == 0) break;
      if (inflateReset(&zs) != Z_OK) {  // a further gzip member follows (bgzip-style files)
        rc = Z_DATA_ERROR;
        break;
      }
    }
  }
  if (rc != Z_STREAM_END) {
    inflateEnd(&zs);
    return false;
  }
  inflateEnd(&zs);
  buf.swap(out);
  return true;
}

bool load(const std::string& path, std::string& buf) {
  buf.clear();
  if (!read_file(path, buf)) {
    g_io_err = "cannot open " + path;
    return false;
  }
  if (!gunzip(buf)) {
    g_io_err = "corrupt gzip stream in " + path;
    return false;
  }
  return true;
}

// ---------------------------------------------------------------------------------------
// COO -> CSC: rows sorted within a column, duplicated (row, col) entries summed (what
// as(readMM(.), "CsparseMatrix") gives)
// ---------------------------------------------------------------------------------------
struct CooPart {
  std::vector<int32_t> ri;
  std::vector<int64_t> ci;
  std::vector<double> v;
};

void coo_to_csc(int32_t P, int64_t N, const std::vector<const CooPart*>& parts, plaidgpu_spmat* m) {
  size_t nz = 0;
  for (const CooPart* c : parts) nz += c->ri.size();
  m->P = P;
  m->N = N;
  std::vector<int64_t> cnt((size_t)N + 1, 0);
  for (const CooPart* c : parts)
    for (int64_t col : c->ci) {
      if (col < 0 || col >= N) abort();  // callers range-check every column index: never reached
      ++cnt[(size_t)col + 1];
    }
  for (int64_t j = 0; j < N; ++j) cnt[(size_t)j + 1] += cnt[(size_t)j];
  std::vector<int32_t> rows(nz);
  std::vector<double> vals(nz);
  {
    std::vector<int64_t> fill(cnt.begin(), cnt.end() - 1);
    for (const CooPart* c : parts) {  // parts in file order: the bucket fill is stable
      const size_t n = c->ri.size();
      for (size_t k = 0; k < n; ++k) {
        const int64_t d = fill[(size_t)c->ci[k]]++;
        rows[(size_t)d] = c->ri[k];
        vals[(size_t)d] = c->v[k];
      }
    }
  }
  m->p.assign((size_t)N + 1, 0);
  m->i.clear();
  m->x.clear();
  m->i.reserve(nz);
  m->x.reserve(nz);
  std::vector<std::pair<int32_t, double>> col;
  for (int64_t j = 0; j < N; ++j) {
    const int64_t a = cnt[(size_t)j], b = cnt[(size_t)j + 1];
    bool sorted = true;
    for (int64_t k = a + 1; k < b; ++k)
      if (rows[(size_t)k] <= rows[(size_t)k - 1]) {
        sorted = false;
        break;
      }
    if (sorted) {
      m->i.insert(m->i.end(), rows.begin() + a, rows.begin() + b);
      m->x.insert(m->x.end(), vals.begin() + a, vals.begin() + b);
    } else {
      col.clear();
      for (int64_t k = a; k < b; ++k) col.emplace_back(rows[(size_t)k], vals[(size_t)k]);
      std::stable_sort(col.begin(), col.end(), [](const std::pair<int32_t, double>& l, const std::pair<int32_t, double>& r) {
        return l.first < r.first;
      });
      for (size_t k = 0; k < col.size(); ++k) {
        if (k > 0 && col[k].first == col[k - 1].first) {
          m->x.back() += col[k].second;
        } else {
          m->i.push_back(col[k].first);
          m->x.push_back(col[k].second);
        }
      }
    }
    m->p[(size_t)j + 1] = (int32_t)m->i.size();
  }
}

// ---------------------------------------------------------------------------------------
// Matrix Market
// ---------------------------------------------------------------------------------------
inline const char* skip_ws(const char* s, const char* e) {
  while (s < e && (*s == ' ' || *s == '\t' || *s == '\r')) ++s;
  return s;
}

inline bool parse_i64(const char*& s, const char* e, int64_t* out) {
  s = skip_ws(s, e);
  if (s >= e) return false;
  bool neg = false;
  if (*s == '-' || *s == '+') neg = (*s++ == '-');
  if (s >= e || *s < '0' || *s > '9') return false;
  int64_t v = 0;
  while (s < e && *s >= '0' && *s <= '9') v = v * 10 + (*s++ - '0');
  *out = neg ? -v : v;
  return true;
}

// counts are small integers in nearly every file: fast path for digits only, strtod for the rest
inline bool parse_f64(const char*& s, const char* e, double* out) {
  s = skip_ws(s, e);
  if (s >= e) return false;
  const char* q = s;
  int64_t v = 0;
  int nd = 0;
  while (q < e && *q >= '0' && *q <= '9' && nd < 15) {
    v = v * 10 + (*q++ - '0');
    ++nd;
  }
  if (nd > 0 && (q == e || *q == '\n' || *q == '\r' || *q == ' ' || *q == '\t')) {
    *out = (double)v;
    s = q;
    return true;
  }
  char tmp[64];
  size_t n = 0;
  q = s;
  while (q < e && *q != '\n' && *q != '\r' && *q != ' ' && *q != '\t' && n + 1 < sizeof(tmp)) tmp[n++] = *q++;
  tmp[n] = 0;
  char* endp = nullptr;
  *out = strtod(tmp, &endp);
  if (endp == tmp) return false;
  s = q;
  return true;
}

int parse_mtx(const std::string& buf, plaidgpu_spmat* m) {
  const char* s = buf.data();
  const char* e = s + buf.size();
  auto line_end = [&](const char* q) {
    const void* nl = memchr(q, '\n', (size_t)(e - q));
    return nl ? (const char*)nl : e;
  };
  if (buf.compare(0, 14, "%%MatrixMarket") != 0) return io_fail("not a Matrix Market file (missing %%MatrixMarket banner)");
  const char* le = line_end(s);
  std::string banner(s, le);
  for (char& ch : banner) ch = (char)tolower((unsigned char)ch);
  if (banner.find("coordinate") == std::string::npos) return io_fail("Matrix Market: only the coordinate format is supported");
  const bool pattern = banner.find("pattern") != std::string::npos;
  const bool symmetric = banner.find("symmetric") != std::string::npos;
  if (banner.find("complex") != std::string::npos || banner.find("hermitian") != std::string::npos ||
      banner.find("skew") != std::string::npos)
    return io_fail("Matrix Market: complex / hermitian / skew-symmetric matrices are not supported");
  s = le < e ? le + 1 : e;
  while (s < e && (*s == '%' || *s == '\n' || *s == '\r')) {  // comments and blank lines
    le = line_end(s);
    s = le < e ? le + 1 : e;
  }
  int64_t P = 0, N = 0, nz = 0;
  if (!parse_i64(s, e, &P) || !parse_i64(s, e, &N) || !parse_i64(s, e, &nz)) return io_fail("Matrix Market: bad size line");
  if (P <= 0 || P > 0x7fffffff || N < 0 || nz < 0) return io_fail("Matrix Market: dimensions out of range");
  if (symmetric && P != N) return io_fail("Matrix Market: a symmetric matrix must be square");  // mirrored entries would leave the matrix
  le = line_end(s);
  s = le < e ? le + 1 : e;
  // the entry lines are independent: parse them on several host threads, one contiguous slice of the
  // text each (cut at line ends), and concatenate the slices in file order
  struct Slice : CooPart {
    const char *s, *e;
    int64_t lines = 0;
    std::string err;
  };
  const size_t body = (size_t)(e - s);
  unsigned nt = std::thread::hardware_concurrency();
  if (nt == 0) nt = 1;
  nt = (unsigned)std::min<size_t>(std::min<unsigned>(nt, 32), body / (4u << 20) + 1);
  std::vector<Slice> sl(nt);
  {
    const char* cur = s;
    for (unsigned t = 0; t < nt; ++t) {
      sl[t].s = cur;
      const char* stop = (t + 1 == nt) ? e : std::min(e, s + body / nt * (t + 1));
      if (stop < cur) stop = cur;
      if (stop < e) stop = line_end(stop) < e ? line_end(stop) + 1 : e;
      sl[t].e = stop;
      cur = stop;
    }
  }
  auto work = [&](Slice* w) {
    const char* q = w->s;
    const char* qe = w->e;
    const size_t guess = (size_t)(qe - q) / 8 + 16;
    w->ri.reserve(guess);
    w->ci.reserve(guess);
    w->v.reserve(guess);
    while (q < qe) {
      while (q < qe && (*q == '\n' || *q == '\r' || *q == ' ' || *q == '\t')) ++q;
      if (q >= qe) break;
      int64_t r = 0, c = 0;
      double val = 1.0;
      if (!parse_i64(q, qe, &r) || !parse_i64(q, qe, &c) || (!pattern && !parse_f64(q, qe, &val))) {
        w->err = "malformed entry";
        return;
      }
      if (r < 1 || r > P || c < 1 || c > N) {
        w->err = "index out of range";
        return;
      }
      ++w->lines;
      w->ri.push_back((int32_t)(r - 1));
      w->ci.push_back(c - 1);
      w->v.push_back(val);
      if (symmetric && r != c) {
        if (c > P || r > N) {  // cannot happen for a square matrix; kept as a guard on the mirrored index
          w->err = "index out of range";
          return;
        }
        w->ri.push_back((int32_t)(c - 1));
        w->ci.push_back(r - 1);
        w->v.push_back(val);
      }
      const void* nl = memchr(q, '\n', (size_t)(qe - q));
      q = nl ? (const char*)nl + 1 : qe;
    }
  };
  if (nt == 1) {
    work(&sl[0]);
  } else {
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t) th.emplace_back(work, &sl[t]);
    for (std::thread& t : th) t.join();
  }
  int64_t lines = 0;
  size_t total = 0;
  for (const Slice& w : sl) {
    if (!w.err.empty()) return io_fail("Matrix Market: " + w.err + " (entry " + std::to_string(lines + w.lines + 1) + ")");
    lines += w.lines;
    total += w.ri.size();
  }
  if (lines < nz) return io_fail("Matrix Market: truncated file, " + std::to_string(lines) + " of " + std::to_string(nz) + " entries");
  if (lines > nz) return io_fail("Matrix Market: more entries than the size line declares");
  if (total > 0x7fffffffull) return io_fail("Matrix Market: more than 2^31-1 entries (dgCMatrix limit)");
  std::vector<const CooPart*> parts;
  for (const Slice& w : sl) parts.push_back(&w);
  coo_to_csc((int32_t)P, N, parts, m);
  return PLAIDGPU_OK;
}

// one name per line: column `col` (0-based) of a tab-separated file, or the last column present
void read_names(const std::string& buf, int col, std::vector<std::string>& out) {
  size_t pos = 0;
  while (pos < buf.size()) {
    size_t eol = buf.find('\n', pos);
    if (eol == std::string::npos) eol = buf.size();
    size_t end = eol;
    if (end > pos && buf[end - 1] == '\r') --end;
    if (end > pos) {
      size_t f0 = pos;
      int k = 0;
      while (k < col) {
        const size_t t = buf.find('\t', f0);
        if (t == std::string::npos || t >= end) break;
        f0 = t + 1;
        ++k;
      }
      size_t f1 = buf.find('\t', f0);
      if (f1 == std::string::npos || f1 > end) f1 = end;
      out.emplace_back(buf, f0, f1 - f0);
    }
    pos = eol + 1;
  }
}

bool exists(const std::string& p) {
  FILE* f = fopen(p.c_str(), "rb");
  if (f) fclose(f);
  return f != nullptr;
}

// ---------------------------------------------------------------------------------------
// R serialisation (XDR), the subset a saved dgCMatrix uses
// ---------------------------------------------------------------------------------------
struct Sexp;
using SexpP = std::shared_ptr<Sexp>;
struct Sexp {
  int type = 254;  // NILVALUE
  std::string str;                 // CHARSXP / SYMSXP name
  std::vector<int32_t> ints;       // LGLSXP / INTSXP
  std::vector<double> reals;       // REALSXP
  std::vector<SexpP> items;        // STRSXP / VECSXP elements, pairlist values
  std::vector<std::string> tags;   // pairlist tags (parallel to items)
  SexpP attr;                      // pairlist of attributes (S4 slots live here)
  const Sexp* get(const char* tag) const {
    if (!attr) return nullptr;
    for (size_t k = 0; k < attr->items.size(); ++k)
      if (attr->tags[k] == tag) return attr->items[k].get();
    return nullptr;
  }
};

struct XdrReader {
  const unsigned char* b;
  size_t n, o = 0;
  std::vector<SexpP> refs;
  bool ok = true;
  std::string why;

  bool need(size_t k) {
    if (o + k > n) {
      ok = false;
      why = "unexpected end of the serialised stream";
      return false;
    }
    return true;
  }
  int32_t i32() {
    if (!need(4)) return 0;
    const uint32_t v = ((uint32_t)b[o] << 24) | ((uint32_t)b[o + 1] << 16) | ((uint32_t)b[o + 2] << 8) | b[o + 3];
    o += 4;
    return (int32_t)v;
  }
  double f64() {
    if (!need(8)) return 0.0;
    uint64_t v = 0;
    for (int k = 0; k < 8; ++k) v = (v << 8) | b[o + k];
    o += 8;
    double d;
    memcpy(&d, &v, 8);
    return d;
  }
  int64_t len() {
    int64_t v = i32();
    if (v == -1) {  // long vector: upper, lower
      const int64_t hi = i32();
      const int64_t lo = (uint32_t)i32();
      v = (hi << 32) + lo;
    }
    if (v < 0) {
      ok = false;
      why = "negative vector length";
      return 0;
    }
    return v;
  }
  SexpP fail(const std::string& msg) {
    if (ok) {
      ok = false;
      why = msg;
    }
    return std::make_shared<Sexp>();
  }

  SexpP item(int depth = 0) {
    auto nil = [] { return std::make_shared<Sexp>(); };
    if (!ok) return nil();
    if (depth > 200) return fail("serialised object nested too deeply");
    int32_t flags = i32();
    int type = flags & 0xff;
    bool has_attr = (flags & 0x200) != 0, has_tag = (flags & 0x400) != 0;
    switch (type) {
      case 254: case 253: case 242: case 241: case 251: case 252:  // NIL, global / empty / base env, missing, unbound
        return nil();
      case 255: {  // REFSXP
        int32_t idx = flags >> 8;
        if (idx == 0) idx = i32();
        if (idx < 1 || (size_t)idx > refs.size()) return fail("bad reference index");
        return refs[(size_t)idx - 1];
      }
      case 1: {  // SYMSXP
        SexpP nm = item(depth + 1);
        auto s = std::make_shared<Sexp>();
        s->type = 1;
        s->str = nm->str;
        refs.push_back(s);
        return s;
      }
      case 249: case 250: case 247: {  // namespace / package / persistent: a STRSXP without header
        i32();
        const int32_t cnt = i32();
        auto s = std::make_shared<Sexp>();
        s->type = 16;
        for (int32_t k = 0; k < cnt && ok; ++k) s->items.push_back(item(depth + 1));
        refs.push_back(s);
        return s;
      }
      case 2: case 6: case 240: case 239: {  // pairlists (walked iteratively along the cdr)
        auto s = std::make_shared<Sexp>();
        s->type = 2;
        while (ok) {
          if (type == 240 || type == 239) has_attr = true;
          if (has_attr) item(depth + 1);  // attributes of the cons cell itself: not needed
          std::string tag;
          if (has_tag) tag = item(depth + 1)->str;
          s->tags.push_back(tag);
          s->items.push_back(item(depth + 1));
          flags = i32();
          type = flags & 0xff;
          has_attr = (flags & 0x200) != 0;
          has_tag = (flags & 0x400) != 0;
          if (type == 254) break;
          if (type != 2 && type != 6 && type != 240 && type != 239) return fail("unexpected pairlist tail");
        }
        return s;
      }
      case 9: {  // CHARSXP
        const int32_t cnt = i32();
        auto s = std::make_shared<Sexp>();
        s->type = 9;
        if (cnt >= 0) {
          if (!need((size_t)cnt)) return nil();
          s->str.assign((const char*)b + o, (size_t)cnt);
          o += (size_t)cnt;
        } else {
          s->str = "NA";
        }
        return s;
      }
      case 238: {  // ALTREP: info pairlist (class symbol, package, type), state, attributes
        SexpP info = item(depth + 1), state = item(depth + 1), attr = item(depth + 1);
        const std::string cls = info->items.empty() ? "" : info->items[0]->str;
        auto s = std::make_shared<Sexp>();
        if (cls == "compact_intseq" || cls == "compact_realseq") {
          const bool real_state = !state->reals.empty();
          if ((real_state ? state->reals.size() : state->ints.size()) < 3) return fail("bad compact sequence");
          const double cnt = real_state ? state->reals[0] : state->ints[0];
          const double start = real_state ? state->reals[1] : state->ints[1];
          const double inc = real_state ? state->reals[2] : state->ints[2];
          if (!(cnt >= 0.0) || cnt > 2147483647.0) return fail("compact sequence length out of range");  // a dgCMatrix slot never exceeds 2^31-1
          if (cls == "compact_intseq") {
            s->type = 13;
            for (int64_t k = 0; k < (int64_t)cnt; ++k) s->ints.push_back((int32_t)(start + inc * (double)k));
          } else {
            s->type = 14;
            for (int64_t k = 0; k < (int64_t)cnt; ++k) s->reals.push_back(start + inc * (double)k);
          }
        } else if (cls.compare(0, 5, "wrap_") == 0) {  // state = list(x, meta): x is the plain vector
          if (state->items.empty()) return fail("bad ALTREP wrapper");
          *s = *state->items[0];
        } else {
          return fail("unsupported ALTREP class " + cls);
        }
        if (attr->type == 2) s->attr = attr;
        return s;
      }
      case 25: {  // S4SXP: only attributes
        auto s = std::make_shared<Sexp>();
        s->type = 25;
        if (has_attr) s->attr = item(depth + 1);
        return s;
      }
      default:
        break;
    }
    auto s = std::make_shared<Sexp>();
    s->type = type;
    if (type == 10 || type == 13) {
      const int64_t cnt = len();
      if (cnt < 0 || (uint64_t)cnt > (uint64_t)(n - o) / 4) return fail("unexpected end of the stream (a vector is longer than what is left)");
      if (!need((size_t)cnt * 4)) return nil();
      s->ints.resize((size_t)cnt);
      for (int64_t k = 0; k < cnt; ++k) s->ints[(size_t)k] = i32();
    } else if (type == 14) {
      const int64_t cnt = len();
      if (cnt < 0 || (uint64_t)cnt > (uint64_t)(n - o) / 8) return fail("unexpected end of the stream (a vector is longer than what is left)");
      if (!need((size_t)cnt * 8)) return nil();
      s->reals.resize((size_t)cnt);
      for (int64_t k = 0; k < cnt; ++k) s->reals[(size_t)k] = f64();
    } else if (type == 16 || type == 19 || type == 20) {
      const int64_t cnt = len();
      if (!need((size_t)cnt)) return nil();  // every element takes at least 4 bytes: cheap sanity bound
      s->items.reserve((size_t)cnt);
      for (int64_t k = 0; k < cnt && ok; ++k) s->items.push_back(item(depth + 1));
    } else if (type == 24) {
      const int64_t cnt = len();
      if (!need((size_t)cnt)) return nil();
      o += (size_t)cnt;
    } else {
      return fail("unsupported SEXP type " + std::to_string(type));
    }
    if (has_attr) s->attr = item(depth + 1);
    return s;
  }
};

bool class_is_dgc(const Sexp* obj) {
  if (!obj || obj->type != 25) return false;
  const Sexp* cls = obj->get("class");
  return cls && !cls->items.empty() && cls->items[0]->str == "dgCMatrix";
}

int dgc_to_spmat(const Sexp* obj, plaidgpu_spmat* m) {
  const Sexp *si = obj->get("i"), *sp = obj->get("p"), *sx = obj->get("x"), *sd = obj->get("Dim"), *sn = obj->get("Dimnames");
  if (!si || !sp || !sx || !sd || sd->ints.size() != 2) return io_fail("dgCMatrix without i / p / x / Dim slots");
  m->P = sd->ints[0];
  m->N = sd->ints[1];
  if (m->P <= 0 || m->N < 0 || sp->ints.size() != (size_t)m->N + 1) return io_fail("dgCMatrix: inconsistent Dim and p");
  if (si->ints.size() != sx->reals.size() || (int64_t)si->ints.size() != (int64_t)sp->ints.back())
    return io_fail("dgCMatrix: inconsistent i, x and p");
  for (size_t j = 0; j + 1 < sp->ints.size(); ++j)
    if (sp->ints[j] > sp->ints[j + 1] || sp->ints[j] < 0) return io_fail("dgCMatrix: p is not non-decreasing");
  for (int32_t r : si->ints)
    if (r < 0 || r >= m->P) return io_fail("dgCMatrix: row index out of range");
  m->p = sp->ints;
  m->i = si->ints;
  m->x = sx->reals;
  if (sn && sn->items.size() == 2) {
    for (int d = 0; d < 2; ++d) {
      const Sexp* names = sn->items[(size_t)d].get();
      std::vector<std::string>& dst = d == 0 ? m->rownames : m->colnames;
      if (names && names->type == 16)
        for (const SexpP& c : names->items) dst.push_back(c->str);
    }
  }
  return PLAIDGPU_OK;
}

}  // namespace

extern "C" {

const char* plaidgpu_io_error(void) { return g_io_err.c_str(); }

int plaidgpu_spmat_read_mtx(const char* path, plaidgpu_spmat** out) try {
  if (!path || !out) return io_fail("null argument");
  std::string buf;
  if (!load(path, buf)) return PLAIDGPU_ERR_ARG;
  std::unique_ptr<plaidgpu_spmat> m(new plaidgpu_spmat);
  const int rc = parse_mtx(buf, m.get());
  if (rc) return rc;
  *out = m.release();
  return PLAIDGPU_OK;
} catch (const std::bad_alloc&) {
  return PLAIDGPU_ERR_NOMEM;
} catch (...) {
  return PLAIDGPU_ERR_ARG;
}

int plaidgpu_spmat_read_10x(const char* dir, plaidgpu_spmat** out) try {
  if (!dir || !out) return io_fail("null argument");
  const std::string d(dir);
  auto pick = [&](std::initializer_list<const char*> names) -> std::string {
    for (const char* nm : names)
      if (exists(d + "/" + nm)) return d + "/" + nm;
    return std::string();
  };
  const std::string fm = pick({"matrix.mtx.gz", "matrix.mtx"});
  const std::string ff = pick({"features.tsv.gz", "features.tsv", "genes.tsv.gz", "genes.tsv"});
  const std::string fb = pick({"barcodes.tsv.gz", "barcodes.tsv"});
  if (fm.empty()) return io_fail("no matrix.mtx[.gz] in " + d);
  plaidgpu_spmat* m = nullptr;
  int rc = plaidgpu_spmat_read_mtx(fm.c_str(), &m);
  if (rc) return rc;
  std::unique_ptr<plaidgpu_spmat> hold(m);
  std::string buf;
  if (!ff.empty()) {
    if (!load(ff, buf)) return PLAIDGPU_ERR_ARG;
    read_names(buf, 1, m->rownames);  // gene symbols (Read10X gene.column = 2); the only column if there is one
    if ((int64_t)m->rownames.size() != m->P) return io_fail("features file has " + std::to_string(m->rownames.size()) + " rows, matrix has " + std::to_string(m->P));
  }
  if (!fb.empty()) {
    if (!load(fb, buf)) return PLAIDGPU_ERR_ARG;
    read_names(buf, 0, m->colnames);
    if ((int64_t)m->colnames.size() != m->N) return io_fail("barcodes file has " + std::to_string(m->colnames.size()) + " rows, matrix has " + std::to_string(m->N));
  }
  *out = hold.release();
  return PLAIDGPU_OK;
} catch (const std::bad_alloc&) {
  return PLAIDGPU_ERR_NOMEM;
} catch (...) {
  return PLAIDGPU_ERR_ARG;
}

int plaidgpu_spmat_read_rda(const char* path, const char* object, plaidgpu_spmat** out) try {
  if (!path || !out) return io_fail("null argument");
  std::string buf;
  if (!load(path, buf)) return PLAIDGPU_ERR_ARG;
  size_t o = 0;
  bool rda = false;
  if (buf.compare(0, 5, "RDX3\n") == 0 || buf.compare(0, 5, "RDX2\n") == 0) {
    o = 5;
    rda = true;
  } else if (buf.compare(0, 3, "RDA") == 0 || buf.compare(0, 3, "RDB") == 0) {
    return io_fail("only XDR (save(ascii = FALSE)) files are supported");
  }
  if (buf.compare(o, 2, "X\n") != 0) return io_fail(std::string(path) + ": not an R save() / saveRDS() XDR stream (xz / bzip2 compression is not supported)");
  XdrReader r{(const unsigned char*)buf.data(), buf.size()};
  r.o = o + 2;
  const int32_t version = r.i32();
  r.i32();
  r.i32();
  if (version == 3) {
    const int32_t k = r.i32();
    if (k < 0 || !r.need((size_t)k)) return io_fail("bad native-encoding header");
    r.o += (size_t)k;
  } else if (version != 2) {
    return io_fail("unsupported serialisation version " + std::to_string(version));
  }
  SexpP top = r.item();
  if (!r.ok) return io_fail(std::string(path) + ": " + r.why);
  const Sexp* found = nullptr;
  if (rda) {  // save(): a tagged pairlist of the saved objects
    if (top->type != 2) return io_fail("save() file without an object list");
    for (size_t k = 0; k < top->items.size(); ++k) {
      if (object && *object && top->tags[k] != object) continue;
      if (class_is_dgc(top->items[k].get())) {
        found = top->items[k].get();
        break;
      }
      if (object && *object) return io_fail(std::string("object '") + object + "' is not a dgCMatrix");
    }
  } else if (class_is_dgc(top.get())) {  // saveRDS(): the object itself
    found = top.get();
  }
  if (!found) return io_fail(object && *object ? std::string("no object named '") + object + "'" : std::string("no dgCMatrix in ") + path);
  std::unique_ptr<plaidgpu_spmat> m(new plaidgpu_spmat);
  const int rc = dgc_to_spmat(found, m.get());
  if (rc) return rc;
  *out = m.release();
  return PLAIDGPU_OK;
} catch (const std::bad_alloc&) {
  return PLAIDGPU_ERR_NOMEM;
} catch (...) {
  return PLAIDGPU_ERR_ARG;
}

void plaidgpu_spmat_free(plaidgpu_spmat* m) { delete m; }

int plaidgpu_spmat_view(const plaidgpu_spmat* m, plaidgpu_matrix* M) try {
  if (!m || !M) return PLAIDGPU_ERR_ARG;
  memset(M, 0, sizeof(*M));
  M->kind = PLAIDGPU_CSC;
  M->location = PLAIDGPU_HOST;
  M->P = m->P;
  M->N = m->N;
  M->p = m->p.data();
  M->i = m->i.data();
  M->x = m->x.data();
  return PLAIDGPU_OK;
} catch (const std::bad_alloc&) {
  return PLAIDGPU_ERR_NOMEM;
} catch (...) {
  return PLAIDGPU_ERR_ARG;
}

int64_t plaidgpu_spmat_nnz(const plaidgpu_spmat* m) { return m ? (int64_t)m->i.size() : 0; }
int64_t plaidgpu_spmat_num_rownames(const plaidgpu_spmat* m) { return m ? (int64_t)m->rownames.size() : 0; }
int64_t plaidgpu_spmat_num_colnames(const plaidgpu_spmat* m) { return m ? (int64_t)m->colnames.size() : 0; }
const char* plaidgpu_spmat_rowname(const plaidgpu_spmat* m, int64_t k) {
  return (m && k >= 0 && k < (int64_t)m->rownames.size()) ? m->rownames[(size_t)k].c_str() : nullptr;
}
const char* plaidgpu_spmat_colname(const plaidgpu_spmat* m, int64_t k) {
  return (m && k >= 0 && k < (int64_t)m->colnames.size()) ? m->colnames[(size_t)k].c_str() : nullptr;
}

}  // extern "C"
