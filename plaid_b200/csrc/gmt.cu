// Gene-set ingestion (scope row f2): read.gmt + gmt2mat of the reference (R/gmt-utils.R:99-125, 19-66) as
// host C++ behind the C ABI, so that a GMT file becomes the gene x set incidence matrix (and the row
// alignment with X) without the R-level loops that take ~50 s for 50k sets (reference
// experiments/benchmark/benchmark-plaid.R:42-43).  Pure host code: text parsing is not GPU work.
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/plaidgpu.h"

struct plaidgpu_gmt {
  std::vector<std::string> set_names;   // gmt2mat column order
  std::vector<std::string> gene_names;  // gmt2mat row order
  std::vector<int32_t> Gp, Gi;          // CSC genes x sets, rows sorted within a column
  std::unordered_map<std::string, int32_t> gene_pos;
};

namespace {

struct RawSet {
  std::string name;
  std::vector<std::string> genes;  // unique, file order; "" and "NA" dropped (R/gmt-utils.R:117)
};

void parse(const char* text, size_t len, std::vector<RawSet>& sets) {
  size_t pos = 0;
  std::unordered_map<std::string, char> seen;
  while (pos < len) {
    size_t eol = pos;
    while (eol < len && text[eol] != '\n') ++eol;
    size_t end = eol;
    if (end > pos && text[end - 1] == '\r') --end;
    // utils::read.csv(sep = "!", comment.char = "#")[, 1]  (:108): '#' ends the line wherever it stands, '!' is the
    // column separator (only the first column is kept) and a line that is empty afterwards is skipped.  Deviation
    // kept: read.csv's quote = "\"" (a double quote would start a field in which '!', '#' and line ends are literal)
    for (size_t i = pos; i < end; ++i)
      if (text[i] == '#' || text[i] == '!') {
        end = i;
        break;
      }
    if (end > pos) {
      RawSet s;
      int field = 0;
      size_t f0 = pos;
      seen.clear();
      for (size_t i = pos; i <= end; ++i) {
        // fields are tab separated; genes may additionally be separated by blanks (:115)
        const bool sep = (i == end) || text[i] == '\t' || (field >= 2 && text[i] == ' ');
        if (!sep) continue;
        if (field == 0) {
          s.name.assign(text + f0, i - f0);
        } else if (field >= 2) {
          std::string g(text + f0, i - f0);
          if (!g.empty() && g != "NA" && seen.emplace(g, 1).second) s.genes.push_back(std::move(g));
        }
        if (i < end && text[i] == '\t' && field < 2) ++field; else if (field >= 2) field = 2; else ++field;
        f0 = i + 1;
      }
      sets.push_back(std::move(s));
    }
    pos = eol + 1;
  }
}

int build(std::vector<RawSet>& raw, plaidgpu_gmt* g) {
  // gmt <- gmt[order(-sapply(gmt, length))]; gmt <- gmt[!duplicated(names(gmt))]   (:25-26)
  std::vector<size_t> ord(raw.size());
  for (size_t k = 0; k < ord.size(); ++k) ord[k] = k;
  std::stable_sort(ord.begin(), ord.end(), [&](size_t a, size_t b) { return raw[a].genes.size() > raw[b].genes.size(); });
  std::unordered_map<std::string, char> names_seen;
  std::vector<size_t> keep;
  for (size_t k : ord)
    if (names_seen.emplace(raw[k].name, 1).second) keep.push_back(k);
  // bg <- names(sort(table(unlist(gmt)), decreasing = TRUE))   (:30-31): names sorted, then stable by count
  std::unordered_map<std::string, int32_t> cnt;
  for (size_t k : keep)
    for (const std::string& s : raw[k].genes) ++cnt[s];
  std::vector<std::pair<std::string, int32_t>> bg(cnt.begin(), cnt.end());
  std::sort(bg.begin(), bg.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
  std::stable_sort(bg.begin(), bg.end(), [](const auto& a, const auto& b) { return a.second > b.second; });
  std::unordered_map<std::string, int32_t> pos;
  pos.reserve(bg.size() * 2);
  for (size_t k = 0; k < bg.size(); ++k) pos.emplace(bg[k].first, (int32_t)k);
  // incidence + final row order by decreasing membership count (:62); counts equal `cnt` here because
  // read.gmt already made the members of a set unique, so the stable re-sort is the identity
  const size_t S = keep.size(), P = bg.size();
  g->set_names.resize(S);
  g->gene_names.resize(P);
  for (size_t k = 0; k < P; ++k) g->gene_names[k] = bg[k].first;
  g->Gp.assign(S + 1, 0);
  size_t nnz = 0;
  for (size_t k : keep) nnz += raw[k].genes.size();
  if (nnz > 0x7fffffffu) return PLAIDGPU_ERR_ARG;
  g->Gi.resize(nnz);
  size_t o = 0;
  for (size_t j = 0; j < S; ++j) {
    const RawSet& s = raw[keep[j]];
    g->set_names[j] = s.name;
    const size_t b = o;
    for (const std::string& m : s.genes) g->Gi[o++] = pos[m];
    std::sort(g->Gi.begin() + b, g->Gi.begin() + o);
    g->Gp[j + 1] = (int32_t)o;
  }
  g->gene_pos = std::move(pos);
  return PLAIDGPU_OK;
}

}  // namespace

extern "C" {

int plaidgpu_gmt_from_buffer(const char* text, int64_t len, plaidgpu_gmt** out) try {
  if (!text || len < 0 || !out) return PLAIDGPU_ERR_ARG;
  *out = nullptr;
  plaidgpu_gmt* g = new (std::nothrow) plaidgpu_gmt();
  if (!g) return PLAIDGPU_ERR_NOMEM;
  std::vector<RawSet> raw;
  parse(text, (size_t)len, raw);
  const int rc = build(raw, g);
  if (rc != PLAIDGPU_OK) {
    delete g;
    return rc;
  }
  *out = g;
  return PLAIDGPU_OK;
} catch (const std::bad_alloc&) {
  return PLAIDGPU_ERR_NOMEM;
} catch (...) {
  return PLAIDGPU_ERR_ARG;
}

int plaidgpu_gmt_read(const char* path, plaidgpu_gmt** out) try {
  if (!path || !out) return PLAIDGPU_ERR_ARG;
  FILE* f = fopen(path, "rb");
  if (!f) return PLAIDGPU_ERR_ARG;
  std::string buf;
  char tmp[1 << 16];
  size_t n;
  while ((n = fread(tmp, 1, sizeof(tmp), f)) > 0) buf.append(tmp, n);
  fclose(f);
  return plaidgpu_gmt_from_buffer(buf.data(), (int64_t)buf.size(), out);
} catch (const std::bad_alloc&) {
  return PLAIDGPU_ERR_NOMEM;
} catch (...) {
  return PLAIDGPU_ERR_ARG;
}

void plaidgpu_gmt_free(plaidgpu_gmt* g) { delete g; }
int64_t plaidgpu_gmt_num_sets(const plaidgpu_gmt* g) { return g ? (int64_t)g->set_names.size() : 0; }
int64_t plaidgpu_gmt_num_genes(const plaidgpu_gmt* g) { return g ? (int64_t)g->gene_names.size() : 0; }
int64_t plaidgpu_gmt_nnz(const plaidgpu_gmt* g) { return g ? (int64_t)g->Gi.size() : 0; }
const char* plaidgpu_gmt_set_name(const plaidgpu_gmt* g, int64_t k) {
  return (g && k >= 0 && k < (int64_t)g->set_names.size()) ? g->set_names[(size_t)k].c_str() : nullptr;
}
const char* plaidgpu_gmt_gene_name(const plaidgpu_gmt* g, int64_t k) {
  return (g && k >= 0 && k < (int64_t)g->gene_names.size()) ? g->gene_names[(size_t)k].c_str() : nullptr;
}
int plaidgpu_gmt_csc(const plaidgpu_gmt* g, int32_t* Gp, int32_t* Gi) try {
  if (!g || !Gp || (!Gi && !g->Gi.empty())) return PLAIDGPU_ERR_ARG;
  memcpy(Gp, g->Gp.data(), g->Gp.size() * sizeof(int32_t));
  if (!g->Gi.empty()) memcpy(Gi, g->Gi.data(), g->Gi.size() * sizeof(int32_t));
  return PLAIDGPU_OK;
} catch (const std::bad_alloc&) {
  return PLAIDGPU_ERR_NOMEM;
} catch (...) {
  return PLAIDGPU_ERR_ARG;
}
int plaidgpu_gmt_rowmap(const plaidgpu_gmt* g, const char* const* xnames, int32_t P, int32_t* rowmap) try {
  if (!g || !xnames || !rowmap || P < 0) return PLAIDGPU_ERR_ARG;
  std::unordered_map<std::string, char> seen;
  seen.reserve((size_t)P * 2);
  for (int32_t r = 0; r < P; ++r) {
    rowmap[r] = -1;
    if (!xnames[r]) continue;
    std::string n(xnames[r]);
    if (!seen.emplace(n, 1).second) continue;  // first occurrence of a duplicated rowname wins (R/plaid.R:71)
    auto it = g->gene_pos.find(n);
    if (it != g->gene_pos.end()) rowmap[r] = it->second;
  }
  return PLAIDGPU_OK;
} catch (const std::bad_alloc&) {
  return PLAIDGPU_ERR_NOMEM;
} catch (...) {
  return PLAIDGPU_ERR_ARG;
}

}  // extern "C"
