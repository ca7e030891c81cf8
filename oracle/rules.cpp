// ORACLE -- TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (see oracle/README.md).
//
// CPU restatement of the SparseConvNet (scn) rulebook algorithms that
// /root/reference/uresnet/models/uresnet_sparse.py:19-24 reaches through
// `import sparseconvnet as scn` (uresnet_sparse.py:9).  scn itself is an un-vendored,
// unpinned third-party dependency that is absent from /root/reference and from this image,
// so this file restates its *published algorithm* from the description in SURVEY.md
// Appendix A (items 1-6); the call sites that anchor it are uresnet_sparse.py:20 (InputLayer,
// mode 3), :21 (SubmanifoldConvolution k=3) and :22 (UNet -> Convolution/Deconvolution
// k=2,s=2).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference
// legs may load this library; the product path never does.
//
// Algorithm shape follows scn's CPU metadata code: one std::unordered_map per batch sample
// from coordinate to row, sequential insertion in row order, row ids handed out in order of
// first occurrence, per-offset rule lists of (input row, output row).
//
// Build: g++ -O2 -shared -fPIC -o oracle/liboracle_rules.so oracle/rules.cpp
#include <cstdint>
#include <cstring>
#include <unordered_map>
#include <vector>

namespace {

struct PointHash {
  size_t operator()(const std::vector<int64_t>& p) const {
    // FNV-1a over the coordinate words
    uint64_t h = 1469598103934665603ull;
    for (int64_t v : p) {
      h ^= (uint64_t)v;
      h *= 1099511628211ull;
    }
    return (size_t)h;
  }
};
using Grid = std::unordered_map<std::vector<int64_t>, int32_t, PointHash>;

// one grid per batch sample, resized on demand (Appendix A.1: sample ids may be sparse)
struct Grids {
  std::vector<Grid> g;
  Grid& at(int64_t b) {
    if ((size_t)b >= g.size()) g.resize(b + 1);
    return g[b];
  }
  const Grid* find(int64_t b) const {
    if (b < 0 || (size_t)b >= g.size()) return nullptr;
    return &g[b];
  }
};

void build_grids(const int64_t* coords, int64_t n, int D, Grids& gs) {
  std::vector<int64_t> p(D);
  for (int64_t r = 0; r < n; ++r) {
    const int64_t* c = coords + r * (D + 1);
    for (int d = 0; d < D; ++d) p[d] = c[d];
    gs.at(c[D])[p] = (int32_t)r;
  }
}

}  // namespace

extern "C" {

// InputLayer, mode 3 (uresnet_sparse.py:20).  coords [n, D+1] int64 (last col = sample).
// row(p) = number of distinct (sample, coord) seen before p's first occurrence (App. A.2).
// Outputs: in2row[n]; active[nActive, D+1] (caller allocates n rows).  Returns nActive.
int64_t orc_input_rows(const int64_t* coords, int64_t n, int D, int32_t* in2row, int64_t* active) {
  Grids gs;
  std::vector<int64_t> p(D);
  int64_t ctr = 0;
  for (int64_t i = 0; i < n; ++i) {
    const int64_t* c = coords + i * (D + 1);
    for (int d = 0; d < D; ++d) p[d] = c[d];
    Grid& g = gs.at(c[D]);
    auto it = g.find(p);
    if (it == g.end()) {
      g.emplace(p, (int32_t)ctr);
      std::memcpy(active + ctr * (D + 1), c, sizeof(int64_t) * (D + 1));
      in2row[i] = (int32_t)ctr++;
    } else {
      in2row[i] = it->second;
    }
  }
  return ctr;
}

// SubmanifoldConvolution rules, filter k^D (uresnet_sparse.py:21; App. A.3).
// For every active output site, probe out + (d - k/2); offset index enumerates the k^D box
// with dimension D-1 fastest.  rules are returned as parallel arrays sorted by offset:
// rule_off[K+1] (CSR offsets), rule_in[], rule_out[] (caller allocates n*K).  Returns #pairs.
int64_t orc_subm_rules(const int64_t* active, int64_t n, int D, int k, int64_t* rule_off, int32_t* rule_in,
                       int32_t* rule_out) {
  Grids gs;
  build_grids(active, n, D, gs);
  int K = 1;
  for (int d = 0; d < D; ++d) K *= k;
  std::vector<std::vector<int32_t>> rin(K), rout(K);
  std::vector<int64_t> p(D);
  for (int64_t o = 0; o < n; ++o) {
    const int64_t* c = active + o * (D + 1);
    const Grid* g = gs.find(c[D]);
    for (int f = 0; f < K; ++f) {
      int rem = f;
      for (int d = D - 1; d >= 0; --d) {
        p[d] = c[d] + (rem % k) - k / 2;
        rem /= k;
      }
      auto it = g->find(p);
      if (it != g->end()) {
        rin[f].push_back(it->second);
        rout[f].push_back((int32_t)o);
      }
    }
  }
  int64_t tot = 0;
  for (int f = 0; f < K; ++f) {
    rule_off[f] = tot;
    std::memcpy(rule_in + tot, rin[f].data(), rin[f].size() * sizeof(int32_t));
    std::memcpy(rule_out + tot, rout[f].data(), rout[f].size() * sizeof(int32_t));
    tot += (int64_t)rin[f].size();
  }
  rule_off[K] = tot;
  return tot;
}

// Convolution with filter == stride == s (uresnet_sparse.py:22, downsample=[2,2]; App. A.4).
// child p -> parent p / s per dimension; offset = row-major index of p % s (dim D-1 fastest).
// Output rows are handed out in discovery order while walking input rows 0..n-1 (the
// deterministic stand-in for scn's hash-iteration order, which is unspecified).
// Outputs: parent[n], kofs[n], coarse[m, D+1] (caller allocates n rows).  Returns m.
int64_t orc_down_rules(const int64_t* active, int64_t n, int D, int s, int32_t* parent, int32_t* kofs,
                       int64_t* coarse) {
  Grids gs;
  std::vector<int64_t> p(D);
  int64_t ctr = 0;
  for (int64_t i = 0; i < n; ++i) {
    const int64_t* c = active + i * (D + 1);
    int f = 0;
    for (int d = 0; d < D; ++d) {
      p[d] = c[d] / s;
      f = f * s + (int)(c[d] % s);
    }
    Grid& g = gs.at(c[D]);
    auto it = g.find(p);
    int32_t row;
    if (it == g.end()) {
      row = (int32_t)ctr;
      g.emplace(p, row);
      for (int d = 0; d < D; ++d) coarse[ctr * (D + 1) + d] = p[d];
      coarse[ctr * (D + 1) + D] = c[D];
      ++ctr;
    } else {
      row = it->second;
    }
    parent[i] = row;
    kofs[i] = f;
  }
  return ctr;
}

}  // extern "C"
