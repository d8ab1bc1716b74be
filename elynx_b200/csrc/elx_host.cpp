// Host-side helpers: tree walk compilation and the per-component eigensystems.
#include <algorithm>
#include <cmath>
#include <cstdio>

#include "elx_internal.hpp"

namespace elx {

bool compile_walk(int n_nodes, const int32_t *parent, const int32_t *leaf_row, WalkProgram &out, int &n_leaves,
                  std::string &msg) {
  const int N = n_nodes;
  if (N < 1) { msg = "tree: no nodes"; return false; }
  if (parent[0] != -1) { msg = "tree: parent[0] must be -1 (root first, preorder)"; return false; }
  std::vector<int> n_child(N, 0), first_child(N, -1), next_sib(N, -1), last_child(N, -1);
  {
    // preorder check: parent[n] must lie on the path root..n-1
    std::vector<int> path;
    path.push_back(0);
    for (int n = 1; n < N; ++n) {
      int p = parent[n];
      if (p < 0 || p >= n) { msg = "tree: parent[n] must be in [0, n) (preorder)"; return false; }
      while (!path.empty() && path.back() != p) path.pop_back();
      if (path.empty()) { msg = "tree: nodes are not in preorder"; return false; }
      path.push_back(n);
      if (first_child[p] < 0) first_child[p] = n; else next_sib[last_child[p]] = n;
      last_child[p] = n;
      n_child[p]++;
    }
  }
  int T = 0;
  for (int n = 0; n < N; ++n) if (n_child[n] == 0) ++T;
  std::vector<char> seen(T, 0);
  for (int n = 0; n < N; ++n) {
    if (n_child[n] == 0) {
      int r = leaf_row[n];
      if (r < 0 || r >= T || seen[r]) { msg = "tree: leaf_row of the leaves must be a permutation of 0..T-1"; return false; }
      seen[r] = 1;
    } else if (leaf_row[n] != -1) {
      msg = "tree: leaf_row must be -1 for internal nodes";
      return false;
    }
  }
  n_leaves = T;
  std::vector<int64_t> size(N, 1);
  for (int n = N - 1; n >= 1; --n) size[parent[n]] += size[n];

  out.recs.clear();
  out.recs.reserve(N);
  std::vector<int> free_slots;
  int high = 1;  // slot 0 = root states
  auto alloc = [&]() {
    if (!free_slots.empty()) { int s = free_slots.back(); free_slots.pop_back(); return s; }
    return high++;
  };
  struct Frame { int node, src, dst; std::vector<int> kids; size_t next; };
  std::vector<Frame> st;
  auto enter = [&](int v, int src, bool is_last) {
    Frame f;
    f.node = v; f.src = src; f.next = 0;
    if (n_child[v] == 0) {
      f.dst = -1;
    } else {
      f.dst = is_last ? src : alloc();
      for (int c = first_child[v]; c >= 0; c = next_sib[c]) f.kids.push_back(c);
      // heaviest child last (stable otherwise)
      size_t h = 0;
      for (size_t i = 1; i < f.kids.size(); ++i) if (size[f.kids[i]] > size[f.kids[h]]) h = i;
      int heavy = f.kids[h];
      f.kids.erase(f.kids.begin() + (long)h);
      f.kids.push_back(heavy);
    }
    WalkRec r;
    r.node = v; r.src = (int16_t)src; r.dst = (int16_t)f.dst; r.leaf_row = leaf_row[v]; r.pad = 0;
    out.recs.push_back(r);
    st.push_back(std::move(f));
  };
  enter(0, 0, true);
  while (!st.empty()) {
    Frame &f = st.back();
    if (f.next < f.kids.size()) {
      int c = f.kids[f.next];
      bool last = (f.next + 1 == f.kids.size());
      f.next++;
      int src = f.dst;
      enter(c, src, last);
    } else {
      // a slot allocated here (dst != src) is released unless the last child took it over: the last child
      // writes into f.dst itself, so the slot is simply dead once this subtree is done.
      if (f.dst >= 0 && f.dst != f.src) free_slots.push_back(f.dst);
      st.pop_back();
    }
  }
  out.n_slots = high;
  if (high > 32000) { msg = "tree: too many live slots"; return false; }
  return true;
}

void compile_pair_walk(int n_nodes, const int32_t *parent, const int32_t *leaf_row, PairProgram &out) {
  const int N = n_nodes;
  std::vector<std::vector<int>> kids(N);
  for (int n = 1; n < N; ++n) kids[parent[n]].push_back(n);
  std::vector<int64_t> size(N, 1);
  for (int n = N - 1; n >= 1; --n) size[parent[n]] += size[n];
  for (int n = 0; n < N; ++n)
    std::stable_sort(kids[n].begin(), kids[n].end(), [&](int a, int b) { return size[a] < size[b]; });

  out.recs.clear();
  out.recs.reserve((size_t)N / 2 + 2);
  std::vector<int> free_slots;
  int high = 0;
  auto alloc = [&]() {
    if (!free_slots.empty()) { int s = free_slots.back(); free_slots.pop_back(); return s; }
    return high++;
  };
  // the groups of v: consecutive pairs of its size-sorted children; with an odd count the lightest child stands alone,
  // so that the heaviest child is always in the last group
  auto n_groups = [&](int v) { return (int)(kids[v].size() + 1) / 2; };
  auto group = [&](int v, int gi, int &x, int &y) {
    const std::vector<int> &k = kids[v];
    const int odd = (int)k.size() & 1;
    if (odd && gi == 0) { x = k[0]; y = -1; return; }
    const int i = 2 * gi - odd;
    x = k[i]; y = k[i + 1];
  };
  auto emit = [&](int A, int B, int src, int free_after_read) {
    PairRec r;
    r.nodeA = A; r.nodeB = B; r.src = (int16_t)src;
    if (free_after_read >= 0) free_slots.push_back(free_after_read);  // the kernel reads src before it writes dstA/dstB
    r.dstA = (int8_t)((!kids[A].empty() && n_groups(A) > 1) ? alloc() : -1);
    r.dstB = (int8_t)((B >= 0 && !kids[B].empty()) ? alloc() : -1);
    r.leafA = kids[A].empty() ? leaf_row[A] : -1;
    r.leafB = (B >= 0 && kids[B].empty()) ? leaf_row[B] : -1;
    out.recs.push_back(r);
    return r;
  };
  struct Frame { int v, src_in, vs, gi, stage, A, B, dstA, dstB; };
  std::vector<Frame> st;
  {
    PairRec r = emit(0, -1, -1, -1);
    if (!kids[0].empty()) st.push_back(Frame{0, -1, r.dstA, 0, 0, 0, 0, 0, 0});
  }
  while (!st.empty()) {
    Frame &f = st.back();
    if (f.stage == 0) {
      if (f.gi == n_groups(f.v)) { st.pop_back(); continue; }
      int x, y;
      group(f.v, f.gi, x, y);
      int A = x, B = y;
      if (y >= 0 && kids[x].empty() && !kids[y].empty()) { A = y; B = x; }  // the internal member continues in registers
      const bool last = f.gi + 1 == n_groups(f.v);
      const int src = f.gi == 0 ? f.src_in : f.vs;
      PairRec r = emit(A, B, src, last ? f.vs : -1);
      f.A = A; f.B = B; f.dstA = r.dstA; f.dstB = r.dstB;
      f.stage = 1;
    } else if (f.stage == 1) {
      f.stage = 2;
      if (!kids[f.A].empty()) {
        Frame c{f.A, -1, f.dstA, 0, 0, 0, 0, 0, 0};
        st.push_back(c);  // invalidates f
      }
    } else {
      f.stage = 0;
      f.gi++;
      if (f.B >= 0 && !kids[f.B].empty()) {
        Frame c{f.B, f.dstB, f.dstB, 0, 0, 0, 0, 0, 0};
        st.push_back(c);
      }
    }
  }
  out.n_slots = high;
}

void compile_quad_walk(int n_nodes, const int32_t *parent, const int32_t *leaf_row, QuadProgram &out, int max_frontier) {
  const int N = n_nodes;
  constexpr int WMAX = 4, INF = 1 << 29;
  const int W = max_frontier < 1 ? 1 : (max_frontier > WMAX ? WMAX : max_frontier);  // frontier nodes per cap (4: quad mode, 2: duo mode)
  std::vector<std::vector<int>> kids(N);
  for (int n = 1; n < N; ++n) kids[parent[n]].push_back(n);
  std::vector<int64_t> size(N, 1);
  for (int n = N - 1; n >= 1; --n) size[parent[n]] += size[n];
  for (int n = 0; n < N; ++n)
    std::stable_sort(kids[n].begin(), kids[n].end(), [&](int a, int b) { return size[a] < size[b]; });

  // ---- tree DP.  "Open cap" = the cap that contains the branch above v.  cost[v][w]: fewest draws inside v's subtree
  // when the open cap has w frontier nodes at or below v; nn[v][w]: how many nodes of the open cap lie at or below v.
  // w == 1 is either "v is a frontier node" (closed: v's own caps are counted) or, for a unary v, "v is summed out".
  struct Choice {
    std::vector<int8_t> kw[WMAX + 1];    // v summed out: frontier share of every kid (kids[] order)
    std::vector<int8_t> ckw;             // v closed: share and cap number of every kid (caps = runs of kids)
    std::vector<int32_t> cgrp;
    bool closed1 = true;                 // the w == 1 entry is the closed one
  };
  std::vector<Choice> ch(N);
  std::vector<int> cost((size_t)N * (W + 1), INF), nn((size_t)N * (W + 1), 0);
  auto C_ = [&](int v, int w) -> int & { return cost[(size_t)v * (W + 1) + w]; };
  auto N_ = [&](int v, int w) -> int & { return nn[(size_t)v * (W + 1) + w]; };
  struct Part { int cost = INF, nodes = 0; std::vector<int8_t> kw; };
  for (int v = N - 1; v >= 0; --v) {
    const std::vector<int> &k = kids[v];
    if (k.empty()) { C_(v, 1) = 0; N_(v, 1) = 1; continue; }
    // v summed out: all kids share the open cap
    if ((int)k.size() <= W) {
      Part comb[WMAX + 1];
      comb[0].cost = 0;
      for (size_t i = 0; i < k.size(); ++i) {
        Part nx[WMAX + 1];
        for (int w0 = 0; w0 <= W; ++w0) {
          if (comb[w0].cost >= INF) continue;
          for (int w = 1; w0 + w <= W; ++w) {
            if (C_(k[i], w) >= INF) continue;
            const int val = comb[w0].cost + C_(k[i], w);
            Part &t = nx[w0 + w];
            if (val < t.cost) { t.cost = val; t.nodes = comb[w0].nodes + N_(k[i], w); t.kw = comb[w0].kw; t.kw.push_back((int8_t)w); }
          }
        }
        for (int w = 0; w <= W; ++w) comb[w] = nx[w];
      }
      for (int w = 1; w <= W; ++w)
        if (comb[w].cost < INF && comb[w].nodes + 1 <= QUAD_MAX_CAP) {
          C_(v, w) = comb[w].cost; N_(v, w) = comb[w].nodes + 1; ch[v].kw[w] = comb[w].kw;
        }
    }
    // v closed: its kids are packed, in order, into caps of at most W frontier nodes (state = frontier size of the cap
    // being filled).  Back-pointers instead of per-state copies of the choices: linear in the number of kids, also for a
    // star with 10^5 leaves.
    struct Step { int8_t prev, w; bool fresh; };  // how state s after kid i was reached
    std::vector<Step> steps((k.size() + 1) * (W + 1), Step{-1, 0, false});
    int dcost[WMAX + 1], dnodes[WMAX + 1];
    for (int s = 0; s <= W; ++s) { dcost[s] = INF; dnodes[s] = 0; }
    dcost[0] = 0;
    for (size_t i = 0; i < k.size(); ++i) {
      int ncost[WMAX + 1], nnodes[WMAX + 1];
      for (int s = 0; s <= W; ++s) { ncost[s] = INF; nnodes[s] = 0; }
      Step *st_i = &steps[(i + 1) * (W + 1)];
      for (int s0 = 0; s0 <= W; ++s0) {
        if (dcost[s0] >= INF) continue;
        for (int w = 1; w <= W; ++w) {
          if (C_(k[i], w) >= INF) continue;
          if (s0 > 0 && s0 + w <= W && dnodes[s0] + N_(k[i], w) <= QUAD_MAX_CAP) {  // joins the cap being filled
            const int val = dcost[s0] + C_(k[i], w);
            if (val < ncost[s0 + w]) { ncost[s0 + w] = val; nnodes[s0 + w] = dnodes[s0] + N_(k[i], w); st_i[s0 + w] = Step{(int8_t)s0, (int8_t)w, false}; }
          }
          const int val = dcost[s0] + C_(k[i], w) + 1;  // starts a new cap
          if (val < ncost[w]) { ncost[w] = val; nnodes[w] = N_(k[i], w); st_i[w] = Step{(int8_t)s0, (int8_t)w, true}; }
        }
      }
      for (int s = 0; s <= W; ++s) { dcost[s] = ncost[s]; dnodes[s] = nnodes[s]; }
    }
    int best = 1;
    for (int s = 1; s <= W; ++s) if (dcost[s] < dcost[best]) best = s;
    {
      std::vector<int8_t> kw(k.size());
      std::vector<char> fresh(k.size());
      int s = best;
      for (size_t i = k.size(); i-- > 0;) {
        const Step &t = steps[(i + 1) * (W + 1) + s];
        kw[i] = t.w; fresh[i] = t.fresh ? 1 : 0;
        s = t.prev;
      }
      ch[v].ckw = kw;
      ch[v].cgrp.assign(k.size(), 0);
      int grp = -1;
      for (size_t i = 0; i < k.size(); ++i) { if (fresh[i]) ++grp; ch[v].cgrp[i] = grp; }
    }
    struct { int cost; } dpb{dcost[best]};
    if (dpb.cost < C_(v, 1)) { C_(v, 1) = dpb.cost; N_(v, 1) = 1; ch[v].closed1 = true; } else ch[v].closed1 = false;
  }

  // ---- records, depth first (light caps and light frontier nodes first), slots
  out.recs.clear();
  std::vector<int> free_slots;
  int high = 0;
  auto alloc = [&]() {
    if (!free_slots.empty()) { int s = free_slots.back(); free_slots.pop_back(); return s; }
    return high++;
  };
  struct Frame { int v, slot, gi, n_groups; std::vector<std::pair<int, int>> pending; size_t pi; size_t kid0 = 0; };  // kid0: first kid of cap gi
  std::vector<Frame> st;
  auto n_groups = [&](int v) { return ch[v].cgrp.empty() ? 0 : (int)ch[v].cgrp.back() + 1; };
  st.push_back(Frame{0, alloc(), 0, n_groups(0), {}, 0, 0});
  struct Visit { int x, w, par; };
  std::vector<Visit> vs;
  while (!st.empty()) {
    Frame &f = st.back();
    if (f.pi < f.pending.size()) {
      const std::pair<int, int> nb = f.pending[f.pi++];
      st.push_back(Frame{nb.first, nb.second, 0, n_groups(nb.first), {}, 0, 0});  // invalidates f
      continue;
    }
    if (f.gi == f.n_groups) { st.pop_back(); continue; }
    QuadRec r;
    r.root = f.v; r.src = (int16_t)f.slot; r.n_cap = 0;
    for (int q = 0; q < 4; ++q) { r.out[q] = -1; r.dst[q] = -1; r.leaf[q] = -1; }
    int nf = 0;
    vs.clear();
    const std::vector<int> &k = kids[f.v];
    size_t kid1 = f.kid0;  // the caps of a node are runs of its kids
    while (kid1 < k.size() && ch[f.v].cgrp[kid1] == f.gi) ++kid1;
    for (size_t i = kid1; i-- > f.kid0;) vs.push_back(Visit{k[i], ch[f.v].ckw[i], -1});
    f.kid0 = kid1;
    std::vector<int> boundary;
    while (!vs.empty()) {
      const Visit t = vs.back();
      vs.pop_back();
      QuadCapNode &cn = r.cap[r.n_cap];
      cn.node = t.x; cn.par = (int8_t)t.par; cn.fr = -1;
      const int local = r.n_cap++;
      const bool is_leaf = kids[t.x].empty();
      if (is_leaf || (t.w == 1 && ch[t.x].closed1)) {
        cn.fr = (int8_t)nf;
        r.out[nf] = t.x;
        r.leaf[nf] = is_leaf ? leaf_row[t.x] : -1;
        if (!is_leaf) boundary.push_back(nf);
        ++nf;
      } else {
        const std::vector<int> &kk = kids[t.x];
        for (size_t i = kk.size(); i-- > 0;) vs.push_back(Visit{kk[i], ch[t.x].kw[t.w][i], local});
      }
    }
    f.gi++;
    if (f.gi == f.n_groups) free_slots.push_back(f.slot);  // the record reads src before it writes dst
    std::stable_sort(boundary.begin(), boundary.end(), [&](int a, int b) { return size[r.out[a]] < size[r.out[b]]; });
    f.pending.clear();
    f.pi = 0;
    for (int q : boundary) {
      const int s = alloc();
      r.dst[q] = (int8_t)s;
      f.pending.push_back(std::make_pair(r.out[q], s));
    }
    out.recs.push_back(r);
  }
  out.n_slots = high;
}

bool eigen_reversible(int K, const double *pi, const double *exch, double *lambda, double *A, double *B,
                      std::string &msg) {
  std::vector<double> S((size_t)K * K), V((size_t)K * K, 0.0), sq(K);
  double emax = 0;
  for (int i = 0; i < K * K; ++i) emax = std::max(emax, std::fabs(exch[i]));
  for (int i = 0; i < K; ++i) {
    if (!(pi[i] > 0) || !std::isfinite(pi[i])) {
      msg = "model: stationary frequencies must be > 0 for the eigen path (non-reversible models unsupported)";
      return false;
    }
    sq[i] = std::sqrt(pi[i]);
  }
  for (int i = 0; i < K; ++i)
    for (int j = 0; j < K; ++j) {
      if (i == j) continue;
      double a = exch[i * K + j], b = exch[j * K + i];
      if (!std::isfinite(a) || a < 0) { msg = "model: exchangeabilities must be finite and >= 0"; return false; }
      if (std::fabs(a - b) > 1e-10 * (emax > 0 ? emax : 1)) {
        msg = "model: exchangeability matrix must be symmetric (non-reversible models unsupported)";
        return false;
      }
    }
  // Q = setDiagonal(E diag(pi)) (RateMatrix.hs:87-103); S = D^{1/2} Q D^{-1/2}
  for (int i = 0; i < K; ++i) {
    double rs = 0;
    for (int j = 0; j < K; ++j)
      if (j != i) rs += std::fabs(exch[i * K + j] * pi[j]);
    for (int j = 0; j < K; ++j)
      S[(size_t)i * K + j] = (i == j) ? -rs : sq[i] * 0.5 * (exch[i * K + j] + exch[j * K + i]) * sq[j];
    V[(size_t)i * K + i] = 1.0;
  }
  // cyclic Jacobi
  for (int sweep = 0; sweep < 100; ++sweep) {
    double off = 0, diag = 0;
    for (int i = 0; i < K; ++i)
      for (int j = 0; j < K; ++j) (i == j ? diag : off) += S[(size_t)i * K + j] * S[(size_t)i * K + j];
    if (off <= 1e-34 * (diag > 0 ? diag : 1) || off == 0) break;
    for (int p = 0; p < K - 1; ++p)
      for (int q = p + 1; q < K; ++q) {
        double apq = S[(size_t)p * K + q];
        if (apq == 0) continue;
        double app = S[(size_t)p * K + p], aqq = S[(size_t)q * K + q];
        double theta = (aqq - app) / (2 * apq);
        double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1));
        double c = 1 / std::sqrt(t * t + 1), s = t * c;
        for (int k = 0; k < K; ++k) {
          double akp = S[(size_t)k * K + p], akq = S[(size_t)k * K + q];
          S[(size_t)k * K + p] = c * akp - s * akq;
          S[(size_t)k * K + q] = s * akp + c * akq;
        }
        for (int k = 0; k < K; ++k) {
          double apk = S[(size_t)p * K + k], aqk = S[(size_t)q * K + k];
          S[(size_t)p * K + k] = c * apk - s * aqk;
          S[(size_t)q * K + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < K; ++k) {
          double vkp = V[(size_t)k * K + p], vkq = V[(size_t)k * K + q];
          V[(size_t)k * K + p] = c * vkp - s * vkq;
          V[(size_t)k * K + q] = s * vkp + c * vkq;
        }
      }
  }
  for (int k = 0; k < K; ++k) {
    double l = S[(size_t)k * K + k];
    // a rate matrix has eigenvalues <= 0; clamp rounding noise on the zero eigenvalue
    lambda[k] = l > 0 ? 0.0 : l;
  }
  for (int i = 0; i < K; ++i)
    for (int k = 0; k < K; ++k) {
      A[(size_t)i * K + k] = V[(size_t)i * K + k] / sq[i];
      B[(size_t)k * K + i] = V[(size_t)i * K + k] * sq[i];
    }
  return true;
}

}  // namespace elx
