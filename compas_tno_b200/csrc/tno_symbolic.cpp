// Host-side symbolic analysis (see tno_symbolic.hpp).
#include "tno_symbolic.hpp"

#include <algorithm>
#include <cstdlib>
#include <cmath>
#include <functional>
#include <numeric>
#include <stdexcept>

namespace tno {

namespace {

using Adj = std::vector<std::vector<int>>;

Adj build_adjacency(int n, const std::vector<std::pair<int, int>>& edges) {
  Adj adj(n);
  for (auto& e : edges) {
    if (e.first == e.second) continue;
    adj[e.first].push_back(e.second);
    adj[e.second].push_back(e.first);
  }
  for (auto& a : adj) {
    std::sort(a.begin(), a.end());
    a.erase(std::unique(a.begin(), a.end()), a.end());
  }
  return adj;
}

// Exact minimum (external) degree on the explicit elimination graph.  The graphs here have a few
// thousand vertices of degree <= 8, so the quadratic selection loop is irrelevant next to one pinv.
// tie_break 0: first vertex of minimum degree; 1: among those, the one whose elimination creates the least fill
// (minimum deficiency); 2: minimum deficiency among the vertices within two of the minimum degree (degree breaks ties).
// fill_out (optional) receives nnz(L) of the ordering, diagonal included.
std::vector<int> order_mindeg(const Adj& adj0, const std::vector<int>& subset, int tie_break = 0, int64_t* fill_out = nullptr) {
  const int n = (int)adj0.size();
  std::vector<char> in(n, 0), done(n, 0);
  for (int v : subset) in[v] = 1;
  Adj adj(n);
  for (int v : subset)
    for (int u : adj0[v])
      if (in[u]) adj[v].push_back(u);
  std::vector<int> order;
  order.reserve(subset.size());
  std::vector<int> merged;
  int64_t nnz = 0;
  auto deficiency = [&](int v) {
    // pairs of neighbours of v that are not adjacent yet
    const std::vector<int>& nb = adj[v];
    int64_t miss = 0;
    for (size_t a = 0; a < nb.size(); ++a)
      for (size_t b = a + 1; b < nb.size(); ++b)
        if (!std::binary_search(adj[nb[a]].begin(), adj[nb[a]].end(), nb[b])) ++miss;
    return miss;
  };
  for (size_t step = 0; step < subset.size(); ++step) {
    int best = -1;
    size_t bestdeg = (size_t)-1;
    for (int v : subset) {
      if (done[v]) continue;
      if (adj[v].size() < bestdeg) {
        bestdeg = adj[v].size();
        best = v;
      }
    }
    if (tie_break >= 1 && bestdeg > 1) {
      const size_t slack = tie_break == 2 ? 2 : 0;
      int64_t bestmiss = -1;
      size_t missdeg = 0;
      for (int v : subset) {
        if (done[v] || adj[v].size() > bestdeg + slack) continue;
        const int64_t miss = deficiency(v);
        if (bestmiss < 0 || miss < bestmiss || (miss == bestmiss && adj[v].size() < missdeg)) {
          bestmiss = miss;
          missdeg = adj[v].size();
          best = v;
          if (miss == 0 && missdeg == bestdeg) break;
        }
      }
    }
    const int v = best;
    done[v] = 1;
    order.push_back(v);
    nnz += 1 + (int64_t)adj[v].size();
    const std::vector<int> nb = adj[v];
    for (int u : nb) {
      merged.clear();
      std::set_union(adj[u].begin(), adj[u].end(), nb.begin(), nb.end(), std::back_inserter(merged));
      merged.erase(std::remove_if(merged.begin(), merged.end(), [&](int w) { return w == u || w == v; }), merged.end());
      adj[u] = merged;
    }
    adj[v].clear();
  }
  if (fill_out) *fill_out = nnz;
  return order;
}

// Geometric nested dissection: split the vertex set at the coordinate median of its longer extent,
// take the vertices of the first half that touch the second half as separator, recurse, separator last.
void order_nested(const Adj& adj, const double* xy, std::vector<int> verts, int leaf, std::vector<int>* out) {
  if ((int)verts.size() <= leaf) {
    std::vector<int> o = order_mindeg(adj, verts);
    out->insert(out->end(), o.begin(), o.end());
    return;
  }
  double lo[2] = {1e300, 1e300}, hi[2] = {-1e300, -1e300};
  for (int v : verts)
    for (int a = 0; a < 2; ++a) {
      lo[a] = std::min(lo[a], xy[2 * v + a]);
      hi[a] = std::max(hi[a], xy[2 * v + a]);
    }
  const int ax = (hi[0] - lo[0] >= hi[1] - lo[1]) ? 0 : 1;
  std::sort(verts.begin(), verts.end(), [&](int a, int b) {
    if (xy[2 * a + ax] != xy[2 * b + ax]) return xy[2 * a + ax] < xy[2 * b + ax];
    if (xy[2 * a + 1 - ax] != xy[2 * b + 1 - ax]) return xy[2 * a + 1 - ax] < xy[2 * b + 1 - ax];
    return a < b;
  });
  // cut between distinct coordinates closest to the median so that a grid line is not split
  size_t half = verts.size() / 2;
  size_t cut = half;
  {
    size_t up = half, dn = half;
    while (up < verts.size() && xy[2 * verts[up] + ax] == xy[2 * verts[up - 1] + ax]) ++up;
    while (dn > 0 && xy[2 * verts[dn] + ax] == xy[2 * verts[dn - 1] + ax]) --dn;
    if (dn == 0 && up == verts.size()) {
      cut = half;  // all on one line: plain split
    } else if (dn == 0) {
      cut = up;
    } else if (up == verts.size()) {
      cut = dn;
    } else {
      cut = (up - half <= half - dn) ? up : dn;
    }
  }
  const int n = (int)adj.size();
  std::vector<char> side(n, 0);  // 1 = first half, 2 = second half
  for (size_t i = 0; i < verts.size(); ++i) side[verts[i]] = (i < cut) ? 1 : 2;
  std::vector<int> A, Bv, S;
  for (size_t i = 0; i < verts.size(); ++i) {
    const int v = verts[i];
    if (side[v] == 2) {
      Bv.push_back(v);
      continue;
    }
    bool touches = false;
    for (int u : adj[v])
      if (side[u] == 2) {
        touches = true;
        break;
      }
    (touches ? S : A).push_back(v);
  }
  if (A.empty() || Bv.empty()) {  // degenerate split: fall back
    std::vector<int> o = order_mindeg(adj, verts);
    out->insert(out->end(), o.begin(), o.end());
    return;
  }
  // sub-problems must not see the separator or the other side: order_mindeg / recursion restrict to subsets
  order_nested(adj, xy, A, leaf, out);
  order_nested(adj, xy, Bv, leaf, out);
  std::vector<int> so = order_mindeg(adj, S);
  out->insert(out->end(), so.begin(), so.end());
}

// pattern of L for the permuted matrix; returns parent
void symbolic_factor(const Adj& padj, std::vector<std::vector<int>>* cols, std::vector<int>* parent) {
  const int n = (int)padj.size();
  cols->assign(n, {});
  parent->assign(n, -1);
  std::vector<std::vector<int>> children(n);
  std::vector<int> tmp;
  for (int j = 0; j < n; ++j) {
    std::vector<int>& s = (*cols)[j];
    for (int i : padj[j])
      if (i > j) s.push_back(i);
    for (int c : children[j]) {
      const std::vector<int>& cs = (*cols)[c];
      tmp.clear();
      // cs is sorted, first entry is j itself
      std::set_union(s.begin(), s.end(), cs.begin() + 1, cs.end(), std::back_inserter(tmp));
      s.swap(tmp);
    }
    if (!s.empty()) {
      (*parent)[j] = s[0];
      children[s[0]].push_back(j);
    }
  }
}

}  // namespace

int Symbolic::pos(int row, int col) const {
  const int* b = rowidx.data() + colptr[col];
  const int* e = rowidx.data() + colptr[col + 1];
  if (row == col) return colptr[col];
  const int* it = std::lower_bound(b + 1, e, row);
  if (it == e || *it != row) return -1;
  return (int)(it - rowidx.data());
}

void analyse(int ni, const std::vector<std::pair<int, int>>& edges, int ordering, const double* coords,
             bool build_update_lists, bool build_dot_lists, Symbolic* S, int chain_width) {
  S->ni = ni;
  Adj adj = build_adjacency(ni, edges);
  S->nnz_a = ni;
  for (auto& a : adj) S->nnz_a += (int)a.size();
  S->nnz_a = (S->nnz_a - ni) / 2 + ni;

  std::vector<int> all(ni);
  std::iota(all.begin(), all.end(), 0);
  std::vector<int> order;
  if (ordering == ORD_NATURAL) {
    order = all;
  } else if (ordering == ORD_NESTED && coords != nullptr) {
    order_nested(adj, coords, all, 12, &order);
  } else {
    // minimum degree with three tie-breaking / look-aside rules: keep the ordering with the smallest factor
    int64_t fbest = 0;
    order = order_mindeg(adj, all, 0, &fbest);
    for (int rule = 1; rule <= 2; ++rule) {
      int64_t f = 0;
      std::vector<int> o = order_mindeg(adj, all, rule, &f);
      if (f < fbest) {
        fbest = f;
        order.swap(o);
      }
    }
  }

  auto permuted_adj = [&](const std::vector<int>& ord) {
    std::vector<int> inv(ni);
    for (int r = 0; r < ni; ++r) inv[ord[r]] = r;
    Adj p(ni);
    for (int r = 0; r < ni; ++r) {
      for (int u : adj[ord[r]]) p[r].push_back(inv[u]);
      std::sort(p[r].begin(), p[r].end());
    }
    return p;
  };

  // first pass: elimination tree of the fill-reducing order, then relabel by a postorder so that
  // every subtree is a contiguous block of columns (fill is unchanged)
  std::vector<std::vector<int>> cols;
  std::vector<int> parent;
  symbolic_factor(permuted_adj(order), &cols, &parent);
  {
    std::vector<std::vector<int>> children(ni);
    std::vector<int> roots;
    for (int j = 0; j < ni; ++j) (parent[j] < 0 ? roots : children[parent[j]]).push_back(j);
    std::vector<int> post;
    post.reserve(ni);
    std::vector<std::pair<int, size_t>> stack;
    for (int r : roots) {
      stack.push_back({r, 0});
      while (!stack.empty()) {
        auto& top = stack.back();
        if (top.second < children[top.first].size()) {
          int c = children[top.first][top.second++];
          stack.push_back({c, 0});
        } else {
          post.push_back(top.first);
          stack.pop_back();
        }
      }
    }
    // ... then stable-sort the postorder by level (height above the leaves).  Children still precede
    // parents, so fill and tree are unchanged, and every level set becomes a contiguous block of columns
    // (and, row-wise, a contiguous block of rows): the level-synchronous kernels need no index lists.
    std::vector<int> lvl(ni, 0);
    for (int j = 0; j < ni; ++j)
      if (parent[j] >= 0) lvl[parent[j]] = std::max(lvl[parent[j]], lvl[j] + 1);
    std::stable_sort(post.begin(), post.end(), [&](int a, int b) { return lvl[a] < lvl[b]; });
    std::vector<int> neworder(ni);
    for (int r = 0; r < ni; ++r) neworder[r] = order[post[r]];
    order.swap(neworder);
  }
  symbolic_factor(permuted_adj(order), &cols, &parent);

  // ---- chain partition of the narrow top of the tree ------------------------------------------------
  // Levels whose width drops to <= chain_width form the "narrow region".  Inside it a chain is a maximal
  // tree path v1 -> v2 -> ... where v(i+1) = parent(v(i)) and v(i) is the ONLY child of v(i+1) inside the
  // region.  The rows of a chain only reach (a) columns outside the region or in descendant chains and
  // (b) earlier columns of the same chain, so a chain is processed as one dense triangular block.
  // Chains are renumbered contiguously, ordered by their depth in the tree of chains ("chain level").
  S->wide_levels = 0;
  S->wide_rows = ni;
  S->chain_ptr.assign(1, ni);
  S->chain_clevel.clear();
  S->clevel_ptr.assign(1, 0);
  if (chain_width > 0 && ni > 0) {
    std::vector<int> lvl(ni, 0);
    for (int j = 0; j < ni; ++j)
      if (parent[j] >= 0) lvl[parent[j]] = std::max(lvl[parent[j]], lvl[j] + 1);
    const int nlev = *std::max_element(lvl.begin(), lvl.end()) + 1;
    std::vector<int> width(nlev, 0);
    for (int j = 0; j < ni; ++j) width[lvl[j]]++;
    int Lc = nlev;
    for (int l = 1; l < nlev; ++l)
      if (width[l] <= chain_width) {
        Lc = l;
        break;
      }
    int wide_rows = 0;
    for (int l = 0; l < Lc; ++l) wide_rows += width[l];
    if (Lc < nlev) {
      // nodes are level sorted: the region is [wide_rows, ni)
      std::vector<int> nchild(ni, 0), only_child(ni, -1);
      for (int j = wide_rows; j < ni; ++j)
        if (parent[j] >= 0) {
          nchild[parent[j]]++;
          only_child[parent[j]] = j;
        }
      // a chain block holds at most chain_max rows (the on-chip kernel addresses chain rows with 6 bits)
      static const char* env_max = getenv("TNO_CHAIN_MAX");
      const int chain_max = env_max ? std::max(1, atoi(env_max)) : 64;
      std::vector<int> chain_of(ni, -1), chain_first, chain_len, clev;
      for (int j = wide_rows; j < ni; ++j) {  // ascending = children first
        if (nchild[j] == 1 && chain_len[chain_of[only_child[j]]] < chain_max) {
          const int c = chain_of[only_child[j]];
          chain_of[j] = c;
          chain_len[c]++;
        } else {
          int cl = 0;
          // chain level = 1 + max over the chains of the in-region children
          for (int c2 = wide_rows; c2 < j; ++c2)
            if (parent[c2] == j) cl = std::max(cl, clev[chain_of[c2]] + 1);
          chain_of[j] = (int)chain_first.size();
          chain_first.push_back(j);
          chain_len.push_back(1);
          clev.push_back(cl);
        }
      }
      // ---- merge short chains into their parent chain --------------------------------------------------
      // Every chain level costs the kernels a fixed number of dependent, barrier-separated steps per sweep however
      // short its chains are.  A chain of at most `merge_len` rows is therefore absorbed by the chain above it: the
      // merged block is the dense triangle over a small SUBTREE of the elimination tree instead of a path (independent
      // branches inside it are structural zeros, which the dense factorisation and the explicit inverse keep zero).
      {
        static const char* env = getenv("TNO_CHAIN_MERGE");
        const int merge_len = env ? atoi(env) : 8;
        const int nch0 = (int)chain_first.size();
        std::vector<int> clast(nch0, -1), cpar(nch0, -1), rep(nch0), glen(chain_len);
        for (int j = wide_rows; j < ni; ++j) clast[chain_of[j]] = j;
        for (int c = 0; c < nch0; ++c) cpar[c] = parent[clast[c]] >= 0 ? chain_of[parent[clast[c]]] : -1;
        std::iota(rep.begin(), rep.end(), 0);
        std::vector<int> byl(nch0);
        std::iota(byl.begin(), byl.end(), 0);
        std::stable_sort(byl.begin(), byl.end(), [&](int a, int b) { return clev[a] > clev[b]; });   // parents first
        for (int c : byl) {
          if (cpar[c] < 0 || merge_len <= 0) continue;
          const int g = rep[cpar[c]];
          if (chain_len[c] <= merge_len && glen[g] + chain_len[c] <= std::min(48, chain_max)) {
            rep[c] = g;
            glen[g] += chain_len[c];
          }
        }
        bool merged = false;
        for (int c = 0; c < nch0; ++c) merged |= rep[c] != c;
        if (merged) {
          std::vector<int> newid(nch0, -1);
          int nn = 0;
          for (int c = 0; c < nch0; ++c)
            if (rep[c] == c) newid[c] = nn++;
          for (int j = wide_rows; j < ni; ++j) chain_of[j] = newid[rep[chain_of[j]]];
          chain_first.assign(nn, -1);
          chain_len.assign(nn, 0);
          for (int j = wide_rows; j < ni; ++j) {
            if (chain_first[chain_of[j]] < 0) chain_first[chain_of[j]] = j;
            chain_len[chain_of[j]]++;
          }
          // chain levels of the merged chain tree (children have smaller node indices than their parents)
          clev.assign(nn, 0);
          for (int j = wide_rows; j < ni; ++j)
            if (parent[j] >= 0 && chain_of[parent[j]] != chain_of[j])
              clev[chain_of[parent[j]]] = std::max(clev[chain_of[parent[j]]], clev[chain_of[j]] + 1);
        }
      }
      const int nch = (int)chain_first.size();
      std::vector<int> corder(nch);
      std::iota(corder.begin(), corder.end(), 0);
      std::stable_sort(corder.begin(), corder.end(), [&](int a, int b) { return clev[a] < clev[b]; });
      // members of each chain in path order (ascending node index within a chain is path order)
      std::vector<std::vector<int>> members(nch);
      for (int j = wide_rows; j < ni; ++j) members[chain_of[j]].push_back(j);
      std::vector<int> neworder(order.begin(), order.begin() + wide_rows);
      S->chain_ptr.clear();
      S->chain_clevel.clear();
      for (int ci : corder) {
        S->chain_ptr.push_back((int)neworder.size());
        S->chain_clevel.push_back(clev[ci]);
        for (int j : members[ci]) neworder.push_back(order[j]);
      }
      S->chain_ptr.push_back((int)neworder.size());
      const int ncl = S->chain_clevel.empty() ? 0 : S->chain_clevel.back() + 1;
      S->clevel_ptr.assign(ncl + 1, 0);
      for (int cl : S->chain_clevel) S->clevel_ptr[cl + 1]++;
      for (int l = 0; l < ncl; ++l) S->clevel_ptr[l + 1] += S->clevel_ptr[l];
      S->wide_levels = Lc;
      S->wide_rows = wide_rows;
      order.swap(neworder);
      symbolic_factor(permuted_adj(order), &cols, &parent);
    }
  }

  S->perm = order;
  S->iperm.assign(ni, 0);
  for (int r = 0; r < ni; ++r) S->iperm[order[r]] = r;
  S->parent = parent;
  S->colptr.assign(ni + 1, 0);
  S->rowidx.clear();
  S->max_colcount = 0;
  S->factor_flops = 0;
  for (int j = 0; j < ni; ++j) {
    S->rowidx.push_back(j);
    S->rowidx.insert(S->rowidx.end(), cols[j].begin(), cols[j].end());
    S->colptr[j + 1] = (int)S->rowidx.size();
    const int c = (int)cols[j].size() + 1;
    S->max_colcount = std::max(S->max_colcount, c);
    S->factor_flops += (int64_t)c * c;
  }

  // level sets (height above the leaves)
  S->level.assign(ni, 0);
  for (int j = 0; j < ni; ++j)
    if (parent[j] >= 0) S->level[parent[j]] = std::max(S->level[parent[j]], S->level[j] + 1);
  S->nlevels = ni ? (*std::max_element(S->level.begin(), S->level.end()) + 1) : 0;
  S->level_ptr.assign(S->nlevels + 1, 0);
  for (int j = 0; j < ni; ++j) S->level_ptr[S->level[j] + 1]++;
  for (int l = 0; l < S->nlevels; ++l) S->level_ptr[l + 1] += S->level_ptr[l];
  S->level_cols.assign(ni, 0);
  {
    std::vector<int> fill(S->level_ptr.begin(), S->level_ptr.end() - 1);
    for (int j = 0; j < ni; ++j) S->level_cols[fill[S->level[j]]++] = j;
  }

  // row-wise view of the strictly lower part
  S->rowptr.assign(ni + 1, 0);
  for (int j = 0; j < ni; ++j)
    for (int p = S->colptr[j] + 1; p < S->colptr[j + 1]; ++p) S->rowptr[S->rowidx[p] + 1]++;
  for (int i = 0; i < ni; ++i) S->rowptr[i + 1] += S->rowptr[i];
  S->rowcol.assign(S->rowptr[ni], 0);
  S->rowpos.assign(S->rowptr[ni], 0);
  {
    std::vector<int> fill(S->rowptr.begin(), S->rowptr.end() - 1);
    for (int j = 0; j < ni; ++j)
      for (int p = S->colptr[j] + 1; p < S->colptr[j + 1]; ++p) {
        const int i = S->rowidx[p];
        S->rowcol[fill[i]] = j;
        S->rowpos[fill[i]] = p;
        fill[i]++;
      }
  }

  if (build_update_lists) {
    S->upd_ptr.assign(ni + 1, 0);
    S->upd_pos.clear();
    for (int j = 0; j < ni; ++j) {
      const int b = S->colptr[j] + 1, e = S->colptr[j + 1];
      for (int pa = b; pa < e; ++pa) {
        const int ca = S->rowidx[pa];
        // rows rowidx[pa..e) all appear in column ca (elimination theory); walk both sorted lists
        int t = S->colptr[ca];
        for (int pb = pa; pb < e; ++pb) {
          const int r = S->rowidx[pb];
          while (S->rowidx[t] != r) {
            ++t;
            if (t >= S->colptr[ca + 1]) throw std::runtime_error("symbolic: update target missing");
          }
          S->upd_pos.push_back(t);
        }
      }
      S->upd_ptr[j + 1] = (int64_t)S->upd_pos.size();
    }
  }

  if (build_dot_lists) {
    const int nnz = (int)S->rowidx.size();
    S->dot_ptr.assign(nnz + 1, 0);
    S->dot_a.clear();
    S->dot_b.clear();
    S->dot_c.clear();
    // entry e = L(i, j), i >= j:  pairs over columns c < j present in both row i and row j
    for (int j = 0; j < ni; ++j) {
      for (int p = S->colptr[j]; p < S->colptr[j + 1]; ++p) {
        const int i = S->rowidx[p];
        int a = S->rowptr[i], ae = S->rowptr[i + 1];
        int b = S->rowptr[j], be = S->rowptr[j + 1];
        while (a < ae && b < be) {
          const int ca = S->rowcol[a], cb = S->rowcol[b];
          if (ca >= j || cb >= j) break;
          if (ca == cb) {
            S->dot_a.push_back(S->rowpos[a]);
            S->dot_b.push_back(S->rowpos[b]);
            S->dot_c.push_back(ca);
            ++a;
            ++b;
          } else if (ca < cb) {
            ++a;
          } else {
            ++b;
          }
        }
        S->dot_ptr[p + 1] = (int64_t)S->dot_a.size();
      }
    }
  }
}

}  // namespace tno
