// Symbolic analysis of  -(Ci^T diag(q) Ci)  for one form diagram (host, once per plan).
//
// The sparsity pattern is the free-free adjacency of the form diagram and is identical for every
// candidate q, so ordering, elimination tree, pattern of L, level sets and all scatter / update
// index lists are computed once here and then only *values* change per candidate.
// Reference counterpart: scipy.sparse.linalg.splu inside xyz_from_q
// (src/compas_tno/algorithms/equilibrium.py:230-234) redoes all of this on every call.
#pragma once
#include <cstdint>
#include <vector>

namespace tno {

enum Ordering { ORD_AUTO = 0, ORD_NATURAL = 1, ORD_MINDEG = 2, ORD_NESTED = 3 };

struct Symbolic {
  int ni = 0;                    // order of the matrix (free vertices)
  std::vector<int> perm;         // perm[row]  = free position (0..ni-1, ascending vertex order) placed at that row
  std::vector<int> iperm;        // iperm[free position] = row
  std::vector<int> colptr;       // ni+1, CSC of L, diagonal first in every column, rows ascending
  std::vector<int> rowidx;       // nnz_l
  std::vector<int> parent;       // elimination tree (-1 = root)
  std::vector<int> level;        // height above the leaves; columns of one level are independent
  std::vector<int> level_ptr;    // nlevels+1
  std::vector<int> level_cols;   // columns sorted by level
  int nlevels = 0;
  // chain partition of the narrow top of the tree (see tno_symbolic.cpp).  Rows [0, wide_rows) belong to the
  // wide levels [0, wide_levels) and are level ordered; rows of chain c are [chain_ptr[c], chain_ptr[c+1]),
  // chains are sorted by chain level, clevel_ptr indexes chains.
  int wide_levels = 0, wide_rows = 0;
  std::vector<int> chain_ptr, chain_clevel, clevel_ptr;
  int nnz_a = 0;                 // lower triangle incl. diagonal
  int max_colcount = 0;
  int64_t factor_flops = 0;      // sum_j c_j^2 (c_j = off-diagonal count + 1)
  // right-looking update targets: for column j with off-diagonals p1..pc (positions colptr[j]+1..),
  // for a in 1..c, b in a..c:  L[upd_pos[t++]] -= L[p_b] * L[p_a]
  std::vector<int64_t> upd_ptr;  // ni+1
  std::vector<int> upd_pos;
  // pull ("dot product") form used by the on-chip kernel: for every entry e of L (position order),
  // pairs (pa, pb) with  L[e] = A[e] - sum L[pa] * L[pb]
  std::vector<int64_t> dot_ptr;  // nnz_l+1
  std::vector<int> dot_a, dot_b, dot_c;   // dot_c: the common column of the pair
  // row-wise view of L (CSR, strictly lower part): for the backward / pull-form forward solve
  std::vector<int> rowptr;       // ni+1
  std::vector<int> rowcol;       // column of each strictly-lower entry, ascending
  std::vector<int> rowpos;       // its position in the CSC arrays

  int pos(int row, int col) const;  // position of L(row, col) in the CSC arrays, -1 if structurally zero
};

// adjacency given as an edge list between free positions (duplicates and self loops tolerated).
// coords (2 per free position) are only used by the nested-dissection ordering and may be null.
void analyse(int ni, const std::vector<std::pair<int, int>>& edges, int ordering, const double* coords,
             bool build_update_lists, bool build_dot_lists, Symbolic* out, int chain_width = 8);

}  // namespace tno
