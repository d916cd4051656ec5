"""CPU: the host-side symbolic analysis (ordering, elimination tree, pattern of L, chain partition)."""
import numpy as np
import pytest

from compas_tno_b200 import _lib
from compas_tno_b200.problems import cross_form_arrays


def _free_graph(n_div):
    xy, edges, fixed = cross_form_arrays(n_div)
    fixed = set(fixed)
    free = [v for v in range(len(xy)) if v not in fixed]
    pos = {v: i for i, v in enumerate(free)}
    e = [(pos[u], pos[v]) for u, v in edges if u not in fixed and v not in fixed]
    return len(free), np.array(e, dtype=np.int32)


def _dense_symbolic(n, edges, perm):
    """Boolean Cholesky of the permuted pattern (the definition the library must reproduce)."""
    inv = np.empty(n, int)
    inv[perm] = np.arange(n)
    A = np.eye(n, dtype=bool)
    for u, v in edges:
        A[inv[u], inv[v]] = A[inv[v], inv[u]] = True
    L = np.tril(A)
    for j in range(n):
        rows = np.nonzero(L[j + 1:, j])[0] + j + 1
        for a in rows:
            L[rows[rows >= a], a] = True
    return L


@pytest.mark.parametrize("n_div", [4, 6, 10])
@pytest.mark.parametrize("ordering,chain_width", [(2, 0), (2, 4), (1, 4), (2, 64)])
def test_pattern_matches_boolean_cholesky(n_div, ordering, chain_width):
    n, edges = _free_graph(n_div)
    S = _lib.symbolic_analyse(n, edges, ordering=ordering, chain_width=chain_width)
    perm = S["perm"]
    assert sorted(perm.tolist()) == list(range(n))                      # a permutation
    L = _dense_symbolic(n, edges, perm)
    got = np.zeros((n, n), dtype=bool)
    for j in range(n):
        rows = S["rowidx"][S["colptr"][j]:S["colptr"][j + 1]]
        assert rows[0] == j and np.all(np.diff(rows) > 0)              # diagonal first, rows ascending
        got[rows, j] = True
    assert np.array_equal(got, L)
    assert S["info"]["nnz_l"] == int(L.sum())
    # elimination tree: parent = first off-diagonal row; children precede parents
    for j in range(n):
        rows = S["rowidx"][S["colptr"][j]:S["colptr"][j + 1]]
        assert S["parent"][j] == (rows[1] if len(rows) > 1 else -1)
        assert S["parent"][j] == -1 or S["parent"][j] > j


@pytest.mark.parametrize("n_div", [6, 10, 20])
def test_level_order_and_chain_partition(n_div):
    n, edges = _free_graph(n_div)
    S = _lib.symbolic_analyse(n, edges, ordering=2, chain_width=4)
    parent, info = S["parent"], S["info"]
    level = np.zeros(n, int)
    for j in range(n):
        if parent[j] >= 0:
            level[parent[j]] = max(level[parent[j]], level[j] + 1)
    assert info["etree_height"] == level.max() + 1
    wide = info["wide_rows"]
    assert np.all(np.diff(level[:wide]) >= 0)                           # wide rows are level ordered
    assert np.all(level[:wide] < info["wide_levels"]) and np.all(level[wide:] >= info["wide_levels"])
    cp = S["chain_ptr"]
    assert cp[0] == wide and cp[-1] == n and np.all(np.diff(cp) > 0)    # chains tile the narrow region
    chain_of = np.full(n, -1)
    for c in range(len(cp) - 1):
        chain_of[cp[c]:cp[c + 1]] = c
        for r in range(cp[c], cp[c + 1] - 1):
            # a chain is a tree path, or (short chains merged into their parent chain) a small subtree: every row but
            # the last has its parent inside the same block, further on
            assert cp[c] <= r < parent[r] < cp[c + 1]
    # rows of a chain only reach outside columns below the chain or earlier columns of the same chain
    for j in range(n):
        for i in S["rowidx"][S["colptr"][j] + 1:S["colptr"][j + 1]]:
            if chain_of[i] >= 0 and chain_of[j] >= 0 and chain_of[i] != chain_of[j]:
                assert cp[chain_of[j] + 1] <= cp[chain_of[i]]          # j's chain lies entirely before i's chain


def test_fill_is_reduced_by_minimum_degree():
    n, edges = _free_graph(20)
    nat = _lib.symbolic_analyse(n, edges, ordering=1)["info"]["nnz_l"]
    md = _lib.symbolic_analyse(n, edges, ordering=2)["info"]["nnz_l"]
    assert md < 0.6 * nat and md <= 4143         # SURVEY.md section 8 quotes 4143 for SuperLU's MMD on this graph


def test_ordering_quality_at_the_survey_sizes():
    """The best-of-three minimum-degree ordering matches or beats the fill SURVEY.md section 8 measured with
    SuperLU's MMD (764 / 1738 / 4143 / 21594 at discretisation 10 / 14 / 20 / 40)."""
    for n_div, mmd in ((10, 764), (14, 1738), (20, 4143), (40, 21594)):
        n, edges = _free_graph(n_div)
        got = _lib.symbolic_analyse(n, edges, ordering=2)["info"]["nnz_l"]
        assert got <= 1.015 * mmd, (n_div, got, mmd)
