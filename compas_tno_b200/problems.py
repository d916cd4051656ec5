"""Host-side carrier types for the batched hot path.

``Problem`` / ``FixedProblem`` mirror the attribute names of the reference's
dataclasses (reference: src/compas_tno/problems/problems.py:39-224, :435-551) so
that (a) an object built by the reference itself can be handed to
``compas_tno_b200.TnoPlan`` unchanged (duck typing on the same attribute names)
and (b) a problem can be built without ``compas_tna`` from plain arrays.

Construction is one-time host pre-processing (numpy / scipy / LAPACK) and is not
part of the accelerated path.  The arithmetic of every evaluation happens in
the CUDA library behind ``include/tno_b200.h``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Callable, List, Optional, Sequence

import numpy as np
from scipy.linalg import qr as _pivoted_qr
from scipy.sparse import csr_matrix, diags
from scipy.sparse import vstack as _svstack

__all__ = [
    "Problem",
    "FixedProblem",
    "connectivity_matrix",
    "find_independents_QR",
    "cross_form_arrays",
]


def connectivity_matrix(edges: Sequence[Sequence[int]], n: Optional[int] = None) -> csr_matrix:
    """Edge-vertex incidence matrix: row e = (u, v) holds -1 at u and +1 at v.

    Restates ``compas.matrices.connectivity_matrix`` (not vendored in the reference;
    call site src/compas_tno/problems/problems.py:314).
    """
    e = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
    m = e.shape[0]
    if n is None:
        n = int(e.max()) + 1 if m else 0
    rows = np.repeat(np.arange(m), 2)
    cols = e.reshape(-1)
    vals = np.tile(np.array([-1.0, 1.0]), m)
    return csr_matrix((vals, (rows, cols)), shape=(m, n))


def find_independents_QR(E: np.ndarray, tol: Optional[float] = None) -> List[int]:
    """Rank-revealing pivoted QR pick of the independent edges.

    Same rule as src/compas_tno/algorithms/independents.py:126-154: the columns the
    pivoting pushes past the numerical rank are the independents.  The choice is
    LAPACK tie-break dependent, so parity is defined *given* ``ind``.
    """
    if not tol:
        tol = 1e-12
    _, R, piv = _pivoted_qr(E, pivoting=True)
    dr = np.abs(np.diag(R))
    rank = int(np.sum(dr > tol * dr[0]))
    return sorted(int(c) for c in piv[rank:])


@dataclass
class Problem:
    """All matrices of one form diagram.  Field names follow the reference."""

    q: Optional[np.ndarray] = None
    m: Optional[int] = None
    n: Optional[int] = None
    ni: Optional[int] = None
    nb: Optional[int] = None
    E: Optional[np.ndarray] = None
    C: Any = None
    Ct: Any = None
    Ci: Any = None
    Cit: Any = None
    Cb: Any = None
    U: Any = None
    V: Any = None
    W: Any = None
    P: Optional[np.ndarray] = None
    free: List[int] = field(default_factory=list)
    fixed: List[int] = field(default_factory=list)
    ph: Optional[np.ndarray] = None
    lb: Optional[np.ndarray] = None
    ub: Optional[np.ndarray] = None
    lb0: Optional[np.ndarray] = None
    ub0: Optional[np.ndarray] = None
    s: Optional[np.ndarray] = None
    X: Optional[np.ndarray] = None
    y0: Optional[np.ndarray] = None
    free_x: List[int] = field(default_factory=list)
    free_y: List[int] = field(default_factory=list)
    Citx: Any = None
    City: Any = None
    qmin: Optional[np.ndarray] = None
    qmax: Optional[np.ndarray] = None
    ind: List[int] = field(default_factory=list)
    k: Optional[int] = None
    dep: List[int] = field(default_factory=list)
    B: Optional[np.ndarray] = None
    d: Optional[np.ndarray] = None
    d0: Optional[np.ndarray] = None
    variables: List[str] = field(default_factory=list)
    features: List[str] = field(default_factory=list)
    constraints: List[str] = field(default_factory=list)
    thk: Optional[float] = None
    rho: Optional[float] = None
    min_lb: float = 0.0
    envelope: Any = None
    b: Optional[np.ndarray] = None
    px0: Optional[np.ndarray] = None
    py0: Optional[np.ndarray] = None
    pzv: Optional[np.ndarray] = None
    pz0: Optional[np.ndarray] = None
    Edinv: Any = None
    Ed: Any = None
    Ei: Any = None
    # callables + start vector, written by the set-up (reference setup.py:466-473)
    fobj: Optional[Callable] = None
    fgrad: Optional[Callable] = None
    fconstr: Optional[Callable] = None
    fjac: Optional[Callable] = None
    x0: Optional[np.ndarray] = None
    bounds: Optional[list] = None
    f0: Optional[float] = None
    g0: Optional[np.ndarray] = None

    @classmethod
    def from_arrays(
        cls,
        xyz,
        edges,
        fixed,
        loads=None,
        lb=None,
        ub=None,
        target=None,
        q=None,
        qmin: float = -1e4,
        qmax: float = 1e-8,
    ) -> "Problem":
        """Build the base problem from plain arrays.

        Follows src/compas_tno/problems/problems.py:226-392 step by step (vertex order
        = row order of ``xyz``, edge order = row order of ``edges``, ``free`` ascending,
        ``fixed`` in the order given = ``form.supports()`` order).
        """
        xyz = np.array(xyz, dtype=np.float64).reshape(-1, 3)
        edges = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
        n = xyz.shape[0]
        m = edges.shape[0]
        fixed = [int(i) for i in fixed]
        nb = len(fixed)
        fixed_set = set(fixed)
        free = [i for i in range(n) if i not in fixed_set]
        ni = len(free)

        P = np.zeros((n, 3)) if loads is None else np.array(loads, dtype=np.float64).reshape(n, 3)
        col = lambda a, fill: (np.full((n, 1), fill) if a is None else np.array(a, dtype=np.float64).reshape(n, 1))  # noqa: E731
        lb_ = col(lb, -1e30)
        ub_ = col(ub, +1e30)
        s_ = col(target, 0.0)
        s_[np.abs(s_) < 1e-6] = 0.0  # problems.py:294-295
        q_ = -np.ones((m, 1)) if q is None else np.array(q, dtype=np.float64).reshape(m, 1)

        C = connectivity_matrix(edges, n)
        Ci = C[:, free]
        Cb = C[:, fixed]
        Ct = C.transpose().tocsr()
        Cit = Ci.transpose().tocsr()
        uvw = C @ xyz
        U = diags(uvw[:, 0])
        V = diags(uvw[:, 1])
        Citx = Ct[free, :]
        City = Ct[free, :]
        E = _svstack((Citx @ U, City @ V)).toarray()

        x = xyz[:, [0]].copy()
        y = xyz[:, [1]].copy()
        prob = cls(
            q=q_, m=m, n=n, ni=ni, nb=nb, E=E, C=C, Ct=Ct, Ci=Ci, Cit=Cit, Cb=Cb, U=U, V=V,
            P=P, free=free, fixed=fixed, ph=np.vstack([P[free, 0:1], P[free, 1:2]]),
            lb=lb_, ub=ub_, lb0=lb_, ub0=ub_, s=s_, X=xyz.copy(), y0=y,
            free_x=list(free), free_y=list(free), Citx=Citx, City=City,
            qmin=np.full((m, 1), float(qmin)), qmax=np.full((m, 1), float(qmax)),
            ind=list(range(m)), k=m, dep=[], B=np.identity(m) if m <= 4096 else None,
            d=np.zeros((m, 1)),
        )
        prob.x0 = x  # plan x-coordinates until the set-up overwrites it
        return prob


@dataclass
class FixedProblem(Problem):
    """Fixed-plan problem: q = B q_ind + d (reference problems.py:435-551)."""

    @classmethod
    def from_arrays(cls, xyz, edges, fixed, loads=None, lb=None, ub=None, target=None, q=None,
                    qmin: float = -1e4, qmax: float = 1e-8, ind=None, tol=None) -> "FixedProblem":
        base = Problem.from_arrays(xyz, edges, fixed, loads, lb, ub, target, q, qmin, qmax)
        return cls.from_problem(base, ind=ind, tol=tol)

    @classmethod
    def from_problem(cls, base: Problem, ind=None, tol=None) -> "FixedProblem":
        m = base.m
        if ind is None:
            ind = find_independents_QR(base.E, tol=tol)
        ind = [int(i) for i in ind]
        k = len(ind)
        ind_set = set(ind)
        dep = [e for e in range(m) if e not in ind_set]
        # problems.py:516-528 : Edinv = -pinv(E[:, dep]); B[dep] = Edinv E[:, ind]; d[dep] = -Edinv ph
        rcond = tol if tol else 1e-17
        Ed = base.E[:, dep]
        Edinv = -csr_matrix(np.linalg.pinv(Ed, rcond=rcond))
        Ei = csr_matrix(base.E[:, ind])
        B = np.zeros((m, k))
        B[dep] = (Edinv @ Ei).toarray()
        B[ind] = np.identity(k)
        d = np.zeros((m, 1))
        d[dep] = -(Edinv @ base.ph)
        kw = {f: getattr(base, f) for f in base.__dataclass_fields__}
        kw.update(ind=ind, k=k, dep=dep, B=B, d=d, d0=d.copy(), Edinv=Edinv, Ed=Ed, Ei=Ei)
        return cls(**kw)


def cross_form_arrays(n_div: int, span: float = 10.0):
    """Plan geometry of the orthogonal cross form diagram (``FormDiagram.create_cross``).

    ``compas_tna`` is not vendored in the reference; this generator follows the
    reference fixture ``data/crossvault_form.json`` (validated identical for n = 14 in
    tests/test_oracle.py): an (n+1) x (n+1) grid of vertices with the grid edges
    plus the two corner-to-corner diagonals, supports at the four corners.

    Returns (xy (nv, 2), edges (m, 2) int, fixed list).  Vertex numbering is row-major
    in (iy, ix); the fixture's numbering differs, only the *sets* are compared.
    """
    n = int(n_div)
    h = span / n
    idx = lambda ix, iy: iy * (n + 1) + ix  # noqa: E731
    xy = np.array([[ix * h, iy * h] for iy in range(n + 1) for ix in range(n + 1)], dtype=np.float64)
    edges = []
    for iy in range(n + 1):
        for ix in range(n + 1):
            if ix < n:
                edges.append((idx(ix, iy), idx(ix + 1, iy)))
            if iy < n:
                edges.append((idx(ix, iy), idx(ix, iy + 1)))
    for i in range(n):
        edges.append((idx(i, i), idx(i + 1, i + 1)))
        edges.append((idx(i, n - i), idx(i + 1, n - i - 1)))
    fixed = [idx(0, 0), idx(n, 0), idx(0, n), idx(n, n)]
    return xy, np.array(edges, dtype=np.int64), fixed
