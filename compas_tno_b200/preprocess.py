"""One-time pre-processing on the GPU, batched over plan geometries.

The reference turns a form diagram into a fixed-plan problem on the host (``FixedProblem.from_formdiagram``,
``problems/problems.py:435-551``): independent edges by a pivoted QR of the equilibrium matrix
(``algorithms/independents.py:126-154``), then ``B[dep] = -pinv(E[:, dep]) E[:, ind]``, ``d[dep] = pinv(E[:, dep]) ph``
(``problems.py:516-528``).  A Monte-Carlo over the plan geometry (BASELINE configuration 4) repeats the dense
pseudo-inverse for every sample.  ``batched_B_d`` does it for G geometries of one topology in one launch sequence on the
device (``tno_preprocess_batch``: blocked Householder QR, trailing updates on the FP64 tensor pipe); the set of independent
edges stays the one of the base problem, and the least-squares residual returned per geometry tells whether it still fits
(it does for perturbations that keep the alignments of the diagram, e.g. grid-line spacings of the cross form).

torch supplies the device buffers and the stream only.  No CPU fallback.
"""
from __future__ import annotations

import copy
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib


@dataclass
class PreprocessResult:
    B: "object"          # torch (G, m, k) float64, device
    d: "object"          # torch (G, m) float64, device
    residual: np.ndarray  # (G,) largest entry of the least-squares residual of Ed X = [-Ei | ph]
    rdiag_min: np.ndarray  # (G,) smallest |r_jj| of the QR factor of E[:, dep]
    rdiag_max: np.ndarray  # (G,)
    xmax: np.ndarray     # (G,) largest |entry| of the solution

    def fits(self, tol=1e-9):
        """True where the frozen independents still describe the geometry: E[:, dep] of full rank and zero residual."""
        scale = np.maximum(self.xmax, 1.0) * np.maximum(self.rdiag_max, 1e-300)
        return (self.residual <= tol * scale) & (self.rdiag_min > 1e-12 * self.rdiag_max)


def topology_arrays(problem):
    """(edges (m, 2) int32, free_index (n,) int32) of a Problem: C[e, u] = -1, C[e, v] = +1; free vertices ascending."""
    from scipy.sparse import csr_matrix
    Cm = csr_matrix(problem.C)
    m, n = Cm.shape
    edges = np.zeros((m, 2), np.int32)
    for e in range(m):
        idx = Cm.indices[Cm.indptr[e]:Cm.indptr[e + 1]]
        val = Cm.data[Cm.indptr[e]:Cm.indptr[e + 1]]
        if len(idx) != 2 or sorted(val) != [-1.0, 1.0]:
            raise ValueError("connectivity row %d is not (-1, +1)" % e)
        edges[e, 0] = idx[int(np.argmin(val))]
        edges[e, 1] = idx[int(np.argmax(val))]
    free_index = -np.ones(n, np.int32)
    free = np.asarray(problem.free, dtype=int)
    free_index[free] = np.arange(len(free), dtype=np.int32)
    return edges, free_index


def batched_B_d(problem, XY, ph=None, use_mma=True, device=0) -> PreprocessResult:
    """B (G, m, k) and d (G, m) of ``problem``'s topology and independent edges for the plan geometries XY (G, n, 2).

    ``ph`` (2 ni,) or (G, 2 ni): horizontal loads [px_free ; py_free]; default: the problem's own ``ph``."""
    import torch
    lib = _lib.load()
    dev = torch.device("cuda", device)
    edges, free_index = topology_arrays(problem)
    m, n = edges.shape[0], len(free_index)
    ni = int((free_index >= 0).sum())
    ind = np.ascontiguousarray(np.asarray(problem.ind, dtype=np.int32).reshape(-1))
    k = len(ind)
    XYt = XY if isinstance(XY, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(XY, dtype=np.float64)))
    XYt = XYt.to(dev, dtype=torch.float64).contiguous()
    if XYt.dim() != 3 or XYt.shape[1] != n or XYt.shape[2] != 2:
        raise ValueError("XY must be (G, %d, 2)" % n)
    G = XYt.shape[0]
    if ph is None:
        ph = np.asarray(problem.ph, dtype=np.float64).reshape(-1)
    pht = ph if isinstance(ph, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(ph, dtype=np.float64)))
    pht = pht.to(dev, dtype=torch.float64).contiguous()
    if pht.numel() == 2 * ni:
        ph_stride = 0
    elif pht.numel() == G * 2 * ni:
        ph_stride = 2 * ni
    else:
        raise ValueError("ph must have 2 ni or G x 2 ni entries")
    B = torch.empty((G, m, k), dtype=torch.float64, device=dev)
    d = torch.empty((G, m), dtype=torch.float64, device=dev)
    info = torch.empty((G, 4), dtype=torch.float64, device=dev)
    nbytes = int(lib.tno_preprocess_work_bytes(n, m, ni, k, G))
    if nbytes < 0:
        raise ValueError("bad sizes")
    work = torch.empty((nbytes + 7) // 8, dtype=torch.float64, device=dev)
    p = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))  # noqa: E731
    with torch.cuda.device(dev):
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(lib.tno_preprocess_batch(n, m, ni, k, p(np.ascontiguousarray(edges)), p(free_index), p(ind),
                                            C.c_void_p(XYt.data_ptr()), C.c_void_p(pht.data_ptr()), ph_stride, G,
                                            C.c_void_p(B.data_ptr()), C.c_void_p(d.data_ptr()), C.c_void_p(info.data_ptr()),
                                            C.c_void_p(work.data_ptr()), nbytes, 1 if use_mma else 0, stream))
    ih = info.cpu().numpy()
    del work
    return PreprocessResult(B=B, d=d, residual=ih[:, 0].copy(), rdiag_min=ih[:, 1].copy(), rdiag_max=ih[:, 2].copy(),
                            xmax=ih[:, 3].copy())


def problem_with_geometry(problem, xy, B, d):
    """A copy of ``problem`` on the plan geometry ``xy`` (n, 2) with the given B (m, k) and d (m,): the fields the
    evaluation reads (X, x0 / y0, U, V, E, B, d, d0) follow the new coordinates, everything else is shared."""
    from scipy.sparse import diags, vstack
    P = copy.copy(problem)
    xy = np.asarray(xy, dtype=np.float64)
    X = np.array(problem.X, dtype=np.float64, copy=True)
    X[:, :2] = xy
    P.X = X
    if getattr(problem, "x0", None) is not None and np.asarray(problem.x0).size == X.shape[0]:
        P.x0 = xy[:, 0].reshape(-1, 1).copy()
        P.y0 = xy[:, 1].reshape(-1, 1).copy()
    Cm = problem.C
    u, v = Cm @ xy[:, 0], Cm @ xy[:, 1]
    P.U, P.V = diags(np.asarray(u).ravel()), diags(np.asarray(v).ravel())
    if getattr(problem, "Citx", None) is not None:
        P.E = vstack((problem.Citx @ P.U, problem.City @ P.V)).toarray()
    P.B = np.array(B, dtype=np.float64, copy=True)
    P.d = np.array(d, dtype=np.float64, copy=True).reshape(-1, 1)
    P.d0 = P.d.copy()
    return P


def symmetric_grid_perturbations(xy, n_div, G, sigma=0.02, seed=0, span=10.0):
    """G plan geometries of the cross form diagram whose grid-line positions are perturbed symmetrically about both axes
    and equally in x and y (a construction tolerance of the spacing): the diagonals stay straight and the diagram keeps its
    degree of static indeterminacy, so the independent edges of the regular grid remain valid."""
    xy = np.asarray(xy, dtype=np.float64)
    h = span / n_div
    ix = np.rint(xy[:, 0] / h).astype(int)
    iy = np.rint(xy[:, 1] / h).astype(int)
    g = np.arange(n_div + 1) * h
    rng = np.random.default_rng(seed)
    out = np.empty((G,) + xy.shape)
    for s in range(G):
        dl = rng.normal(0.0, sigma * h, n_div + 1)
        dl[0] = dl[-1] = 0.0
        dl = 0.5 * (dl - dl[::-1])
        gp = g + dl
        out[s, :, 0] = gp[ix]
        out[s, :, 1] = gp[iy]
    return out
