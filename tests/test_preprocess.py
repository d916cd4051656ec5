"""GPU pre-processing: B and d of a fixed-plan problem for a batch of plan geometries, against the
reference recipe problems/problems.py:516-528 (numpy pinv on the host)."""
import numpy as np
import pytest

from compas_tno_b200.problems import FixedProblem, Problem, cross_form_arrays
from compas_tno_b200 import preprocess as pre


def _cross_problem(n_div, loads=False, seed=0):
    xy, edges, fixed = cross_form_arrays(n_div)
    n = len(xy)
    xyz = np.column_stack([xy, np.zeros(n)])
    P = np.zeros((n, 3))
    P[:, 2] = -1.0
    base = Problem.from_arrays(xyz, edges, fixed, loads=P)
    fp = FixedProblem.from_problem(base)
    if loads:   # horizontal loads the diagram can take: ph in the range of E
        rng = np.random.default_rng(seed)
        fp.ph = base.E @ rng.uniform(-1.0, 0.0, size=(base.m, 1))
    return fp, xy, edges, fixed


def _reference_B_d(problem, xy, ph):
    """problems.py:516-528 on the host for the plan geometry xy, with the problem's frozen independents."""
    from scipy.sparse import diags, vstack
    u, v = problem.C @ xy[:, 0], problem.C @ xy[:, 1]
    E = vstack((problem.Citx @ diags(np.asarray(u).ravel()), problem.City @ diags(np.asarray(v).ravel()))).toarray()
    ind = list(problem.ind)
    dep = [e for e in range(problem.m) if e not in set(ind)]
    Edinv = -np.linalg.pinv(E[:, dep], rcond=1e-17)
    B = np.zeros((problem.m, len(ind)))
    B[dep] = Edinv @ E[:, ind]
    B[ind] = np.identity(len(ind))
    d = np.zeros(problem.m)
    d[dep] = -(Edinv @ np.asarray(ph).reshape(-1))
    return B, d


def test_topology_arrays_reproduce_the_connectivity():
    fp, xy, edges, fixed = _cross_problem(6)
    e2, free_index = pre.topology_arrays(fp)
    assert np.array_equal(e2, np.asarray(edges, dtype=np.int32))
    assert list(np.nonzero(free_index < 0)[0]) == sorted(fixed)
    assert np.array_equal(free_index[np.asarray(fp.free)], np.arange(fp.ni))


def test_symmetric_grid_perturbations_keep_the_rank_of_E():
    """CPU: the perturbation family of the geometry Monte-Carlo keeps the number of independents; a generic one does not."""
    fp, xy, edges, fixed = _cross_problem(6)
    XY = pre.symmetric_grid_perturbations(xy, 6, 3, sigma=0.05, seed=1)
    r0 = np.linalg.matrix_rank(fp.E, tol=1e-9)
    for g in range(3):
        assert np.abs(XY[g] - xy).max() > 1e-3
        P2 = pre.problem_with_geometry(fp, XY[g], fp.B, fp.d)
        assert np.linalg.matrix_rank(P2.E, tol=1e-9) == r0
    rng = np.random.default_rng(0)
    xyg = xy + rng.normal(0, 0.02, xy.shape)
    assert np.linalg.matrix_rank(pre.problem_with_geometry(fp, xyg, fp.B, fp.d).E, tol=1e-9) > r0


@pytest.fixture(scope="module")
def gpu():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch


@pytest.mark.gpu
@pytest.mark.parametrize("n_div,loads", [(6, False), (10, True), (14, False)])
def test_batched_B_d_matches_the_host_pseudo_inverse(gpu, n_div, loads):
    fp, xy, edges, fixed = _cross_problem(n_div, loads=loads)
    G = 6
    XY = pre.symmetric_grid_perturbations(xy, n_div, G, sigma=0.03, seed=3)
    XY[0] = xy
    for use_mma in (True, False):
        res = pre.batched_B_d(fp, XY, use_mma=use_mma)
        B, d = res.B.cpu().numpy(), res.d.cpu().numpy()
        assert res.fits().all(), (res.residual, res.rdiag_min, res.rdiag_max)
        for g in range(G):
            Bh, dh = _reference_B_d(fp, XY[g], fp.ph)
            # tolerance: 1e-9 of the largest entry (cond(E[:, dep]) ~ 1e2; measured ~1e-13)
            assert np.abs(B[g] - Bh).max() <= 1e-9 * max(1.0, np.abs(Bh).max()), (use_mma, g, np.abs(B[g] - Bh).max())
            assert np.abs(d[g] - dh).max() <= 1e-9 * max(1.0, np.abs(dh).max()), (use_mma, g, np.abs(d[g] - dh).max())
        if use_mma:
            Bm, dm = B.copy(), d.copy()
    assert np.abs(B - Bm).max() <= 1e-11 * max(1.0, np.abs(B).max()) and np.abs(d - dm).max() <= 1e-11 * max(1.0, np.abs(d).max())
    # the unperturbed geometry reproduces the problem's own B, d (made by the host recipe when the problem was built)
    assert np.abs(B[0] - fp.B).max() <= 1e-9 * np.abs(fp.B).max()


@pytest.mark.gpu
def test_a_geometry_that_breaks_the_independents_is_reported(gpu):
    fp, xy, edges, fixed = _cross_problem(10)
    rng = np.random.default_rng(5)
    XY = np.stack([xy, xy + rng.normal(0, 0.02, xy.shape)])
    res = pre.batched_B_d(fp, XY)
    assert res.fits()[0] and not res.fits()[1]
    assert res.residual[1] > 1e-4


@pytest.mark.gpu
def test_dome_fixture_with_radial_perturbation(gpu):
    """Dome fixture (data/analysis.json arrays in the golden case): the stored B, d are reproduced, and hoop radii scaled ring
    by ring keep the independents valid."""
    from golden_io import load_case
    P, objective, xs, _ = load_case("dome_q_zb_reac_min")
    xy = np.asarray(P.X, dtype=np.float64)[:, :2]
    c = xy.mean(axis=0)
    r = np.linalg.norm(xy - c, axis=1)
    rings = np.unique(np.round(r, 6))
    rng = np.random.default_rng(2)
    XY = np.stack([xy] * 3)
    for g in (1, 2):
        f = 1.0 + 0.02 * rng.uniform(-1, 1, len(rings))
        f[-1] = 1.0
        scale = f[np.searchsorted(rings, np.round(r, 6))]
        XY[g] = c + (xy - c) * scale[:, None]
    ph = np.zeros(2 * P.ni)
    res = pre.batched_B_d(P, XY, ph=ph)
    assert res.fits().all(), (res.residual, res.rdiag_min)
    B = res.B.cpu().numpy()
    assert np.abs(B[0] - P.B).max() <= 1e-8 * np.abs(P.B).max()
    for g in (1, 2):
        Bh, _ = _reference_B_d(_with_cit(P), XY[g], ph)
        assert np.abs(B[g] - Bh).max() <= 1e-8 * np.abs(Bh).max()


def _with_cit(P):
    if getattr(P, "Citx", None) is None:
        from scipy.sparse import csr_matrix
        Ct = csr_matrix(P.C).transpose().tocsr()
        P.Citx = Ct[list(P.free), :]
        P.City = Ct[list(P.free), :]
    return P


@pytest.mark.gpu
def test_evaluation_on_a_gpu_preprocessed_geometry_matches_the_oracle(gpu):
    """End of the chain: B, d made on the GPU -> plan -> batched evaluation, against the oracle on the problem whose B, d
    come from the host recipe."""
    from oracle import callbacks_np as onp
    from golden_io import rel_err
    from compas_tno_b200.plan import TnoPlan
    from compas_tno_b200 import workloads as W
    P0, x_star = W.crossvault_problem(10, n_states=4)
    obj = "min"
    xs = W.sample_candidates(P0, x_star, 8, sigma=0.02)
    xy0 = np.asarray(P0.X)[:, :2]
    XY = pre.symmetric_grid_perturbations(xy0, 10, 2, sigma=0.02, seed=9)
    res = pre.batched_B_d(_with_cit(P0), XY, ph=np.zeros(2 * P0.ni))
    assert res.fits().all()
    for g in range(2):
        Bh, dh = _reference_B_d(P0, XY[g], np.zeros(2 * P0.ni))
        Pg = pre.problem_with_geometry(P0, XY[g], res.B[g].cpu().numpy(), res.d[g].cpu().numpy())
        Ph = pre.problem_with_geometry(P0, XY[g], Bh, dh)
        plan = TnoPlan(Pg, objective=obj)
        got = plan.evaluate_host(xs)
        ref = onp.evaluate_batch(xs, Ph, obj)
        ok = got["status"] == 0
        assert ok.any()
        for key, tol in (("z", 1e-10), ("g", 1e-10), ("f", 1e-10), ("grad", 1e-8), ("J", 1e-8)):
            assert rel_err(got[key][ok], ref[key][ok]) <= tol, (g, key, rel_err(got[key][ok], ref[key][ok]))
        plan.close()
