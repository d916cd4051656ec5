"""Batched multi-start driver: every start reaches the optimum scipy's SLSQP finds when it is driven by
the oracle callbacks (the reference's solvers/scipy.py:172 call), and the two QP kernel variants agree."""
import numpy as np
import pytest

from golden_io import load_case


@pytest.fixture(scope="module")
def tno():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from compas_tno_b200 import _lib
    return _lib.load()


def _bounds(P):
    from compas_tno_b200.multistart import variable_bounds
    return variable_bounds(P)


def test_variable_bounds_follow_the_reference_layout():
    """CPU: q_ind in [qmin, qmax], zb in [lb, ub] at the supports (problems/setup.py:168-208)."""
    P, objective, xs, _ = load_case("cross6_q_zb_min")
    lb, ub = _bounds(P)
    assert lb.shape == ub.shape == (xs.shape[1],)
    k = P.k
    assert np.all(lb[:k] == np.asarray(P.qmin).reshape(-1)[np.asarray(P.ind)]) and np.all(ub[:k] <= 1e-6)
    fixed = np.asarray(P.fixed, dtype=int)
    assert np.allclose(lb[k:], np.asarray(P.lb).reshape(-1)[fixed]) and np.allclose(ub[k:], np.asarray(P.ub).reshape(-1)[fixed])
    assert np.all(lb < ub)


def _slsqp_with_oracle(P, objective, x0, lb, ub, iters=500):
    from scipy.optimize import fmin_slsqp
    from oracle import callbacks_np as onp
    M = onp.clone_problem(P)
    fo, fg = onp.objective_selector(objective)

    def both(fn):
        return lambda x: (onp.constr_wrapper(x, M), fn(x, M))[1]
    out = fmin_slsqp(both(fo), x0, f_ieqcons=lambda x: onp.constr_wrapper(x, M).ravel(),
                     fprime=lambda x: np.asarray(both(fg)(x)).ravel(),
                     fprime_ieqcons=lambda x: np.asarray(onp.sensitivities_wrapper(x, M)), bounds=list(zip(lb, ub)),
                     iter=iters, iprint=0, full_output=True, acc=1e-9)
    assert out[3] == 0, out[4]
    return out[0], out[1]


@pytest.mark.gpu
@pytest.mark.parametrize("case,n_starts,spread", [("cross6_q_zb_min", 96, 0.10), ("cross14_q_zb_min", 64, 0.05)])
def test_every_start_converges_to_the_slsqp_optimum(tno, case, n_starts, spread):
    from compas_tno_b200.plan import TnoPlan
    from compas_tno_b200.multistart import MultiStartSolver
    P, objective, xs, _ = load_case(case)
    lb, ub = _bounds(P)
    x_ref, f_ref = _slsqp_with_oracle(P, objective, xs[0].copy(), lb, ub)
    rng = np.random.default_rng(11)
    X0 = xs[0] * (1.0 + spread * rng.uniform(-1, 1, size=(n_starts, xs.shape[1])))
    X0[0] = xs[0]
    plan = TnoPlan(P, objective=objective)
    res = MultiStartSolver(plan, lb, ub, acc=1e-8, max_iter=200).solve(X0)
    assert res.success.all(), "starts that did not converge: %s" % np.nonzero(~res.success)[0]
    assert np.all(res.gmin >= -1e-6)                                      # feasible
    assert np.max(np.abs(res.f - f_ref)) <= 1e-6 * max(1.0, abs(f_ref))   # same optimum, every start
    assert res.best >= 0 and abs(res.fopt - f_ref) <= 1e-6 * max(1.0, abs(f_ref))
    # the tensor-pipe variant of the QP kernel takes the same path
    res2 = MultiStartSolver(plan, lb, ub, acc=1e-8, max_iter=200, use_mma=True).solve(X0[:16])
    assert res2.success.all()
    assert np.max(np.abs(res2.f - f_ref)) <= 1e-6 * max(1.0, abs(f_ref))


@pytest.mark.gpu
def test_qp_kernel_against_a_host_interior_point_solution(tno):
    """One QP of the driver's form, random but feasible, against scipy's SLSQP on the same QP."""
    import ctypes as C
    import torch
    from scipy.optimize import minimize
    from compas_tno_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(5)
    S, nv, ng = 5, 7, 40
    dev = torch.device("cuda", 0)
    Hs, cs, As, bs, xs_ = [], [], [], [], []
    for s in range(S):
        Q = rng.normal(size=(nv, nv))
        Hs.append(Q @ Q.T + nv * np.eye(nv))
        cs.append(rng.normal(size=nv) * 3)
        As.append(rng.normal(size=(ng, nv)))
        bs.append(rng.uniform(0.1, 2.0, size=ng))
        xs_.append(rng.uniform(-0.5, 0.5, size=nv))
    lb, ub = -np.ones(nv), np.ones(nv)
    tt = lambda a: torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float64))).to(dev)  # noqa: E731
    H, c, A, b, x = tt(Hs), tt(cs), tt(As), tt(bs), tt(xs_)
    mt = ng + 2 * nv
    d, lam = torch.empty((S, nv), dtype=torch.float64, device=dev), torch.empty((S, mt), dtype=torch.float64, device=dev)
    info, vold = torch.empty((S, 8), dtype=torch.float64, device=dev), torch.empty((S, nv), dtype=torch.float64, device=dev)
    work = torch.empty((S, 4, mt), dtype=torch.float64, device=dev)
    p = lambda t_: C.c_void_p(t_.data_ptr())  # noqa: E731
    lbd, ubd = tt(lb), tt(ub)
    for use_mma in (0, 1):
        _lib.check(lib.tno_sqp_qp_solve(S, nv, ng, 0, p(H), p(c), p(b), p(A), None, p(x), p(lbd), p(ubd), None, p(d), p(lam),
                                         p(info), p(vold), p(work), 60, 1e-11, use_mma, None))
        torch.cuda.synchronize()
        dn, ln, inf = d.cpu().numpy(), lam.cpu().numpy(), info.cpu().numpy()
        assert np.all(inf[:, 5] == 1.0), inf
        for s in range(S):
            cons = [{"type": "ineq", "fun": lambda v, s=s: As[s] @ v + bs[s], "jac": lambda v, s=s: As[s]}]
            ref = minimize(lambda v, s=s: 0.5 * v @ Hs[s] @ v + cs[s] @ v, np.zeros(nv), jac=lambda v, s=s: Hs[s] @ v + cs[s],
                           constraints=cons, bounds=list(zip(lb - xs_[s], ub - xs_[s])), method="SLSQP",
                           options=dict(ftol=1e-14, maxiter=500))
            assert ref.success
            assert np.max(np.abs(dn[s] - ref.x)) < 1e-6
            # KKT: stationarity with the returned multipliers, complementarity
            Aall = np.vstack([As[s], np.eye(nv), -np.eye(nv)])
            ball = np.concatenate([bs[s], xs_[s] - lb, ub - xs_[s]])
            assert np.max(np.abs(Hs[s] @ dn[s] + cs[s] - Aall.T @ ln[s])) < 1e-7
            assert np.all(Aall @ dn[s] + ball > -1e-8) and np.all(ln[s] >= 0)
            assert np.max(np.abs(ln[s] * (Aall @ dn[s] + ball))) < 1e-6
            assert np.allclose(vold.cpu().numpy()[s], cs[s] - As[s].T @ ln[s][:ng], atol=1e-9)
