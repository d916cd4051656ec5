"""GPU: parity cases added in round 2 (VERDICT r1 "what's weak" 1-4).

* hundreds of oracle-checked candidates of the headline workload, spread over the batch: first and last (ragged) wave of
  the persistent on-chip grid, across the chunk boundaries of the interleaved family;
* eight oracle-checked candidates at discretisation 40;
* candidates the kernels FLAG (status != 0) still carry the reference's numbers: the reference's pivoted LU returns a
  solution for an indefinite but non-singular matrix, the un-pivoted L D L^T returns the same one up to pivot growth;
* pz_scale together with the lambdv variable; envelopexy without the xyb variable, gradient only.

Everything goes through the C ABI (TnoPlan -> tno_eval_batch / tno_eval_batch_host); block-wise and element-wise checks
from golden_io.
"""
import numpy as np
import pytest

from golden_io import assert_blockwise, elem_err, load_case, rel_err

pytestmark = pytest.mark.gpu

TOL_VAL = 1e-10
TOL_JAC = 1e-8


@pytest.fixture(scope="module")
def tno():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import compas_tno_b200 as pkg
    return pkg


@pytest.fixture(scope="module")
def disc20():
    """Headline workload + 264 oracle evaluations at indices spread over a ragged batch (computed once, ~10 s)."""
    from compas_tno_b200 import workloads as W
    from oracle import callbacks_np as onp
    P, xs = W.crossvault_problem(20, n_states=6)
    N = 2 * 296 * 8 + 77                      # not a multiple of the 296-CTA persistent grid, nor of 128 / 1024
    X = W.sample_candidates(P, xs, N, seed=20261020, basis=P.candidate_basis)
    rng = np.random.default_rng(3)
    idx = np.unique(np.concatenate([
        np.arange(0, 64),                      # first wave of the persistent grid
        np.arange(N - 96, N),                  # last, ragged wave
        np.arange(1024 - 12, 1024 + 12),       # chunk boundaries of the interleaved family (TNO_CHUNK_MAX = 1024 below)
        np.arange(4096 - 12, 4096 + 12),
        rng.choice(N, 64, replace=False)]))
    ref = onp.evaluate_batch(X[idx], P, "min")
    return P, X, idx, ref


@pytest.mark.parametrize("kernel", ["onchip", "interleaved"])
def test_disc20_oracle_candidates_spread_over_a_ragged_batch(tno, disc20, kernel, monkeypatch):
    import torch
    from compas_tno_b200.plan import TnoPlan
    P, X, idx, ref = disc20
    assert len(idx) >= 256
    if kernel == "interleaved":
        monkeypatch.setenv("TNO_CHUNK_MAX", "1024")   # 4 813 candidates = 4 full chunks + a ragged one
    plan = TnoPlan(P, objective="min", kernel=kernel)
    res = plan.evaluate_device(torch.from_numpy(X).cuda())
    torch.cuda.synchronize()
    if kernel == "interleaved":
        assert plan.info["chunk"] == 1024
    sel = torch.from_numpy(idx).cuda()
    got = {k: res[k][sel].cpu().numpy() for k in ("f", "grad", "g", "J", "z", "q", "gmin", "status")}
    assert np.all(got["status"] == 0)
    assert rel_err(got["q"], ref["q"]) < TOL_VAL
    assert rel_err(got["z"], ref["z"]) < TOL_VAL and elem_err(got["z"], ref["z"]) < 10 * TOL_VAL
    assert rel_err(got["f"], ref["f"]) < TOL_VAL
    assert np.max(np.abs(got["f"] - ref["f"]) / np.abs(ref["f"])) < TOL_VAL          # per candidate
    assert rel_err(got["grad"], ref["grad"]) < TOL_JAC
    for j in range(0, len(idx), 16):                                                   # per candidate, block by block
        assert_blockwise(got["g"][j], ref["g"][j], got["J"][j], ref["J"][j], P, TOL_VAL, TOL_JAC, "candidate %d" % idx[j])
    assert_blockwise(got["g"], ref["g"], got["J"], ref["J"], P, TOL_VAL, TOL_JAC, "all")
    assert np.allclose(got["gmin"], ref["g"].min(axis=1), rtol=1e-9, atol=1e-9)
    plan.close()


def test_disc40_eight_oracle_candidates(tno):
    """Discretisation 40 (BASELINE configuration 4 geometry), eight candidates incl. the ends of a ragged batch."""
    import torch
    from compas_tno_b200 import workloads as W
    from compas_tno_b200.plan import TnoPlan
    from oracle import callbacks_np as onp
    P, xstar = W.crossvault_problem(40, n_states=3)
    N = 333
    X = W.sample_candidates(P, xstar, N, seed=8, basis=P.candidate_basis)
    idx = np.array([0, 1, 31, 32, 127, 128, N - 2, N - 1])
    plan = TnoPlan(P, objective="min")
    res = plan.evaluate_device(torch.from_numpy(X).cuda())
    torch.cuda.synchronize()
    sel = torch.from_numpy(idx).cuda()
    got = {k: res[k][sel].cpu().numpy() for k in ("f", "grad", "g", "J", "z", "status")}
    ref = onp.evaluate_batch(X[idx], P, "min")
    assert np.all(got["status"] == 0)
    assert rel_err(got["z"], ref["z"]) < TOL_VAL
    assert np.max(np.abs(got["f"] - ref["f"]) / np.abs(ref["f"])) < TOL_VAL
    assert rel_err(got["grad"], ref["grad"]) < TOL_JAC
    assert_blockwise(got["g"], ref["g"], got["J"], ref["J"], P, TOL_VAL, TOL_JAC, "disc40")


def _min_pivot(P, q, perm=None):
    """Smallest pivot of the un-pivoted L D L^T of -(Ci^T Q Ci) in natural order (dense, test sizes only) and the
    smallest eigenvalue: how far from positive definite the candidate is."""
    from scipy.sparse import diags
    A = -(P.Cit @ diags(q) @ P.Ci).toarray()
    return float(np.linalg.eigvalsh(A).min()), float(np.linalg.cond(A))


@pytest.mark.parametrize("kernel", ["onchip", "interleaved", "interleaved_tpc"])
def test_flagged_candidates_still_match_the_oracle(tno, kernel):
    """status != 0 is a warning, not a different result: the reference's LU runs on an indefinite matrix too.  The flagged
    candidates' z, g, f are compared with the oracle at a tolerance scaled by the conditioning of their matrix."""
    from compas_tno_b200.plan import TnoPlan
    from oracle import callbacks_np as onp
    P, objective, xs, _ = load_case("cross14_q_zb_min")
    rng = np.random.default_rng(41)
    k = P.k
    X = np.repeat(xs[:1], 24, axis=0) * (1.0 + 0.02 * rng.uniform(-1, 1, size=(24, xs.shape[1])))
    for j in range(0, 24, 2):                 # every other candidate: one or two independent edges pushed into tension
        cols = rng.choice(k, size=1 + (j // 2) % 2, replace=False)
        X[j, cols] = np.abs(X[j, cols]) * rng.uniform(0.2, 1.5)
    X[22, :k] *= -1.0                          # tension everywhere: the matrix is negative definite, every pivot < 0
    plan = TnoPlan(P, objective=objective, kernel=kernel)
    got = plan.evaluate_host(X)
    flagged = np.nonzero(got["status"] & 1)[0]
    assert len(flagged) >= 6 and 22 in flagged, got["status"]
    assert np.all(got["status"][1::2] == 0)
    checked = 0
    for j in flagged:
        ref = onp.evaluate(X[j], onp.clone_problem(P), objective)
        mineig, cond = _min_pivot(P, ref["q"])
        assert mineig < 0 or cond > 1e10, (j, mineig)          # the flag is right: not positive definite
        if not np.all(np.isfinite(ref["z"])) or cond > 1e9:
            continue                                            # singular to working precision: nothing to compare
        tol = max(1e-9, 1e-15 * cond * 1e3)                     # pivot growth of the un-pivoted factorisation
        ez, eg = rel_err(got["z"][j], ref["z"]), rel_err(got["g"][j], ref["g"])
        ef = abs(got["f"][j] - ref["f"]) / max(1.0, abs(ref["f"]))
        print("flagged candidate %d: min eig %.3e cond %.2e  err z %.1e g %.1e f %.1e J %.1e (kernel %s)" % (
            j, mineig, cond, ez, eg, ef, rel_err(got["J"][j], ref["J"]), kernel))
        assert ez < tol and eg < tol and ef < tol, (j, ez, eg, ef, tol)
        assert rel_err(got["J"][j], ref["J"]) < 100 * tol
        checked += 1
    assert checked >= 4
    plan.close()


@pytest.mark.parametrize("kernel", ["interleaved", "onchip"])
def test_pz_scale_with_the_lambdv_variable(tno, kernel):
    """pz = pz_scale * (lambdv * pzv + pz0), so d pz / d lambdv = pz_scale * pzv.  The lambdv column of J
    (envelope and reac_bounds rows) must carry the factor: oracle on a Problem whose pzv, pz0 are scaled."""
    import copy
    from compas_tno_b200.plan import TnoPlan
    from oracle import callbacks_np as onp
    P, objective, xs, _ = load_case("dome_q_zb_lambdv")
    plan = TnoPlan(P, objective=objective, kernel=kernel)
    scale = np.array([0.8, 1.0, 1.25])
    got = plan.evaluate_host(xs[:3], pz_scale=scale)
    base = plan.evaluate_host(xs[:3])
    assert np.array_equal(got["J"][1], base["J"][1])            # scale 1 changes nothing
    assert not np.allclose(got["J"][0], base["J"][0])
    for j in range(3):
        Pj = copy.copy(P)
        Pj.pzv, Pj.pz0 = P.pzv * scale[j], P.pz0 * scale[j]
        Pj.P = P.P.copy()
        Pj.P[:, 2] *= scale[j]
        ref = onp.evaluate(xs[j], onp.clone_problem(Pj), objective)
        assert rel_err(got["z"][j], ref["z"]) < TOL_VAL
        assert abs(got["f"][j] - ref["f"]) <= TOL_VAL * max(1.0, abs(ref["f"]))
        assert rel_err(got["grad"][j], ref["grad"]) < TOL_JAC
        assert_blockwise(got["g"][j], ref["g"], got["J"][j], ref["J"], P, TOL_VAL, TOL_JAC, "pz_scale %g" % scale[j])
    plan.close()


def test_pz_scale_with_lambdv_and_reac_bounds(tno):
    """Same with the reaction rows: the dRz / d lambdv term of jacobian.py:355-373 carries the scale too."""
    import copy
    from compas_tno_b200.plan import TnoPlan
    from oracle import callbacks_np as onp
    P, objective, xs, _ = load_case("dome_q_zb_lambdv")
    Pr = copy.copy(P)
    src = load_case("dome_q_zb_reac_min")[0]
    Pr.constraints = ["funicular", "envelope", "reac_bounds"]
    Pr.b = src.b.copy()
    for kernel in ("interleaved", "onchip"):
        plan = TnoPlan(Pr, objective=objective, kernel=kernel)
        scale = np.array([0.9, 1.1])
        got = plan.evaluate_host(xs[:2], pz_scale=scale)
        for j in range(2):
            Pj = copy.copy(Pr)
            Pj.pzv, Pj.pz0 = Pr.pzv * scale[j], Pr.pz0 * scale[j]
            Pj.P = Pr.P.copy()
            Pj.P[:, 2] *= scale[j]
            ref = onp.evaluate(xs[j], onp.clone_problem(Pj), objective)
            assert_blockwise(got["g"][j], ref["g"], got["J"][j], ref["J"], Pr, TOL_VAL, TOL_JAC, (kernel, j))
        plan.close()


def test_envelopexy_without_xyb_gradient_only(tno):
    """with the envelopexy constraint but no xyb variable the thrust gradient has no dx/dq_ind term
    (derivatives.py:257-301 keeps dxdq = 0 on a fixed plan) and must not read the Dx / Dy panels, which are only solved
    when J is requested.  grad alone == grad next to J == oracle."""
    import copy
    from compas_tno_b200.plan import TnoPlan
    from oracle import callbacks_np as onp
    src, objective, xs_src, _ = load_case("cross14_q_xyb_zb_min")
    P = copy.copy(src)
    P.variables = ["q", "zb"]
    P.constraints = ["funicular", "envelopexy", "envelope"]
    k, nb = P.k, P.nb
    X = np.hstack([xs_src[:, :k], xs_src[:, k + 2 * nb:]])      # drop the xyb block
    rng = np.random.default_rng(5)
    X = np.repeat(X[:1], 9, axis=0) * (1.0 + 0.02 * rng.uniform(-1, 1, size=(9, X.shape[1])))
    plan = TnoPlan(P, objective=objective)
    assert plan.info["kernel"] == 1 and plan.nvar == k + nb
    # poison the scratch with a full evaluation of very different candidates first
    plan.evaluate_host(X[::-1] * 1.3)
    only_grad = plan.evaluate_host(X, outputs=("grad",))
    full = plan.evaluate_host(X)
    assert np.array_equal(only_grad["grad"], full["grad"])
    f_only = plan.evaluate_host(X, outputs=("f", "grad", "g"))
    assert np.array_equal(f_only["grad"], full["grad"])
    ref = onp.evaluate_batch(X[:4], P, objective)
    assert rel_err(full["grad"][:4], ref["grad"]) < TOL_JAC
    assert rel_err(full["f"][:4], ref["f"]) < TOL_VAL
    assert_blockwise(full["g"][:4], ref["g"], full["J"][:4], ref["J"], P, TOL_VAL, TOL_JAC, "envelopexy, fixed supports")
    plan.close()


def test_argument_errors_are_value_errors(tno):
    """TNO_EINVAL (plain argument error) and TNO_EUNSUPPORTED (combination not provided) are told apart."""
    import torch
    from compas_tno_b200.plan import TnoPlan
    P, objective, xs, _ = load_case("cross6_q_zb_min")
    plan = TnoPlan(P, objective=objective)
    x = torch.from_numpy(xs).cuda()
    buf = torch.empty(len(xs) * plan.ng * plan.nvar + 1, dtype=torch.float64, device="cuda")
    J = buf[1:].view(len(xs), plan.ng, plan.nvar)              # 8-byte aligned only
    with pytest.raises(ValueError):
        plan.evaluate_device(x, ("J",), out={"J": J})
    with pytest.raises(NotImplementedError):
        plan.evaluate_host(xs, load_dir=np.tile([1.0, 0.0], (len(xs), 1)))
    plan.close()
