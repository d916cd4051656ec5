"""CPU: the numpy restatement (oracle/) against the reference's own outputs and known answers."""
import numpy as np
import pytest

from golden_io import CASES, load_case, rel_err
from oracle import callbacks_np as onp


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    P, objective, xs, exp = load_case(name)
    got = onp.evaluate_batch(xs, P, objective)
    # the restatement runs the same scipy calls in the same order: expect ~rounding, assert a hard 1e-12
    assert rel_err(got["q"], exp["q"]) < 1e-13
    assert rel_err(got["X"], exp["X"]) < 1e-12
    assert rel_err(got["g"], exp["g"]) < 1e-12
    assert rel_err(got["f"], exp["f"]) < 1e-12
    assert rel_err(got["grad"], exp["grad"]) < 1e-12
    assert got["J"].shape == exp["J"].shape
    assert rel_err(got["J"], exp["J"]) < 1e-11


def _known(golden_dir):
    import os
    return np.load(os.path.join(golden_dir, "fixtures_known_answers.npz"))


@pytest.mark.parametrize("tag", ["cross14", "fan", "dome", "freeform"])
def test_fixture_known_answers(tag, golden_dir):
    """z = xyz_from_q(q), q = B q_ind + d and the load path stored in the reference's data/*.json."""
    from compas_tno_b200.problems import FixedProblem
    K = _known(golden_dir)
    g = lambda a: K["%s_%s" % (tag, a)]  # noqa: E731
    P = FixedProblem.from_arrays(g("xyz"), g("edges"), g("fixed"), g("P"), g("lb"), g("ub"), g("target"), g("q"), ind=g("ind"))
    q = g("q").reshape(-1, 1)
    Xf = onp.xyz_from_q(q, P.P[P.free], P.X[P.fixed], P.Ci, P.Cit, P.Cb)
    assert np.max(np.abs(Xf - g("xyz")[P.free])) < 5e-13
    q2 = onp.q_from_variables(q[P.ind], P.B, P.d)
    assert rel_err(q2, q) < 5e-12
    assert np.max(np.abs(P.E @ q - P.ph)) < 1e-10
    # load path  sum |q| l^2   (stored as attributes.loadpath)
    uvw = P.C @ g("xyz")
    lp = float(np.sum(np.abs(g("q")) * np.sum(uvw ** 2, axis=1)))
    assert abs(lp - float(g("loadpath"))) / float(g("loadpath")) < 1e-12


def test_dome_reactions_and_envelope(golden_dir):
    from compas_tno_b200.problems import FixedProblem
    from oracle.envelopes_np import DomeEnvelopeNP
    K = _known(golden_dir)
    g = lambda a: K["dome_%s" % a]  # noqa: E731
    P = FixedProblem.from_arrays(g("xyz"), g("edges"), g("fixed"), g("P"), g("lb"), g("ub"), g("target"), g("q"), ind=g("ind"))
    q = g("q")
    R = (P.Cb.T @ (q[:, None] * (P.C @ g("xyz")))) - P.P[P.fixed]
    assert rel_err(R, g("react")) < 1e-12
    t = float(g("thk"))
    assert t == float(g("fopt")) == float(g("xopt")[-1])
    env = DomeEnvelopeNP((5.0, 5.0), 5.0, t)
    ub, lb = env.compute_bounds(g("xyz")[:, 0], g("xyz")[:, 1], t)
    assert np.max(np.abs(ub.ravel() - g("ub"))) < 1e-13
    assert np.max(np.abs(lb.ravel() - g("lb"))) < 1e-13
    assert np.max(np.abs(env.compute_bound_react(g("xyz")[:, 0], g("xyz")[:, 1], t, P.fixed) - g("b"))) < 1e-13
    # variable layout [q_ind ; zb ; t] is bit exact
    assert np.array_equal(g("xopt")[: P.k], q[P.ind])
    assert np.array_equal(g("xopt")[P.k: P.k + P.nb], g("xyz")[P.fixed, 2])


def test_crossvault_envelope_closed_form(golden_dir):
    from oracle.envelopes_np import CrossVaultEnvelopeNP
    K = _known(golden_dir)
    for tag in ("cross14", "fan"):
        xyz = K[tag + "_xyz"]
        env = CrossVaultEnvelopeNP((0.0, 10.0), (0.0, 10.0), 0.5)
        ub, lb = env.compute_bounds(xyz[:, 0], xyz[:, 1], 0.5)
        assert np.max(np.abs(ub.ravel() - K[tag + "_ub"])) < 1e-13
        assert np.max(np.abs(lb.ravel() - K[tag + "_lb"])) < 1e-13
        assert np.max(np.abs(env.compute_middle(xyz[:, 0], xyz[:, 1]).ravel() - K[tag + "_target"])) < 1e-13


@pytest.mark.parametrize("cls_args", [("dome", (5.0, 5.0), 5.0), ("cross", (0.0, 10.0), (0.0, 10.0))])
def test_envelope_thickness_derivatives_fd(cls_args):
    from oracle.envelopes_np import CrossVaultEnvelopeNP, DomeEnvelopeNP
    rng = np.random.default_rng(3)
    x = rng.uniform(2.5, 7.5, 50)   # inside the dome footprint (r < R)
    y = rng.uniform(2.5, 7.5, 50)
    env = DomeEnvelopeNP(cls_args[1], cls_args[2], 0.4) if cls_args[0] == "dome" else CrossVaultEnvelopeNP(cls_args[1], cls_args[2], 0.4)
    t, h = 0.4, 1e-6
    dub, dlb = env.compute_bounds_derivatives(x, y, t)[:2]
    up, lp = env.compute_bounds(x, y, t + h)
    um, lm = env.compute_bounds(x, y, t - h)
    ok = (lp.ravel() > 0) & (lm.ravel() > 0)
    assert np.max(np.abs((up - um) / (2 * h) - dub)) < 1e-7
    assert np.max(np.abs(((lp - lm) / (2 * h) - dlb).ravel()[ok])) < 1e-7


def test_cross_form_generator_matches_fixture(golden_dir):
    from compas_tno_b200.problems import cross_form_arrays
    K = _known(golden_dir)
    xy, edges, fixed = cross_form_arrays(14, 10.0)
    fx = K["cross14_xyz"][:, :2]
    key = lambda p: (round(float(p[0]), 9), round(float(p[1]), 9))  # noqa: E731
    assert {key(p) for p in xy} == {key(p) for p in fx}
    eset = lambda pts, es: {frozenset((key(pts[u]), key(pts[v]))) for u, v in es}  # noqa: E731
    assert eset(xy, edges) == eset(fx, K["cross14_edges"])
    assert {key(xy[i]) for i in fixed} == {key(fx[i]) for i in K["cross14_fixed"]}


def test_oracle_jacobian_vs_finite_differences():
    """Third opinion: analytic J against central differences of g (reference ships d_fconstr for this)."""
    P, objective, xs, _ = load_case("cross6_q_zb_min")
    M = onp.clone_problem(P)
    x = xs[1]
    J = onp.sensitivities_wrapper(x, M)
    h = 1e-6
    for c in range(len(x)):
        e = np.zeros_like(x)
        e[c] = h
        fd = (onp.constr_wrapper(x + e, M) - onp.constr_wrapper(x - e, M)) / (2 * h)
        assert np.max(np.abs(fd - J[:, c])) < 2e-5 * max(1.0, np.max(np.abs(J[:, c])))
