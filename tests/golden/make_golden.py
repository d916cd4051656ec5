"""Generate tests/golden/*.npz by running the UNMODIFIED reference callbacks.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

For every case it stores the problem arrays (so the test never needs the
reference's data/ directory), the evaluation points x and what the reference's
own ``constr_wrapper`` / ``sensitivities_wrapper`` / objective / gradient returned
for them, called in the order the solver calls them.

Storage trick (asserted here, undone in tests/golden_io.py): the funicular rows
of J are exactly [B 0; -B 0] (+ the d0 column for lambdh) and the second envelope
block is exactly minus the first, so only the n x nvar block ``J_env`` and the
2nb x nvar block ``J_reac`` are stored.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)

from compas_tno_b200.problems import FixedProblem, cross_form_arrays  # noqa: E402
from oracle import fixtures, refstub  # noqa: E402
from oracle.envelopes_np import CrossVaultEnvelopeNP, DomeEnvelopeNP  # noqa: E402

SEED = 20261019


def reference_problem(ref, P, variables, constraints, thk, envelope=None, extra=None):
    """Instantiate the reference's FixedProblem dataclass from our host arrays."""
    names = ("q m n ni nb E C Ct Ci Cit Cb U V P free fixed ph lb ub lb0 ub0 s X y0 free_x free_y "
             "Citx City qmin qmax ind k dep B d d0 Edinv Ed Ei").split()
    kw = {}
    for name in names:
        v = getattr(P, name)
        kw[name] = v.copy() if isinstance(v, np.ndarray) else v
    M = ref.FixedProblem(**kw)
    M.x0 = P.X[:, [0]].copy()
    M.variables = list(variables)
    M.constraints = list(constraints)
    M.features = ["fixed"]
    M.thk = thk
    M.min_lb = 0.0
    M.envelope = envelope
    for key, val in (extra or {}).items():
        setattr(M, key, np.array(val, dtype=np.float64, copy=True))
    if extra and "stiff" in extra:
        M.Ecomp_method = "simplified"      # the only method set-up accepts (problems/setup.py:146-157)
    return M


def sample_points(rng, P, form, n_pts, extra_tail, sigma=0.05):
    """x* from the fixture followed by perturbed candidates."""
    q_star = form["q"][P.ind]
    zb_star = form["xyz"][P.fixed, 2]
    lbb = form["lb"][P.fixed]
    ubb = form["ub"][P.fixed]
    lo = np.where(lbb < -0.5, zb_star - 0.05, lbb)
    xs = []
    for j in range(n_pts):
        if j == 0:
            qi, zb = q_star.copy(), zb_star.copy()
        else:
            qi = q_star * (1.0 + sigma * rng.uniform(-1, 1, size=q_star.shape))
            zb = rng.uniform(lo, ubb)
        xs.append(np.concatenate([qi, zb, extra_tail(j, rng)]))
    return np.array(xs)


def run_case(ref, name, form, P, variables, constraints, objective, xs, thk, envelope=None, extra=None):
    M = reference_problem(ref, P, variables, constraints, thk, envelope, extra)
    fobj, fgrad = ref.objective_selector(objective)
    m, n, nb, k = P.m, P.n, P.nb, P.k
    out = dict(g=[], J_env=[], J_reac=[], f=[], grad=[], X=[], q=[], mineig=[])
    for x in xs:
        g = ref.constr_wrapper(x.copy(), M)
        J = ref.sensitivities_wrapper(x.copy(), M)
        f = fobj(x.copy(), M)
        grad = np.asarray(fgrad(x.copy(), M)).ravel()
        nvar = J.shape[1]
        assert nvar == len(x), (nvar, len(x))
        # structural identities that let us store J compactly
        fun = np.zeros((2 * m, nvar))
        fun[:m, :k] = P.B
        fun[m:, :k] = -P.B
        if "lambdh" in variables:
            c = variables_offset(variables, P, "lambdh")
            fun[:m, c] = extra["d0"].ravel()
            fun[m:, c] = -extra["d0"].ravel()
        assert np.array_equal(J[: 2 * m], fun)
        nxy = 4 * n if "envelopexy" in constraints else 0
        if nxy:
            out.setdefault("J_envxy", []).append(J[2 * m: 2 * m + nxy].copy())
            J = np.vstack([J[: 2 * m], J[2 * m + nxy:]])
        env = J[2 * m: 2 * m + n]
        neg = J[2 * m + n: 2 * m + 2 * n].copy()
        if "t" in variables:
            c = variables_offset(variables, P, "t")
            # t column is [-dlb ; +dub]: not a negated pair; store the ub half separately
            out.setdefault("J_dub", []).append(neg[:, c].copy())
            neg[:, c] = -env[:, c]
        assert np.array_equal(neg, -env)
        out["g"].append(g)
        out["J_env"].append(env.copy())
        out["J_reac"].append(J[2 * m + 2 * n:].copy())
        out["f"].append(f)
        out["grad"].append(grad)
        out["X"].append(M.X.copy())
        out["q"].append(M.q.ravel().copy())
        Aneg = -(M.Cit @ __import__("scipy.sparse").sparse.diags(M.q.ravel()) @ M.Ci).toarray()
        out["mineig"].append(np.linalg.eigvalsh(Aneg).min())
    data = {key: np.array(val) for key, val in out.items()}
    data.update(
        x=xs, xyz=form["xyz"], edges=form["edges"], fixed=np.array(P.fixed), loads=form["P"],
        lb=form["lb"], ub=form["ub"], target=form["target"], q_form=form["q"], ind=np.array(P.ind),
        B=P.B, d=P.d.ravel(), thk=np.float64(thk),
        variables=np.array(variables), constraints=np.array(constraints), objective=np.array(objective),
    )
    for key, val in (extra or {}).items():
        data["extra_" + key] = np.asarray(val, dtype=np.float64)
    if envelope is not None:
        data["envelope"] = np.array(type(envelope).__name__.replace("NP", ""))
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **data)
    print("%-28s N=%d nvar=%d ng=%d  min eig(-A) %.3e..%.3e  min g %.3e  %.0f KB" % (
        name, len(xs), xs.shape[1], len(data["g"][0]), data["mineig"].min(), data["mineig"].max(),
        data["g"].min(), os.path.getsize(path) / 1024))


def variables_offset(variables, P, name):
    pos = P.k
    for v in ("xyb", "zb", "t", "lambdh", "lambdv"):
        if v == name:
            return pos
        if v in variables:
            pos += 2 * P.nb if v == "xyb" else (P.nb if v == "zb" else 1)
    raise KeyError(name)


def main():
    ref = refstub.load_reference()
    rng = np.random.default_rng(SEED)
    none = lambda j, rng: np.zeros(0)  # noqa: E731

    # ---- known answers straight from the fixtures -------------------------------------------
    known = {}
    for tag, fname in (("cross14", "crossvault_form.json"), ("fan", "fan_form.json"),
                       ("dome", "analysis.json"), ("freeform", "form.json")):
        form = fixtures.load_form(fname)
        for key, val in form.items():
            known["%s_%s" % (tag, key)] = val
    np.savez_compressed(os.path.join(HERE, "fixtures_known_answers.npz"), **known)

    # ---- A: cross vault fixture, [q, zb], funicular + envelope, min thrust ------------------
    form = fixtures.load_form("crossvault_form.json")
    P = FixedProblem.from_arrays(form["xyz"], form["edges"], form["fixed"], form["P"], form["lb"], form["ub"],
                                 form["target"], form["q"], ind=form["ind"])
    xs = sample_points(rng, P, form, 4, none)
    run_case(ref, "cross14_q_zb_min", form, P, ["q", "zb"], ["funicular", "envelope"], "min", xs, 0.5)

    run_case(ref, "cross14_q_zb_bestfit", form, P, ["q", "zb"], ["funicular", "envelope"], "bestfit", xs[:3], 0.5)
    run_case(ref, "cross14_q_zb_loadpath", form, P, ["q", "zb"], ["funicular", "envelope"], "loadpath", xs[:3], 0.5)
    centre = form["xyz"][:, :2].mean(axis=0)
    out = form["xyz"][P.fixed, :2] - centre
    dXb = np.column_stack([out / np.linalg.norm(out, axis=1, keepdims=True), -0.5 * np.ones(P.nb)])
    dXb[1] *= 0.3                        # not symmetric on purpose
    stiff = np.random.default_rng(SEED + 77).uniform(0.01, 0.05, size=(P.m, 1))
    run_case(ref, "cross14_q_zb_ecomp_nonlinear", form, P, ["q", "zb"], ["funicular", "envelope"], "Ecomp-nonlinear", xs[:3], 0.5,
             extra={"dXb": dXb, "stiff": stiff})
    # moving supports: x = [q_ind | xb | yb | zb], x / y limits of +-0.25 m around the form diagram
    xyb0 = form["xyz"][P.fixed, :2].flatten("F")
    rxy = np.random.default_rng(SEED + 78)
    xs_xyb = np.array([np.concatenate([x[: P.k], xyb0 + (0.0 if j == 0 else 1.0) * rxy.uniform(-0.08, 0.08, size=xyb0.shape), x[P.k:]])
                       for j, x in enumerate(xs[:3])])
    lim = {"xlimits": np.column_stack([form["xyz"][:, 0] - 0.25, form["xyz"][:, 0] + 0.25]),
           "ylimits": np.column_stack([form["xyz"][:, 1] - 0.25, form["xyz"][:, 1] + 0.25])}
    run_case(ref, "cross14_q_xyb_zb_min", form, P, ["q", "xyb", "zb"], ["funicular", "envelopexy", "envelope"], "min", xs_xyb, 0.5,
             extra=lim)
    run_case(ref, "cross14_q_zb_feasibility", form, P, ["q", "zb"], ["funicular", "envelope"], "feasibility", xs[:2], 0.5)

    # ---- B: cross vault, thickness variable -------------------------------------------------
    env = CrossVaultEnvelopeNP((0.0, 10.0), (0.0, 10.0), 0.5)
    tail = lambda j, rng: np.array([0.5 if j == 0 else rng.uniform(0.3, 0.5)])  # noqa: E731
    xs = sample_points(rng, P, form, 3, tail)
    run_case(ref, "cross14_q_zb_t", form, P, ["q", "zb", "t"], ["funicular", "envelope"], "t", xs, 0.5, envelope=env)
    # moving supports and a variable thickness together: the parametric envelope is evaluated at the moved x, y
    xs_xt = np.array([np.concatenate([x[: P.k], xyb0 + (0.0 if j == 0 else 1.0) * rxy.uniform(-0.08, 0.08, size=xyb0.shape), x[P.k:]])
                      for j, x in enumerate(xs)])
    run_case(ref, "cross14_q_xyb_zb_t", form, P, ["q", "xyb", "zb", "t"], ["funicular", "envelopexy", "envelope"], "t", xs_xt, 0.5,
             envelope=env, extra=lim)

    # ---- C: fan fixture ---------------------------------------------------------------------
    form = fixtures.load_form("fan_form.json")
    P = FixedProblem.from_arrays(form["xyz"], form["edges"], form["fixed"], form["P"], form["lb"], form["ub"],
                                 form["target"], form["q"], ind=form["ind"])
    xs = sample_points(rng, P, form, 3, none)
    run_case(ref, "fan_q_zb_max", form, P, ["q", "zb"], ["funicular", "envelope"], "max", xs, 0.5)

    # ---- D: dome fixture with reaction bounds -----------------------------------------------
    form = fixtures.load_form("analysis.json")
    P = FixedProblem.from_arrays(form["xyz"], form["edges"], form["fixed"], form["P"], form["lb"], form["ub"],
                                 form["target"], form["q"], ind=form["ind"])
    thk = float(form["thk"])
    xs = sample_points(rng, P, form, 3, none)
    run_case(ref, "dome_q_zb_reac_min", form, P, ["q", "zb"], ["funicular", "envelope", "reac_bounds"], "min",
             xs, thk, extra={"b": form["b"]})

    xyb0d = form["xyz"][P.fixed, :2].flatten("F")
    rxyd = np.random.default_rng(SEED + 79)
    xs_d = np.array([np.concatenate([x[: P.k], xyb0d + (0.0 if j == 0 else 1.0) * rxyd.uniform(-0.05, 0.05, size=xyb0d.shape), x[P.k:]])
                     for j, x in enumerate(xs[:2])])
    limd = {"xlimits": np.column_stack([form["xyz"][:, 0] - 0.3, form["xyz"][:, 0] + 0.3]),
            "ylimits": np.column_stack([form["xyz"][:, 1] - 0.3, form["xyz"][:, 1] + 0.3]), "b": form["b"]}
    run_case(ref, "dome_q_xyb_zb_reac_min", form, P, ["q", "xyb", "zb"], ["funicular", "envelopexy", "envelope", "reac_bounds"], "min",
             xs_d, thk, extra=limd)
    run_case(ref, "dome_q_zb_loadpath", form, P, ["q", "zb"], ["funicular", "envelope"], "loadpath", xs[:2], thk)
    out = form["xyz"][P.fixed, :2] - np.array([5.0, 5.0])
    dXb = np.column_stack([out / np.linalg.norm(out, axis=1, keepdims=True), np.linspace(-0.4, 0.2, P.nb)])
    run_case(ref, "dome_q_zb_ecomp_linear", form, P, ["q", "zb"], ["funicular", "envelope"], "Ecomp-linear", xs[:2], thk,
             extra={"dXb": dXb})
    run_case(ref, "dome_q_zb_bestfit", form, P, ["q", "zb"], ["funicular", "envelope"], "bestfit", xs[:2], thk)

    # ---- E: dome, thickness variable, x* = stored optimum -----------------------------------
    env = DomeEnvelopeNP((5.0, 5.0), 5.0, thk)
    tail = lambda j, rng: np.array([thk if j == 0 else rng.uniform(0.15, 0.5)])  # noqa: E731
    xs = sample_points(rng, P, form, 3, tail)
    assert np.array_equal(xs[0], form["xopt"]), "xopt layout [q_ind; zb; t] must be bit exact"
    run_case(ref, "dome_q_zb_t", form, P, ["q", "zb", "t"], ["funicular", "envelope"], "t", xs, thk, envelope=env)

    # ---- F: dome, horizontal load multiplier + reaction bounds ------------------------------
    weight = np.abs(form["P"][:, 2])
    theta = 0.3
    px0 = (np.cos(theta) * weight).reshape(-1, 1)
    py0 = (np.sin(theta) * weight).reshape(-1, 1)
    loads_h = form["P"].copy()
    loads_h[:, 0] = px0.ravel()
    loads_h[:, 1] = py0.ravel()
    Ph = FixedProblem.from_arrays(form["xyz"], form["edges"], form["fixed"], loads_h, form["lb"], form["ub"],
                                  form["target"], form["q"], ind=form["ind"])
    tail = lambda j, rng: np.array([0.02 if j == 0 else rng.uniform(0.0, 0.05)])  # noqa: E731
    xs = sample_points(rng, Ph, form, 3, tail)
    form_h = dict(form, P=loads_h)
    run_case(ref, "dome_q_zb_lambdh_reac", form_h, Ph, ["q", "zb", "lambdh"], ["funicular", "envelope", "reac_bounds"],
             "max_load", xs, thk, extra={"b": form["b"], "px0": px0, "py0": py0, "d0": Ph.d0})

    # ---- G: dome, vertical load multiplier --------------------------------------------------
    r = np.hypot(form["xyz"][:, 0] - 5.0, form["xyz"][:, 1] - 5.0)
    pzv = np.where(r < 1.5, -1.0, 0.0).reshape(-1, 1)
    pz0 = form["P"][:, [2]].copy()
    tail = lambda j, rng: np.array([1.0 if j == 0 else rng.uniform(0.0, 3.0)])  # noqa: E731
    xs = sample_points(rng, P, form, 3, tail)
    run_case(ref, "dome_q_zb_lambdv", form, P, ["q", "zb", "lambdv"], ["funicular", "envelope"], "max_load",
             xs, thk, extra={"pzv": pzv, "pz0": pz0})

    # ---- H: free-form mesh, many supports, k = 360 ------------------------------------------
    form = fixtures.load_form("form.json")
    P = FixedProblem.from_arrays(form["xyz"], form["edges"], form["fixed"], form["P"], form["lb"], form["ub"],
                                 form["target"], form["q"], ind=form["ind"])
    xs = sample_points(rng, P, form, 2, none, sigma=0.02)
    run_case(ref, "freeform_q_zb_min", form, P, ["q", "zb"], ["funicular", "envelope"], "min", xs, 0.5)

    # ---- I: synthetic cross form n = 6 (tiny; python-loop friendly) -------------------------
    xy, edges, fixed = cross_form_arrays(6, 10.0)
    envc = CrossVaultEnvelopeNP((0.0, 10.0), (0.0, 10.0), 0.5)
    ub, lb = envc.compute_bounds(xy[:, 0], xy[:, 1], 0.5)
    zt = envc.compute_middle(xy[:, 0], xy[:, 1])
    xyz = np.column_stack([xy, zt.ravel()])
    area = np.full(len(xy), (10.0 / 6) ** 2)
    loads = np.zeros((len(xy), 3))
    loads[:, 2] = -20.0 * 0.5 * area
    P6 = FixedProblem.from_arrays(xyz, edges, fixed, loads, lb.ravel(), ub.ravel(), zt.ravel())
    from scipy.optimize import lsq_linear
    res = lsq_linear(np.vstack([1e4 * P6.E, np.identity(P6.m)]), np.concatenate([np.zeros(P6.E.shape[0]), np.ones(P6.m)]),
                     bounds=(0, np.inf), method="trf")
    qstar = -res.x
    form6 = dict(xyz=xyz, edges=edges, fixed=np.array(fixed), P=loads, lb=lb.ravel(), ub=ub.ravel(), target=zt.ravel(), q=qstar)
    xs = sample_points(rng, P6, form6, 3, none)
    run_case(ref, "cross6_q_zb_min", form6, P6, ["q", "zb"], ["funicular", "envelope"], "min", xs, 0.5)


if __name__ == "__main__":
    main()
