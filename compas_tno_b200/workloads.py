"""Synthetic form diagrams and candidate batches for benchmarks and tests.

``compas_tna`` (FormDiagram.create_cross / create_circular_radial_spaced, Envelope.apply_*) is not
vendored with the reference, so the plan geometry follows the reference *fixtures*
(``data/crossvault_form.json``: validated identical for n = 14; ``data/analysis.json``: 16 meridians x 20
equally spaced hoops) and loads / bounds are documented synthetic stand-ins:
  * pz_i = -rho * t * A_i with rho = 20 and A_i the plan tributary area (NOT claimed equal to
    compas_tna's self-weight lumping),
  * ub / lb from the closed-form envelopes of ``envelopes.py``.
Everything here is one-time host set-up (numpy / scipy); the hot path is in the CUDA library.
"""
from __future__ import annotations

import numpy as np

from .envelopes import CrossVaultEnvelope, DomeEnvelope
from .problems import FixedProblem, cross_form_arrays

SEED = 20261019


def radial_form_arrays(n_hoops: int = 20, n_parallels: int = 16, center=(5.0, 5.0), radius: float = 5.0):
    """Radial / circular form diagram: 1 apex + n_hoops rings of n_parallels vertices, equally spaced radii.
    Meridian segments and hoop edges; the outermost ring is the support ring and has no hoop edges
    (in the reference fixture those are ``_is_edge: false``)."""
    xc, yc = center
    pts = [(xc, yc)]
    for h in range(1, n_hoops + 1):
        r = radius * h / n_hoops
        for p in range(n_parallels):
            a = 2.0 * np.pi * p / n_parallels
            pts.append((xc + r * np.cos(a), yc + r * np.sin(a)))
    vid = lambda h, p: 1 + (h - 1) * n_parallels + (p % n_parallels)  # noqa: E731
    edges = []
    for p in range(n_parallels):
        edges.append((0, vid(1, p)))
    for h in range(1, n_hoops + 1):
        for p in range(n_parallels):
            if h < n_hoops:
                edges.append((vid(h, p), vid(h + 1, p)))
                edges.append((vid(h, p), vid(h, p + 1)))
    fixed = [vid(n_hoops, p) for p in range(n_parallels)]
    return np.array(pts, dtype=np.float64), np.array(edges, dtype=np.int64), fixed


def _tributary_areas(xy, edges):
    """Plan tributary area per vertex: half the length-weighted star (good enough for a synthetic load)."""
    n = len(xy)
    a = np.zeros(n)
    L = np.linalg.norm(xy[edges[:, 0]] - xy[edges[:, 1]], axis=1)
    np.add.at(a, edges[:, 0], 0.25 * L * L)
    np.add.at(a, edges[:, 1], 0.25 * L * L)
    return a


def compressive_basis(problem: FixedProblem, n_states: int = 12, seed: int = SEED):
    """``n_states`` compression-only equilibrium states q = B y + d of the fixed plan.

    On the cross form n^2 of the 2n(n+2) edges are structurally forced to q = 0 under vertical load
    (SURVEY.md section 8d), so a random perturbation of q_ind produces tension and an indefinite
    Ci^T Q Ci.  Compressive states form a convex cone; we pick points of it with small LPs in the k
    independent parameters:   max sum_e c_e t_e   s.t.  (B y + d)_e <= -t_e,  0 <= t_e <= u_e
    with random caps u_e in [0.5, 1.5] (t_e stays 0 exactly on the forced-zero edges).
    Convex combinations of the returned states are again compressive equilibrium states.
    Returns Y (n_states, k)."""
    from scipy.optimize import linprog
    from scipy.sparse import csr_matrix, hstack, identity, vstack
    rng = np.random.default_rng(seed)
    m, k = problem.m, problem.k
    B = csr_matrix(problem.B)
    d = problem.d.ravel()
    # stage 1: which edges can be strictly compressed at all?   max sum t  s.t.  B y + d + t <= 0, 0 <= t <= 1
    res = linprog(np.concatenate([np.zeros(k), -np.ones(m)]), A_ub=hstack([B, identity(m, format="csr")], format="csr"),
                  b_ub=-d, bounds=[(-1e3, 1e3)] * k + [(0.0, 1.0)] * m, method="highs")
    if res.status != 0:
        raise RuntimeError("no compressive state found: %s" % res.message)
    active = res.x[k:] > 1e-7
    # stage 2: well-conditioned states:  min s  s.t.  -s <= (B y + d)_e <= -u_e (active), <= 0 (forced-zero edges)
    ones = csr_matrix(np.ones((m, 1)))
    A_ub = vstack([hstack([B, csr_matrix((m, 1))]), hstack([-B, -ones])], format="csr")
    states = []
    for _ in range(n_states):
        u = np.where(active, rng.uniform(0.5, 1.5, size=m), 0.0)
        b_ub = np.concatenate([-d - u, d])
        c = np.concatenate([np.zeros(k), [1.0]])
        res = linprog(c, A_ub=A_ub, b_ub=b_ub, bounds=[(None, None)] * k + [(0.0, None)], method="highs")
        if res.status != 0:
            raise RuntimeError("no compressive state found: %s" % res.message)
        states.append(res.x[:k])
    return np.array(states)


def compressive_state(problem: FixedProblem, seed: int = SEED):
    """Mean of a few basis states: (q_ind*, q*)."""
    Y = compressive_basis(problem, 4, seed)
    y = Y.mean(axis=0)
    q = problem.B @ y.reshape(-1, 1) + problem.d
    return y, q.ravel()


def crossvault_problem(n_div: int = 20, thk: float = 0.5, span: float = 10.0, rho: float = 20.0,
                       variables=("q", "zb"), constraints=("funicular", "envelope"), n_states: int = 12):
    """Cross vault on the orthogonal cross form diagram of discretisation ``n_div``.
    Returns (problem, x_star) with x_star = [q_ind*, zb*(, t)]."""
    xy, edges, fixed = cross_form_arrays(n_div, span)
    env = CrossVaultEnvelope((0.0, span), (0.0, span), thk)
    ub, lb = env.compute_bounds(xy[:, 0], xy[:, 1], thk)
    mid = env.compute_middle(xy[:, 0], xy[:, 1])
    xyz = np.column_stack([xy, mid.ravel()])
    h = span / n_div
    on_x = (np.isclose(xy[:, 0], 0) | np.isclose(xy[:, 0], span)).astype(float)
    on_y = (np.isclose(xy[:, 1], 0) | np.isclose(xy[:, 1], span)).astype(float)
    area = h * h * (1 - 0.5 * on_x) * (1 - 0.5 * on_y)
    loads = np.zeros((len(xy), 3))
    loads[:, 2] = -rho * thk * area
    P = FixedProblem.from_arrays(xyz, edges, fixed, loads, lb.ravel(), ub.ravel(), mid.ravel())
    P.variables, P.constraints, P.features = list(variables), list(constraints), ["fixed"]
    P.thk, P.envelope = thk, env
    zb = mid.ravel()[P.fixed]
    return _finish(P, zb, mid.ravel(), variables, thk, n_states)


def dome_problem(n_hoops: int = 20, n_parallels: int = 16, thk: float = 0.5, radius: float = 5.0, rho: float = 20.0,
                 variables=("q", "zb"), constraints=("funicular", "envelope"), n_states: int = 12):
    """Hemispherical dome on the radial form diagram.  Returns (problem, x_star)."""
    c = (radius, radius)
    xy, edges, fixed = radial_form_arrays(n_hoops, n_parallels, c, radius)
    env = DomeEnvelope(c, radius, thk)
    ub, lb = env.compute_bounds(xy[:, 0], xy[:, 1], thk)
    mid = env.compute_middle(xy[:, 0], xy[:, 1])
    xyz = np.column_stack([xy, mid.ravel()])
    loads = np.zeros((len(xy), 3))
    loads[:, 2] = -rho * thk * _tributary_areas(xy, edges)
    P = FixedProblem.from_arrays(xyz, edges, fixed, loads, lb.ravel(), ub.ravel(), mid.ravel())
    P.variables, P.constraints, P.features = list(variables), list(constraints), ["fixed"]
    P.thk, P.envelope = thk, env
    if "reac_bounds" in constraints:
        P.b = env_bound_react(env, xy, fixed, thk)
    zb = np.full(len(fixed), 0.0)
    return _finish(P, zb, mid.ravel(), variables, thk, n_states)


def dome_lambdh_problem(n_hoops: int = 20, n_parallels: int = 16, thk: float = 0.5, radius: float = 5.0, rho: float = 20.0,
                        theta: float = 0.3, lambd: float = 0.02, constraints=("funicular", "envelope", "reac_bounds"),
                        n_states: int = 12):
    """Dome under a horizontal load multiplier: px0, py0 = (cos, sin)(theta) x |pz|, q = B q_ind + lambdh d0
    (variables q, zb, lambdh).  The compressive basis is that of the vertically loaded dome (d = 0); x* carries
    lambdh = ``lambd``.  Returns (problem, x_star)."""
    P0, x0 = dome_problem(n_hoops, n_parallels, thk, radius, rho, ("q", "zb"), constraints, n_states)
    w = np.abs(P0.P[:, 2])
    loads = P0.P.copy()
    loads[:, 0], loads[:, 1] = np.cos(theta) * w, np.sin(theta) * w
    Ph = FixedProblem.from_arrays(P0.X, _edges_of(P0), P0.fixed, loads, P0.lb.ravel(), P0.ub.ravel(), P0.s.ravel(), ind=P0.ind)
    Ph.variables, Ph.constraints, Ph.features = ["q", "zb", "lambdh"], list(constraints), ["fixed"]
    Ph.thk, Ph.envelope = thk, P0.envelope
    Ph.px0, Ph.py0 = loads[:, [0]].copy(), loads[:, [1]].copy()
    if getattr(P0, "b", None) is not None:
        Ph.b = P0.b
    Ph.candidate_basis = P0.candidate_basis
    # d0 (zero on the independent edges, problems.py:516-528) is a poor equilibrium state on its own; candidates
    # use q_ind = y - lambdh * shift, i.e. q = B y + lambdh * (d0 - B B^+ d0): the minimum-norm particular solution
    Ph.lambdh_shift = np.linalg.lstsq(Ph.B, Ph.d0.ravel(), rcond=None)[0]
    xq = x0[: Ph.k] - lambd * Ph.lambdh_shift
    Ph.q = Ph.B @ xq.reshape(-1, 1) + lambd * Ph.d0
    return Ph, np.concatenate([xq, x0[Ph.k:], [lambd]])


def dome_direction_problem(n_hoops: int = 20, n_parallels: int = 16, thk: float = 0.5, radius: float = 5.0, rho: float = 20.0,
                           lambd: float = 0.02, constraints=("funicular", "envelope", "reac_bounds"), n_states: int = 12):
    """Dome for a sweep over horizontal load *directions* (BASELINE configuration 5):
    two load bases, theta = 0 (px0 = |pz|, py0 = 0, d0) and theta = 90 degrees (px0_b = 0, py0_b = |pz|, d0_b), so that
    a candidate with direction (c, s) sees d0(theta) = c d0 + s d0_b.  Returns (problem, x_star) with x_star for
    theta = 0; ``direction_candidates`` builds the batch."""
    Px, x0 = dome_lambdh_problem(n_hoops, n_parallels, thk, radius, rho, 0.0, lambd, constraints, n_states)
    Py, _ = dome_lambdh_problem(n_hoops, n_parallels, thk, radius, rho, 0.5 * np.pi, lambd, constraints, n_states)
    Py.px0[:] = 0.0  # cos(pi / 2) leaves 6e-17 |pz| behind
    Px.py0[:] = 0.0
    Px.d0_b, Px.px0_b, Px.py0_b = Py.d0.copy(), Py.px0.copy(), Py.py0.copy()
    Px.lambdh_shift_b = Py.lambdh_shift
    return Px, x0


def direction_candidates(problem, x_star, N: int, seed: int = SEED, sigma: float = 0.05):
    """(X, load_dir): N candidates with directions theta_j = 2 pi j / N; q_ind carries the shift of its direction so
    that q = B y + lambdh (c d0 + s d0_b) stays a compression-only state."""
    k = problem.k
    theta = 2.0 * np.pi * np.arange(N) / N
    load_dir = np.ascontiguousarray(np.column_stack([np.cos(theta), np.sin(theta)]))
    x0 = np.asarray(x_star, dtype=np.float64).copy()
    x0[:k] += x0[-1] * problem.lambdh_shift
    X = sample_candidates(problem, x0, N, seed=seed, sigma=sigma, basis=problem.candidate_basis)
    shift = load_dir[:, [0]] * problem.lambdh_shift[None, :] + load_dir[:, [1]] * problem.lambdh_shift_b[None, :]
    X[:, :k] -= X[:, [-1]] * shift
    return np.ascontiguousarray(X), load_dir


def lambdh_candidates(problem, x_star, N: int, seed: int = SEED, sigma: float = 0.05):
    """sample_candidates for a lambdh problem: convex combinations of the compressive basis, lambdh perturbed by
    sigma, and the q_ind shift that goes with each candidate's lambdh."""
    k = problem.k
    x0 = np.asarray(x_star, dtype=np.float64).copy()
    x0[:k] += x0[-1] * problem.lambdh_shift
    X = sample_candidates(problem, x0, N, seed=seed, sigma=sigma, basis=problem.candidate_basis)
    X[:, :k] -= X[:, [-1]] * problem.lambdh_shift[None, :]
    return np.ascontiguousarray(X)


def _edges_of(P):
    """(m, 2) vertex pairs (u, v) of the connectivity matrix rows: C[e, u] = -1, C[e, v] = +1."""
    C = P.C.tocsr()
    out = np.zeros((C.shape[0], 2), dtype=np.int64)
    for e in range(C.shape[0]):
        cols = C.indices[C.indptr[e]: C.indptr[e + 1]]
        vals = C.data[C.indptr[e]: C.indptr[e + 1]]
        out[e, 0], out[e, 1] = cols[vals < 0][0], cols[vals > 0][0]
    return out


def _finish(P, zb, mid, variables, thk, n_states):
    """Basis of compressive states, each scaled so that its highest node touches the middle surface's crown
    (heights above the supports scale like 1 / |q|); x* = their mean.  Stores the basis on the problem."""
    from scipy.sparse import diags
    from scipy.sparse.linalg import splu
    Y = compressive_basis(P, n_states)
    crown = float(mid.max() - zb.mean())
    Xb = P.X[P.fixed].copy()
    Xb[:, 2] = zb
    for i in range(len(Y)):
        q = (P.B @ Y[i].reshape(-1, 1) + P.d).ravel()
        Q = diags(q)
        zi = splu((P.Cit @ Q @ P.Ci).tocsc()).solve(P.P[P.free, 2] - (P.Cit @ Q @ P.Cb) @ Xb[:, 2])
        rise = float(zi.max() - zb.mean())
        if rise > 0 and not np.any(np.abs(P.d) > 0):
            Y[i] *= rise / crown
    P.candidate_basis = Y
    y = Y.mean(axis=0)
    P.q = P.B @ y.reshape(-1, 1) + P.d
    x_star = np.concatenate([y, zb] + ([np.array([thk])] if ("t" in variables or "n" in variables) else []))
    return P, x_star


def env_bound_react(env, xy, fixed, thk, floor: float = 0.1):
    """Reaction bounds b = (t / 2) x unit radial (the closed form the dome fixture follows), with
    every component kept at least ``floor`` x t / 2 in magnitude: a support on a coordinate axis would otherwise get a zero
    bound, whose constraint row  |b| - rise |R_xy / R_z|  admits no asymmetric state at all -- a synthetic sweep over
    perturbed candidates would then be infeasible by construction.  (Synthetic stand-in; the fixtures keep their own b.)"""
    d = xy[fixed] - np.array([env.xc, env.yc])
    r = np.linalg.norm(d, axis=1, keepdims=True)
    b = 0.5 * thk * d / r
    return np.where(np.abs(b) < floor * 0.5 * thk, np.where(b < 0, -1.0, 1.0) * floor * 0.5 * thk, b)


def sample_candidates(problem, x_star, N: int, seed: int = SEED, sigma: float = 0.05, basis=None):
    """Candidate batch (adapted): q_ind_j = s_j * sum_i w_ji Y_i with Dirichlet weights
    over the compressive basis states ``basis`` (default: x* only) and s_j ~ U(1 - sigma, 1 + sigma), which keeps
    every candidate a compression-only equilibrium state; zb ~ U(lb_b, ub_b) where the envelope is defined
    there, else x*'s value; trailing scalar variables (t, lambda) perturbed by sigma."""
    rng = np.random.default_rng(seed)
    k, nb = problem.k, problem.nb
    x_star = np.asarray(x_star, dtype=np.float64)
    X = np.repeat(x_star[None, :], N, axis=0)
    if basis is not None and len(basis) > 1:
        Wt = rng.dirichlet(np.ones(len(basis)), size=N)
        X[:, :k] = Wt @ np.asarray(basis)
    X[:, :k] *= 1.0 + sigma * rng.uniform(-1, 1, size=(N, 1))
    if "zb" in problem.variables:
        lo = problem.lb.ravel()[problem.fixed]
        hi = problem.ub.ravel()[problem.fixed]
        ok = (hi > lo) & (lo > -0.5)
        zb = rng.uniform(np.where(ok, lo, 0.0), np.where(ok, hi, 1.0), size=(N, nb))
        X[:, k:k + nb] = np.where(ok[None, :], zb, X[:, k:k + nb])
    if X.shape[1] > k + nb:
        X[:, k + nb:] *= 1.0 + sigma * rng.uniform(-1, 1, size=(N, X.shape[1] - k - nb))
    return np.ascontiguousarray(X)


def margin_rows(problem, plan, q_frac: float = 0.1):
    """Coefficient of the margin s in every row of g (0 = the row must merely hold):  envelope rows 1 (metres), the
    reac_bounds rows with a non-zero bound 1, and the  qmax - q  rows of the edges whose force density can move with the
    variables  q_frac x median|q|  per metre of margin -- so that a state with margin s keeps those edges in compression
    by that much and stays compression-only under small perturbations.  Edges that are structurally force-free (zero row
    of B and zero d: n^2 of them on the cross form) cannot carry slack and get 0."""
    m, ng = plan.m, plan.ng
    coef = np.zeros(ng)
    nfun = 2 * m if "funicular" in plan.constraints else 0
    coef[nfun:] = 1.0
    if "reac_bounds" in plan.constraints:
        b = np.abs(np.asarray(problem.b, dtype=np.float64))
        coef[ng - 2 * plan.nb:] = (np.concatenate([b[:, 0], b[:, 1]]) > 1e-9).astype(np.float64)
    if nfun:
        B = np.asarray(problem.B, dtype=np.float64)
        q0 = np.abs(np.asarray(problem.q, dtype=np.float64).ravel())
        movable = (np.abs(B).max(axis=1) > 1e-12) & (q0 > 1e-9 * max(q0.max(), 1e-300))
        coef[m:2 * m] = np.where(movable, q_frac * float(np.median(q0[movable])) if movable.any() else 0.0, 0.0)
    return coef


def max_margin_state(plan, x0, iters: int = 60, verbose: bool = False, rows=None, step: float = 0.1):
    """A strictly feasible variable vector for sweeps that should contain feasible candidates (multi-start, Monte-Carlo,
    thickness sweeps): maximise the smallest weighted constraint margin,

        max s   s.t.   g_r(x) - coef_r s >= 0  for every row r      (coef = ``rows``, see margin_rows; 0 = plain g_r >= 0)

    by sequential linear programming with a trust region: each iteration evaluates g and the dense J of the current
    point on the GPU (``plan.evaluate_host``, the same callables the reference hands to its solvers, solvers/scipy.py:172),
    solves  max s  s.t.  g + J dx - s coef >= 0, |dx_i| <= delta scale_i  with HiGHS, and accepts the step when the
    true margin grows (else the region shrinks).  The linear programme is always feasible, so an infeasible start is fine.
    One-time host set-up of a workload, not part of the timed path.  Returns (x, s): s > 0 means every constraint holds,
    the weighted ones with that margin."""
    from scipy.optimize import linprog
    from scipy.sparse import csr_matrix, hstack
    x = np.asarray(x0, dtype=np.float64).ravel().copy()
    nvar = plan.nvar
    if rows is None:
        rows = np.zeros(plan.ng)
        rows[(2 * plan.m if "funicular" in plan.constraints else 0):] = 1.0
    coef = np.asarray(rows, dtype=np.float64)
    wrows = coef > 0
    other = ~wrows
    scale = np.where(np.abs(x) > 1e-3, np.abs(x), 1.0)

    def ev(xv):
        r = plan.evaluate_host(xv.reshape(1, -1), ("g", "J", "status"))
        g = r["g"][0]
        ok = r["status"][0] == 0 and bool(np.all(np.isfinite(g)))
        merit = -np.inf
        if ok:
            viol = float(g[other].min()) if other.any() else 0.0
            mrg = float((g[wrows] / coef[wrows]).min())
            merit = mrg if viol >= -1e-9 else min(mrg, 0.0) + 1e3 * viol
        return g, r["J"][0], merit

    g, J, merit = ev(x)
    delta = step
    col_s = csr_matrix(coef.reshape(-1, 1))
    for it in range(iters):
        if not np.isfinite(merit):
            break
        # -(J dx) + s coef <= g
        A = hstack([csr_matrix(-J), col_s], format="csr")
        c = np.concatenate([np.zeros(nvar), [-1.0]])
        bnds = [(-delta * sc, delta * sc) for sc in scale] + [(None, None)]
        res = linprog(c, A_ub=A, b_ub=g, bounds=bnds, method="highs")
        if res.status != 0:
            break
        xn = x + res.x[:nvar]
        gn, Jn, mn = ev(xn)
        if verbose:
            print("  max_margin it %d delta %.3g: merit %.6g -> %.6g (LP predicts %.6g)" % (it, delta, merit, mn, res.x[nvar]))
        if mn > merit + 1e-12:
            gain = mn - merit
            x, g, J, merit = xn, gn, Jn, mn
            delta = min(delta * 1.5, 0.5)
            if gain < 1e-7 * max(1.0, abs(merit)):
                break
        else:
            delta *= 0.4
            if delta < 1e-6:
                break
    return x, float(merit)


def sample_around(problem, x_center, N: int, seed: int = SEED, sigma_lo: float = 1e-4, sigma_hi: float = 0.1):
    """Candidates scattered around a (feasible) centre with a ladder of relative perturbation sizes, log-uniform in
    [sigma_lo, sigma_hi]: the tight ones stay feasible, the wide ones leave the envelope, so a sweep over them has a
    feasible FRACTION (and a best feasible candidate that is not the centre itself).  Every variable is perturbed
    independently: x_j = x_c * (1 + sigma_j * xi_j), xi ~ U(-1, 1)."""
    rng = np.random.default_rng(seed)
    xc = np.asarray(x_center, dtype=np.float64).ravel()
    sig = np.exp(rng.uniform(np.log(sigma_lo), np.log(sigma_hi), size=(N, 1)))
    X = xc[None, :] * (1.0 + sig * rng.uniform(-1, 1, size=(N, len(xc))))
    return np.ascontiguousarray(X)
