"""Development prototype (CPU, numpy) of the multi-start SQP iteration the device driver implements
(compas_tno_b200/multistart.py + csrc/tno_sqp.cu): damped-BFGS SQP, the QP subproblem solved by a primal-dual interior
point method on the dense Jacobian, L1 merit line search.  Driven by the oracle callbacks; compares the optimum with
scipy's SLSQP (what the reference calls, solvers/scipy.py:172).

usage: python tools/sqp_proto.py [case]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))


def qp_ipm(H, c, A, b, iters=60, tol=1e-10):
    """min 1/2 d'Hd + c'd  s.t.  A d + b >= 0   (Mehrotra predictor-corrector).  Returns d, lam, ok."""
    m, n = A.shape
    d = np.zeros(n)
    s = np.maximum(b, 1.0)
    lam = np.ones(m)
    scale = 1.0 + np.abs(c).max()
    for it in range(iters):
        rd = H @ d + c - A.T @ lam
        rp = A @ d + b - s
        mu = s @ lam / m
        if max(np.abs(rd).max() / scale, np.abs(rp).max(), mu) < tol:
            return d, lam, True, it
        w = lam / s
        M = H + A.T @ (w[:, None] * A)
        L = np.linalg.cholesky(M + 1e-14 * np.eye(n) * np.trace(M) / n)

        def solve(rc):
            # rows: H dd - A' dl = -rd ; A dd - ds = -rp ; s dl + lam ds = rc
            # ds = A dd + rp ; dl = (rc - lam ds) / s
            rhs = -rd + A.T @ ((rc - lam * rp) / s)
            dd = np.linalg.solve(L.T, np.linalg.solve(L, rhs))
            ds = A @ dd + rp
            dl = (rc - lam * ds) / s
            return dd, ds, dl

        def maxstep(v, dv):
            neg = dv < 0
            return min(1.0, float((-v[neg] / dv[neg]).min())) if neg.any() else 1.0

        dd, ds, dl = solve(-s * lam)
        ap, ad = maxstep(s, ds), maxstep(lam, dl)
        mu_aff = (s + ap * ds) @ (lam + ad * dl) / m
        sigma = (mu_aff / mu) ** 3
        dd, ds, dl = solve(sigma * mu - s * lam - ds * dl)
        ap, ad = 0.995 * maxstep(s, ds), 0.995 * maxstep(lam, dl)
        ap, ad = min(ap, 1.0), min(ad, 1.0)
        d += ap * dd
        s += ap * ds
        lam += ad * dl
    return d, lam, False, iters


def sqp(fun, x0, lb, ub, iters=100, acc=1e-6, verbose=False):
    """fun(x) -> f, grad, g, J  (g >= 0 feasible)."""
    n = x0.size
    x = np.clip(x0.copy(), lb, ub)
    f, gr, g, J = fun(x)
    H = np.eye(n)
    rho = 0.0
    I = np.eye(n)
    for it in range(iters):
        A = np.vstack([J, I, -I])
        b = np.concatenate([g, x - lb, ub - x])
        d, lam, ok, nit = qp_ipm(H, gr, A, b)
        lam_g = lam[: g.size]
        viol = np.maximum(0.0, -g).sum()
        rho = max(rho, 1.01 * lam.max()) if it == 0 else max(0.5 * (rho + lam.max()), lam.max())
        phi0 = f + rho * viol
        dphi = gr @ d - rho * viol
        # SLSQP's stopping rule: |grad.d| + sum lam |g| small and constraints satisfied
        if abs(gr @ d) + lam @ np.abs(b) < acc and viol < acc:
            return x, f, it, True
        alpha = 1.0
        for ls in range(12):
            xn = x + alpha * d
            fn, grn, gn, Jn = fun(xn)
            if np.isfinite(fn) and fn + rho * np.maximum(0.0, -gn).sum() <= phi0 + 0.1 * alpha * dphi:
                break
            alpha *= 0.5 if ls else 0.3
        sk = xn - x
        yk = (grn - Jn.T @ lam_g) - (gr - J.T @ lam_g)
        Hs = H @ sk
        sHs = sk @ Hs
        sy = sk @ yk
        if sy < 0.2 * sHs:
            th = 0.8 * sHs / (sHs - sy)
            yk = th * yk + (1 - th) * Hs
            sy = sk @ yk
        if sHs > 0 and sy > 0:
            H = H - np.outer(Hs, Hs) / sHs + np.outer(yk, yk) / sy
        if verbose:
            print("it %3d f %.9f viol %.2e |d| %.2e alpha %.3f qp_it %d ok %d rho %.2e" % (it, f, viol, np.abs(d).max(), alpha, nit, ok, rho))
        x, f, gr, g, J = xn, fn, grn, gn, Jn
    return x, f, iters, False


def main():
    from scipy.optimize import fmin_slsqp
    from golden_io import load_case
    from oracle import callbacks_np as onp
    case = sys.argv[1] if len(sys.argv) > 1 else "cross6_q_zb_min"
    P, objective, xs, _ = load_case(case)
    M = onp.clone_problem(P)
    fo, fg = onp.objective_pair(objective) if hasattr(onp, "objective_pair") else (onp.f_min_thrust, onp.gradient_fmin)

    def fun(x):
        g = onp.constr_wrapper(x, M).ravel().copy()
        f = float(fo(x, M))
        gr = np.asarray(fg(x, M)).ravel().copy()
        J = np.asarray(onp.sensitivities_wrapper(x, M)).copy()
        return f, gr, g, J

    k = P.k
    lb = np.array([-1e4] * k + [float(P.lb[i, 0]) for i in P.fixed])
    ub = np.array([1e-8] * k + [float(P.ub[i, 0]) for i in P.fixed])
    x0 = xs[0].copy()
    bounds = list(zip(lb, ub))
    out = fmin_slsqp(lambda x: fun(x)[0], x0, f_ieqcons=lambda x: fun(x)[2], fprime=lambda x: fun(x)[1],
                     fprime_ieqcons=lambda x: fun(x)[3], bounds=bounds, iter=500, iprint=0, full_output=True, acc=1e-9)
    print("SLSQP : f = %.10f its %d mode %d" % (out[1], out[2], out[3]))
    rng = np.random.default_rng(0)
    for trial in range(4):
        xs0 = x0 * (1.0 + (0.05 * trial) * rng.uniform(-1, 1, x0.size))
        x, f, it, ok = sqp(fun, xs0, lb, ub, iters=200, acc=1e-8, verbose=(trial == 0 and "-v" in sys.argv))
        print("SQP   : f = %.10f its %d ok %d  |x - x_slsqp| %.2e  rel df %.2e" % (f, it, ok, np.abs(x - out[0]).max(), abs(f - out[1]) / abs(out[1])))


if __name__ == "__main__":
    main()
