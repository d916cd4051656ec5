"""Drop-in replacements for the four callables the reference stores on ``Problem``.

Same names, argument order and return shapes as the reference's
``constr_wrapper`` (src/compas_tno/problems/constraints.py:20),
``sensitivities_wrapper`` (problems/jacobian.py:53), ``f_min_thrust`` / ``f_max_thrust`` /
``f_reduce_thk`` / ``f_tight_crosssection`` (problems/objectives.py:91,151,415,438) and their
gradients (problems/derivatives.py:223,336,183,203).  Each call evaluates a batch of one on the
GPU; the four callbacks of one solver iterate share a single launch sequence through a one-entry
cache keyed on the bytes of x (the reference factorises the same matrix three times instead,
SURVEY.md Appendix D.8).

``accelerate(problem, objective)`` performs what ``set_up_general_optimisation`` does at
src/compas_tno/problems/setup.py:466-469, with these shims instead of the numpy callbacks.
"""
from __future__ import annotations

import numpy as np

from .plan import TnoPlan

__all__ = [
    "accelerate", "constr_wrapper", "sensitivities_wrapper", "f_min_thrust", "f_max_thrust", "gradient_fmin",
    "gradient_fmax", "f_reduce_thk", "gradient_reduce_thk", "f_tight_crosssection", "gradient_tight_crosssection",
    "f_bestfit", "gradient_bestfit", "f_loadpath_general", "gradient_loadpath", "f_complementary_energy", "gradient_complementary_energy",
    "f_complementary_energy_nonlinear", "gradient_complementary_energy_nonlinear", "f_constant", "gradient_feasibility", "objective_selector", "evaluate_batch",
]

_ATTR = "_tno_b200_state"


class _State:
    def __init__(self, problem, objective, device, kernel):
        self.plan = TnoPlan(problem, objective=objective, device=device, kernel=kernel)
        self.objective = objective
        self.key = None
        self.res = None
        self.warned = False

    def eval1(self, x, problem):
        x = np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(-1))
        key = x.tobytes()
        if key != self.key:
            self.res = self.plan.evaluate_host(x.reshape(1, -1), ("f", "grad", "g", "J", "z", "q", "status"))
            self.key = key
            # the reference's callbacks leave q and the updated geometry on the problem (constraints.py:90-95)
            try:
                problem.q = self.res["q"][0].reshape(-1, 1).copy()
                problem.X[:, 2] = self.res["z"][0]
                if "envelopexy" in self.plan.constraints:
                    # x and y of the moved geometry come back inside g: rows x - xlim0 and y - ylim0
                    g, m, n = self.res["g"][0], self.plan.m, self.plan.n
                    r0 = 2 * m if "funicular" in self.plan.constraints else 0
                    problem.X[:, 0] = g[r0: r0 + n] + np.asarray(problem.xlimits)[:, 0]
                    problem.X[:, 1] = g[r0 + 2 * n: r0 + 3 * n] + np.asarray(problem.ylimits)[:, 0]
            except (AttributeError, TypeError, ValueError) as exc:
                # a frozen / read-only problem object cannot take the state write-back; the returned values are unaffected
                if not self.warned:
                    import warnings
                    warnings.warn("tno_b200: problem.q / problem.X not updated after an evaluation (%s: %s)" % (type(exc).__name__, exc))
                    self.warned = True
        return self.res


def _state(problem, objective=None):
    if isinstance(problem, list):  # the solvers pass args=[problem]
        problem = problem[0]
    st = getattr(problem, _ATTR, None)
    if st is None or (objective is not None and st.objective != objective):
        device = getattr(problem, "_tno_b200_device", 0)
        kernel = getattr(problem, "_tno_b200_kernel", "auto")
        st = _State(problem, objective if objective is not None else "min", device, kernel)
        setattr(problem, _ATTR, st)
    return st, problem


def constr_wrapper(variables, M):
    st, M = _state(M)
    return st.eval1(variables, M)["g"][0].copy()


def sensitivities_wrapper(variables, M):
    st, M = _state(M)
    return st.eval1(variables, M)["J"][0].copy()


def _objective(name):
    def fobj(variables, problem):
        st, problem = _state(problem, name)
        return float(st.eval1(variables, problem)["f"][0])

    def fgrad(variables, problem):
        st, problem = _state(problem, name)
        return st.eval1(variables, problem)["grad"][0].copy()

    return fobj, fgrad


f_min_thrust, gradient_fmin = _objective("min")
f_max_thrust, gradient_fmax = _objective("max")
f_reduce_thk, gradient_reduce_thk = _objective("t")
f_tight_crosssection, gradient_tight_crosssection = _objective("n")
f_bestfit, gradient_bestfit = _objective("bestfit")
f_constant, gradient_feasibility = _objective("feasibility")
f_loadpath_general, gradient_loadpath = _objective("loadpath")
f_complementary_energy, gradient_complementary_energy = _objective("Ecomp-linear")
f_complementary_energy_nonlinear, gradient_complementary_energy_nonlinear = _objective("Ecomp-nonlinear")

_SELECT = {
    "min": (f_min_thrust, gradient_fmin),
    "max": (f_max_thrust, gradient_fmax),
    "t": (f_reduce_thk, gradient_reduce_thk),
    "s": (f_tight_crosssection, gradient_tight_crosssection),
    "n": (f_tight_crosssection, gradient_tight_crosssection),
    "max_load": (f_tight_crosssection, gradient_tight_crosssection),
    "bestfit": (f_bestfit, gradient_bestfit),
    "target": (f_bestfit, gradient_bestfit),
    "feasibility": (f_constant, gradient_feasibility),
    "loadpath": (f_loadpath_general, gradient_loadpath),
    "Ecomp-linear": (f_complementary_energy, gradient_complementary_energy),
    "Ecomp-nonlinear": (f_complementary_energy_nonlinear, gradient_complementary_energy_nonlinear),
}


def objective_selector(objective):
    """Same contract as problems/objectives.py:27-88, restricted to the accelerated objectives."""
    if objective not in _SELECT:
        print("Please, provide a valid objective for the optimisation")
        raise NotImplementedError
    return _SELECT[objective]


def accelerate(problem, objective="min", device=0, kernel="auto"):
    """Install the GPU callbacks on ``problem`` (reference setup.py:466-469) and return it."""
    problem._tno_b200_device = device
    problem._tno_b200_kernel = kernel
    if hasattr(problem, _ATTR):
        delattr(problem, _ATTR)
    objective = {"s": "n", "max_load": "n", "target": "bestfit"}.get(objective, objective)
    fobj, fgrad = objective_selector(objective)
    _state(problem, objective)
    problem.fobj, problem.fgrad = fobj, fgrad
    problem.fconstr, problem.fjac = constr_wrapper, sensitivities_wrapper
    return problem


def evaluate_batch(problem, xs, objective="min", outputs=("f", "grad", "g", "J", "z", "status"), device=0):
    """The batched entry point the reference does not have: all callbacks for every row of xs."""
    st, problem = _state(problem, {"s": "n", "max_load": "n", "target": "bestfit"}.get(objective, objective))
    return st.plan.evaluate_host(xs, outputs)
