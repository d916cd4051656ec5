"""CPU: the synthetic benchmark workloads (compas_tno_b200/workloads.py) produce compression-only equilibrium
states, i.e. candidates every kernel family can factor without pivoting."""
import numpy as np
import pytest
from scipy.sparse import diags

from compas_tno_b200 import workloads as W


def _min_eig(P, q):
    A = -(P.Cit @ diags(q) @ P.Ci).toarray()
    return float(np.linalg.eigvalsh(A).min())


@pytest.mark.parametrize("n_div", [6, 10])
def test_crossvault_candidates_are_compressive(n_div):
    P, xs = W.crossvault_problem(n_div, n_states=4)
    X = W.sample_candidates(P, xs, 16, seed=3, basis=P.candidate_basis)
    assert X.shape == (16, P.k + P.nb)
    for x in X:
        q = (P.B @ x[: P.k].reshape(-1, 1) + P.d).ravel()
        assert q.max() <= 1e-9
        assert _min_eig(P, q) > 0.0
        assert np.abs(P.E @ q.reshape(-1, 1) - P.ph).max() < 1e-9      # horizontal equilibrium of the fixed plan


def test_dome_lambdh_candidates_are_compressive_and_in_equilibrium():
    P, xs = W.dome_lambdh_problem(8, 8, n_states=4)
    assert P.variables == ["q", "zb", "lambdh"] and len(xs) == P.k + P.nb + 1
    X = W.lambdh_candidates(P, xs, 12, seed=5)
    for x in X:
        lam = x[-1]
        q = (P.B @ x[: P.k].reshape(-1, 1) + lam * P.d0).ravel()
        assert q.max() <= 1e-9
        assert _min_eig(P, q) > 0.0
        # E q = lambdh * ph : the horizontal loads scale with the multiplier
        assert np.abs(P.E @ q.reshape(-1, 1) - lam * P.ph).max() < 1e-8


def test_direction_sweep_candidates_follow_their_direction():
    P, xs = W.dome_direction_problem(8, 8, n_states=4)
    X, load_dir = W.direction_candidates(P, xs, 8, seed=5)
    assert np.allclose(np.hypot(load_dir[:, 0], load_dir[:, 1]), 1.0)
    w = np.abs(P.P[P.free, 2:3])
    for x, (c, s) in zip(X, load_dir):
        lam = x[-1]
        q = (P.B @ x[: P.k].reshape(-1, 1) + lam * (c * P.d0 + s * P.d0_b)).ravel()
        assert q.max() <= 1e-9
        assert _min_eig(P, q) > 0.0
        ph = np.vstack([c * w, s * w])
        assert np.abs(P.E @ q.reshape(-1, 1) - lam * ph).max() < 1e-8
