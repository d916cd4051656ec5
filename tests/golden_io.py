"""Load a golden case (tests/golden/*.npz) back into a host Problem + expected outputs."""
from __future__ import annotations

import os

import numpy as np

from compas_tno_b200.problems import FixedProblem

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

CASES = [
    "cross6_q_zb_min",
    "cross14_q_zb_min",
    "cross14_q_zb_t",
    "fan_q_zb_max",
    "dome_q_zb_reac_min",
    "dome_q_zb_t",
    "dome_q_zb_lambdh_reac",
    "dome_q_zb_lambdv",
    "freeform_q_zb_min",
    "cross14_q_zb_bestfit", "cross14_q_zb_feasibility", "cross14_q_zb_loadpath", "dome_q_zb_loadpath",
    "cross14_q_zb_ecomp_nonlinear", "dome_q_zb_ecomp_linear", "cross14_q_xyb_zb_min", "cross14_q_xyb_zb_t", "dome_q_xyb_zb_reac_min",
    "dome_q_zb_bestfit",
]


def var_offset(variables, k, nb, name):
    pos = k
    for v in ("xyb", "zb", "t", "lambdh", "lambdv"):
        if v == name:
            return pos
        if v in variables:
            pos += 2 * nb if v == "xyb" else (nb if v == "zb" else 1)
    raise KeyError(name)


def load_case(name, envelope_factory=None):
    """Return (problem, objective, xs, expected dict with full dense J)."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    P = FixedProblem.from_arrays(z["xyz"], z["edges"], z["fixed"], z["loads"], z["lb"], z["ub"], z["target"],
                                 z["q_form"], ind=z["ind"])
    # B, d as the golden generator had them (pinv can differ in the last bits across LAPACK builds)
    P.B = z["B"].copy()
    P.d = z["d"].reshape(-1, 1).copy()
    P.d0 = P.d.copy()
    P.variables = [str(v) for v in z["variables"]]
    P.constraints = [str(c) for c in z["constraints"]]
    P.features = ["fixed"]
    P.thk = float(z["thk"])
    for key in z.files:
        if key.startswith("extra_"):
            val = z[key]
            attr = key[len("extra_"):]
            setattr(P, attr, val.copy())
    if "extra_stiff" in z.files:
        P.Ecomp_method = "simplified"
    if "envelope" in z.files:
        kind = str(z["envelope"])
        if envelope_factory is None:
            from oracle.envelopes_np import CrossVaultEnvelopeNP, DomeEnvelopeNP
            envelope_factory = {"CrossVaultEnvelope": lambda t: CrossVaultEnvelopeNP((0.0, 10.0), (0.0, 10.0), t),
                                "DomeEnvelope": lambda t: DomeEnvelopeNP((5.0, 5.0), 5.0, t)}
        P.envelope = envelope_factory[kind](P.thk)
    objective = str(z["objective"])
    xs = z["x"]

    m, n, nb, k = P.m, P.n, P.nb, P.k
    N, nvar = xs.shape
    J = []
    for j in range(N):
        fun = np.zeros((2 * m, nvar))
        fun[:m, :k] = P.B
        fun[m:, :k] = -P.B
        if "lambdh" in P.variables:
            c = var_offset(P.variables, k, nb, "lambdh")
            fun[:m, c] = z["extra_d0"].ravel()
            fun[m:, c] = -z["extra_d0"].ravel()
        env = z["J_env"][j]
        neg = -env
        if "t" in P.variables:
            c = var_offset(P.variables, k, nb, "t")
            neg[:, c] = z["J_dub"][j]
        mid = [z["J_envxy"][j]] if "J_envxy" in z.files else []
        J.append(np.vstack([fun] + mid + [env, neg, z["J_reac"][j]]))
    expected = dict(g=z["g"], J=np.array(J), f=np.asarray(z["f"]).reshape(len(z["x"])), grad=z["grad"], X=z["X"], q=z["q"], mineig=z["mineig"])
    return P, objective, xs, expected


def rel_err(a, b):
    """max |a - b| / max |b|  -- the relative measure used for all parity statements."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = np.max(np.abs(b)) if b.size else 1.0
    if scale == 0.0:
        scale = 1.0
    return float(np.max(np.abs(a - b)) / scale) if b.size else 0.0


def row_blocks(P):
    """Row ranges of g / J by constraint block, in the reference's stacking order (problems/constraints.py:104-155):
    funicular (2m), envelopexy (4n), envelope (2n), reac_bounds (2nb).  The funicular and envelope blocks are further
    split into their two halves (q - qmin | qmax - q, z - lb | ub - z): the halves have very different magnitudes."""
    m, n, nb = P.m, P.n, P.nb
    out, pos = [], 0
    for name in ("funicular", "envelopexy", "envelope", "reac_bounds"):
        if name not in P.constraints:
            continue
        if name == "funicular":
            out += [("funicular:q-qmin", pos, pos + m), ("funicular:qmax-q", pos + m, pos + 2 * m)]
            pos += 2 * m
        elif name == "envelopexy":
            out += [("envelopexy:x", pos, pos + 2 * n), ("envelopexy:y", pos + 2 * n, pos + 4 * n)]
            pos += 4 * n
        elif name == "envelope":
            out += [("envelope:z-lb", pos, pos + n), ("envelope:ub-z", pos + n, pos + 2 * n)]
            pos += 2 * n
        else:
            out += [("reac_bounds", pos, pos + 2 * nb)]
            pos += 2 * nb
    return out


def elem_err(a, b, floor_frac=1e-3, min_scale=0.0):
    """Element-wise relative error with an absolute floor:  max_i |a_i - b_i| / max(|b_i|, floor), floor =
    max(floor_frac * max|b|, min_scale).  A small entry cannot hide behind the largest one of its block (the norm-wise
    rel_err lets it); min_scale is the floor of blocks that are zero up to rounding (see block_err)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if not b.size:
        return 0.0
    floor = max(floor_frac * float(np.max(np.abs(b))), min_scale)
    if floor == 0.0:
        return float(np.max(np.abs(a - b)))
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))


def block_err(a, b, min_scale):
    """Norm-wise relative error of one block, max|a - b| / max(max|b|, min_scale).  min_scale (1e-6 of the whole array's
    largest entry) keeps blocks that are zero up to rounding -- dx/dq_ind on a fixed plan is 1e-15 -- from being compared
    digit by digit, while a block five orders of magnitude below the largest one (g_reac ~ 0.1 next to q - qmin ~ 1e4) is
    still measured on its own scale."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if not b.size:
        return 0.0
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), min_scale, 1e-300))


def assert_blockwise(got_g, ref_g, got_J, ref_J, P, tol_val, tol_jac, what=""):
    """g and J against the reference block by block: norm-wise on the block's own scale and element-wise (absolute floor =
    1e-3 of the block's largest entry).  g / J have the candidate axis first or not at all."""
    g_floor = 1e-6 * float(np.max(np.abs(ref_g))) if np.size(ref_g) else 0.0
    J_floor = 1e-6 * float(np.max(np.abs(ref_J))) if (ref_J is not None and np.size(ref_J)) else 0.0
    for name, lo, hi in row_blocks(P):
        gg, rg = np.asarray(got_g)[..., lo:hi], np.asarray(ref_g)[..., lo:hi]
        assert block_err(gg, rg, g_floor) < tol_val, (what, "g", name, block_err(gg, rg, g_floor))
        assert elem_err(gg, rg, 1e-3, g_floor) < tol_val * 10, (what, "g elementwise", name, elem_err(gg, rg, 1e-3, g_floor))
        if got_J is not None:
            gj, rj = np.asarray(got_J)[..., lo:hi, :], np.asarray(ref_J)[..., lo:hi, :]
            assert block_err(gj, rj, J_floor) < tol_jac, (what, "J", name, block_err(gj, rj, J_floor))
            assert elem_err(gj, rj, 1e-3, J_floor) < tol_jac * 10, (what, "J elementwise", name, elem_err(gj, rj, 1e-3, J_floor))
