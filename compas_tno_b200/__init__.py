"""compas_tno_b200 -- B200-native batched evaluation of compas_tno's optimisation callbacks.

Scope: equilibrium z(q), constraints g, objective f, grad f and the dense
constraint Jacobian for N candidate variable vectors sharing one form diagram, in fp64, on sm_100a.
"""
from .problems import FixedProblem, Problem, cross_form_arrays, find_independents_QR  # noqa: F401
from .envelopes import CrossVaultEnvelope, DomeEnvelope  # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):
    # the CUDA-backed pieces are imported lazily so that host-only tooling can import the package
    if name == "TnoPlan":
        from .plan import TnoPlan
        return TnoPlan
    if name in ("accelerate", "evaluate_batch"):
        from . import callbacks
        return getattr(callbacks, name)
    raise AttributeError(name)
