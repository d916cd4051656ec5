"""Parametric envelopes handed to the CUDA library as (kind, params).

The reference calls ``compas_tna.envelope.Envelope.compute_bounds`` /
``compute_bounds_derivatives`` (src/compas_tno/problems/constraints.py:123-126,
jacobian.py:289); that package is not vendored, so the closed forms recovered from
the reference fixtures are used.  The host methods below
exist for set-up code (initial ub / lb); the per-candidate evaluation for a
variable thickness happens on the device (csrc/tno_interleaved.cu: envelope_eval).
"""
from __future__ import annotations

import numpy as np

from . import _lib


class _RadialEnvelope:
    is_parametric = True
    kind = _lib.ENV_FIXED

    def __init__(self, xc, yc, radius, thickness, lb_outside=-1.0, rho=20.0):
        self.xc, self.yc, self.radius = float(xc), float(yc), float(radius)
        self.thickness = float(thickness)
        self.lb_outside = float(lb_outside)
        self.rho = rho

    def abi_params(self):
        return [self.xc, self.yc, self.radius, self.lb_outside, 0.0, 0.0, 0.0, 0.0]

    def _d2(self, x, y):
        raise NotImplementedError

    def compute_bounds(self, x, y, thk):
        t = float(np.asarray(thk).reshape(-1)[0])
        d2 = self._d2(np.asarray(x, float).reshape(-1), np.asarray(y, float).reshape(-1))
        re, ri = self.radius + t / 2, self.radius - t / 2
        ub = np.sqrt(re * re - d2)
        ins = ri * ri - d2
        lb = np.where(ins > 0, np.sqrt(np.where(ins > 0, ins, 1.0)), self.lb_outside)
        return ub.reshape(-1, 1), lb.reshape(-1, 1)

    def compute_bounds_derivatives(self, x, y, thk):
        t = float(np.asarray(thk).reshape(-1)[0])
        ub, lb = self.compute_bounds(x, y, t)
        d2 = self._d2(np.asarray(x, float).reshape(-1), np.asarray(y, float).reshape(-1))
        re, ri = self.radius + t / 2, self.radius - t / 2
        ins = ri * ri - d2
        dub = re / (2 * ub.ravel())
        dlb = np.where(ins > 0, -ri / (2 * np.where(ins > 0, lb.ravel(), 1.0)), 0.0)
        return dub.reshape(-1, 1), dlb.reshape(-1, 1)

    def compute_middle(self, x, y):
        d2 = self._d2(np.asarray(x, float).reshape(-1), np.asarray(y, float).reshape(-1))
        return np.sqrt(np.maximum(self.radius ** 2 - d2, 0.0)).reshape(-1, 1)


class DomeEnvelope(_RadialEnvelope):
    """Hemispherical dome: ub / lb = sqrt((R +- t/2)^2 - r^2)."""

    kind = _lib.ENV_DOME

    def __init__(self, center=(5.0, 5.0), radius=5.0, thickness=0.5, lb_outside=-1.0, rho=20.0):
        super().__init__(center[0], center[1], radius, thickness, lb_outside, rho)

    def _d2(self, x, y):
        return (x - self.xc) ** 2 + (y - self.yc) ** 2


class CrossVaultEnvelope(_RadialEnvelope):
    """Cross vault of the reference fixtures (springing angle 30 degrees)."""

    kind = _lib.ENV_CROSSVAULT

    def __init__(self, x_span=(0.0, 10.0), y_span=(0.0, 10.0), thickness=0.5, lb_outside=-1.0, rho=20.0):
        radius = 0.5 * (x_span[1] - x_span[0]) / np.cos(np.radians(30.0))
        super().__init__(0.5 * (x_span[0] + x_span[1]), 0.5 * (y_span[0] + y_span[1]), radius, thickness, lb_outside, rho)

    def _d2(self, x, y):
        dx, dy = x - self.xc, y - self.yc
        d = np.where(np.abs(dy) >= np.abs(dx), dx, dy)
        return d * d


def abi_envelope(envelope):
    """(kind, params[8]) of an envelope object; duck-typed so that the oracle's / a user's class works too."""
    if envelope is None:
        return _lib.ENV_FIXED, [0.0] * 8
    if hasattr(envelope, "abi_params"):
        return envelope.kind, envelope.abi_params()
    name = type(envelope).__name__.lower()
    kind = _lib.ENV_DOME if "dome" in name else _lib.ENV_CROSSVAULT if "cross" in name else None
    if kind is None or not all(hasattr(envelope, a) for a in ("xc", "yc", "radius")):
        raise NotImplementedError("envelope %r has no closed form known to the CUDA library" % type(envelope).__name__)
    return kind, [float(envelope.xc), float(envelope.yc), float(envelope.radius),
                  float(getattr(envelope, "lb_outside", -1.0)), 0.0, 0.0, 0.0, 0.0]
