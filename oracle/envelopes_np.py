"""Closed-form masonry envelopes (numpy).  TEST INFRASTRUCTURE.

``compas_tna.envelope`` is NOT vendored in the reference and is unpinned
(requirements.txt:1-3); the reference only *calls* it
(src/compas_tno/problems/constraints.py:123-126, jacobian.py:289,
bounds_update.py:63-64,128-129).  These closed forms were recovered from the
reference fixtures and are pinned by them at one
thickness each (tests/test_oracle.py): dome to 3e-14, cross vault to 2e-14.
The thickness derivatives are analytic derivatives of those forms; they are
pinned only against central finite differences.  PARITY UNPINNED against the
real compas_tna arithmetic.
"""
from __future__ import annotations

import numpy as np


class DomeEnvelopeNP:
    """Hemispherical dome, middle radius R, centre (xc, yc), thickness t.

    ub = sqrt((R + t/2)^2 - r^2);  lb = sqrt((R - t/2)^2 - r^2) where defined,
    else ``lb_outside`` (the fixtures store -1.0 there).
    """

    is_parametric = True

    def __init__(self, center=(5.0, 5.0), radius=5.0, thickness=0.5, lb_outside=-1.0, rho=20.0):
        self.xc, self.yc = float(center[0]), float(center[1])
        self.radius = float(radius)
        self.thickness = float(thickness)
        self.lb_outside = float(lb_outside)
        self.rho = rho

    def _r2(self, x, y):
        x = np.asarray(x, dtype=np.float64).reshape(-1)
        y = np.asarray(y, dtype=np.float64).reshape(-1)
        return (x - self.xc) ** 2 + (y - self.yc) ** 2

    def compute_bounds(self, x, y, thk):
        t = float(np.asarray(thk).reshape(-1)[0])
        r2 = self._r2(x, y)
        re, ri = self.radius + t / 2, self.radius - t / 2
        ub = np.sqrt(re * re - r2)
        inside = ri * ri - r2
        lb = np.where(inside > 0.0, np.sqrt(np.where(inside > 0.0, inside, 1.0)), self.lb_outside)
        return ub.reshape(-1, 1), lb.reshape(-1, 1)

    def compute_bounds_derivatives(self, x, y, thk):
        t = float(np.asarray(thk).reshape(-1)[0])
        ub, lb = self.compute_bounds(x, y, t)
        r2 = self._r2(x, y)
        re, ri = self.radius + t / 2, self.radius - t / 2
        dub = (re / (2.0 * ub.reshape(-1))).reshape(-1, 1)
        inside = ri * ri - r2
        dlb = np.where(inside > 0.0, -ri / (2.0 * np.where(inside > 0.0, lb.reshape(-1), 1.0)), 0.0).reshape(-1, 1)
        return dub, dlb

    def compute_bound_react(self, x, y, thk, fixed):
        t = float(np.asarray(thk).reshape(-1)[0])
        x = np.asarray(x, dtype=np.float64).reshape(-1)[fixed] - self.xc
        y = np.asarray(y, dtype=np.float64).reshape(-1)[fixed] - self.yc
        r = np.hypot(x, y)
        return (t / 2) * np.column_stack([x / r, y / r])


class CrossVaultEnvelopeNP:
    """Cross vault of the legacy fixtures: span L on [x0, x0+L]^2, springing angle 30 deg.

    rho = (L/2)/cos(30 deg); delta = x - xc where |y - yc| >= |x - xc| else y - yc;
    ub = sqrt((rho + t/2)^2 - delta^2); lb = sqrt((rho - t/2)^2 - delta^2).
    """

    is_parametric = True

    def __init__(self, x_span=(0.0, 10.0), y_span=(0.0, 10.0), thickness=0.5, lb_outside=-1.0, rho=20.0):
        self.xc = 0.5 * (x_span[0] + x_span[1])
        self.yc = 0.5 * (y_span[0] + y_span[1])
        self.radius = 0.5 * (x_span[1] - x_span[0]) / np.cos(np.radians(30.0))
        self.thickness = float(thickness)
        self.lb_outside = float(lb_outside)
        self.rho = rho

    def _delta2(self, x, y):
        dx = np.asarray(x, dtype=np.float64).reshape(-1) - self.xc
        dy = np.asarray(y, dtype=np.float64).reshape(-1) - self.yc
        d = np.where(np.abs(dy) >= np.abs(dx), dx, dy)
        return d * d

    def compute_middle(self, x, y):
        return np.sqrt(self.radius ** 2 - self._delta2(x, y)).reshape(-1, 1)

    def compute_bounds(self, x, y, thk):
        t = float(np.asarray(thk).reshape(-1)[0])
        d2 = self._delta2(x, y)
        re, ri = self.radius + t / 2, self.radius - t / 2
        ub = np.sqrt(re * re - d2)
        inside = ri * ri - d2
        lb = np.where(inside > 0.0, np.sqrt(np.where(inside > 0.0, inside, 1.0)), self.lb_outside)
        return ub.reshape(-1, 1), lb.reshape(-1, 1)

    def compute_bounds_derivatives(self, x, y, thk):
        t = float(np.asarray(thk).reshape(-1)[0])
        ub, lb = self.compute_bounds(x, y, t)
        d2 = self._delta2(x, y)
        re, ri = self.radius + t / 2, self.radius - t / 2
        dub = (re / (2.0 * ub.reshape(-1))).reshape(-1, 1)
        inside = ri * ri - d2
        dlb = np.where(inside > 0.0, -ri / (2.0 * np.where(inside > 0.0, lb.reshape(-1), 1.0)), 0.0).reshape(-1, 1)
        return dub, dlb
