"""Import the *unmodified* reference callbacks from /root/reference/src.

TEST INFRASTRUCTURE (golden-vector generation only; never shipped, never timed).

The reference (compas_tno 0.3.0) imports ``compas``, ``compas_tna``, ``compas_fd``
and ``cvxpy`` at module level; none of them is installed here and none of them is
*called* by the hot-path callbacks.  We
register empty placeholder modules so that the import succeeds, then hand back
the reference's own functions.  This only works where /root/reference exists
(the build container); the GPU box uses the committed golden vectors instead.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_SRC = os.environ.get("TNO_REFERENCE_SRC", "/root/reference/src")
REFERENCE_DATA = os.environ.get("TNO_REFERENCE_DATA", "/root/reference/data")


class _Placeholder:
    """Accepts any construction / attribute access; used for type names only."""

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        return _Placeholder()

    def __call__(self, *a, **k):
        return _Placeholder()


def _never(*a, **k):
    raise RuntimeError("third-party stub called on the hot path: oracle assumption violated")


def _register(name, **attrs):
    if name in sys.modules and not getattr(sys.modules[name], "__tno_stub__", False):
        return  # a real package is installed; leave it alone
    mod = types.ModuleType(name)
    mod.__tno_stub__ = True
    mod.__dict__.update(attrs)
    sys.modules[name] = mod


_GEOMETRY_NAMES = (
    "Point Frame Scale Translation Vector Line Polyline Plane Rotation Transformation "
    "centroid_points cross_vectors length_vector subtract_vectors add_vectors scale_vector "
    "normalize_vector norm_vector dot_vectors angle_vectors distance_point_point "
    "closest_point_on_line_xy distance_point_line_xy distance_point_point_xy "
    "closest_point_in_cloud intersection_segment_segment_xy intersection_line_line_xy "
    "sort_points_xy delaunay_triangulation is_point_on_line_xy mirror_points_line rotate_points "
    "is_point_in_convex_polygon_xy midpoint_point_point"
).split()


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "compas_tno"))


def load_reference():
    """Return a namespace with the reference's hot-path callables."""
    if not available():
        raise RuntimeError("reference sources not present at %s" % REFERENCE_SRC)

    _register("compas")
    _register("compas.data", Data=object)
    _register("compas.datastructures", Mesh=_Placeholder)
    _register("compas.matrices", connectivity_matrix=_never)
    _register(
        "compas.geometry",
        **{n: (_Placeholder if n[0].isupper() else _never) for n in _GEOMETRY_NAMES},
    )
    _register("compas.colors", Color=_Placeholder)
    _register("compas.utilities", geometric_key=_never, geometric_key_xy=_never)
    _register("compas.itertools", pairwise=_never)
    _register("compas_tna")
    _register("compas_tna.diagrams", FormDiagram=_Placeholder, ForceDiagram=_Placeholder)
    _register("compas_tna.envelope", Envelope=_Placeholder)
    _register("compas_fd")
    _register("compas_fd.solvers", fd_numpy=_never)
    _register(
        "cvxpy",
        Minimize=_never,
        Problem=_Placeholder,
        diag=_never,
        matrix_frac=_never,
        Variable=_Placeholder,
    )
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)

    import compas_tno.algorithms as alg  # noqa: E402
    import compas_tno.problems as prb  # noqa: E402

    ns = types.SimpleNamespace(
        Problem=prb.Problem,
        FixedProblem=prb.FixedProblem,
        constr_wrapper=prb.constr_wrapper,
        sensitivities_wrapper=prb.sensitivities_wrapper,
        f_min_thrust=prb.f_min_thrust,
        f_max_thrust=prb.f_max_thrust,
        gradient_fmin=prb.gradient_fmin,
        gradient_fmax=prb.gradient_fmax,
        f_constant=prb.f_constant,
        gradient_feasibility=prb.gradient_feasibility,
        f_loadpath_general=prb.f_loadpath_general,
        gradient_loadpath=prb.gradient_loadpath,
        f_complementary_energy=prb.f_complementary_energy,
        gradient_complementary_energy=prb.gradient_complementary_energy,
        f_complementary_energy_nonlinear=prb.f_complementary_energy_nonlinear,
        gradient_complementary_energy_nonlinear=prb.gradient_complementary_energy_nonlinear,
        f_bestfit=prb.f_bestfit,
        gradient_bestfit=prb.gradient_bestfit,
        f_reduce_thk=prb.f_reduce_thk,
        gradient_reduce_thk=prb.gradient_reduce_thk,
        f_tight_crosssection=prb.f_tight_crosssection,
        gradient_tight_crosssection=prb.gradient_tight_crosssection,
        objective_selector=prb.objective_selector,
        d_fconstr=prb.d_fconstr,
        d_fobj=prb.d_fobj,
        q_from_variables=alg.q_from_variables,
        xyz_from_q=alg.xyz_from_q,
        find_independents=alg.find_independents,
        compute_reactions=alg.compute_reactions,
    )
    return ns
