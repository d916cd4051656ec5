"""ctypes binding of include/tno_b200.h.  Fails loudly when the CUDA library is missing."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libtno_b200.so")

ABI_VERSION = 2

# masks / enums (mirror include/tno_b200.h)
VAR_Q, VAR_ZB, VAR_T, VAR_LAMBDH, VAR_LAMBDV, VAR_XYB = 1, 2, 4, 8, 16, 32
CON_FUNICULAR, CON_ENVELOPE, CON_REAC_BOUNDS, CON_ENVELOPEXY = 1, 2, 4, 8
OBJ_NONE, OBJ_MIN, OBJ_MAX, OBJ_T, OBJ_NEG_LAST, OBJ_BESTFIT, OBJ_FEASIBILITY, OBJ_LOADPATH = 0, 1, 2, 3, 4, 5, 6, 7
OBJ_ECOMP_LINEAR, OBJ_ECOMP_NONLINEAR = 8, 9
ENV_FIXED, ENV_DOME, ENV_CROSSVAULT = 0, 1, 2
KERNEL_AUTO, KERNEL_INTERLEAVED, KERNEL_ONCHIP, KERNEL_INTERLEAVED_TPC = 0, 1, 2, 3
STATUS_OK, STATUS_NOT_PD, STATUS_NONFINITE = 0, 1, 2

EXPORTS = [
    "tno_abi_version", "tno_last_error", "tno_plan_create", "tno_plan_destroy", "tno_plan_query",
    "tno_plan_symbolic", "tno_eval_batch", "tno_eval_batch_host", "tno_plan_launch_count",
    "tno_plan_profile", "tno_plan_profile_read", "tno_symbolic_analyse", "tno_fetch_jacobian_host", "tno_device_peak", "tno_plan_fetch_stats",
    "tno_sqp_qp_solve", "tno_sqp_bfgs_update", "tno_preprocess_batch", "tno_preprocess_work_bytes",
]

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


class ProblemDesc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("n", C.c_int32), ("m", C.c_int32), ("nb", C.c_int32), ("k", C.c_int32),
        ("edges", _ip), ("fixed", _ip),
        ("B", _dp), ("d", _dp), ("P", _dp), ("X", _dp), ("lb", _dp), ("ub", _dp), ("s", _dp),
        ("qmin", _dp), ("qmax", _dp), ("b", _dp), ("px0", _dp), ("py0", _dp), ("pzv", _dp), ("pz0", _dp),
        ("variables", C.c_uint32), ("constraints", C.c_uint32), ("objective", C.c_int32),
        ("envelope_kind", C.c_int32), ("env_params", C.c_double * 8), ("thk", C.c_double),
        ("device", C.c_int32), ("kernel", C.c_int32), ("ordering", C.c_int32), ("reserved", C.c_int32),
        ("dXb", _dp), ("stiff", _dp), ("d_b", _dp), ("px0_b", _dp), ("py0_b", _dp),
        ("xlimits", _dp), ("ylimits", _dp),
    ]


class PlanInfo(C.Structure):
    _fields_ = [
        ("n", C.c_int32), ("m", C.c_int32), ("ni", C.c_int32), ("nb", C.c_int32), ("k", C.c_int32),
        ("nvar", C.c_int32), ("ng", C.c_int32), ("nrhs", C.c_int32), ("nnz_a", C.c_int32), ("nnz_l", C.c_int32),
        ("etree_height", C.c_int32), ("max_colcount", C.c_int32), ("kernel", C.c_int32), ("chunk", C.c_int32),
        ("scratch_bytes", C.c_int64), ("factor_flops", C.c_int64), ("solve_flops", C.c_int64), ("smem_bytes", C.c_int64),
        ("last_kernel", C.c_int32), ("onchip_ctas_per_sm", C.c_int32),
    ]

    def as_dict(self):
        return {name: int(getattr(self, name)) for name, _ in self._fields_}


class BatchInputs(C.Structure):
    _fields_ = [("pz_scale", C.c_void_p), ("load_dir", C.c_void_p), ("bounds", C.c_void_p)]


class Outputs(C.Structure):
    _fields_ = [("f", C.c_void_p), ("grad", C.c_void_p), ("g", C.c_void_p), ("J", C.c_void_p),
                ("z", C.c_void_p), ("q", C.c_void_p), ("gmin", C.c_void_p), ("status", C.c_void_p)]


_lib = None


def load():
    """Load libtno_b200.so, building it first when it is missing or when the sources changed since it was built
    (content hash recorded beside the library, see _build.needs_build; TNO_REBUILD=1 forces a build,
    TNO_NO_AUTOBUILD=1 loads whatever is there).  No fallback."""
    global _lib
    if _lib is not None:
        return _lib
    from ._build import build, needs_build
    if os.environ.get("TNO_REBUILD") == "1":
        build(force=True)
    elif not os.path.exists(LIB_PATH) or (os.environ.get("TNO_NO_AUTOBUILD") != "1" and needs_build()):
        build(force=False)
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("compas_tno_b200: CUDA library %s is missing and could not be built; "
                           "there is no CPU fallback" % LIB_PATH)
    path = LIB_PATH
    variant = os.environ.get("TNO_LIB_VARIANT")   # developer knob: a library built by tools/build_variant.py (A/B timing)
    if variant:
        path = os.path.join(os.path.dirname(LIB_PATH), "variants", "libtno_b200_%s.so" % variant)
        if not os.path.exists(path):
            raise RuntimeError("compas_tno_b200: variant library %s is missing" % path)
    lib = C.CDLL(path)
    lib.tno_abi_version.restype = C.c_int
    lib.tno_last_error.restype = C.c_char_p
    lib.tno_plan_create.argtypes = [C.POINTER(ProblemDesc), C.POINTER(C.c_void_p)]
    lib.tno_plan_create.restype = C.c_int
    lib.tno_plan_destroy.argtypes = [C.c_void_p]
    lib.tno_plan_destroy.restype = None
    lib.tno_plan_query.argtypes = [C.c_void_p, C.POINTER(PlanInfo)]
    lib.tno_plan_query.restype = C.c_int
    lib.tno_plan_symbolic.argtypes = [C.c_void_p, _ip, _ip, _ip, _ip, _ip]
    lib.tno_plan_symbolic.restype = C.c_int
    lib.tno_eval_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.POINTER(BatchInputs),
                                   C.POINTER(Outputs), C.c_void_p]
    lib.tno_eval_batch.restype = C.c_int
    lib.tno_eval_batch_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.POINTER(BatchInputs), C.POINTER(Outputs)]
    lib.tno_eval_batch_host.restype = C.c_int
    lib.tno_fetch_jacobian_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    lib.tno_fetch_jacobian_host.restype = C.c_int
    lib.tno_plan_launch_count.argtypes = [C.c_void_p]
    lib.tno_plan_launch_count.restype = C.c_int64
    lib.tno_plan_profile.argtypes = [C.c_void_p, C.c_int]
    lib.tno_plan_profile.restype = C.c_int
    lib.tno_plan_profile_read.argtypes = [C.c_void_p, C.c_int32, C.c_char_p, _dp, C.POINTER(C.c_int64)]
    lib.tno_plan_profile_read.restype = C.c_int
    lib.tno_symbolic_analyse.argtypes = [C.c_int32, C.c_int32, _ip, C.c_int32, C.c_int32, _ip, _ip, _ip, C.c_int32, _ip, _ip,
                                         C.c_int32, _ip]
    lib.tno_symbolic_analyse.restype = C.c_int
    lib.tno_device_peak.argtypes = [C.c_int32, C.c_int32, _dp]
    lib.tno_device_peak.restype = C.c_int
    lib.tno_plan_fetch_stats.argtypes = [C.c_void_p, _dp]
    lib.tno_plan_fetch_stats.restype = C.c_int
    vp = C.c_void_p
    lib.tno_sqp_qp_solve.argtypes = [C.c_int64, C.c_int32, C.c_int32, C.c_int32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp,
                                     vp, C.c_int32, C.c_double, C.c_int32, vp]
    lib.tno_sqp_qp_solve.restype = C.c_int
    lib.tno_sqp_bfgs_update.argtypes = [C.c_int64, C.c_int32, C.c_int32, C.c_int32, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.tno_sqp_bfgs_update.restype = C.c_int
    lib.tno_preprocess_work_bytes.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64]
    lib.tno_preprocess_work_bytes.restype = C.c_int64
    lib.tno_preprocess_batch.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, _ip, _ip, _ip, vp, vp, C.c_int64, C.c_int64,
                                         vp, vp, vp, vp, C.c_int64, C.c_int32, vp]
    lib.tno_preprocess_batch.restype = C.c_int
    if lib.tno_abi_version() != ABI_VERSION:
        raise RuntimeError("libtno_b200.so ABI %d != binding ABI %d" % (lib.tno_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


class TnoError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        msg = load().tno_last_error().decode("utf-8", "replace")
        if rc == -5:   # TNO_EUNSUPPORTED: the reference raises NotImplementedError for its unsupported combinations
            raise NotImplementedError("tno_b200: " + msg)
        if rc == -1:   # TNO_EINVAL: a plain argument error
            raise ValueError("tno_b200: " + msg)
        raise TnoError("tno_b200 error %d: %s" % (rc, msg))


def device_peak(what="fp64_fma", device=0):
    """TFLOP/s of the FP64 FMA pipe ("fp64_fma") or the FP64 tensor-core instruction ("fp64_mma"), measured in place."""
    v = C.c_double(0.0)
    check(load().tno_device_peak(int(device), {"fp64_fma": 0, "fp64_mma": 1}[what], C.byref(v)))
    return float(v.value)


def symbolic_analyse(n_free, edges, ordering=0, chain_width=4):
    """Host-only symbolic analysis (no GPU needed): dict(perm, colptr, rowidx, parent, chain_ptr, info)."""
    import numpy as np
    lib = load()
    e = np.ascontiguousarray(np.asarray(edges, dtype=np.int32).reshape(-1, 2))
    info = np.zeros(8, np.int32)
    p = lambda a: a.ctypes.data_as(_ip)  # noqa: E731
    check(lib.tno_symbolic_analyse(n_free, len(e), p(e), ordering, chain_width, None, None, None, 0, None, None, 0, p(info)))
    perm = np.zeros(n_free, np.int32)
    colptr = np.zeros(n_free + 1, np.int32)
    rowidx = np.zeros(int(info[0]), np.int32)
    parent = np.zeros(n_free, np.int32)
    chain_ptr = np.zeros(int(info[5]) + 1, np.int32)
    check(lib.tno_symbolic_analyse(n_free, len(e), p(e), ordering, chain_width, p(perm), p(colptr), p(rowidx), len(rowidx),
                                   p(parent), p(chain_ptr), len(chain_ptr), p(info)))
    keys = ("nnz_l", "etree_height", "max_colcount", "wide_levels", "wide_rows", "chains", "chain_levels", "factor_flops")
    return dict(perm=perm, colptr=colptr, rowidx=rowidx, parent=parent, chain_ptr=chain_ptr,
                info=dict(zip(keys, (int(v) for v in info))))
