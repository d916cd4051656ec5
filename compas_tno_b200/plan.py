"""TnoPlan: one symbolic analysis per form diagram, batched evaluation on the GPU.

Host-side mirror of what the reference's solver loop does with a ``Problem``
(src/compas_tno/solvers/solver.py:126-129 fetches fobj / fgrad / fconstr / fjac and
calls them once per iterate); here one call evaluates N iterates.
torch is used only for device tensor handles, streams and pinned host buffers.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, Optional, Sequence

import numpy as np

from . import _lib
from .envelopes import abi_envelope

_VAR_BITS = {"q": _lib.VAR_Q, "xyb": _lib.VAR_XYB, "zb": _lib.VAR_ZB, "t": _lib.VAR_T, "n": _lib.VAR_T,
             "lambdh": _lib.VAR_LAMBDH, "lambdv": _lib.VAR_LAMBDV}
_CON_BITS = {"funicular": _lib.CON_FUNICULAR, "envelopexy": _lib.CON_ENVELOPEXY, "envelope": _lib.CON_ENVELOPE,
             "reac_bounds": _lib.CON_REAC_BOUNDS}
_OBJ = {None: _lib.OBJ_NONE, "none": _lib.OBJ_NONE, "min": _lib.OBJ_MIN, "max": _lib.OBJ_MAX, "t": _lib.OBJ_T,
        "n": _lib.OBJ_NEG_LAST, "s": _lib.OBJ_NEG_LAST, "max_load": _lib.OBJ_NEG_LAST,
        "bestfit": _lib.OBJ_BESTFIT, "target": _lib.OBJ_BESTFIT, "feasibility": _lib.OBJ_FEASIBILITY,
        "loadpath": _lib.OBJ_LOADPATH, "Ecomp-linear": _lib.OBJ_ECOMP_LINEAR,
        "Ecomp-nonlinear": _lib.OBJ_ECOMP_NONLINEAR}
_KERNELS = {"auto": _lib.KERNEL_AUTO, "interleaved": _lib.KERNEL_INTERLEAVED, "onchip": _lib.KERNEL_ONCHIP,
            "interleaved_tpc": _lib.KERNEL_INTERLEAVED_TPC}
_ORDERINGS = {"auto": 0, "natural": 1, "mindeg": 2, "nested": 3}

ALL_OUTPUTS = ("f", "grad", "g", "J", "z", "q", "gmin", "status")


def _f64(a, shape=None):
    if a is None:
        return None
    out = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    if shape is not None:
        out = out.reshape(shape)
    return out


def _edges_from_C(Cmat):
    """(m, 2) int32 (u, v) with C[e, u] = -1 and C[e, v] = +1, from a scipy sparse or dense C."""
    Cd = Cmat.tocoo() if hasattr(Cmat, "tocoo") else None
    if Cd is None:
        from scipy.sparse import coo_matrix
        Cd = coo_matrix(np.asarray(Cmat))
    m = Cd.shape[0]
    edges = np.full((m, 2), -1, dtype=np.int32)
    for r, c, v in zip(Cd.row, Cd.col, Cd.data):
        if v < 0:
            edges[r, 0] = c
        elif v > 0:
            edges[r, 1] = c
    if (edges < 0).any():
        raise ValueError("connectivity matrix rows must hold exactly one -1 and one +1")
    return edges


class TnoPlan:
    """Everything shared by all candidates of one form diagram, resident on one GPU.

    ``problem`` is any object with the reference's ``Problem`` attribute names
    (reference src/compas_tno/problems/problems.py:39-224): C, fixed, B, d (d0), P, X, lb, ub, s,
    qmin, qmax, variables, constraints, features and, as applicable, b, px0, py0, pzv, pz0, envelope, thk.
    """

    def __init__(self, problem, objective: Optional[str] = "min", device: int = 0, kernel: str = "auto",
                 ordering: str = "auto"):
        self._lib = _lib.load()
        self._handle = C.c_void_p()
        M = problem
        variables = list(M.variables or [])
        constraints = list(M.constraints or [])
        features = list(M.features or [])
        if "fixed" not in features:
            raise NotImplementedError("only the fixed-plan problem (features=['fixed']) is accelerated")
        bad = [v for v in variables if v not in _VAR_BITS] + [c for c in constraints if c not in _CON_BITS]
        if bad:
            raise NotImplementedError("not accelerated: %s" % bad)
        if "q" not in variables:
            variables = ["q"] + variables
        if objective not in _OBJ:
            raise NotImplementedError("objective %r is not accelerated" % (objective,))
        self.variables, self.constraints, self.objective = variables, constraints, objective

        n, m, nb, k = int(M.n), int(M.m), int(M.nb), int(M.k)
        keep = []  # keep numpy buffers alive until tno_plan_create returns

        def dptr(a, shape):
            if a is None:
                return None
            arr = _f64(a, shape)
            keep.append(arr)
            return arr.ctypes.data_as(C.POINTER(C.c_double))

        edges = np.ascontiguousarray(_edges_from_C(M.C), dtype=np.int32)
        fixed = np.ascontiguousarray(np.asarray(M.fixed, dtype=np.int32))
        keep += [edges, fixed]
        d_src = M.d0 if ("lambdh" in variables and getattr(M, "d0", None) is not None) else M.d
        env_kind, env_params = abi_envelope(getattr(M, "envelope", None)) if ("t" in variables or "n" in variables) \
            else (_lib.ENV_FIXED, [0.0] * 8)

        D = _lib.ProblemDesc()
        D.abi_version = _lib.ABI_VERSION
        D.n, D.m, D.nb, D.k = n, m, nb, k
        D.edges = edges.ctypes.data_as(C.POINTER(C.c_int32))
        D.fixed = fixed.ctypes.data_as(C.POINTER(C.c_int32))
        D.B = dptr(M.B, (m, k))
        D.d = dptr(d_src, (m,))
        D.P = dptr(M.P, (n, 3))
        D.X = dptr(M.X, (n, 3))
        D.lb = dptr(getattr(M, "lb", None), (n,))
        D.ub = dptr(getattr(M, "ub", None), (n,))
        D.s = dptr(getattr(M, "s", None), (n,))
        D.qmin = dptr(np.broadcast_to(np.asarray(M.qmin, dtype=np.float64).reshape(-1), (m,)), (m,))
        D.qmax = dptr(np.broadcast_to(np.asarray(M.qmax, dtype=np.float64).reshape(-1), (m,)), (m,))
        D.b = dptr(getattr(M, "b", None), (nb, 2)) if "reac_bounds" in constraints else None
        D.px0 = dptr(getattr(M, "px0", None), (n,)) if "lambdh" in variables else None
        D.py0 = dptr(getattr(M, "py0", None), (n,)) if "lambdh" in variables else None
        D.pzv = dptr(getattr(M, "pzv", None), (n,)) if "lambdv" in variables else None
        D.pz0 = dptr(getattr(M, "pz0", None), (n,)) if "lambdv" in variables else None
        if "envelopexy" in constraints:
            D.xlimits = dptr(getattr(M, "xlimits", None), (n, 2))
            D.ylimits = dptr(getattr(M, "ylimits", None), (n, 2))
        # second horizontal load basis (per-candidate load directions): d0_b, px0_b, py0_b next to d0, px0, py0
        if "lambdh" in variables and getattr(M, "d0_b", None) is not None:
            D.d_b = dptr(M.d0_b, (m,))
            D.px0_b = dptr(M.px0_b, (n,))
            D.py0_b = dptr(M.py0_b, (n,))
        if objective in ("Ecomp-linear", "Ecomp-nonlinear"):
            D.dXb = dptr(getattr(M, "dXb", None), (nb, 3))
        if objective == "Ecomp-nonlinear":
            if getattr(M, "Ecomp_method", "simplified") != "simplified":
                raise NotImplementedError("Ecomp_method 'complete' is not available (neither in problems/setup.py:156)")
            D.stiff = dptr(getattr(M, "stiff", None), (m,))
        D.variables = sum({_VAR_BITS[v] for v in variables})
        D.constraints = sum({_CON_BITS[c] for c in constraints})
        D.objective = _OBJ[objective]
        D.envelope_kind = env_kind
        for i, v in enumerate(env_params):
            D.env_params[i] = v
        D.thk = float(np.asarray(M.thk).reshape(-1)[0]) if getattr(M, "thk", None) is not None else 0.0
        D.device = int(device)
        D.kernel = _KERNELS[kernel]
        D.ordering = _ORDERINGS[ordering]
        _lib.check(self._lib.tno_plan_create(C.byref(D), C.byref(self._handle)))
        del keep
        self.device = int(device)
        info = _lib.PlanInfo()
        _lib.check(self._lib.tno_plan_query(self._handle, C.byref(info)))
        self._info = info
        self.n, self.m, self.ni, self.nb, self.k = n, m, int(info.ni), nb, k
        self.nvar, self.ng = int(info.nvar), int(info.ng)

    # ------------------------------------------------------------------------------------------
    @property
    def info(self) -> Dict[str, int]:
        info = _lib.PlanInfo()
        _lib.check(self._lib.tno_plan_query(self._handle, C.byref(info)))
        return info.as_dict()

    @property
    def fetch_stats(self) -> Dict[str, float]:
        """Split and measured rates of the most recent Jacobian fetch to host memory (tno_plan_fetch_stats)."""
        st = (C.c_double * 4)()
        _lib.check(self._lib.tno_plan_fetch_stats(self._handle, st))
        return {"host_fill_fraction": float(st[0]), "host_fill_gb_per_s": float(st[1]), "dma_gb_per_s": float(st[2]),
                "host_threads": int(st[3])}

    @property
    def launch_count(self) -> int:
        return int(self._lib.tno_plan_launch_count(self._handle))

    @property
    def bytes_per_eval(self) -> int:
        """Algorithmic bytes of one evaluation: x in; z, f, grad, g, dense J out."""
        return 8 * (self.nvar + self.n + 1 + self.nvar + self.ng + self.ng * self.nvar)

    def profile(self, enable: bool = True):
        """Start (and reset) or stop per-kernel CUDA-event timing inside the library."""
        _lib.check(self._lib.tno_plan_profile(self._handle, 1 if enable else 0))

    def profile_read(self):
        """{kernel name: (total_ms, launches)} of the kernels timed since ``profile(True)``."""
        kmax = 32
        names = C.create_string_buffer(32 * kmax)
        ms = (C.c_double * kmax)()
        cnt = (C.c_int64 * kmax)()
        nk = self._lib.tno_plan_profile_read(self._handle, kmax, names, ms, cnt)
        if nk < 0:
            _lib.check(nk)
        out = {}
        for i in range(nk):
            if cnt[i]:
                out[names.raw[32 * i:32 * i + 32].split(b"\0")[0].decode()] = (float(ms[i]), int(cnt[i]))
        return out

    def symbolic(self):
        i = self._info
        perm = np.zeros(i.ni, np.int32)
        colptr = np.zeros(i.ni + 1, np.int32)
        rowidx = np.zeros(i.nnz_l, np.int32)
        parent = np.zeros(i.ni, np.int32)
        level = np.zeros(i.ni, np.int32)
        p = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))  # noqa: E731
        _lib.check(self._lib.tno_plan_symbolic(self._handle, p(perm), p(colptr), p(rowidx), p(parent), p(level)))
        return dict(perm=perm, colptr=colptr, rowidx=rowidx, parent=parent, level=level)

    def close(self):
        if getattr(self, "_handle", None) and self._handle.value:
            self._lib.tno_plan_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _shape(self, name, N):
        return {"f": (N,), "grad": (N, self.nvar), "g": (N, self.ng), "J": (N, self.ng, self.nvar),
                "z": (N, self.n), "q": (N, self.m), "gmin": (N,), "status": (N,)}[name]

    # ------------------------------------------------------------------------------------------
    def evaluate_host(self, x, outputs: Iterable[str] = ALL_OUTPUTS, pz_scale=None, load_dir=None, bounds=None) -> Dict[str, np.ndarray]:
        """Host buffers in, host buffers out (tno_eval_batch_host); no torch involved.
        ``pz_scale`` (N,): vertical load multiplier per candidate; ``load_dir`` (N, 2): (cos, sin) of the horizontal
        load direction per candidate (problems with ``lambdh`` and a second load basis ``d0_b, px0_b, py0_b``);
        ``bounds`` (N, 2n) = [ub | lb] per candidate replacing the problem's envelope in g (no thickness variable)."""
        x = _f64(x)
        if x.ndim == 1:
            x = x.reshape(1, -1)
        N, ld = x.shape
        if ld != self.nvar:
            raise ValueError("x has %d columns, the problem has %d variables" % (ld, self.nvar))
        res, O = {}, _lib.Outputs()
        for name in outputs:
            arr = np.empty(self._shape(name, N), dtype=np.int32 if name == "status" else np.float64)
            res[name] = arr
            setattr(O, name, arr.ctypes.data)
        inp = _lib.BatchInputs()
        pz = _f64(pz_scale, (N,)) if pz_scale is not None else None
        ld2 = _f64(load_dir, (N, 2)) if load_dir is not None else None
        if pz is not None:
            inp.pz_scale = pz.ctypes.data
        if ld2 is not None:
            inp.load_dir = ld2.ctypes.data
        bd = _f64(bounds, (N, 2 * self.n)) if bounds is not None else None
        if bd is not None:
            inp.bounds = bd.ctypes.data
        _lib.check(self._lib.tno_eval_batch_host(self._handle, x.ctypes.data, N, ld, C.byref(inp), C.byref(O)))
        return res

    def evaluate_device(self, x, outputs: Iterable[str] = ALL_OUTPUTS, pz_scale=None, out=None, stream=None, load_dir=None,
                        bounds=None):
        """Device tensors in, device tensors out (tno_eval_batch); asynchronous on the current torch stream."""
        import torch
        if not x.is_cuda or x.dtype != torch.float64 or x.dim() != 2 or x.stride(1) != 1:
            raise ValueError("x must be a 2-D float64 CUDA tensor with unit inner stride")
        if x.device.index != self.device:
            raise ValueError("x lives on cuda:%s, the plan on cuda:%d" % (x.device.index, self.device))
        N = x.shape[0]
        if x.shape[1] != self.nvar:
            raise ValueError("x has %d columns, the problem has %d variables" % (x.shape[1], self.nvar))
        res, O = {}, _lib.Outputs()
        for name in outputs:
            t = None if out is None else out.get(name)
            if t is None:
                t = torch.empty(self._shape(name, N), device=x.device,
                                dtype=torch.int32 if name == "status" else torch.float64)
            elif tuple(t.shape) != self._shape(name, N) or not t.is_contiguous():
                raise ValueError("bad preallocated output %r" % name)
            res[name] = t
            setattr(O, name, t.data_ptr())
        opt = _lib.BatchInputs()
        if pz_scale is not None:
            opt.pz_scale = pz_scale.data_ptr()
        if load_dir is not None:
            if tuple(load_dir.shape) != (N, 2) or load_dir.dtype != torch.float64 or not load_dir.is_contiguous() or not load_dir.is_cuda:
                raise ValueError("load_dir must be a contiguous (N, 2) float64 CUDA tensor")
            opt.load_dir = load_dir.data_ptr()
        if bounds is not None:
            if tuple(bounds.shape) != (N, 2 * self.n) or bounds.dtype != torch.float64 or not bounds.is_contiguous() or not bounds.is_cuda:
                raise ValueError("bounds must be a contiguous (N, 2n) float64 CUDA tensor [ub | lb]")
            opt.bounds = bounds.data_ptr()
        st = torch.cuda.current_stream(x.device).cuda_stream if stream is None else stream
        _lib.check(self._lib.tno_eval_batch(self._handle, x.data_ptr(), N, x.stride(0), C.byref(opt), C.byref(O),
                                            C.c_void_p(st)))
        return res

    def evaluate_batch(self, x_host, to_host: Sequence[str] = ("f", "gmin", "status"),
                       on_device: Sequence[str] = ("grad", "g", "J", "z"), pz_scale=None, buffers=None, load_dir=None,
                       pipeline: int = 4, bounds=None):
        """End-to-end call: pinned host x -> device, evaluate, ``to_host`` outputs -> pinned host.

        Outputs named in ``on_device`` are materialised in HBM (for an on-device consumer) and returned as
        CUDA tensors; ``to_host`` ones come back as pinned CPU tensors.  Synchronises before returning.
        Large batches are evaluated in ``pipeline`` slices so that the device-to-host copy of one slice (on a second
        stream) overlaps the evaluation of the next; results are identical to the unsliced call.
        """
        import torch
        dev = torch.device("cuda", self.device)
        xh = x_host if isinstance(x_host, torch.Tensor) else torch.from_numpy(_f64(x_host))
        if xh.dim() == 1:
            xh = xh.reshape(1, -1)
        N = xh.shape[0]
        buffers = buffers if buffers is not None else {}
        xd = buffers.get("x_dev")
        if xd is None or xd.shape != xh.shape:
            xd = torch.empty(xh.shape, dtype=torch.float64, device=dev)
            buffers["x_dev"] = xd
        xd.copy_(xh, non_blocking=True)

        def to_dev(a):
            if a is None:
                return None
            t = a if isinstance(a, torch.Tensor) else torch.from_numpy(_f64(a))
            return t.to(dev, non_blocking=True)

        pzd, ldd, bdd = to_dev(pz_scale), to_dev(load_dir), to_dev(bounds)
        names = tuple(dict.fromkeys(tuple(to_host) + tuple(on_device)))
        out = buffers.get("out")
        if out is None or any(k not in out or tuple(out[k].shape) != self._shape(k, N) for k in names):
            out = {k: torch.empty(self._shape(k, N), device=dev, dtype=torch.int32 if k == "status" else torch.float64)
                   for k in names}
            buffers["out"] = out
        host = {}
        for name in to_host:
            h = buffers.get("h_" + name)
            if h is None or h.shape != out[name].shape:
                h = torch.empty(out[name].shape, dtype=out[name].dtype, pin_memory=True)
                buffers["h_" + name] = h
            host[name] = h
        d2h_bytes = sum(host[k].numel() * host[k].element_size() for k in host)
        nslices = pipeline if (pipeline > 1 and N >= 256 * pipeline and d2h_bytes >= (64 << 20)) else 1
        main = torch.cuda.current_stream(dev)

        def fetch(name, i0, i1, stream):
            """device -> pinned host for one output slice; J goes through tno_fetch_jacobian_host (constant rows filled by
            host threads instead of crossing PCIe) unless its funicular rows vary with a per-candidate load direction"""
            if name == "J" and ldd is None and i1 > i0:
                _lib.check(self._lib.tno_fetch_jacobian_host(self._handle, out["J"][i0:i1].data_ptr(), host["J"][i0:i1].data_ptr(),
                                                      i1 - i0, C.c_void_p(stream.cuda_stream)))
            else:
                with torch.cuda.stream(stream):
                    host[name][i0:i1].copy_(out[name][i0:i1], non_blocking=True)

        if nslices == 1:
            self.evaluate_device(xd, names, pz_scale=pzd, out=out, load_dir=ldd, bounds=bdd)
            for name in to_host:
                fetch(name, 0, N, main)
        else:
            copy_stream = buffers.get("copy_stream")
            if copy_stream is None:
                copy_stream = torch.cuda.Stream(device=dev)
                buffers["copy_stream"] = copy_stream
            copy_stream.wait_stream(main)          # earlier users of the host / device buffers on either stream
            step = -(-N // nslices)
            for i0 in range(0, N, step):
                i1 = min(N, i0 + step)
                self.evaluate_device(xd[i0:i1], names, out={k: out[k][i0:i1] for k in names},
                                     pz_scale=None if pzd is None else pzd[i0:i1],
                                     load_dir=None if ldd is None else ldd[i0:i1],
                                     bounds=None if bdd is None else bdd[i0:i1])
                done = torch.cuda.Event()
                done.record(main)
                copy_stream.wait_event(done)
                for name in sorted(to_host, key=lambda k: k == "J"):   # small outputs first, J (with its host fill) last
                    fetch(name, i0, i1, copy_stream)
            main.wait_stream(copy_stream)
        main.synchronize()
        ret = {name: (host[name] if name in host else out[name]) for name in names}
        return ret
