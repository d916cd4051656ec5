"""Read the reference's serialized FormDiagrams (``/root/reference/data/*.json``).

TEST INFRASTRUCTURE (golden-vector generation only).  Produces plain arrays in
the vertex / edge order the reference's ``Problem.from_formdiagram`` would see
(src/compas_tno/problems/problems.py:250-300): vertices in dict order, edges in
``edgedata`` order skipping ``_is_edge: false``, supports in vertex order.
"""
from __future__ import annotations

import ast
import json
import os

import numpy as np

from .refstub import REFERENCE_DATA


def _load_json(name):
    path = os.path.join(REFERENCE_DATA, name)
    text = open(path).read()
    try:
        return json.loads(text)
    except json.JSONDecodeError:
        # data/analysis.json is truncated right after '"message": ' 
        return json.loads(text + '""}}')


def _mesh_block(doc, name):
    """Return (vertex dict, edgedata dict, defaults_v, defaults_e, attributes)."""
    if name == "analysis.json":
        doc = doc["form"]
    if "data" in doc and "vertex" in doc["data"]:
        d = doc["data"]
    elif "value" in doc and "vertex" in doc["value"]:
        d = doc["value"]
    else:
        d = doc
    dva = d.get("default_vertex_attributes", d.get("dva", {}))
    dea = d.get("default_edge_attributes", d.get("dea", {}))
    return d["vertex"], d["edgedata"], dva, dea, d.get("attributes", {})


def load_form(name):
    """Arrays of one fixture: dict with xyz, edges, fixed, P, lb, ub, target, q, ind, b, r, attrs."""
    doc = _load_json(name)
    vertex, edgedata, dva, dea, attrs = _mesh_block(doc, name)
    keys = list(vertex.keys())
    k_i = {key: i for i, key in enumerate(keys)}
    n = len(keys)

    def vget(key, a, default=0.0):
        v = vertex[key].get(a, dva.get(a, default))
        return default if v is None else v

    xyz = np.array([[vget(k, "x"), vget(k, "y"), vget(k, "z")] for k in keys], dtype=np.float64)
    P = np.array([[vget(k, "px"), vget(k, "py"), vget(k, "pz")] for k in keys], dtype=np.float64)
    lb = np.array([vget(k, "lb", np.nan) for k in keys], dtype=np.float64)
    ub = np.array([vget(k, "ub", np.nan) for k in keys], dtype=np.float64)
    target = np.array([vget(k, "target") for k in keys], dtype=np.float64)
    is_sup = [bool(vget(k, "is_fixed", False)) or bool(vget(k, "is_anchor", False)) or bool(vget(k, "is_support", False)) for k in keys]
    fixed = [i for i in range(n) if is_sup[i]]
    react = np.array([[vget(k, "_rx"), vget(k, "_ry"), vget(k, "_rz")] for k in keys], dtype=np.float64)[fixed]
    bvals = []
    for i in fixed:
        bv = vertex[keys[i]].get("b", None)
        bvals.append(bv if bv is not None else [np.nan, np.nan])
    b = np.array(bvals, dtype=np.float64).reshape(len(fixed), 2)

    edges, q, is_ind = [], [], []
    for ekey, ed in edgedata.items():
        if not ed.get("_is_edge", dea.get("_is_edge", True)):
            continue
        u, v = ast.literal_eval(ekey)
        edges.append((k_i[str(u)], k_i[str(v)]))
        q.append(ed.get("q", dea.get("q", -1.0)))
        is_ind.append(bool(ed.get("is_ind", dea.get("is_ind", False))))
    edges = np.array(edges, dtype=np.int64)
    q = np.array(q, dtype=np.float64)
    ind = [e for e, f in enumerate(is_ind) if f]

    out = dict(xyz=xyz, edges=edges, fixed=np.array(fixed, dtype=np.int64), P=P, lb=lb, ub=ub,
               target=target, q=q, ind=np.array(ind, dtype=np.int64), b=b, react=react)
    for a in ("loadpath", "thk"):
        if a in attrs and attrs[a] is not None:
            out[a] = np.float64(attrs[a])
    if name == "analysis.json":
        opt = doc["optimiser"]
        od = opt.get("data", opt)
        for key in ("xopt", "x0"):
            if od.get(key) is not None:
                out[key] = np.array(od[key], dtype=np.float64).reshape(-1)
        if od.get("fopt") is not None:
            out["fopt"] = np.float64(od["fopt"])
    return out
