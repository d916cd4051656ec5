"""CPU: the C-ABI library builds, loads and exports every symbol include/tno_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "tno_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tno_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    from compas_tno_b200 import _lib
    assert sorted(_lib.EXPORTS) == _declared_functions()


def test_library_loads_and_exports_all_symbols():
    from compas_tno_b200 import _lib
    lib = _lib.load()
    assert os.path.exists(_lib.LIB_PATH)
    for name in _declared_functions():
        assert hasattr(lib, name), name
    assert lib.tno_abi_version() == _lib.ABI_VERSION


def test_struct_sizes_match_header_layout(tmp_path):
    """sizeof() as gcc sees the header == sizeof() of the ctypes mirror."""
    import subprocess
    from compas_tno_b200 import _lib
    src = tmp_path / "sz.c"
    src.write_text('#include "tno_b200.h"\n#include <stdio.h>\nint main(void){printf("%zu %zu %zu %zu\\n",'
                   'sizeof(tno_problem_desc),sizeof(tno_plan_info),sizeof(tno_outputs),sizeof(tno_batch_inputs));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    assert sizes == [ctypes.sizeof(_lib.ProblemDesc), ctypes.sizeof(_lib.PlanInfo), ctypes.sizeof(_lib.Outputs),
                     ctypes.sizeof(_lib.BatchInputs)]


def test_no_cpu_fallback_without_device():
    """Without a GPU the product path must fail loudly, not fall back to numpy."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from golden_io import load_case
    from compas_tno_b200 import _lib
    from compas_tno_b200.plan import TnoPlan
    P, objective, xs, _ = load_case("cross6_q_zb_min")
    with pytest.raises(_lib.TnoError):
        TnoPlan(P, objective=objective)


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "compas_tno_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), fn


def test_unsupported_combinations_are_rejected_before_any_device_work():
    """Argument validation mirrors the reference's NotImplementedError for combinations its callbacks do not cover; it
    runs before the device is touched, so it is testable on a machine without a GPU."""
    import copy
    import numpy as np
    from golden_io import load_case
    from compas_tno_b200.plan import TnoPlan
    P, objective, xs, _ = load_case("cross14_q_zb_min")

    def variant(**kw):
        Q = copy.copy(P)
        for key, val in kw.items():
            setattr(Q, key, val)
        return Q

    lim = np.column_stack([P.X[:, 0] - 1.0, P.X[:, 0] + 1.0])
    from compas_tno_b200.envelopes import CrossVaultEnvelope
    env = CrossVaultEnvelope((0.0, 10.0), (0.0, 10.0), 0.5)
    U, A = NotImplementedError, ValueError       # unsupported combination (TNO_EUNSUPPORTED) / plain argument error (TNO_EINVAL)
    cases = [
        (variant(variables=["q", "zb", "t"], envelope=env), "bestfit", U),                  # gradient_bestfit knows q and zb only
        (variant(variables=["q", "zb", "t"], envelope=env), "loadpath", U),
        (variant(variables=["q", "zb", "t"]), "t", A),                                      # no parametric envelope on the problem
        (variant(variables=["q", "zb"]), "Ecomp-linear", A),                                # no dXb on the problem
        (variant(variables=["q", "xyb", "zb"], constraints=["funicular", "envelope"]), "min", U),          # xyb needs envelopexy too
        (variant(variables=["q", "xyb", "zb"], constraints=["funicular", "envelopexy", "envelope"]), "min", A),  # no xlimits
        (variant(variables=["q", "xyb", "zb"], constraints=["funicular", "envelopexy", "envelope"], xlimits=lim, ylimits=lim), "bestfit", U),
        (variant(variables=["q", "zb", "tub"]), "min", U),                                  # not accelerated at all
        (variant(constraints=["funicular", "envelope", "displ_map"]), "min", U),
        (variant(features=["sym"]), "min", U),
        (P, "Ecomp-complete", U),
    ]
    for prob, obj, exc in cases:
        with pytest.raises(exc):
            TnoPlan(prob, objective=obj)
