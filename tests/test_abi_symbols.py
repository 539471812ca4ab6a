"""The C-ABI library loads and exports every symbol include/socm_b200.h declares (no compute)."""
import ctypes
import os
import re

from helpers import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "socm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(socm_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    from soc_matching_b200 import _lib, build
    build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/socm_b200.h but not exported"
    assert set(names) == set(_lib.PROTOTYPES), set(names) ^ set(_lib.PROTOTYPES)


def test_version_and_errors_without_gpu():
    from soc_matching_b200 import _lib
    lib = _lib.load()
    assert lib.socm_abi_version() == 1
    # argument validation happens before any CUDA call: NULL setting -> SOCM_ERR_INVALID + message
    rc = lib.socm_target_gemm_f32(None, None, 4, 2, 3, 16, None, 12, None)
    assert rc == 1
    assert b"NULL" in lib.socm_last_error()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "soc_matching_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f"{f} mentions the oracle"


def test_cpu_tensors_are_refused():
    import pytest
    import torch
    import soc_matching_b200 as sb
    sde = sb.DoubleWell(device="cpu", dim=3, kappa=torch.ones(3), nu=torch.ones(3), sigma=torch.eye(3))
    sde.initialize_models()
    with pytest.raises(sb._lib.SocmError):
        sb.stochastic_trajectories(sde, torch.zeros(4, 3), torch.linspace(0, 1, 5), 1.0)
