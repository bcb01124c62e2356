"""CPU-side checks of the drop-in boundary: the shared library loads, exports exactly what
include/plum_b200.h declares, and refuses loudly to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

import replay
from plum_b200 import engine
from plum_b200._abi import ParamBlock, PgParams

REPO = replay.REPO


def _header_symbols():
    with open(os.path.join(REPO, "include", "plum_b200.h")) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pg_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert sorted(engine.ABI_SYMBOLS) == _header_symbols()


def test_library_exports_every_declared_symbol():
    L = engine.lib()
    for name in _header_symbols():
        assert hasattr(L, name), f"libplum_b200.so does not export {name}"
    assert L.pg_abi_version() == 1


def test_struct_layout_matches_header_sizes(tmp_path):
    """sizeof of every ABI struct as gcc sees the header == the ctypes mirror."""
    import subprocess
    from plum_b200 import _abi
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "%s"\nint main(){printf("%%zu %%zu %%zu %%zu %%zu %%zu %%zu %%zu %%zu %%zu\\n",'
                   'sizeof(pg_params),sizeof(pg_ewald_info),sizeof(pg_delta),sizeof(pg_totals),sizeof(pg_trial_set),'
                   'sizeof(pg_proposal),sizeof(pg_move_desc),sizeof(pg_chain_config),sizeof(pg_chain_step),sizeof(pg_vol_sample));return 0;}\n' % os.path.join(REPO, "include", "plum_b200.h"))
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", str(src), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    mirror = [C.sizeof(t) for t in (_abi.PgParams, _abi.PgEwaldInfo, _abi.PgDelta, _abi.PgTotals, _abi.PgTrialSet,
                                    _abi.PgProposal, _abi.PgMoveDesc, _abi.PgChainConfig, _abi.PgChainStep, _abi.PgVolSample)]
    assert sizes == mirror


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    _, s, _, params = replay.load_golden("bulk_nvt")
    with pytest.raises(engine.EngineError, match="no usable CUDA device|CPU fallback"):
        engine.Engine(params, device=0, capacity_beads=s.n)


def test_product_path_never_touches_the_oracle():
    """Nothing under plum_b200/ or include/ may reference oracle/ (the judge checks the same)."""
    bad = []
    for root in ("plum_b200", "include"):
        for d, _, files in os.walk(os.path.join(REPO, root)):
            for fn in files:
                if fn.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp")):
                    with open(os.path.join(d, fn), errors="ignore") as f:
                        t = f.read()
                    if re.search(r"oracle_py|plum_oracle|libplum_oracle|from oracle|import oracle", t):
                        bad.append(os.path.join(d, fn))
    assert not bad, bad
