"""C-ABI surface: the shared library loads, exports every symbol include/pcr_cuda.h declares, and the product path
fails loudly (no CPU fallback) — CPU only, no compute calls."""
import ctypes
import os
import re
import pytest
from simpleslam_b200 import capi, registers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "pcr_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pcr_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    declared = _declared_symbols()
    assert sorted(capi.SYMBOLS) == declared
    L = capi.lib()
    for s in declared:
        assert hasattr(L, s), "missing export: " + s


def test_struct_layouts_match_header_sizes():
    # field-by-field mirrors of pcr_params / pcr_stats / pcr_loam_iter_log (natural alignment, same order as the header)
    assert ctypes.sizeof(capi.LoamIterLog) == 16 * 8 + 36 * 8 + 6 * 8 + 6 * 8 + 8 + 4 + 4
    assert ctypes.sizeof(capi.Stats) == 4 * 4 + 7 * 8 + 8 + 4 + 4 + 4 + 4 + 4 + 4 + 8
    p = capi.default_params(capi.PCR_NDT)
    assert p.method == capi.PCR_NDT and p.cores == 4
    assert p.loam_max_iters == 8 and abs(p.loam_plane_thresh - 0.2) < 1e-7
    assert p.ndt_search == capi.PCR_NDT_DIRECT7 and p.ndt_max_iters == 35 and p.ndt_min_points == 6
    assert p.vgicp_k == 20 and p.vgicp_max_iters == 64 and p.vgicp_optimizer == capi.PCR_LSQ_LM and p.vgicp_lm_init_lambda == 1e-9


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.PcrError) as e:
        capi.Context(capi.PCR_LOAM)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)
    with pytest.raises(capi.PcrError):
        registers.make_register("ndt")


def test_unknown_register_type_raises_like_reference():
    # frontend/src/LidarOdometry.cpp:49-53: unknown config string -> runtime_error
    with pytest.raises(RuntimeError, match="is not exist"):
        registers.make_register("icp")


def test_logger_callback_receives_errors():
    """a logger callback in the C ABI (the reference logs through its spdlog singleton); default = stderr"""
    import torch
    got = []
    capi.set_logger(lambda level, msg: got.append((level, msg)))
    try:
        if torch.cuda.is_available():
            c = capi.Context(capi.PCR_LOAM)
            with pytest.raises(capi.PcrError):
                c.align([[0.0, 0.0, 0.0]], [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]])   # no target set
            c.close()
            assert any(level == 3 and "no target" in msg for level, msg in got), got
        else:
            with pytest.raises(capi.PcrError):
                capi.Context(capi.PCR_LOAM)
            assert any(level == 3 and "no CPU fallback" in msg for level, msg in got), got
    finally:
        capi.set_logger(None)
