"""Two register instances on two host threads at the same time (the reference runs the LO register and the loop-closure
register on different threads, SURVEY §8b "Threading"): results must equal the single-threaded ones bit for bit —
no process-global mutable state, one CUDA stream + buffers per context."""
import threading
import numpy as np
import pytest
import data
from simpleslam_b200 import capi

pytestmark = pytest.mark.gpu


def test_two_contexts_two_threads():
    lc, vc = data.loam_case(), data.vgicp_case()
    ref_loam = capi.Context(capi.PCR_LOAM)
    ref_loam.set_target(lc["dst"])
    T_l, conv_l = ref_loam.align(lc["src"], lc["T_guess"])
    ref_loam.close()
    ref_v = capi.Context(capi.PCR_VGICP)
    ref_v.set_target(vc["dst"])
    T_v, conv_v = ref_v.align(vc["src"], vc["T_guess"])
    fit_v = ref_v.fitness()
    ref_v.close()

    errors = []

    def lo_thread():
        try:
            c = capi.Context(capi.PCR_LOAM)
            for _ in range(12):
                T, conv = c.scan2map(lc["src"], lc["dst"], lc["T_guess"])   # index rebuilt per call, like the frontend
                assert conv == conv_l and np.array_equal(T, T_l)
                ds = c.voxel_downsample(lc["raw"], 0.5)
                assert len(ds) == len(lc["src"])
            c.close()
        except Exception as e:  # noqa: BLE001
            errors.append(("lo", repr(e)))

    def lc_thread():
        try:
            c = capi.Context(capi.PCR_VGICP)
            for _ in range(4):
                T, conv = c.scan2map(vc["src"], vc["dst"], vc["T_guess"])
                assert conv == conv_v and np.array_equal(T, T_v) and c.fitness() == fit_v
            c.close()
        except Exception as e:  # noqa: BLE001
            errors.append(("lc", repr(e)))

    ts = [threading.Thread(target=lo_thread), threading.Thread(target=lc_thread)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors
