"""CPU-side checks of the boundary: the C-ABI library builds, loads and exports every symbol that
include/hrt_b200.h declares; the host tables are bit-exact; the product fails loudly without a GPU."""
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    from humanoid_real_time_retarget_b200 import _lib
    return _lib.load()


def test_library_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "hrt_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(hrt_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 15
    from humanoid_real_time_retarget_b200 import EXPORTED_SYMBOLS
    assert declared == set(EXPORTED_SYMBOLS)
    for name in declared:
        assert getattr(built, name) is not None
    assert built.hrt_abi_version() == 1


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "humanoid_real_time_retarget_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f"{f} mentions the oracle"


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu(built):
    import humanoid_real_time_retarget_b200 as hrt
    with pytest.raises(hrt.HrtError):
        hrt.Engine(0)
    with pytest.raises(hrt.HrtError):
        hrt.cal_local_rotation(torch.zeros(2, 21, 4), hrt.robot_config.vtrdyn_parent_indices)
    # the C ABI itself also refuses: no device -> HRT_E_NO_DEVICE, not a silent CPU path
    import ctypes as C
    h = C.c_void_p()
    assert built.hrt_ctx_create(0, C.byref(h)) == -5
    assert b"no CPU path" in built.hrt_last_error_string()


def test_tables_bit_exact(skeletons):
    from humanoid_real_time_retarget_b200 import robot_config as cfg
    from oracle import retarget_oracle as oc
    assert cfg.Hu_v5_DOF_AXIS == oc.HU_V5_DOF_AXIS == [2, 0, 1, 1, 1, 2, 0, 1, 1, 1, 2, 1, 0, 2, 1, 0, 1, 2, 1, 1,
                                                       1, 0, 2, 1, 0, 1, 2, 1, 1, 2]
    assert cfg.Hu_DOF_AXIS == oc.HU_DOF_AXIS and len(cfg.Hu_DOF_AXIS) == 32
    assert cfg.Hu_DOF_LOWER == oc.HU_DOF_LOWER and cfg.Hu_DOF_UPPER == oc.HU_DOF_UPPER
    assert len(cfg.Hu_v5_DOF_LOWER) == 30 == len(cfg.Hu_v5_DOF_UPPER)
    assert cfg.vtrdyn_parent_indices == skeletons["vtrdyn_zero_pose/parents"].tolist()
    assert cfg.vtrdyn_parent_indices == [-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 9, 10, 11, 10, 13, 14, 15, 10, 17, 18, 19]
    assert [str(s) for s in skeletons["vtrdyn_zero_pose/node_names"]] == cfg.VTRDYN_JOINT_NAMES
    # SURVEY appendix A
    assert skeletons["hu_v5_zero_pose/parents"].tolist() == [-1, 0, 1, 2, 3, 4, 0, 6, 7, 8, 9, 0, 11, 12, 13, 14, 15, 16,
                                                             17, 18, 18, 11, 21, 22, 23, 24, 25, 26, 27, 27, 11]
    assert cfg.BODY_23_TO_21 == [0, 1, 2, 3, 5, 6, 7, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22]
    # T2Z is non-identity only at the arm joints
    t2z = skeletons["t2z/vtrdyn"]
    ident = np.array([0, 0, 0, 1], np.float32)
    nonid = [j for j in range(21) if not np.array_equal(t2z[j], ident)]
    assert nonid == [14, 15, 16, 18, 19, 20]
    # asset md5s recorded in BASELINE.md
    assert str(skeletons["hu_v5_zero_pose/md5"]) == "7a3627cd3fa4cb2d3fa5307d86f8d25c"
    assert str(skeletons["vtrdyn_zero_pose/md5"]) == "c3e53235468e32f790fa8f7db17fbd1b"
