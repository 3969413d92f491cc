import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        return {k: z[k] for k in z.files}
    return load


@pytest.fixture(scope="session")
def skeletons():
    from oracle import retarget_oracle as oc
    return oc.load_skeletons()


# ------------------------------------------------------------------------------------------------------------------
# Parity report: the GPU tests record every distance distribution they gate on; the session writes them to
# profiles/parity_r02.json (and to gpurun_out/, the only directory that travels back from the GPU box).
# ------------------------------------------------------------------------------------------------------------------
_PARITY = {}


class _ParityRecorder:
    def record(self, key, stats, **extra):
        entry = dict(stats)
        entry.update(extra)
        _PARITY[key] = entry
        print(f"[parity] {key}: " + ", ".join(f"{k}={v:.3e}" if isinstance(v, float) else f"{k}={v}" for k, v in entry.items()))
        return entry


@pytest.fixture(scope="session")
def parity():
    return _ParityRecorder()


def pytest_sessionfinish(session, exitstatus):
    if not _PARITY:
        return
    import json
    try:
        import torch
        dev = torch.cuda.get_device_name(0) if torch.cuda.is_available() else "cpu"
    except Exception:
        dev = "unknown"
    doc = {"device": dev, "exitstatus": int(exitstatus),
           "columns": "frac_le_1e-5 = fraction of frames whose worst |d dof| <= 1e-5 rad; dof_* in rad over frames; fk_pos_* = "
                      "distance in metres between the link positions of the two angle sets (worst link per frame); geodesic_* = "
                      "rotation angle in rad between the two local rotations (worst joint per frame)",
           "entries": _PARITY}
    for d in ("profiles", "gpurun_out"):
        path = os.path.join(ROOT, d)
        try:
            os.makedirs(path, exist_ok=True)
            with open(os.path.join(path, "parity_r02.json"), "w") as f:
                json.dump(doc, f, indent=1, sort_keys=True)
        except OSError:
            pass
