"""Extract the name -> name joint-mapping tables of the reference (retarget/robot_config/Hu.py:27-105, Hu_v5.py:35-113;
SURVEY.md appendix A: "keep verbatim") into humanoid_real_time_retarget_b200/data/joint_mappings.json.
Dev container only:  python tools/extract_tables.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(__file__))
import ref_shim  # noqa: E402

NAMES = ["SMPL2HU_JOINT_MAPPING", "NOITOM2HU_JOINT_MAPPING", "VTRDYN2HU_JOINT_MAPPING", "VTRDYN_LITE2HU_JOINT_MAPPING"]


def main():
    ref = ref_shim.load()
    out = {"Hu": {n: dict(getattr(ref.hu_cfg, n)) for n in NAMES}, "Hu_v5": {n: dict(getattr(ref.hu_v5_cfg, n)) for n in NAMES}}
    dst = os.path.join(os.path.dirname(__file__), "..", "humanoid_real_time_retarget_b200", "data", "joint_mappings.json")
    with open(dst, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", os.path.normpath(dst), {k: {n: len(v) for n, v in d.items()} for k, d in out.items()})


if __name__ == "__main__":
    main()
