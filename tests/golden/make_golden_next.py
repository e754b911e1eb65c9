"""Golden vectors for the SURVEY §8f rows (frontend loop, loop-closure verification, ScanContext), generated from the CPU
oracle like make_golden.py (PARITY UNPINNED by the reference: these freeze the oracle's own outputs).

    python tests/golden/make_golden_next.py
"""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import pyoracle as orc  # noqa: E402
from oracle import pyfrontend as opf  # noqa: E402
from oracle import pyscancontext as osc  # noqa: E402
from simpleslam_b200 import workloads  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    orc.build()
    seq = workloads.c5_sequence(16)
    # every 3rd point of each scan keeps the fixture small; the loop is run on exactly these clouds
    scans = [np.ascontiguousarray(f["scan"][::3]) for f in seq["frames"]]
    oo = opf.OracleOdometry("loam", threads=4)
    poses = np.stack([oo.step(s, f["stamp"], f["local_odom"]) for s, f in zip(scans, seq["frames"])])
    desc = np.stack([osc.make_scancontext(orc.voxel_downsample(s, 0.5)["points"], 2.0) for s in scans[:4]])
    d01 = osc.distance(desc[0], desc[1])
    d03 = osc.distance(desc[0], desc[3], sector_key_align=True)
    np.savez_compressed(
        os.path.join(OUT, "frontend_small.npz"),
        scans=np.concatenate([s[:, :3] for s in scans]), scan_offsets=np.cumsum([0] + [len(s) for s in scans]),
        stamps=np.array([f["stamp"] for f in seq["frames"]]), local_odom=np.stack([f["local_odom"] for f in seq["frames"]]),
        poses=poses, converged=np.array(oo.converged), n_keyframes=len(oo.kfs), submap_sizes=np.array([len(s) for s in oo.submaps]),
        sc_desc=desc, sc_dist=np.array([d01[0], d03[0]]), sc_shift=np.array([d01[1], d03[1]]))
    print("frontend_small: frames", len(scans), "keyframes", len(oo.kfs), "submaps", len(oo.submaps), "sc", d01, d03)


if __name__ == "__main__":
    main()
