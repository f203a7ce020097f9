"""Generates tests/golden/pair_indoor_8k.npz — a small scan pair with the stage results of the CPU oracle.

PARITY UNPINNED: the reference (FCCF.cpp) ships no golden vectors and cannot be built or run here
(PCL / Eigen / Ceres / FLANN absent), so this fixture is ORACLE-generated, not reference-generated.  What
it pins: (1) the oracle itself against regressions (tests/test_golden_cpu.py re-runs the oracle on the
stored clouds), (2) the CUDA path on the GPU box against the same committed numbers
(tests/test_parity_gpu.py::test_golden_fixture), independently of how the box's libm or numpy round.
The input clouds are stored in the file, so the fixture does not depend on the scene generator either.

    python tests/golden/make_golden.py        (run from the repository root)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from fccf_pcr_b200 import scenes  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402

INT_BLOBS = ["vg1_cell1", "vg1_cnt1", "vg1_cell2", "vg1_cnt2", "vg2_cell1", "vg2_cnt1", "vg2_cell2", "vg2_cnt2", "vox_key1", "vox_cnt1",
             "vox_key2", "vox_cnt2", "grow_label1", "grow_label2", "merge_label1", "merge_label2", "face_id1", "face_id2", "base1", "base2",
             "matches", "n_hyp", "n_centres", "cluster_num", "top_centre0", "top_centre1", "top_centre2", "fv_counts0", "fv_off0",
             "qv_pairs0", "qv_pair_off0"]
FLT_BLOBS = ["face_plane1", "face_plane2", "qv_score0", "top_s10", "top_s20", "top_T0"]


def main():
    n, leaf, seed = 8000, 0.2, 11
    src, tar, Tgt = scenes.make_pair("indoor", n, seed)
    o = Oracle()
    T = o.register(src, tar, leaf)
    out = {"src": src.astype(np.float32), "tar": tar.astype(np.float32), "leaf": np.float32(leaf), "T_oracle": T, "T_ground_truth": Tgt.astype(np.float32)}
    for b in INT_BLOBS + FLT_BLOBS:
        out["blob_" + b] = o.blob(b)
    path = os.path.join(ROOT, "tests", "golden", "pair_indoor_8k.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", "n_hyp", out["blob_n_hyp"], "centres", out["blob_n_centres"])


if __name__ == "__main__":
    main()
