"""The oracle against the committed golden fixture (tests/golden/make_golden.py): a regression pin of the
CPU restatement on stored inputs.  The fixture is oracle-generated — the reference has no golden vectors
and cannot run here (parity unpinned)."""
import os

import numpy as np

from fccf_pcr_b200 import scenes

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pair_indoor_8k.npz")


def test_oracle_reproduces_the_golden_fixture(orc):
    g = np.load(GOLD)
    T = orc.register(g["src"], g["tar"], float(g["leaf"]))
    for name in g.files:
        if not name.startswith("blob_"):
            continue
        want, got = g[name], orc.blob(name[5:])
        assert want.shape == got.shape, name
        if want.dtype.kind in "iu":
            np.testing.assert_array_equal(got, want, err_msg=name)       # integer stages: bit-exact
        else:
            np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-6, err_msg=name)
    assert scenes.rotation_error_deg(T, g["T_oracle"]) <= 1e-4 and scenes.translation_error(T, g["T_oracle"]) <= 1e-5
    # the registration itself is right: close to the ground truth the pair was made with
    assert scenes.rotation_error_deg(T, g["T_ground_truth"]) < 1.0 and scenes.translation_error(T, g["T_ground_truth"]) < 0.08
