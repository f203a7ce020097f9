"""The C-ABI boundary without a GPU: libfccf.so loads, exports every symbol include/fccf.h declares,
its parameter block carries the reference defaults (FCCF.cpp:126-176), and — there being no CPU
fallback — every compute entry point refuses to run without a CUDA device."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "fccf.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fccf_[a-z_0-9]+)\s*\(", src)))


def test_header_and_library_agree(built):
    import fccf_pcr_b200 as fccf

    names = _declared_functions()
    assert len(names) >= 16
    L = C.CDLL(fccf.LIB_PATH)
    for n in names:
        assert hasattr(L, n), "libfccf.so does not export %s" % n
    assert sorted(fccf.EXPORTS) == names          # the ctypes binding covers the whole header
    # no torch / C++ types in the signatures, and nothing from the oracle linked in
    out = subprocess.run(["nm", "-D", "--defined-only", fccf.LIB_PATH], capture_output=True, text=True).stdout
    assert "orc_" not in out
    ldd = subprocess.run(["ldd", fccf.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in ldd and "fccf_oracle" not in ldd


def test_default_params_are_the_reference_globals(built):
    import fccf_pcr_b200 as fccf

    p = fccf.default_params()
    ref = dict(parameter_l1=0.5, parameter_l2=1.0, parameter_k1=5.0, parameter_k2=2.0, normal_vector_threshold1=5.0,
               normal_vector_threshold2=8.0, face_voxel_size=1.0, voxel_point_threshold=5, curvature_threshold=0.05,
               select_plane_number=15, quick_verify_angel_threshold=10.0, quick_verify_distance_threshold=2.0,
               required_optimize_plane=4.0, fine_verify_voxel_size=0.5, fine_verify_number=4,
               included_angle_same_threshold=5.0, included_angle_min_threshold=30.0, included_angle_max_threshold=150.0,
               third_plane_threshold=0.5, third_plane_normal_threshold=5.0, cluster_number_threshold=10,
               cluster_angel_threshold=2.0, cluster_distance_threshold=0.8, seclct_cluster_number=200, rough_threshold_gl=2)
    assert set(ref) == set(fccf.PARAM_FIELDS)
    for k, v in ref.items():
        assert getattr(p, k) == np.float32(v), k
    assert p.emulate_pcl_overflow == 1
    assert C.sizeof(fccf.Params) == 4 * 25 + 4 * 4


def _has_gpu():
    try:
        import fccf_pcr_b200 as fccf

        c = fccf.Context(0)
        c.close()
        return True
    except Exception:
        return False


def test_no_cpu_fallback(built):
    """Without a CUDA device fccf_create returns NULL and every entry point reports NO_DEVICE."""
    import fccf_pcr_b200 as fccf

    if _has_gpu():
        pytest.skip("a CUDA device is present")
    with pytest.raises(fccf.FccfError):
        fccf.Context(0)
    L = fccf.lib()
    T = np.zeros(16, np.float32)
    x = np.zeros((8, 3), np.float32)
    fp = C.POINTER(C.c_float)
    rc = L.fccf_register(None, x.ctypes.data_as(fp), 8, x.ctypes.data_as(fp), 8, C.c_float(0.1), T.ctypes.data_as(fp), None)
    assert rc == 4     # FCCF_ERR_NO_DEVICE
    assert b"no usable CUDA device" in L.fccf_last_error(None)


def test_cli_error_conventions(built, tmp_path):
    """FCCF.cpp:1648-1665: unreadable file -> "Couldn't read file" on stderr, exit 0."""
    import fccf_pcr_b200 as fccf
    from fccf_pcr_b200 import scenes

    r = subprocess.run([fccf.CLI_PATH], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr
    r = subprocess.run([fccf.CLI_PATH, str(tmp_path / "nope.ply"), str(tmp_path / "nope2.ply"), "0.1"], capture_output=True, text=True)
    assert r.returncode == 0 and "Couldn't read file" in r.stderr and r.stdout == ""
    bad = tmp_path / "bad.ply"
    bad.write_text("not a ply\n")
    r = subprocess.run([fccf.CLI_PATH, str(bad), str(bad), "0.1"], capture_output=True, text=True)
    assert r.returncode == 0 and "Couldn't read file" in r.stderr
    if not _has_gpu():
        pts = np.random.default_rng(0).uniform(-1, 1, (100, 3)).astype(np.float32)
        a = tmp_path / "a.ply"
        scenes.write_ply(str(a), pts)
        r = subprocess.run([fccf.CLI_PATH, str(a), str(a), "0.1"], capture_output=True, text=True)
        assert r.returncode == 3 and "no usable CUDA device" in r.stderr      # fails loudly, no CPU path
        assert r.stdout.startswith("Leaf size : 0.1\n")


def test_ply_round_trip(tmp_path):
    from fccf_pcr_b200 import scenes

    pts = np.random.default_rng(1).normal(size=(257, 3)).astype(np.float32)
    for binary in (True, False):
        p = tmp_path / ("b.ply" if binary else "a.ply")
        scenes.write_ply(str(p), pts, binary=binary)
        np.testing.assert_array_equal(scenes.read_ply(str(p)), pts)


def test_scene_generator_is_deterministic():
    from fccf_pcr_b200 import scenes

    a = scenes.make_pair("indoor", 5000, 3)
    b = scenes.make_pair("indoor", 5000, 3)
    np.testing.assert_array_equal(a[0], b[0])
    np.testing.assert_array_equal(a[1], b[1])
    assert a[0].shape == (5000, 3) and a[0].dtype == np.float32
    c = scenes.make_pair("indoor", 5000, 4)
    assert not np.array_equal(a[0], c[0])
    o = scenes.make_pair("outdoor", 20000, 3)
    assert np.abs(o[1]).max() > 50
