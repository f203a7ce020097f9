"""Known-answer tests of the CPU oracle.

The reference ships no tests, golden vectors or fixtures and cannot be built here (PARITY UNPINNED,
see oracle/fccf_oracle.cpp header), so the oracle is pinned against hand-computed cases and against
independent float64 numpy/scipy computations of the same published algorithms.
"""
import math

import numpy as np
import pytest


# ---- (1) pcl::VoxelGrid (FCCF.cpp:1668-1678 / 1377-1387) ------------------------------------
def test_voxelgrid_hand_computed(orc):
    leaf = 0.5
    pts = np.array([
        [0.10, 0.10, 0.10],    # cell (0,0,0)
        [0.40, 0.20, 0.30],    # cell (0,0,0)
        [0.50, 0.10, 0.10],    # exactly on a cell face -> cell (1,0,0)
        [-0.10, 0.10, 0.10],   # cell (-1,0,0)
        [-0.50, 0.10, 0.10],   # exactly on a face -> cell (-1,0,0)
        [-0.51, 0.10, 0.10],   # cell (-2,0,0)
        [0.10, 0.60, 0.10],    # cell (0,1,0)
        [0.10, 0.10, -0.20],   # cell (0,0,-1)
        [0.20, 0.30, 0.20],    # cell (0,0,0)
        [0.90, 0.90, 0.40],    # cell (1,1,0)
    ], np.float32)
    out, cell, cnt = orc.voxelgrid(pts, leaf)
    ijk = np.floor(pts.astype(np.float32) * np.float32(1.0 / leaf)).astype(np.int64)
    mn = ijk.min(0)
    div = ijk.max(0) - mn + 1
    lin = (ijk[:, 0] - mn[0]) + div[0] * ((ijk[:, 1] - mn[1]) + div[1] * (ijk[:, 2] - mn[2]))
    exp_cells = np.unique(lin)
    assert cell.tolist() == exp_cells.tolist()           # ascending linear index: z-major, then y, x
    assert cnt.tolist() == [int((lin == c).sum()) for c in exp_cells]
    assert cnt.sum() == len(pts) and len(out) == 7
    for c, o in zip(exp_cells, out):                     # float32 running sum in index order, then / n
        acc = np.zeros(3, np.float32)
        for p in pts[lin == c]:
            acc = acc + p
        np.testing.assert_array_equal(o, acc / np.float32((lin == c).sum()))
    three = out[[i for i, c in enumerate(exp_cells) if (lin == c).sum() == 3][0]]
    np.testing.assert_allclose(three, [(0.1 + 0.4 + 0.2) / 3, (0.1 + 0.2 + 0.3) / 3, (0.1 + 0.3 + 0.2) / 3], rtol=1e-6)


def test_voxelgrid_empty_nan_and_idempotence(orc):
    out, cell, cnt = orc.voxelgrid(np.zeros((0, 3), np.float32), 0.1)
    assert len(out) == 0
    rng = np.random.default_rng(3)
    pts = rng.uniform(-3, 3, (5000, 3)).astype(np.float32)
    pts[17] = [np.nan, 0, 0]
    pts[99] = [0, np.inf, 0]
    out, cell, cnt = orc.voxelgrid(pts, 0.25)
    assert cnt.sum() == 4998 and np.isfinite(out).all()
    assert (np.diff(cell) > 0).all()
    # second pass at the same leaf (Q2): same number of cells, centroids move by rounding at most
    out2, cell2, cnt2 = orc.voxelgrid(out, 0.25)
    assert len(out2) <= len(out) and len(out) - len(out2) <= 2
    if len(out2) == len(out):
        np.testing.assert_allclose(out2, out, atol=1e-6)


def test_voxelgrid_int32_overflow_bailout(orc):
    # pcl::VoxelGrid warns and returns the input when dx*dy*dz > INT32_MAX (App. A.1)
    pts = np.array([[0, 0, 0], [2000, 2000, 2000], [1, 1, 1]], np.float32)
    out, cell, cnt = orc.voxelgrid(pts, 0.5)
    np.testing.assert_array_equal(out, pts)
    assert cnt.tolist() == [1, 1, 1]


# ---- (2) pcl::octree lattice / growth / DFS order (FCCF.cpp:475-484) --------------------------
def test_octree_first_point_anchor_and_keys(orc):
    p0 = np.array([0.3, -1.7, 2.2], np.float32)
    keys, start, pidx, mn, depth = orc.octree(p0[None, :], 1.0)
    assert depth == 1 and keys.tolist() == [[1, 1, 1]]       # p0 itself has key (1,1,1)
    np.testing.assert_allclose(mn, p0.astype(np.float64) - 1.0)   # box = [p0 - res, p0 + res)
    pts = np.stack([p0, p0 + np.float32(0.999) * np.array([1, 0, 0], np.float32), p0 - np.array([1e-3, 0, 0], np.float32)])
    keys, start, pidx, mn, depth = orc.octree(pts, 1.0)
    # voxel faces lie at p0 + k*res: +0.999 stays in voxel 1, -0.001 falls into voxel 0
    got = {tuple(k): sorted(pidx[start[i]:start[i + 1]].tolist()) for i, k in enumerate(keys.tolist())}
    assert got == {(1, 1, 1): [0, 1], (0, 1, 1): [2]}


def test_octree_growth_and_dfs_order(orc):
    rng = np.random.default_rng(5)
    pts = rng.uniform(-6, 9, (400, 3)).astype(np.float32)
    keys, start, pidx, mn, depth = orc.octree(pts, 1.0)
    V = len(keys)
    # membership: key = (unsigned)((double(p) - min) / res) with the final min
    k_all = np.floor((pts.astype(np.float64) - mn) / 1.0).astype(np.int64)
    assert (k_all >= 0).all() and (k_all < (1 << depth)).all()
    # lattice anchor never changes: (p0 - min) is a whole number of voxels + 1 (p0 sits mid-voxel)
    np.testing.assert_allclose((pts[0].astype(np.float64) - mn) % 1.0, 0.0, atol=1e-12)
    # DFS order = Morton order with x as the most significant bit of every triple
    def morton(k):
        c = 0
        for b in range(depth - 1, -1, -1):
            c = (c << 3) | (((k[0] >> b) & 1) << 2) | (((k[1] >> b) & 1) << 1) | ((k[2] >> b) & 1)
        return c
    codes = [morton(k) for k in keys.tolist()]
    assert codes == sorted(codes) and len(set(codes)) == V
    for v in range(V):
        idx = pidx[start[v]:start[v + 1]]
        assert (np.diff(idx) > 0).all()                      # insertion (ascending index) order
        assert (k_all[idx] == keys[v]).all()
    assert start[V] == len(pts)


# ---- (3) computeMeanAndCovarianceMatrix + eigen33 (FCCF.cpp:490-495) --------------------------
@pytest.mark.parametrize("normal", [(0, 0, 1), (1, 0, 0), (1, 1, 0), (1, 2, 3)])
def test_plane_fit_against_float64_pca(orc, normal):
    rng = np.random.default_rng(11)
    n = np.asarray(normal, float) / np.linalg.norm(normal)
    a = np.cross(n, [0.3, -0.5, 0.8]); a /= np.linalg.norm(a)
    b = np.cross(n, a)
    uv = rng.uniform(-0.5, 0.5, (60, 2))
    pts = (np.array([1.0, -2.0, 0.5]) + uv[:, :1] * a + uv[:, 1:] * b + rng.normal(scale=0.004, size=(60, 1)) * n).astype(np.float32)
    out = orc.plane_fit(pts)
    p64 = pts.astype(np.float64)
    C = np.cov(p64.T, bias=True)
    w, v = np.linalg.eigh(C)
    np.testing.assert_allclose(out[:3], p64.mean(0), rtol=0, atol=2e-6)
    assert abs(abs(float(np.dot(out[3:6], v[:, 0]))) - 1.0) < 1e-4       # float32 raw moments: ~1e-5 rad
    assert abs(np.linalg.norm(out[3:6]) - 1.0) < 1e-6
    assert abs(out[6] - w[0] / w.sum()) < 2e-4 and out[7] == 60


def test_plane_fit_degenerate_matrices(orc):
    # exact plane z = 0 on a lattice: c0 ~ 0 branch -> curvature exactly 0, normal +-z
    g = np.stack(np.meshgrid(np.arange(4), np.arange(4)), -1).reshape(-1, 2).astype(np.float32) * 0.25
    pts = np.concatenate([g, np.zeros((16, 1), np.float32)], 1)
    out = orc.plane_fit(pts)
    assert out[6] == 0.0 and abs(abs(out[5]) - 1.0) < 1e-6
    # rank 1 (points on a line): smallest eigenvalue 0; every row cross product of (C - 0*I) vanishes,
    # and pcl::eigen33 divides the winner by its zero length -> NaN normal (no guard in PCL 1.10)
    line = np.stack([np.linspace(0, 1, 10), np.zeros(10), np.zeros(10)], 1).astype(np.float32)
    out = orc.plane_fit(line)
    assert out[6] == 0.0 and np.isnan(out[3:6]).all()
    # a single repeated point: zero covariance; curvature 0 (trace == 0 branch)
    out = orc.plane_fit(np.ones((8, 3), np.float32))
    assert out[6] == 0.0 or math.isnan(out[6])


# ---- (4) compute_normal_angel (FCCF.cpp:369-377) ---------------------------------------------
def test_normal_angle_values_and_nan_path(orc):
    assert abs(orc.normal_angle((1, 0, 0), (0, 1, 0)) - 90.0) < 1e-5
    assert abs(orc.normal_angle((1, 0, 0), (-1, 0, 0)) - 180.0) < 1e-4
    assert abs(orc.normal_angle((1, 0, 0), (1, 1, 0)) - 45.0) < 1e-4
    # identical vectors: the float quotient may exceed 1 by rounding -> acos -> NaN (Q9), else 0
    vals = [orc.normal_angle(v, v) for v in [(0.6, 0.8, 0.0), (0.1, 0.2, 0.3), (0.577, 0.577, 0.577), (1.0, 0.0, 0.0)]]
    assert all(math.isnan(v) or abs(v) < 0.03 for v in vals)
    assert math.isnan(orc.normal_angle((0, 0, 0), (1, 0, 0)))           # 0/0


# ---- (6) Eigen quaternion <-> matrix, all four branches (FCCF.cpp:1451, 1475) -----------------
def _rot(axis, deg):
    axis = np.asarray(axis, float) / np.linalg.norm(axis)
    a = math.radians(deg)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + math.sin(a) * K + (1 - math.cos(a)) * K @ K


@pytest.mark.parametrize("axis,deg", [((0, 0, 1), 30), ((1, 0, 0), 170), ((0, 1, 0), 175), ((0, 0, 1), 179), ((1, 1, 1), 120), ((1, 2, -1), 200)])
def test_quaternion_round_trip(orc, axis, deg):
    R = _rot(axis, deg).astype(np.float32)
    q = orc.quat_from_matrix(R)                      # (w, x, y, z)
    assert abs(np.linalg.norm(q) - 1) < 1e-6
    R2 = orc.quat_to_matrix(q)
    np.testing.assert_allclose(R2, R, atol=5e-7)
    a = math.radians(deg)
    ax = np.asarray(axis, float) / np.linalg.norm(axis)
    qe = np.array([math.cos(a / 2), *(math.sin(a / 2) * ax)])
    assert min(np.abs(q - qe).max(), np.abs(q + qe).max()) < 1e-6


def test_quaternion_of_non_rotation_is_not_normalised(orc):
    # Q7: toRotationMatrix does not normalise
    R = orc.quat_to_matrix(np.array([2, 0, 0, 0], np.float32))
    np.testing.assert_array_equal(R, np.eye(3, dtype=np.float32))
    R = orc.quat_to_matrix(np.array([1, 1, 0, 0], np.float32))
    np.testing.assert_array_equal(R, np.array([[1, 0, 0], [0, -1, -2], [0, 2, -1]], np.float32))


# ---- fine_verify (FCCF.cpp:785-839) against an independent numpy restatement ------------------
def _fine_verify_numpy(T, s1, s2, res=0.5):
    T = np.asarray(T, np.float32).reshape(4, 4)
    x, y, z = (s2[:, i].astype(np.float32) for i in range(3))
    q = np.stack([x * T[r, 0] + (y * T[r, 1] + (z * T[r, 2] + T[r, 3])) for r in range(3)], 1)
    origin = s1[0].astype(np.float64) - res            # lattice anchored on the first static point
    k1 = np.floor((s1.astype(np.float64) - origin) / res).astype(np.int64)
    k2 = np.floor((q.astype(np.float64) - origin) / res).astype(np.int64)
    d1, d2 = {}, {}
    for k in map(tuple, k1):
        d1[k] = d1.get(k, 0) + 1
    for k in map(tuple, k2):
        d2[k] = d2.get(k, 0) + 1
    rows = sorted((k[0] - 1, k[1] - 1, k[2] - 1, d1[k], d2[k]) for k in d1 if k in d2)
    sim = sum((s + t) * (min(s, t) / max(s, t)) for (_, _, _, s, t) in rows)
    return sim / (len(s1) + len(s2)), np.asarray(rows, np.int32).reshape(-1, 5)


def test_fine_verify_counts_and_score(orc):
    rng = np.random.default_rng(21)
    s1 = rng.uniform(-4, 4, (3000, 3)).astype(np.float32)
    s2 = (s1[rng.permutation(3000)[:2500]] + rng.normal(scale=0.02, size=(2500, 3))).astype(np.float32)
    T = np.eye(4, dtype=np.float32)
    T[:3, :3] = _rot((0, 0, 1), 2.0)
    T[:3, 3] = [0.05, -0.02, 0.01]
    sc, rows = orc.fine_verify(T, s1, s2)
    sc_np, rows_np = _fine_verify_numpy(T, s1, s2)
    rows = rows[np.lexsort((rows[:, 2], rows[:, 1], rows[:, 0]))]
    np.testing.assert_array_equal(rows, rows_np)
    assert abs(sc - sc_np) < 1e-5 * max(1.0, sc_np)


def test_fine_verify_empty_leftover_is_nan(orc):
    sc, rows = orc.fine_verify(np.eye(4, dtype=np.float32), np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32))
    assert math.isnan(sc) and len(rows) == 0            # Q15: 0/0


# ---- quick_verify + the Ceres restatement against scipy (FCCF.cpp:680-783, 178-249) -----------
def _planes(rng, n):
    nr = rng.normal(size=(n, 3)); nr /= np.linalg.norm(nr, axis=1, keepdims=True)
    c = rng.uniform(-5, 5, (n, 3))
    size = rng.integers(50, 400, n).astype(float)
    return np.concatenate([c, nr, size[:, None]], 1)


def test_quick_verify_refinement_matches_scipy(orc):
    from scipy.optimize import least_squares
    from scipy.spatial.transform import Rotation

    rng = np.random.default_rng(8)
    p1 = _planes(rng, 8)
    Rgt = Rotation.from_euler("zyx", [20, -5, 3], degrees=True).as_matrix()
    tgt = np.array([0.7, -0.3, 0.2])
    # plane set 2 = plane set 1 moved by the inverse of (Rgt, tgt)
    p2 = p1.copy()
    p2[:, 3:6] = p1[:, 3:6] @ Rgt
    p2[:, :3] = (p1[:, :3] - tgt) @ Rgt
    # hypothesis = ground truth perturbed by 2 degrees / 5 cm
    dR = Rotation.from_euler("xyz", [1.5, -1.0, 0.8], degrees=True).as_matrix()
    T0 = np.eye(4, dtype=np.float32)
    T0[:3, :3] = dR @ Rgt
    T0[:3, 3] = tgt + [0.05, -0.03, 0.02]
    score, T, pairs, iters = orc.quick_verify(T0, p1.astype(np.float32), p2.astype(np.float32))
    assert len(pairs) == 8 and (pairs[:, 0] == pairs[:, 1]).all() and iters > 0
    total = int(np.float32(p1[:, 6]).sum()) + int(np.float32(p2[:, 6]).sum())
    np.testing.assert_allclose(score, sum(2 * min(a, b) / total for a, b in zip(p1[:, 6], p2[:, 6])), rtol=1e-5)
    # the refined hypothesis must be (numerically) the ground truth: residuals vanish there
    np.testing.assert_allclose(T[:3, :3], Rgt, atol=2e-5)
    np.testing.assert_allclose(T[:3, 3], tgt, atol=2e-4)
    # independent minimiser of the same cost (LidarPlaneFactor, FCCF.cpp:191-194)
    T0d = T0.astype(np.float64)
    n2 = p2[:, 3:6] @ T0d[:3, :3].T
    c2 = p2[:, :3] @ T0d[:3, :3].T + T0d[:3, 3]
    w = np.array([2 * min(a, b) / total for a, b in zip(p1[:, 6], p2[:, 6])])

    def res(x):
        R = Rotation.from_rotvec(x[:3]).as_matrix()
        nn = n2 @ R.T
        cc = c2 @ R.T + x[3:]
        r1 = w * np.linalg.norm(np.cross(p1[:, 3:6], nn), axis=1)
        r2 = w * np.abs((p1[:, 3:6] * p1[:, :3]).sum(1) - (nn * cc).sum(1))
        return np.concatenate([r1, r2])

    sol = least_squares(res, 1e-3 * np.ones(6), xtol=1e-14, ftol=1e-14, gtol=1e-14)
    Rs = Rotation.from_rotvec(sol.x[:3]).as_matrix() @ T0d[:3, :3]
    ts = Rotation.from_rotvec(sol.x[:3]).as_matrix() @ T0d[:3, 3] + sol.x[3:]
    np.testing.assert_allclose(T[:3, :3], Rs, atol=5e-5)
    np.testing.assert_allclose(T[:3, 3], ts, atol=5e-4)


def test_quick_verify_below_four_pairs_leaves_hypothesis_alone(orc):
    rng = np.random.default_rng(9)
    p1 = _planes(rng, 3).astype(np.float32)
    T0 = np.eye(4, dtype=np.float32)
    score, T, pairs, iters = orc.quick_verify(T0, p1, p1)
    np.testing.assert_array_equal(T, T0)
    assert len(pairs) == 3 and iters <= 0


# ---- stdout format (FCCF.cpp:1667, 1687) -----------------------------------------------------
def test_output_format_matches_eigen_default_ioformat(oracle_mod):
    T = np.array([[0.5, -0.25, 0, 1.5], [0.25, 0.5, 0, -0.8], [0, 0, 1, 100.125], [0, 0, 0, 1]], np.float32)
    s = oracle_mod.format_output(0.1, T)
    assert s == ("Leaf size : 0.1\nTransformation: \n"
                 "    0.5   -0.25       0     1.5\n"
                 "   0.25     0.5       0    -0.8\n"
                 "      0       0       1 100.125\n"
                 "      0       0       0       1\n")
    s = oracle_mod.format_output(2, np.eye(4, dtype=np.float32))
    assert s == "Leaf size : 2\nTransformation: \n1 0 0 0\n0 1 0 0\n0 0 1 0\n0 0 0 1\n"
    T = np.eye(4, dtype=np.float32); T[0, 3] = np.float32(1.23456789); T[1, 3] = np.float32(-1e-7)
    s = oracle_mod.format_output(0.25, T)
    assert "1.23457" in s and "-1e-07" in s


# ---- (5)(7) end to end on synthetic pairs with known ground truth ------------------------------
@pytest.mark.parametrize("n,seed,leaf", [(20000, 7, 0.1), (50000, 1, 0.1)])
def test_end_to_end_recovers_ground_truth(orc, n, seed, leaf):
    from fccf_pcr_b200 import scenes

    src, tar, Tgt = scenes.make_pair("indoor", n, seed)
    T = orc.register(src, tar, leaf)
    assert scenes.rotation_error_deg(T, Tgt) < 1.0 and scenes.translation_error(T, Tgt) < 0.08
    # stage invariants
    assert orc.blob("vg1_cnt1").sum() == n and orc.blob("vg1_cnt2").sum() == n
    assert len(orc.blob("face_id1")) == 16 and len(orc.blob("face_id2")) == 16      # Q4: 16 planes, not 15
    b1 = orc.blob("base1").reshape(-1, 3)
    ang = orc.blob("base_angle1")
    assert ((ang > 30) & (ang < 150)).all() and (b1[:, 0] < b1[:, 1]).all() and set(b1[:, 2]) <= {0, 1, 2}
    assert orc.blob("n_hyp").sum() == sum(len(orc.blob("hyp%d" % t)) // 12 for t in range(3))


# ---- (8) degenerate regimes --------------------------------------------------------------------
def test_degenerate_leaf_one_metre_gives_zero_rotation(orc):
    from fccf_pcr_b200 import scenes

    src, tar, _ = scenes.make_pair("indoor", 20000, 7)
    T = orc.register(src, tar, 1.0)        # Q14: no 1 m voxel holds > 5 points
    assert len(orc.blob("face_id1")) == 0 and orc.blob("n_hyp").tolist() == [0, 0, 0]
    assert orc.blob("n_centres").tolist() == [1, 1, 1]
    assert not np.isfinite(T[:3]).all() or np.abs(T[:3, :3]).max() == 0.0


def test_identical_clouds_give_identity(orc):
    from fccf_pcr_b200 import scenes

    src, tar, _ = scenes.make_pair("indoor", 20000, 7)
    T = orc.register(tar, tar, 0.1)
    assert scenes.rotation_error_deg(T, np.eye(4)) < 0.05 and scenes.translation_error(T, np.eye(4)) < 0.01


def test_parameters_follow_the_reference_defaults(orc):
    # FCCF.cpp:126-176 — unknown names are rejected, known ones accepted
    for k, v in dict(parameter_l1=0.5, parameter_k1=5.0, face_voxel_size=1.0, select_plane_number=15, fine_verify_number=4, seclct_cluster_number=200).items():
        orc.set_param(k, v)
    with pytest.raises(KeyError):
        orc.set_param("no_such_parameter", 1.0)


def test_libm_switch_bounds_the_last_ulp_effect(oracle_mod):
    """pcl::eigen33's float atan2/cos/sin: correctly rounded (default) vs this platform's glibc float
    routines.  Decisions and the final transform must not depend on it beyond the parity tolerance."""
    from fccf_pcr_b200 import scenes

    src, tar, _ = scenes.make_pair("indoor", 20000, 7)
    a = oracle_mod.Oracle()
    Ta = a.register(src, tar, 0.1)
    b = oracle_mod.Oracle(libm_float=1)
    try:
        Tb = b.register(src, tar, 0.1)
    finally:
        b.set_param("libm_float", 0)
    for name in ["vox_flag1", "vox_flag2", "merge_label1", "merge_label2", "base1", "base2", "matches", "n_hyp", "n_centres"]:
        assert np.array_equal(a.blob(name), b.blob(name)), name
    pa, pb = a.blob("pvox1"), b.blob("pvox1")
    assert np.abs(pa - pb).max() <= 2.5e-7 and (pa != pb).mean() < 0.05
    assert scenes.rotation_error_deg(Ta, Tb) <= 0.01 and scenes.translation_error(Ta, Tb) <= 1e-3


def test_oracle_switches_do_not_move_the_result(oracle_mod):
    """Three choices of the oracle were made so that the CUDA path can agree with it to the last bit: float libm calls
    of pcl::eigen33 evaluated correctly rounded (instead of the platform's float libm), every Levenberg-Marquardt
    row sum taken in xor-butterfly order (instead of row order), and the Householder reflections of the LM's QR applied
    from one set of column products per column (instead of the textbook loop's three dependent reductions; with it
    reciprocal-multiply back substitution and u*u*u for pow(u, 3)).  This measures what they are worth: with the
    platform float libm (libm_float = 1), plain row-order sums (lm_sequential = 1) and / or the textbook QR
    (lm_textbook = 1) every decision of the pipeline stays the same and the final transform moves by far less than
    the 0.01 degree / 1 mm parity bar."""
    from fccf_pcr_b200 import scenes

    cases = [("indoor", 20000, 7, 0.1, {}), ("indoor", 50000, 1, 0.1, {}), ("indoor", 200000, 2, 0.2, {}),
             ("indoor_rough", 200000, 5, 0.2, dict(third_plane_threshold=0.02, included_angle_same_threshold=30.0, third_plane_normal_threshold=15.0))]
    worst = (0.0, 0.0)
    try:
        for kind, n, seed, leaf, prm in cases:
            src, tar, _ = scenes.make_pair(kind, n, seed)
            ref = oracle_mod.Oracle(**prm)
            T0 = ref.register(src, tar, leaf)
            keep = {nm: ref.blob(nm).copy() for nm in ["merge_label1", "merge_label2", "base1", "base2", "matches", "n_hyp", "n_centres",
                                                       "top_centre0", "top_centre1", "top_centre2", "qv_pairs0", "qv_pairs2"]}
            for libm, seq, tb in ((1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 1)):
                o = oracle_mod.Oracle(libm_float=libm, lm_sequential=seq, lm_textbook=tb, **prm)
                T = o.register(src, tar, leaf)
                for nm, v in keep.items():
                    assert np.array_equal(o.blob(nm), v), (kind, seed, libm, seq, tb, nm)
                de, dt = scenes.rotation_error_deg(T, T0), scenes.translation_error(T, T0)
                worst = (max(worst[0], de), max(worst[1], dt))
                assert de <= 1e-3 and dt <= 1e-4, (kind, seed, libm, seq, tb, de, dt)      # a tenth of the parity bar
                for t in range(3):
                    for Ta, Tb in zip(o.blob("top_T%d" % t).reshape(-1, 4, 4), ref.blob("top_T%d" % t).reshape(-1, 4, 4)):
                        assert scenes.rotation_error_deg(Ta, Tb) <= 1e-3 and scenes.translation_error(Ta, Tb) <= 1e-4
    finally:
        oracle_mod.Oracle(libm_float=0, lm_sequential=0, lm_textbook=0)      # the switches are process-wide
    print("oracle switches: worst final-transform shift %.2e deg, %.2e m" % worst)


def test_stage_functions_known_answers(orc):
    """Hand-checkable answers of the stage entry points of the oracle (select_base, cluster bypass, fusion)."""
    import math

    def plane(n, c=(0, 0, 0), size=100):
        return [*c, *n, size]

    c20, s20 = math.cos(math.radians(20)), math.sin(math.radians(20))
    planes = np.array([plane((1, 0, 0)), plane((0, 1, 0)), plane((c20, s20, 0)), plane((0, 0, 1))], np.float32)
    theta = np.array([0.5, 3.0, 2.0, np.nan])
    orc.hypotheses(planes, theta, np.zeros((0, 7), np.float32), np.zeros(0))
    base = orc.blob("base1").reshape(-1, 3).tolist()
    # (0,1) 90 deg smooth/rough -> type 2; (0,2) 20 deg: outside (30, 150); (0,3) 90 deg with a NaN roughness -> type 3 ("none");
    # (1,2) 70 deg rough/smooth(2.0 <= 2) -> type 2; (1,3), (2,3) -> type 3
    assert base == [[0, 1, 2], [0, 3, 3], [1, 2, 2], [1, 3, 3], [2, 3, 3]]
    np.testing.assert_allclose(orc.blob("base_angle1"), [90, 90, 70, 90, 90], atol=1e-3)
    # Q12: empty pool -> one identity centre; 10 hypotheses pass through untouched
    ten = np.tile(np.array([1, 0, 0, 0, 0.1, 0.2, 0.3], np.float32), (10, 1)) + np.arange(10, dtype=np.float32)[:, None] * 0.01
    nc = orc.cluster(np.concatenate([ten]), [0, 10, 0])
    assert nc.tolist() == [1, 10, 1]
    assert orc.blob("centre0").tolist() == [1, 0, 0, 0, 0, 0, 0] and np.array_equal(orc.blob("centre1").reshape(-1, 7), ten)
    assert orc.blob("cluster_num").tolist() == [0, 200, 0]
    # fusion of a single surviving type returns that hypothesis (rotation rebuilt from its x / y axes)
    T = np.zeros((3, 1, 4, 4), np.float32)
    a = math.radians(30)
    T[0, 0] = [[math.cos(a), -math.sin(a), 0, 1.0], [math.sin(a), math.cos(a), 0, 2.0], [0, 0, 1, 3.0], [0, 0, 0, 1]]
    out = orc.fuse(T, [[0.5], [0], [0]], [[0.25], [0], [0]], [1, 0, 0])
    np.testing.assert_allclose(out, T[0, 0], atol=1e-6)
    tb = orc.blob("type_best").reshape(3, 13)
    assert tb[0, 0] == 2.0 and tb[1, 0] == 0.0            # 0.5/0.5 + 0.25/0.25; empty types score 0
    # Q15: NaN fine scores -> nothing survives the gate -> zero rotation block
    out = orc.fuse(T, [[0.5], [0], [0]], [[np.nan], [0], [0]], [1, 0, 0])
    assert not np.any(np.nan_to_num(out[:3, :3]))
