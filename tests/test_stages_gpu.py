"""Adversarial parity tests of the small stages through their own C-ABI entry points: select_base,
the match loop + computer_transform, transform_cluster and the per-type best / gate / fuse_answer, fed inputs that
whole registrations of the synthetic scenes never produce (NaN normals and roughness, pools of exactly 0 / 10 / 11
hypotheses, cluster_num 0 and 1, a cut-off walk that runs out, NaN translations, 1 / 2 / 3 surviving types, NaN
scores).  Reference: FCCF.cpp:429-468, 841-1018, 1040-1231, 1291-1368, 1546-1606; quirks Q9, Q12, Q13, Q15."""
import numpy as np
import pytest

from fccf_pcr_b200 import scenes

pytestmark = pytest.mark.gpu


def _planes(rng, f, spread=1.0, nan_normal=None):
    """f planes (centroid, near-unit normal, point size) with normals in general position."""
    n = rng.normal(size=(f, 3))
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    n *= rng.uniform(0.97, 1.0, (f, 1))            # face normals are count-weighted means, not re-normalised (Q6)
    c = rng.uniform(-5, 5, (f, 3)) * spread
    size = rng.integers(50, 5000, (f, 1)).astype(np.float64)
    p = np.concatenate([c, n, size], axis=1).astype(np.float32)
    if nan_normal is not None:
        p[nan_normal, 3] = np.nan
    return p


def _room_planes(rng, f):
    """Planes of a box-like scene: normals close to the three axes (many pairs near 90 degrees, many third planes)."""
    axes = np.eye(3)[rng.integers(0, 3, f)] * rng.choice([-1.0, 1.0], (f, 1))
    n = axes + rng.normal(scale=0.03, size=(f, 3))
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    c = rng.uniform(-5, 5, (f, 3))
    size = rng.integers(50, 5000, (f, 1)).astype(np.float64)
    return np.concatenate([c, n, size], axis=1).astype(np.float32)


def _moved(planes, T):
    p = planes.copy().astype(np.float64)
    p[:, :3] = p[:, :3] @ T[:3, :3].T + T[:3, 3]
    p[:, 3:6] = p[:, 3:6] @ T[:3, :3].T
    return p.astype(np.float32)


def _same_f(a, b, tol):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape
    assert np.array_equal(np.isnan(a), np.isnan(b))
    m = ~np.isnan(a)
    if m.any():
        assert np.abs(a[m] - b[m]).max() <= tol * max(1.0, np.abs(b[m]).max())


def test_base_pairs_and_roughness_types(ctx, orc):
    rng = np.random.default_rng(101)
    cases = []
    for f in (0, 1, 2, 5, 16):
        cases.append((_planes(rng, f), rng.uniform(0, 4, f)))
    p = _room_planes(rng, 16); th = rng.uniform(0, 4, 16)
    th[3] = 2.0; th[4] = np.nextafter(2.0, 3.0); th[5] = np.nan        # exactly the threshold (smooth), just above (rough), NaN (type "none", Q9)
    cases.append((p, th))
    p2 = _planes(rng, 12, nan_normal=4); cases.append((p2, rng.uniform(0, 4, 12)))       # a NaN normal: every angle with it is NaN, no pair (Q9)
    p3 = _planes(rng, 8); p3[5] = p3[2]; cases.append((p3, np.zeros(8)))                 # coincident planes: angle 0, outside (30, 150)
    p4 = _planes(rng, 6); p4[1, 3:6] = -p4[0, 3:6]; cases.append((p4, np.full(6, 5.0)))  # opposite normals: 180 degrees
    for planes, theta in cases:
        pairs, ang = ctx.base_pairs(planes, theta)
        orc.hypotheses(planes, theta, np.zeros((0, 7), np.float32), np.zeros(0))
        want = orc.blob("base1").reshape(-1, 3)
        assert np.array_equal(pairs, want), (len(planes), pairs, want)
        np.testing.assert_allclose(ang, orc.blob("base_angle1"), atol=1e-3)


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_hypotheses_stage(ctx, orc, seed):
    rng = np.random.default_rng(200 + seed)
    f1, f2 = (16, 16) if seed < 3 else ((2, 2) if seed == 3 else (7, 11))
    p1 = _room_planes(rng, f1)
    if seed == 3:                                  # exactly two planes per cloud, 90 degrees apart: no third plane exists
        p1[:, 3:6] = np.array([[1, 0.02, 0.01], [0.03, 1, -0.02]], np.float32)
    T = scenes.ground_truth((20.0, 4.0, -3.0), (0.7, -0.4, 0.1))
    p2 = _moved(p1[rng.permutation(f1)][:f2] if f2 <= f1 else np.concatenate([p1, _room_planes(rng, f2 - f1)]), np.linalg.inv(T))
    p2[:, :6] += rng.normal(scale=0.01, size=(len(p2), 6)).astype(np.float32)
    th1, th2 = rng.uniform(0, 4, f1), rng.uniform(0, 4, len(p2))
    if seed == 3:
        th1[:] = 0.5; th2[:] = 0.5
    if seed == 2:
        th1[:] = 0.5; th2[:] = 0.5            # one pool only
        p1[3, 3] = np.nan                      # and a NaN normal among the third-plane candidates
    nh = ctx.hypotheses(p1, th1, p2, th2)
    want = orc.hypotheses(p1, th1, p2, th2)
    assert np.array_equal(nh, want), (nh, want)
    for name in ("base1", "base2", "matches"):
        assert np.array_equal(ctx.blob(name), orc.blob(name)), name
    for t in range(3):
        _same_f(ctx.blob("hyp%d" % t).reshape(-1, 12), orc.blob("hyp%d" % t).reshape(-1, 12), 1e-4)
        _same_f(ctx.blob("hyp_qt%d" % t).reshape(-1, 7), orc.blob("hyp_qt%d" % t).reshape(-1, 7), 1e-4)
    if seed == 3:
        assert nh.sum() > 0 and (ctx.blob("matches").reshape(-1, 3)[:, 2] == 1).all()     # no third plane: the centroid fall-back, one hypothesis per match


def _pool(rng, n, k_clusters, noise_t=0.05, noise_r=0.3, centres=None):
    """n hypotheses (qw qx qy qz tx ty tz) scattered around k cluster centres (translations within noise_t metres,
    rotations within noise_r degrees), quaternions deliberately not unit (Q7)."""
    if n == 0:
        return np.zeros((0, 7), np.float32)
    if centres is None:
        centres = [(rng.normal(size=3), np.radians(rng.uniform(-40, 40)), rng.uniform(-3, 3, 3)) for _ in range(max(1, k_clusters))]
    out = np.zeros((n, 7))
    for i in range(n):
        ax, ang, t = centres[rng.integers(0, len(centres))]
        ax = ax / np.linalg.norm(ax)
        a2 = ang + np.radians(rng.normal(scale=noise_r))
        q = np.concatenate([[np.cos(a2 / 2)], np.sin(a2 / 2) * (ax + rng.normal(scale=0.002, size=3))])
        q *= rng.uniform(0.995, 1.005)
        out[i, :4] = q
        out[i, 4:] = t + rng.normal(scale=noise_t, size=3)
    return out.astype(np.float32)


def _check_cluster(ctx, orc, pools, atol=1e-4):
    qt = np.concatenate(pools, 0)
    nh = np.array([len(p) for p in pools], np.int32)
    nc = ctx.cluster(qt, nh)
    want = orc.cluster(qt, nh)
    assert np.array_equal(nc, want), (nh, nc, want)
    assert np.array_equal(ctx.blob("cluster_num"), orc.blob("cluster_num")), (ctx.blob("cluster_num"), orc.blob("cluster_num"))
    names = set(orc.blob_names())
    for t in range(3):
        for nm in ("cluster_seed_sorted%d" % t, "cluster_size_sorted%d" % t):
            if nm in names:
                assert np.array_equal(ctx.blob(nm), orc.blob(nm)), (nm, nh)
            else:
                assert nh[t] <= 10, nm                 # only pools that bypass clustering have no seed list
        _same_f(ctx.blob("centre%d" % t).reshape(-1, 7), orc.blob("centre%d" % t).reshape(-1, 7), atol)
    return nc


def test_cluster_small_pools_and_bypass(ctx, orc):
    """Q12: an empty pool yields one identity centre, pools of up to cluster_number_threshold (10) pass through,
    11 is the first clustered size; the last hypothesis never seeds."""
    rng = np.random.default_rng(301)
    for sizes in ((0, 0, 0), (0, 10, 11), (1, 0, 0), (10, 10, 10), (11, 11, 11), (11, 0, 12), (2, 300, 0)):
        pools = [_pool(rng, n, 3) for n in sizes]
        nc = _check_cluster(ctx, orc, pools)
        for t, n in enumerate(sizes):
            if n == 0:
                assert nc[t] == 1 and np.array_equal(ctx.blob("centre%d" % t), np.array([1, 0, 0, 0, 0, 0, 0], np.float32))
            elif n <= 10:
                assert nc[t] == n


def test_cluster_num_zero_one_and_cutoff_walk(ctx, orc):
    """Q13: cluster_num = int(200 * pool / total) of 0 and 1 (a small pool beside a large one), and cut-off walks that
    decrement past small clusters until they run out (clusternum < 2) or stop at the first too-small cluster."""
    rng = np.random.default_rng(302)
    big = _pool(rng, 4000, 40)
    for small_n in (11, 15, 21, 45):                       # cluster_num = 0, 0, 1, 2
        _check_cluster(ctx, orc, [big, _pool(rng, small_n, 2), _pool(rng, 12, 12, noise_t=3.0)])
    # one dominant cluster and many singletons: the walk decrements from the big size down through the singletons
    cen = [(np.array([0, 0, 1.0]), 0.3, np.array([1.0, 2.0, 0.5]))]
    dom = _pool(rng, 60, 1, centres=cen)
    lone = _pool(rng, 60, 60, noise_t=4.0, noise_r=20.0)
    mix = np.concatenate([dom, lone])[rng.permutation(120)]
    _check_cluster(ctx, orc, [mix, _pool(rng, 0, 1), _pool(rng, 0, 1)])
    # equal-sized clusters: ties in the exchange sort (Q11) and a walk that emits all of them
    cens = [(rng.normal(size=3), np.radians(10.0 * k), np.array([3.0 * k, 0, 0])) for k in range(8)]
    eq = np.concatenate([_pool(rng, 12, 1, centres=[cc], noise_t=0.01, noise_r=0.05) for cc in cens])
    _check_cluster(ctx, orc, [eq, eq[::-1].copy(), _pool(rng, 0, 1)])


def test_cluster_nan_translations_and_duplicates(ctx, orc):
    """NaN / inf translations never pass the radius test (they stay unallocated singletons that can still seed empty
    clusters), exact duplicates tie on distance 0 and resolve by index."""
    rng = np.random.default_rng(303)
    p = _pool(rng, 200, 5)
    p[7, 4] = np.nan; p[50, 5] = np.inf; p[120, 4:] = np.nan; p[199, 4] = np.nan          # incl. the last one (never a seed)
    q = _pool(rng, 150, 4)
    q[10:20] = q[10]                                                                        # ten identical hypotheses
    q[40:45, :4] = 0.0                                                                      # zero quaternions: rotated axis 0, angle NaN (Q9)
    _check_cluster(ctx, orc, [p, q, _pool(rng, 30, 2)])


def test_cluster_large_random_pools(ctx, orc):
    """Both clustering paths (shared-memory pools up to 2816 hypotheses, global-memory pools beyond) on synthetic pools."""
    rng = np.random.default_rng(304)
    _check_cluster(ctx, orc, [_pool(rng, 2500, 30), _pool(rng, 6000, 60), _pool(rng, 900, 200, noise_t=0.3)])


def _rigid(rng, yaw, t):
    T = np.eye(4, dtype=np.float32)
    T[:3, :3] = scenes.ground_truth((yaw, rng.uniform(-3, 3), rng.uniform(-3, 3)), (0, 0, 0))[:3, :3]
    T[:3, 3] = t
    return T


def test_fuse_surviving_types_and_nan_scores(ctx, orc):
    """FCCF.cpp:1546-1606 + fuse_answer: 1, 2 and 3 types above the 0.8 gate, empty types, ties, all-zero scores and
    NaN fine scores (Q15: every comparison fails, no type survives, the fused matrix degenerates the reference's way)."""
    rng = np.random.default_rng(401)
    k = 4
    base = _rigid(rng, 25.0, (1.5, -0.8, 0.2))
    for trial in range(12):
        T = np.zeros((3, k, 4, 4), np.float32)
        for t in range(3):
            for j in range(k):
                T[t, j] = _rigid(rng, 25.0 + rng.normal(scale=0.3), np.array([1.5, -0.8, 0.2]) + rng.normal(scale=0.02, size=3))
        s1 = rng.uniform(0.1, 1.0, (3, k)).astype(np.float32)
        s2 = rng.uniform(0.05, 0.5, (3, k)).astype(np.float32)
        n_top = np.array([4, 4, 4], np.int32)
        if trial == 1: s1[1:] *= 0.1; s2[1:] *= 0.1                    # only type 0 passes the gate
        if trial == 2: s1[2] *= 0.05; s2[2] *= 0.05                    # two types
        if trial == 3: n_top = np.array([4, 0, 2], np.int32)           # an empty type keeps the identity with score 0
        if trial == 4: n_top = np.array([0, 0, 0], np.int32)           # nothing fine-verified: sums 0, scores 0/0
        if trial == 5: s2[:] = np.nan                                  # Q15
        if trial == 6: s2[1, 2] = np.nan                               # one NaN poisons score2_sum for every type (Q19)
        if trial == 7: s1[:] = 0.0; s2[:] = 0.0
        if trial == 8: s1[:] = 0.5; s2[:] = 0.25; T[:] = base          # exact ties: the first maximum wins
        if trial == 9: n_top = np.array([1, 1, 1], np.int32)
        if trial == 10: s1[0, 0] = np.inf
        if trial == 11: n_top = np.array([4, 4, 4], np.int32); T[1, :, :3, :3] = 0.0     # a degenerate rotation block in one type
        got = ctx.fuse(T, s1, s2, n_top)
        want = orc.fuse(T, s1, s2, n_top)
        assert np.array_equal(np.isnan(got), np.isnan(want)), (trial, got, want)
        np.testing.assert_allclose(np.nan_to_num(got), np.nan_to_num(want), rtol=0, atol=2e-5, err_msg="trial %d" % trial)
        a, b = ctx.blob("type_best").reshape(3, 13), orc.blob("type_best").reshape(3, 13)
        assert np.array_equal(np.isnan(a), np.isnan(b)), trial
        np.testing.assert_allclose(np.nan_to_num(a, posinf=1e30), np.nan_to_num(b, posinf=1e30), rtol=1e-5, atol=1e-6, err_msg="trial %d type_best" % trial)
