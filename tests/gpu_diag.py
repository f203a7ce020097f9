"""Stage-by-stage GPU-vs-oracle diagnostic (not a pytest file): prints one line per intermediate."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fccf_pcr_b200 import scenes, Context
from oracle.oracle import Oracle

INT_BLOBS = ["vg1_cell1", "vg1_cell2", "vg1_cnt1", "vg1_cnt2", "vg2_cell1", "vg2_cell2", "vg2_cnt1", "vg2_cnt2",
             "oct_depth1", "oct_depth2", "vox_key1", "vox_key2", "vox_cnt1", "vox_cnt2", "vox_flag1", "vox_flag2", "vox_pidx1", "vox_pidx2",
             "grow_label1", "grow_label2", "merge_label1", "merge_label2", "n_stage1_faces1", "n_stage1_faces2",
             "face_id1", "face_id2", "face_nvox1", "face_nvox2", "face_vox1", "face_vox2", "face_off1", "face_off2",
             "base1", "base2", "matches", "n_hyp", "n_centres", "cluster_num"]
FLT_BLOBS = ["vg1_xyz1", "vg1_xyz2", "vg2_xyz1", "vg2_xyz2", "cloud_centroid1", "cloud_centroid2", "oct_min1", "oct_min2",
             "vox_plane1", "vox_plane2", "sub1", "sub2", "pvox1", "pvox2", "face_plane1", "face_plane2", "face_theta1", "face_theta2",
             "base_angle1", "base_angle2"]
for t in range(3):
    INT_BLOBS += ["cluster_seed_sorted%d" % t, "cluster_size_sorted%d" % t, "qv_pairs%d" % t, "qv_pair_off%d" % t, "qv_iters%d" % t, "top_centre%d" % t, "fv_off%d" % t]
    FLT_BLOBS += ["hyp%d" % t, "hyp_qt%d" % t, "centre%d" % t, "qv_score%d" % t, "qv_T%d" % t, "top_T%d" % t, "top_s1%d" % t, "top_s2%d" % t]
FLT_BLOBS += ["type_best"]


def sort_counts(o, t):
    c = o.blob("fv_counts%d" % t).reshape(-1, 5)
    off = o.blob("fv_off%d" % t)
    out = []
    for k in range(len(off) - 1):
        r = c[off[k]:off[k + 1]]
        order = np.lexsort((r[:, 2], r[:, 1], r[:, 0]))
        out.append(r[order])
    return np.concatenate(out, 0).reshape(-1) if out else np.zeros(0, np.int32)


def compare(ctx, o, verbose=True):
    bad = 0
    for name in INT_BLOBS + FLT_BLOBS + ["fv_counts0", "fv_counts1", "fv_counts2"]:
        try:
            a = ctx.blob(name)
        except Exception as e:
            print("%-22s GPU blob error: %s" % (name, e)); bad += 1; continue
        try:
            if name.startswith("fv_counts"):
                b = sort_counts(o, int(name[-1]))
            else:
                b = o.blob(name)
        except KeyError:
            print("%-22s (no oracle blob)" % name); continue
        if a.shape != b.shape:
            print("%-22s SHAPE gpu %s oracle %s" % (name, a.shape, b.shape)); bad += 1
            n = min(len(a), len(b)); a = a[:n]; b = b[:n]
            if n == 0: continue
        if a.dtype.kind in "iu":
            nm = int((a != b).sum())
            first = int(np.argmax(a != b)) if nm else -1
            print("%-22s n=%-8d mismatches=%d%s" % (name, len(a), nm, (" first@%d gpu=%d ora=%d" % (first, a[first], b[first])) if nm else ""))
            bad += nm > 0
        else:
            a64 = a.astype(np.float64); b64 = b.astype(np.float64)
            both_nan = np.isnan(a64) & np.isnan(b64)
            d = np.where(both_nan, 0.0, np.abs(a64 - b64))
            d = np.where(np.isnan(d), np.inf, d)
            neq = int(((a != b) & ~both_nan).sum())
            mx = float(d.max()) if len(d) else 0.0
            rel = float((d / np.maximum(np.abs(b64), 1e-30)).max()) if len(d) else 0.0
            print("%-22s n=%-8d not-bit-equal=%-7d max|d|=%.3e maxrel=%.3e" % (name, len(a), neq, mx, rel))
            bad += neq > 0
    return bad


def main():
    cases = [("indoor", 50000, 1, 0.1), ("indoor", 200000, 2, 0.2)]
    prm = {}
    if len(sys.argv) > 1:
        cases = [(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4]))]
        for kv in sys.argv[5:]:
            k, v = kv.split("="); prm[k] = float(v)
    ctx = Context(0, **prm)
    for kind, n, seed, leaf in cases:
        src, tar, Tgt = scenes.make_pair(kind, n, seed)
        o = Oracle(**prm)
        To = o.register(src, tar, leaf)
        t0 = time.time()
        try:
            Tg = ctx.register(src, tar, leaf)
        except Exception as e:
            print("GPU register raised:", e)
            Tg = np.zeros((4, 4), np.float32)
        t1 = time.time()
        tm = ctx.timing
        print("=== %s n=%d seed=%d leaf=%g: gpu wall %.1f ms (h2d %.3f ds %.3f pipe %.3f d2h %.3f total %.3f ms, %d launches); oracle pipe %.1f ms total %.1f ms" %
              (kind, n, seed, leaf, 1e3 * (t1 - t0), tm.h2d_ms, tm.downsample_ms, tm.pipeline_ms, tm.d2h_ms, tm.total_ms, tm.n_launches, 1e3 * o.time_pipeline, 1e3 * o.time_total))
        print("final T gpu:\n", Tg, "\noracle:\n", To)
        print("rot diff deg %.6f trans diff %.6g | vs GT: gpu %.3f deg %.4f m, oracle %.3f deg %.4f m" % (
            scenes.rotation_error_deg(Tg, To), scenes.translation_error(Tg, To), scenes.rotation_error_deg(Tg, Tgt), scenes.translation_error(Tg, Tgt),
            scenes.rotation_error_deg(To, Tgt), scenes.translation_error(To, Tgt)))
        bad = compare(ctx, o)
        # stage-wise inlier counts: the oracle's fine_verify fed with the GPU's own refined transforms
        s1, s2 = ctx.blob("sub1").reshape(-1, 3), ctx.blob("sub2").reshape(-1, 3)
        for t in range(3):
            Tt = ctx.blob("top_T%d" % t).reshape(-1, 4, 4); To_ = o.blob("top_T%d" % t).reshape(-1, 4, 4)
            cg = ctx.blob("fv_counts%d" % t).reshape(-1, 5); off = ctx.blob("fv_off%d" % t)
            for k in range(len(Tt)):
                so, rows = o.fine_verify(Tt[k], s1, s2)
                rows = rows[np.lexsort((rows[:, 2], rows[:, 1], rows[:, 0]))]
                g = cg[off[k]:off[k + 1]]
                print("  type %d top %d: rows gpu %d oracle(on gpu T) %d equal=%s | T diff vs oracle T: %.3e deg %.3e m" % (
                    t, k, len(g), len(rows), np.array_equal(g, rows), scenes.rotation_error_deg(Tt[k], To_[k]) if k < len(To_) else -1, scenes.translation_error(Tt[k], To_[k]) if k < len(To_) else -1))
        print("=== blobs with differences: %d" % bad)
        # repeat timing
        for r in range(3):
            ctx.register(src, tar, leaf)
            tm = ctx.timing
            print("  rerun %d: h2d %.3f ds %.3f pipe %.3f d2h %.3f total %.3f ms" % (r, tm.h2d_ms, tm.downsample_ms, tm.pipeline_ms, tm.d2h_ms, tm.total_ms))


if __name__ == "__main__":
    main()
