"""In-library multi-GPU paths on every visible GPU: fccf_register_batch_multi (one host thread per GPU) against the
single-GPU batch, fccf_score_sharded (8-byte ncclAllReduce) against one context.  python tools/multi_check.py [pairs]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fccf_pcr_b200 as fccf
from fccf_pcr_b200 import scenes

npairs = int(sys.argv[1]) if len(sys.argv) > 1 else 64
nd = fccf.device_count()
print("devices:", nd, flush=True)
base = [scenes.make_pair("indoor", 200000, 100 + i)[:2] for i in range(min(npairs, 16))]
pairs = (base * ((npairs + len(base) - 1) // len(base)))[:npairs]
srcs = [p[0].copy() for p in pairs]; tars = [p[1].copy() for p in pairs]
ctxs = [fccf.Context(d) for d in range(nd)]
hb = ctxs[0].prepare_batch(srcs, tars)
T1 = ctxs[0].register_batch_prepared(hb, 0.2).copy()
ok = True
for n in sorted({1, 2, nd}):
    if n > nd:
        continue
    for _ in range(2):
        Tn, tms = fccf.register_batch_multi(ctxs[:n], hb, 0.2)
    t0 = time.perf_counter()
    for _ in range(3):
        Tn, tms = fccf.register_batch_multi(ctxs[:n], hb, 0.2)
    dt = (time.perf_counter() - t0) / 3
    same = np.array_equal(Tn, T1, equal_nan=True)
    ok = ok and same
    print("register_batch_multi on %d GPU(s): %.2f ms per batch of %d pageable pairs (%.4f ms/registration), identical to one GPU: %s" % (n, dt * 1e3, npairs, dt * 1e3 / npairs, same), flush=True)
# sharded scoring
c0 = ctxs[0]
T0 = c0.register(srcs[0], tars[0], 0.2)
s1 = c0.blob("sub1").reshape(-1, 3).copy(); s2 = c0.blob("sub2").reshape(-1, 3).copy()
rng = np.random.default_rng(5)
H = 20000
hyps = np.tile(np.eye(4, dtype=np.float32), (H, 1, 1))
hyps[:, :3, :3] = T0[:3, :3]
hyps[:, :3, 3] = T0[:3, 3] + rng.uniform(-0.3, 0.3, (H, 3)).astype(np.float32)
hyps[777] = T0; hyps[15000] = T0          # ties: the smaller index must win
ref = c0.score_hypotheses(hyps, s1, s2)
for n in sorted({1, 2, nd}):
    if n > nd:
        continue
    sc, best, idx, used = fccf.score_sharded(ctxs[:n], hyps, s1, s2)
    good = np.array_equal(sc, ref) and idx == int(np.argmax(ref)) and best == float(ref.max())
    ok = ok and good
    print("score_sharded on %d GPU(s): best %.6f at %d (argmax %d), nccl all-reduce used: %s, scores identical: %s" % (n, best, idx, int(np.argmax(ref)), used, np.array_equal(sc, ref)), flush=True)
print("ALL OK" if ok else "MISMATCH")
sys.exit(0 if ok else 1)
