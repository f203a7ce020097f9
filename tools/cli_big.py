"""FCCF command line on a large binary PLY pair (f1: memory-mapped, used in place)."""
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fccf_pcr_b200 as fccf
from fccf_pcr_b200 import scenes

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
leaf = sys.argv[2] if len(sys.argv) > 2 else "0.1"
src, tar, _ = scenes.make_pair("indoor", n, 5)
scenes.write_ply("/tmp/src_big.ply", src)
scenes.write_ply("/tmp/tar_big.ply", tar)
t = time.perf_counter()
r = subprocess.run([fccf.CLI_PATH, "/tmp/src_big.ply", "/tmp/tar_big.ply", leaf], capture_output=True, text=True)
print("wall %.3f s (process start, CUDA context, two registrations: cold + warm)" % (time.perf_counter() - t))
print(r.stdout)
print(r.stderr[-300:])
