"""Lean sequences and their miss path: launches and results of small / small / big / big registrations on one context
against a fresh context (FCCF_NO_LEAN=1 for the comparison).  python tools/lean_diag.py"""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fccf_pcr_b200 as fccf
from fccf_pcr_b200 import scenes
small, ls = scenes.make_pair("indoor", 20000, 7), 0.2
big, lb = scenes.make_pair("indoor", 50000, 1), 0.1
f = fccf.Context(0); Tb1 = f.register(big[0], big[1], lb).copy(); Tb2 = f.register(big[0], big[1], lb).copy(); print('fresh big twice equal', np.array_equal(Tb1, Tb2), f.timing.n_launches); f.close()
c = fccf.Context(0)
c.register(small[0], small[1], ls); print('launches', c.timing.n_launches)
c.register(small[0], small[1], ls); print('launches', c.timing.n_launches)
T2 = c.register(big[0], big[1], lb).copy(); print('launches', c.timing.n_launches, 'equal to fresh', np.array_equal(T2, Tb1), np.abs(T2 - Tb1).max())
for nm in ['n_hyp', 'n_centres', 'top_s20', 'top_s10', 'type_best']:
    print(nm, c.blob(nm)[:12])
T3 = c.register(big[0], big[1], lb).copy(); print('again', np.array_equal(T3, Tb1), c.timing.n_launches)
