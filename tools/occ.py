import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fccf_pcr_b200 as f
c = f.Context(0)
L = f.lib()
L.fccf_debug_int.restype = ctypes.c_int
print("max active clusters:", L.fccf_debug_int(b"vg_fast_max_clusters"))
