import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from tools.gpu_sanity import time_rti
for B in (16384, 65536):
    r = time_rti(B, 40, 32, 1e-12, reps=3, qp_kernel=0)
    print("B %6d v1 ppw32 : %8.2f ms  %9.0f it/s" % (B, r["ms"], r["its_per_s"]), flush=True)
