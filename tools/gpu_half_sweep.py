"""One problem per warp (SEG 32) against two per warp (SEG 16, env QSPUSH_QW_HALF=1) over horizons (development aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from tools.gpu_sanity import time_rti
for N in (20, 31, 32, 40, 47, 48, 55, 63):
    for B in (4096, 16384):
        r = time_rti(B, N, 8, 1e-12, reps=3, qp_kernel=1)
        print("N %3d B %6d half=%s : %8.3f ms %9.0f it/s ok %.3f kipm %.2f" % (N, B, os.environ.get("QSPUSH_QW_HALF", "0"), r["ms"], r["its_per_s"], r["status_ok"], r["qp_iter_mean"]), flush=True)
