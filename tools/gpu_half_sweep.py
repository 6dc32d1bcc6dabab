"""RTI step over horizons (development aid): run once per library variant to compare warp mappings."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from tools.gpu_sanity import time_rti
Ns = [int(a) for a in sys.argv[1:]] or [20, 31, 32, 40, 47, 48, 55, 63]
for N in Ns:
    for B in (4096, 16384):
        r = time_rti(B, N, 8, 1e-12, reps=3, qp_kernel=1)
        print("N %3d B %6d : %8.3f ms %9.0f it/s ok %.3f kipm %.2f" % (N, B, r["ms"], r["its_per_s"], r["status_ok"], r["qp_iter_mean"]), flush=True)
