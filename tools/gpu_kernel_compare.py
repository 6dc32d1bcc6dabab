"""Warp kernel (qp_kernel=1) against the thread kernel (qp_kernel=0) over batch sizes and horizons (development aid)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from tools.gpu_sanity import time_rti
rows = []
for N, Bs in ((40, (1024, 4096, 8192, 16384, 32768, 65536)), (10, (4096, 65536)), (100, (4096, 32768))):
    for B in Bs:
        for kern, ppw in ((1, 8), (0, 32), (0, 8)):
            r = time_rti(B, N, ppw, 1e-12, reps=3, qp_kernel=kern)
            rows.append(dict(N=N, B=B, kernel=kern, ppw=ppw, ms=r["ms"], its=r["its_per_s"]))
            print("N %3d B %6d kernel %d ppw %2d : %8.2f ms  %9.0f it/s  (kipm max %d)" % (N, B, kern, ppw, r["ms"], r["its_per_s"], r["qp_iter_max"]), flush=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "kernel_compare.json"), "w"), indent=1)
