import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from tools.gpu_sanity import time_rti
for B in (4096, 8192, 16384, 32768, 65536, 131072):
    for kern, ppw in ((1, 8), (0, 32), (0, 16)):
        r = time_rti(B, 40, ppw, 1e-12, reps=3, qp_kernel=kern)
        print("B %6d kernel %d ppw %2d : %8.2f ms  %9.0f it/s  (kipm max %d)" % (B, kern, ppw, r["ms"], r["its_per_s"], r["qp_iter_max"]), flush=True)
