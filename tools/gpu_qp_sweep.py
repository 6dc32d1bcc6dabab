"""QP kernel timing sweep (development aid; run under gpurun)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.gpu_sanity import time_rti
out = []
cfgs = [(4096, 32, 1e-12), (4096, 16, 1e-12), (4096, 8, 1e-12), (4096, 8, 1e-8), (16384, 32, 1e-12), (65536, 32, 1e-12)]
if len(sys.argv) > 1:
    cfgs = [tuple(float(v) if 'e' in v else int(v) for v in a.split(',')) for a in sys.argv[1:]]
for B, ppw, tol in cfgs:
    r = time_rti(int(B), 40, int(ppw), tol)
    print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items()}, flush=True)
    out.append(r)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "qp_sweep.json"), "w"), indent=1)
