"""Config 2: 1M (x,u) samples through k_eval_erk4 with device-resident buffers (development aid)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from tests.workloads import gpu_model
from uclv_qs_pushing_matlab_b200.workloads import make_samples_config2
gm = gpu_model("santal"); dev = torch.device("cuda:0")
n = 1 << 20
x, u = make_samples_config2(gm.b, n, knots=gm.S)
xd, ud = torch.from_numpy(x).to(dev), torch.from_numpy(u).to(dev)
Phi = torch.empty(n, 4, dtype=torch.float64, device=dev); A = torch.empty(n, 4, 4, dtype=torch.float64, device=dev); B = torch.empty(n, 4, 2, dtype=torch.float64, device=dev)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
for _ in range(3): gm.eval_erk4_sens_device(xd, ud, 0.05, Phi, A, B)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps): gm.eval_erk4_sens_device(xd, ud, 0.05, Phi, A, B)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print("erk4_sens 1M samples: %.4f ms  %.2f Gsamples/s  %.2f TFLOP/s (2.3 kflop/sample)  %.2f TB/s (272 B/sample)" % (ms, n / ms / 1e6, 2300 * n / ms / 1e9, 272 * n / ms / 1e9))
