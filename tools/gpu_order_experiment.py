"""Development aid: does a planned work-queue order (bin packing by predicted IPM iteration count) shorten the QP launch?
Needs a library built with -DQSPUSH_DEV_ORDER (qsdev_set_order).  Prints the event-timed QP phase for the built-in order
(descending iteration count) and for planned orders."""
import os, sys, ctypes as C, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
import uclv_qs_pushing_matlab_b200 as q
from uclv_qs_pushing_matlab_b200 import _lib as L
from uclv_qs_pushing_matlab_b200.workloads import make_rti_workload, packaged_model

B, N, W, PPW, OV = 4096, 40, 592, 2, 1
gm = packaged_model("santal")
wl = make_rti_workload(B, N, seed=2)

def make():
    s = q.Solver([gm], N, 0.05, B)
    s.set("yref", wl["yref"]); s.set("yref_e", wl["yref_e"]); s.set("u", wl["u_init"]); s.set_int("cold", np.zeros(B, dtype=np.int32))
    tr = np.zeros((N, 6)); tr[:, 0] = 0.01 * (np.arange(N) * 0.05)
    off = np.zeros((B, 6)); off[:, :2] = wl["x0"][:, :2]
    s.set_reference_trajectory(tr, off)
    s.snapshot_guess()
    return s

def time_steps(s, n=20):
    dx0 = torch.from_numpy(wl["x0"]).cuda(); du0 = torch.empty(B, 2, dtype=torch.float64, device="cuda"); dst = torch.empty(B, dtype=torch.int32, device="cuda")
    for _ in range(4): s.step(dx0, 1, du0, dst, restore_guess=True)
    s.sync(); qp = []
    for _ in range(n):
        s.step(dx0, 1, du0, dst, restore_guess=True); s.sync(); qp.append(s.stat("time_qp_sol") * 1e3)
    return float(np.median(qp)), du0.cpu().numpy()

def plan(iters, slack=0):
    order = np.argsort(-iters, kind="stable")
    U = B // PPW
    pred = np.array([iters[order[u * PPW]] for u in range(U)])
    cls = collections.Counter(pred.tolist()); classes = sorted(cls, reverse=True)
    start = {c: 0 for c in classes}; a = 0
    for c in classes: start[c] = a; a += cls[c]
    total = sum((c + OV) * n for c, n in cls.items()); cap = -(-total // W) + slack
    while True:
        n = dict(cls); wl_ = W; groups = []; ok = True
        while sum(n.values()) > 0:
            if wl_ == 0: ok = False; break
            rem = cap; pat = []
            for c in classes:
                m = min(n[c], rem // (c + OV))
                if m > 0: pat.append((c, m)); rem -= m * (c + OV)
            if not pat: ok = False; break
            k = max(1, min(min(n[c] // m for c, m in pat), wl_)); st = 0
            for c, m in pat:
                for _ in range(m): groups.append((st, c, k)); st += c + OV
                n[c] -= k * m
            wl_ -= k
        if ok: break
        cap += 1
    cur = dict(start); out = []
    for st, c, k in sorted(groups, key=lambda g: (g[0], -g[1])):
        for j in range(k):
            u = cur[c]; cur[c] += 1
            out.extend(order[u * PPW:(u + 1) * PPW])
    out = np.array(out, dtype=np.int32)
    assert sorted(out.tolist()) == list(range(B))
    return out, cap

s = make(); t0, u_ref = time_steps(s)
iters = s.get_int("qp_iter")
print("built-in order (descending iteration count): QP %.4f ms  k_ipm %.3f" % (t0, iters.mean()))
fn = L.lib().qsdev_set_order; fn.restype = C.c_int; fn.argtypes = [C.c_void_p, C.c_void_p]
for slack in (0, 1, 2):
    order, cap = plan(iters, slack)
    s2 = make(); fn(s2._h, order.ctypes.data)
    t, u = time_steps(s2)
    print("planned order, capacity %d iteration-times: QP %.4f ms (%.1f %%), u0 identical %s" % (cap, t, 100 * (t / t0 - 1), np.array_equal(u, u_ref)))
rev = np.argsort(-iters, kind="stable").astype(np.int32)
s3 = make(); fn(s3._h, rev.ctypes.data); t, u = time_steps(s3); print("host-set descending order (control): QP %.4f ms" % t)
