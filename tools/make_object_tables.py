"""Generate uclv_qs_pushing_matlab_b200/data/objects.json from the reference's .ply outlines.

Runs ONLY in the build container (needs /root/reference/cad_models).  It pushes each outline through the
product's own host-side ingest (qspush_model_create_from_ply: sortCadPoints + getSpline restated from
PusherSliderModel.m:84-132) and stores the resulting knot vector and control points.  Python's repr()
round-trips doubles exactly, and every value is a float32-representable number (pcread returns single).
The GPU box has no /root/reference, so the package and the tests read this table there.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from uclv_qs_pushing_matlab_b200.capi import Model  # noqa: E402
from uclv_qs_pushing_matlab_b200.object_selection import OBJECT_TABLE  # noqa: E402

CAD = "/root/reference/cad_models"
out = {}
for name, o in OBJECT_TABLE.items():
    flip = name in ("montana", "pulirapid")        # PusherSliderModel.m:107-109
    m = Model.from_ply(os.path.join(CAD, o["pcl_path"]), flip, 3, o["mu_sg"], o["mu_sp"], o["m"], o["tau_max"])
    out[name] = {"pcl_path": o["pcl_path"], "degree": 3, "knots": [float(v) for v in m.S],
                 "ctrl_xy": [[float(a), float(b)] for a, b in m.P], "b": m.b, "c_ellipse": m.c_ellipse}
    print(name, m.n, m.nknots, m.b, m.c_ellipse)
with open(os.path.join(ROOT, "uclv_qs_pushing_matlab_b200", "data", "objects.json"), "w") as f:
    json.dump(out, f, indent=1)
