"""objects_database/object_selection.m:1-46 — physical parameters of the four sliders."""
from __future__ import annotations

from types import SimpleNamespace

# object_selection.m:3-42
OBJECT_TABLE = {
    "santal": dict(mu_sg=0.32, mu_sp=0.19, xwidth=0.068, ywidth=0.082, m=0.2875, tau_max=0.0251,
                   cad_model_path="cad_santal_centered_scaled_rotated_reduced.stl",
                   pcl_path="planar_surface_santal_36_uniformed.ply"),
    "balea": dict(mu_sg=0.35, mu_sp=0.20, xwidth=0.071, ywidth=0.071, m=0.1713, tau_max=0.0042,
                  cad_model_path="Balea_cad_model v1.stl",
                  pcl_path="Balea_cad_model_planar_surface_36.ply"),
    "montana": dict(mu_sg=0.20, mu_sp=0.10, xwidth=0.057, ywidth=0.101, m=0.2467, tau_max=0.0101,
                    cad_model_path="Montana_cad_model.stl",
                    pcl_path="Montana_cad_model_planar_section_34.ply"),
    "pulirapid": dict(mu_sg=0.22, mu_sp=0.1, xwidth=0.13, ywidth=0.23, m=0.500, tau_max=0.0251,
                      cad_model_path="pulirapid_ricarica_simplified.stl",
                      pcl_path="pulirapid_ricarica_test_curvatura2_ply.ply"),
}


def object_selection(obj: str) -> SimpleNamespace:
    """slider = object_selection(obj) — same field names as the reference struct."""
    if obj not in OBJECT_TABLE:
        # object_selection.m:43-45 prints and returns; a Python mirror raises instead of returning nothing
        raise ValueError("Invalid object! Please, chose between: santal, balea, montana, pulirapid")
    o = dict(OBJECT_TABLE[obj])
    o["area"] = o["xwidth"] * o["ywidth"]
    return SimpleNamespace(**o)
