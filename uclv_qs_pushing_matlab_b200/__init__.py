"""uclv_qs_pushing_matlab_b200 — B200-native batched NMPC engine for the quasi-static pusher-slider.

Host-side mirror of the reference's MATLAB interface (same class and method names, argument meaning
and error behaviour) over the C-ABI library libqspush.so (include/qspush.h), whose sm_100a kernels do
all of the arithmetic.  There is no CPU path: without the built library the import fails, and without
a CUDA device every compute call raises QspushError.

    object_selection        objects_database/object_selection.m
    bspline_shape           acados_nmpc/bspline_shape.m
    PusherSliderModel       acados_nmpc/PusherSliderModel.m
    NMPC_controller         acados_nmpc/NMPC_controller.m   (+ acados_ocp: the subset of acados' MATLAB class it drives)
    helper                  helper.m (closed_loop_matlab, open_loop_matlab)
    TrajectoryGenerator     acados_nmpc/TrajectoryGenerator.m
"""
from . import _lib
from ._lib import QspushError
from .capi import Model, Solver, default_ctrl, default_opts, measure_fp64_peak

_lib.lib()  # fail loudly at import time if libqspush.so is missing or does not export the C-ABI

from .object_selection import OBJECT_TABLE, object_selection  # noqa: E402
from .bspline_shape import bspline_shape  # noqa: E402
from .pusher_slider_model import PusherSliderModel  # noqa: E402
from .acados_shim import acados_ocp  # noqa: E402
from .nmpc_controller import NMPC_controller  # noqa: E402
from .helper import helper  # noqa: E402
from .trajectory_generator import TrajectoryGenerator  # noqa: E402
from . import sharding  # noqa: E402

__all__ = ["QspushError", "Model", "Solver", "default_opts", "default_ctrl", "measure_fp64_peak", "OBJECT_TABLE", "object_selection",
           "bspline_shape", "PusherSliderModel", "acados_ocp", "NMPC_controller", "helper", "TrajectoryGenerator",
           "sharding"]
