"""lisec_b200 — the VoxelNet front end of bot15498/Lisec (point->voxel grouping, stacked VFE, dense-grid scatter)
as hand-written sm_100a CUDA behind a C ABI. Importing the package does not load CUDA; creating a Frontend does."""
from . import constants  # noqa: F401
from ._native import LIB_PATH, LisecError  # noqa: F401

__all__ = ["constants", "LisecError", "LIB_PATH", "Frontend"]


def __getattr__(name):
    if name == "Frontend":
        from .frontend import Frontend

        return Frontend
    raise AttributeError(name)
