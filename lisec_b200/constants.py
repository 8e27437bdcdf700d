"""The shapes and caps of the hot path, value for value those of the reference's Constants.py:7-23."""
from __future__ import annotations

# size of voxel (Constants.py:7-9)
voxelx = 0.5
voxely = 0.25
voxelz = 0.25

# number of voxels in the space we care about (Constants.py:12-14): -50..50 m, -50..50 m, 0..2 m
nx = int(100 / voxelx)
ny = int(100 / voxely)
nz = int(2 / voxelz)

# limit of points per voxel (Constants.py:20)
maxPoints = 35

# index of the point axis in every VFE tensor (Constants.py:23)
pointIndex = -2

# VFE output widths of createModel (model_training.py:231-233)
vfe_widths = (16, 32, 64)

# Keras BatchNormalization default epsilon (model_training.py:171 passes no arguments)
bn_epsilon = 1e-3
