"""The HDF5 writer moved into the product (lisec_b200/h5write.py: model.save() needs it); the tests keep this name."""
from lisec_b200.h5write import write_h5  # noqa: F401
