"""feotensor_b200 -- B200-native compute backend behind feotensor's public API.

Host-side mirror of the reference interface (Shape, Coordinate, Axes, Tensor, Matrix, Vector,
ShapeError) over the C ABI of libfeotensor_b200.so (include/feotensor_b200.h): hand-written sm_100a
CUDA kernels, device-resident storage, no CPU fallback.
"""
from ._lib import LIB_PATH, load as load_library
from .core import (BackendError, Context, Coordinate, DeviceStorage, IndexIterator, Matrix, ReferencePanic, Shape,
                   ShapeError, Tensor, Vector, coord, dot_batched, shape)

Axes = list  # axes.rs:1 -- `pub type Axes = Vec<usize>`
DynamicTensor, DynamicMatrix, DynamicVector = Tensor, Matrix, Vector  # the reference's own aliases

__all__ = ["Axes", "BackendError", "Context", "Coordinate", "DeviceStorage", "DynamicMatrix", "DynamicTensor",
           "DynamicVector", "IndexIterator", "LIB_PATH", "Matrix", "ReferencePanic", "Shape", "ShapeError", "Tensor",
           "Vector", "coord", "dot_batched", "load_library", "shape"]
