"""model3d_b200: the ray-tracing hot path of unixpickle/model3d rebuilt for B200 (sm_100a).

The product is ``libm3dgpu.so`` (CUDA kernels behind the C ABI of ``include/m3d.h``);
this package is the thin host-side mirror of the reference's ``model3d`` collision
interface and ``render3d`` renderer interface for that path.
"""
from . import _native  # noqa: F401
from .model3d import (BatchCollisions, ColliderContains, ColliderSolid, MeshCollider,  # noqa: F401
                      MeshSDF, MeshToCollider, MeshToInterpNormalCollider, MeshToSDF, NewColliderSolid,
                      NewColliderSolidHollow, NewColliderSolidInset, Ray, RayCollision,
                      TriangleCollision, UnsupportedError)

from . import render3d  # noqa: F401,E402

__all__ = ["render3d", "MeshCollider", "MeshToCollider", "MeshToInterpNormalCollider", "Ray", "RayCollision",
           "TriangleCollision", "BatchCollisions", "UnsupportedError", "MeshToSDF", "MeshSDF", "ColliderContains",
           "ColliderSolid", "NewColliderSolid", "NewColliderSolidInset", "NewColliderSolidHollow"]
