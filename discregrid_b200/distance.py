"""Host mirror of Discregrid::TriangleMeshDistance (geometry/TriangleMeshDistance.h:93-208) over the C-ABI.

Same surface -- construct from a TriangleMesh or from (vertices, triangles), signed_distance / unsigned_distance
returning Result(distance, nearest_point, nearest_entity, triangle_id) -- but batched: the argument is an (n, 3)
array of points (a single point is promoted) and each Result field is an array.  All queries run in the sm_100a
kernel of csrc/k1_sdf.cu; there is no host implementation.
"""
import ctypes as C
from collections import namedtuple
from enum import IntEnum

import numpy as np

from . import _capi as capi
from .mesh import TriangleMesh


class NearestEntity(IntEnum):          # TriangleMeshDistance.h:75
    V0 = 0
    V1 = 1
    V2 = 2
    E01 = 3
    E12 = 4
    E02 = 5
    F = 6


Result = namedtuple("Result", ["distance", "nearest_point", "nearest_entity", "triangle_id"])   # :80-86


class TriangleMeshDistance:
    def __init__(self, mesh_or_vertices=None, triangles=None):
        self._h = None
        if mesh_or_vertices is not None:
            self.construct(mesh_or_vertices, triangles)

    # TriangleMeshDistance.h:251-267 (the std::vector overload; the raw-pointer overload's 3x over-allocation bug,
    # :232-249, is not reproduced)
    def construct(self, mesh_or_vertices, triangles=None):
        self.close()
        if isinstance(mesh_or_vertices, TriangleMesh):
            V, F = mesh_or_vertices.vertex_data(), mesh_or_vertices.face_data()
        else:
            V, F = mesh_or_vertices, triangles
        V = np.ascontiguousarray(V, dtype=np.float64).reshape(-1, 3)
        F = np.ascontiguousarray(F, dtype=np.uint32).reshape(-1, 3)
        h = C.c_void_p()
        capi.check(capi.lib.dg_mesh_create(capi.ptr(V, capi.F64P), len(V), capi.ptr(F, capi.U32P), len(F), C.byref(h)))
        self._h = h
        self.n_vertices, self.n_triangles = len(V), len(F)
        return self

    @property
    def is_constructed(self):
        return self._h is not None

    @property
    def handle(self):
        if self._h is None:
            # reference: prints "DistanceTriangleMesh error: not constructed." and exit(-1) (:318-321)
            raise capi.DiscregridError(capi.DG_ERR_INVALID, "DistanceTriangleMesh error: not constructed.")
        return self._h

    def close(self):
        if getattr(self, "_h", None) is not None:
            capi.lib.dg_mesh_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self):
        a = (C.c_uint64 * 8)()
        capi.check(capi.lib.dg_mesh_info(self.handle, a))
        return {"n_vertices": a[0], "n_triangles": a[1], "stack_depth": a[2], "watertight_flags": a[3],
                "device_bytes": a[4], "build_us": a[5], "upload_us": a[6]}

    def tree(self):
        """(spheres[n,8], kids[n,2]) in the reference's node numbering (diagnostics)."""
        n = 2 * self.n_triangles - 1
        sph = np.empty((n, 8)); kids = np.empty((n, 2), np.int32)
        capi.check(capi.lib.dg_mesh_tree(self.handle, capi.ptr(sph, capi.F64P), capi.ptr(kids, capi.I32P)))
        return sph, kids

    def pseudonormals(self):
        tri = np.empty((self.n_triangles, 3)); edge = np.empty((self.n_triangles, 3, 3)); vert = np.empty((self.n_vertices, 3))
        capi.check(capi.lib.dg_mesh_pseudonormals(self.handle, capi.ptr(tri, capi.F64P), capi.ptr(edge, capi.F64P),
                                                  capi.ptr(vert, capi.F64P)))
        return tri, edge, vert

    def _query(self, points, signed):
        x = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
        n = len(x)
        dist = np.empty(n); near = np.empty((n, 3)); ent = np.empty(n, np.int32); tri = np.empty(n, np.int32)
        capi.check(capi.lib.dg_mesh_distance(self.handle, capi.ptr(x, capi.F64P), n, int(signed), capi.ptr(dist, capi.F64P),
                                             capi.ptr(near, capi.F64P), capi.ptr(ent, capi.I32P), capi.ptr(tri, capi.I32P)))
        return Result(dist, near, ent, tri)

    def signed_distance(self, points):      # :269-314
        return self._query(points, True)

    def unsigned_distance(self, points):    # :316-334
        return self._query(points, False)


class MeshSignedDistance:
    """The functor GenerateSDF hands to addFunction (cmd/generate_sdf/main.cpp:94-102):
    x -> sign * md.signed_distance(x).distance, sign = -1 for --invert.  CubicLagrangeDiscreteGrid.addFunction
    recognises this type and samples it on the GPU (the reference takes an opaque std::function, which a GPU
    cannot run -- SURVEY F6/H1)."""

    def __init__(self, md, invert=False):
        self.md = md
        self.sign = -1.0 if invert else 1.0

    def __call__(self, points):
        d = self.md.signed_distance(points).distance
        return d if self.sign == 1.0 else self.sign * d
