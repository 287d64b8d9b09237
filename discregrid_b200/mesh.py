"""Host-side triangle mesh container and loaders (input side of the hot path).

`TriangleMesh(path)` mirrors the reference's OBJ reader (discregrid/src/mesh/triangle_mesh.cpp:90-124): only
`v x y z` and `f a b c` lines are used, `a/b/c` face tokens keep the part before the first '/', indices are
1-based, triangles only.  The half-edge structure of the reference mesh class is not needed by the hot path
(TriangleMeshDistance only reads vertex_data()/face_data(), TriangleMeshDistance.h:227-230).
Synthetic closed meshes are provided for benchmarks (the 100,000-triangle bumpy torus).
"""
import numpy as np


class TriangleMesh:
    def __init__(self, path_or_vertices, faces=None):
        if faces is None:
            v, f = _read_obj(path_or_vertices)
        else:
            v, f = path_or_vertices, faces
        self.vertices = np.ascontiguousarray(v, dtype=np.float64).reshape(-1, 3)
        self.faces = np.ascontiguousarray(f, dtype=np.uint32).reshape(-1, 3)

    # reference accessors (triangle_mesh.hpp): vertex_data(), face_data(), nVertices(), nFaces()
    def vertex_data(self):
        return self.vertices

    def face_data(self):
        return self.faces

    def nVertices(self):
        return len(self.vertices)

    def nFaces(self):
        return len(self.faces)

    def exportOBJ(self, path):
        """triangle_mesh.cpp:126-147"""
        with open(path, "w") as fh:
            fh.write("g default\n")
            for p in self.vertices:
                fh.write(f"v {float(p[0])!r} {float(p[1])!r} {float(p[2])!r}\n")
            for t in self.faces:
                fh.write(f"f {t[0] + 1} {t[1] + 1} {t[2] + 1}\n")


def _read_obj(path):
    """the library's OBJ reader (dg_obj_read: the records Discregrid::TriangleMesh(path) accepts, triangle_mesh.cpp:90-124)"""
    import ctypes as C
    from . import _capi as capi
    v, f = capi.F64P(), capi.U32P()
    nv, nf = C.c_uint64(), C.c_uint64()
    rc = capi.lib.dg_obj_read(str(path).encode(), C.byref(v), C.byref(nv), C.byref(f), C.byref(nf))
    if rc == capi.DG_ERR_IO:
        raise FileNotFoundError((capi.lib.dg_last_error() or b"").decode("utf-8", "replace"))
    capi.check(rc)
    try:
        verts = np.ctypeslib.as_array(v, shape=(nv.value, 3)).copy() if nv.value else np.zeros((0, 3))
        faces = np.ctypeslib.as_array(f, shape=(nf.value, 3)).copy() if nf.value else np.zeros((0, 3), np.uint32)
    finally:
        capi.lib.dg_obj_free(v, f)
    if not len(verts) and not len(faces):
        raise ValueError(f"no 'v'/'f' records in {path}")
    return verts, faces


def bumpy_torus(nu=250, nv=200, R=1.0, r0=0.4, amp=0.05, ku=7, kv=5):
    """Deterministic closed, consistently oriented torus with a bumpy minor radius
    r(theta, phi) = r0 + amp*sin(ku*theta)*cos(kv*phi); nu x nv quads -> 2*nu*nv triangles, nu*nv vertices.
    Defaults give exactly 100,000 triangles / 50,000 vertices (BASELINE.md target config)."""
    th = 2.0 * np.pi * np.arange(nu) / nu
    ph = 2.0 * np.pi * np.arange(nv) / nv
    T, P = np.meshgrid(th, ph, indexing="ij")
    r = r0 + amp * np.sin(ku * T) * np.cos(kv * P)
    x = (R + r * np.cos(P)) * np.cos(T)
    y = (R + r * np.cos(P)) * np.sin(T)
    z = r * np.sin(P)
    V = np.stack([x, y, z], -1).reshape(-1, 3)
    i, j = np.meshgrid(np.arange(nu), np.arange(nv), indexing="ij")
    a = (i * nv + j).ravel()
    b = (((i + 1) % nu) * nv + j).ravel()
    c = (((i + 1) % nu) * nv + (j + 1) % nv).ravel()
    d = (i * nv + (j + 1) % nv).ravel()
    F = np.concatenate([np.stack([a, b, c], -1), np.stack([a, c, d], -1)], 0)
    return TriangleMesh(V, F.astype(np.uint32))


def uv_sphere(n_lat=32, n_lon=64, radius=1.0, center=(0.0, 0.0, 0.0)):
    """Closed UV sphere (poles are single vertices), outward orientation."""
    V = [(0.0, 0.0, radius)]
    for a in range(1, n_lat):
        t = np.pi * a / n_lat
        for b in range(n_lon):
            p = 2.0 * np.pi * b / n_lon
            V.append((radius * np.sin(t) * np.cos(p), radius * np.sin(t) * np.sin(p), radius * np.cos(t)))
    V.append((0.0, 0.0, -radius))
    V = np.array(V) + np.array(center)
    F = []
    ring = lambda a, b: 1 + (a - 1) * n_lon + (b % n_lon)
    for b in range(n_lon):
        F.append((0, ring(1, b), ring(1, b + 1)))
    for a in range(1, n_lat - 1):
        for b in range(n_lon):
            F.append((ring(a, b), ring(a + 1, b), ring(a + 1, b + 1)))
            F.append((ring(a, b), ring(a + 1, b + 1), ring(a, b + 1)))
    south = len(V) - 1
    for b in range(n_lon):
        F.append((south, ring(n_lat - 1, b + 1), ring(n_lat - 1, b)))
    return TriangleMesh(V, np.array(F, np.uint32))


def box(lo=(-1.0, -1.0, -1.0), hi=(1.0, 1.0, 1.0)):
    """12-triangle axis-aligned box, outward orientation (not the reference's box.obj ordering)."""
    lo, hi = np.asarray(lo, float), np.asarray(hi, float)
    V = np.array([[x, y, z] for z in (lo[2], hi[2]) for y in (lo[1], hi[1]) for x in (lo[0], hi[0])])
    F = np.array([[0, 2, 1], [1, 2, 3], [4, 5, 6], [5, 7, 6], [0, 1, 4], [1, 5, 4], [2, 6, 3], [3, 6, 7],
                  [0, 4, 2], [2, 4, 6], [1, 3, 5], [3, 7, 5]], np.uint32)
    return TriangleMesh(V, F)
