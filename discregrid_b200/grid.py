"""Host mirror of Discregrid::DiscreteGrid / CubicLagrangeDiscreteGrid over the C-ABI.

Keeps the reference's members (m_domain, m_resolution, m_cell_size, m_inv_cell_size, m_n_cells, m_n_fields,
m_nodes, m_cells, m_cell_map -- discrete_grid.hpp:93-98, cubic_lagrange_discrete_grid.hpp:69-71) as numpy arrays so
that save()/load() stay byte-compatible with the reference's .cdf/.cdm files
(cubic_lagrange_discrete_grid.cpp:678-778), and routes the hot path to the GPU:

  addFunction(MeshSignedDistance)        -> dg_sample_sdf        (K1, csrc/k1_sdf.cu)
  addFunction(DensityMapFunction, pred)  -> dg_density_map       (K3, csrc/k3_density.cu)
  interpolate(field_id, x[, gradient])   -> dg_interpolate_batch (K2, csrc/k2_interp.cu), batched over x

The reference's addFunction takes an opaque std::function; only the two functor types the reference's own tools
build are recognised.  Any other callable raises TypeError -- sample it yourself at nodePositions() and hand the
values to addSampledFunction(); there is no hidden host loop.
"""
import ctypes as C
import struct

import numpy as np

from . import _capi as capi
from .distance import MeshSignedDistance


class DensityMapFunction:
    """The (density_func, predicate) pair GenerateDensityMap passes to addFunction
    (cmd/generate_density_map/main.cpp:96-133): rho0 * int gamma(x+xi) W(xi) dxi over [-h,h]^3 with 16^3 Gauss points."""

    def __init__(self, sdf_field_id=0, smoothing_length=0.1, rest_density=1000.0, no_reduction=False):
        self.field_id, self.h, self.rho0, self.no_reduction = sdf_field_id, float(smoothing_length), float(rest_density), bool(no_reduction)


def grid_desc(domain_min, domain_max, resolution, cell_size=None, inv_cell_size=None):
    """dg_grid_desc; cell sizes computed like the DiscreteGrid ctor (discrete_grid.hpp:22-29) unless given verbatim."""
    d = capi.GridDesc()
    mn = np.ascontiguousarray(domain_min, np.float64); mx = np.ascontiguousarray(domain_max, np.float64)
    res = np.ascontiguousarray(resolution, np.uint32)
    capi.check(capi.lib.dg_grid_init(capi.ptr(mn, capi.F64P), capi.ptr(mx, capi.F64P), capi.ptr(res, capi.U32P), C.byref(d)))
    if cell_size is not None:
        for k in range(3):
            d.cell_size[k] = float(cell_size[k]); d.inv_cell_size[k] = float(inv_cell_size[k])
    return d


def generate_sdf_domain(vertices):
    """GenerateSDF's padded bounding box (cmd/generate_sdf/main.cpp:83-91)."""
    V = np.ascontiguousarray(vertices, np.float64).reshape(-1, 3)
    mn = np.empty(3); mx = np.empty(3)
    capi.check(capi.lib.dg_generate_sdf_domain(capi.ptr(V, capi.F64P), len(V), capi.ptr(mn, capi.F64P), capi.ptr(mx, capi.F64P)))
    return mn, mx


class CubicLagrangeDiscreteGrid:
    def __init__(self, domain_or_filename=None, domain_max=None, resolution=None):
        """CubicLagrangeDiscreteGrid(filename) | CubicLagrangeDiscreteGrid(domain_min, domain_max, resolution)
        (cubic_lagrange_discrete_grid.hpp:12-14)."""
        self.m_nodes, self.m_cells, self.m_cell_map = [], [], []
        self.m_n_fields = 0
        self._fields = {}          # field_id -> device handle (cache; invalidated by addFunction/load/reduceField)
        self._desc = None
        if isinstance(domain_or_filename, str):
            self.load(domain_or_filename)
        elif domain_or_filename is not None:
            self._desc = grid_desc(domain_or_filename, domain_max, resolution)
            self._sync_members()

    # ------------------------------------------------------------------ members / accessors (discrete_grid.hpp:85-98)
    def _sync_members(self):
        d = self._desc
        self.m_domain = (np.array(d.domain_min[:]), np.array(d.domain_max[:]))
        self.m_resolution = np.array(d.resolution[:], np.uint32)
        self.m_cell_size = np.array(d.cell_size[:])
        self.m_inv_cell_size = np.array(d.inv_cell_size[:])
        self.m_n_cells = int(np.prod(self.m_resolution.astype(np.uint64)))

    def domain(self):
        return self.m_domain

    def resolution(self):
        return self.m_resolution

    def cellSize(self):
        return self.m_cell_size

    def invCellSize(self):
        return self.m_inv_cell_size

    def nCells(self):
        return self.m_n_cells

    def nFields(self):
        return self.m_n_fields

    def nNodes(self):
        n = C.c_uint64()
        capi.check(capi.lib.dg_grid_num_nodes(self._desc.resolution, C.byref(n)))
        return n.value

    # discrete_grid.cpp:8-38
    def singleToMultiIndex(self, l):
        n0, n1 = int(self.m_resolution[0]), int(self.m_resolution[1])
        n01 = n0 * n1
        k, t = divmod(int(l), n01)
        j, i = divmod(t, n0)
        return (i, j, k)

    def multiToSingleIndex(self, ijk):
        n0, n1 = int(self.m_resolution[0]), int(self.m_resolution[1])
        return n1 * n0 * int(ijk[2]) + n0 * int(ijk[1]) + int(ijk[0])

    def subdomain(self, l_or_ijk):
        ijk = self.singleToMultiIndex(l_or_ijk) if np.isscalar(l_or_ijk) else l_or_ijk
        origin = self.m_domain[0] + np.array(ijk, np.float64) * self.m_cell_size
        return origin, origin + self.m_cell_size

    def nodePositions(self, l_begin=0, l_end=None):
        """indexToNodePosition for a node range (cubic_lagrange_discrete_grid.cpp:604-665), computed on the GPU."""
        l_end = self.nNodes() if l_end is None else l_end
        x = np.empty((l_end - l_begin, 3))
        capi.check(capi.lib.dg_node_positions(C.byref(self._desc), l_begin, l_end, capi.ptr(x, capi.F64P)))
        return x

    # ------------------------------------------------------------------ addFunction (cubic_lagrange_discrete_grid.cpp:780-899)
    def addFunction(self, func, verbose=False, pred=None):
        if isinstance(func, MeshSignedDistance):
            if pred is not None:
                raise TypeError("a sample predicate is only supported with DensityMapFunction")
            # the whole addFunction in one call: node loop on the GPU, connectivity + cell map written by the library's host threads meanwhile
            coeffs = np.empty(self.nNodes())
            cells = np.empty((self.m_n_cells, 32), np.uint32)
            cmap = np.empty(self.m_n_cells, np.uint32)
            self.last_add_function_ms = np.zeros(6)
            capi.check(capi.lib.dg_add_function_sdf(func.md.handle, C.byref(self._desc), func.sign, capi.ptr(coeffs, capi.F64P),
                                                    capi.ptr(cells, capi.U32P), capi.ptr(cmap, capi.U32P), capi.ptr(self.last_add_function_ms, capi.F64P)))
            self.m_nodes.append(coeffs); self.m_cells.append(cells); self.m_cell_map.append(cmap)
            self.m_n_fields += 1
            return self.m_n_fields - 1
        elif isinstance(func, DensityMapFunction):
            coeffs = np.empty(self.nNodes())
            capi.check(capi.lib.dg_density_map(self._device_field(func.field_id), func.h, func.rho0, int(func.no_reduction), 0,
                                               len(coeffs), capi.ptr(coeffs, capi.F64P)))
        else:
            raise TypeError("addFunction runs on the GPU and only accepts MeshSignedDistance or DensityMapFunction; "
                            "for any other function evaluate it at nodePositions() and call addSampledFunction(values)")
        return self.addSampledFunction(coeffs)

    def addSampledFunction(self, coeffs):
        """Appends a field from node values (what addFunction's loop :806-831 produces), then builds the cell
        table (:833-886, on the GPU) and the identity cell map (:888-891).  Returns the field id (:898)."""
        coeffs = np.ascontiguousarray(coeffs, np.float64)
        if coeffs.shape != (self.nNodes(),):
            raise ValueError(f"expected {self.nNodes()} node values")
        cells = np.empty((self.m_n_cells, 32), np.uint32)
        capi.check(capi.lib.dg_build_cells(self._desc.resolution, 0, self.m_n_cells, capi.ptr(cells, capi.U32P)))
        self.m_nodes.append(coeffs)
        self.m_cells.append(cells)
        self.m_cell_map.append(np.arange(self.m_n_cells, dtype=np.uint32))
        self.m_n_fields += 1
        return self.m_n_fields - 1

    # ------------------------------------------------------------------ reduceField (cubic_lagrange_discrete_grid.cpp:1065-1174)
    def reduceField(self, field_id, pred):
        """Sparsifies a field.  pred(x[n,3], values[n]) -> bool[n] is the reference's Predicate, vectorised; nodes, cells and the
        cell map end up exactly as the reference leaves them (dg_reduce_field)."""
        nodes = np.ascontiguousarray(self.m_nodes[field_id], np.float64).copy()
        keep = np.asarray(pred(self.nodePositions(0, len(nodes)), nodes), bool) & (nodes != np.finfo(np.float64).max)   # :1073
        keep = np.ascontiguousarray(keep, np.uint8)
        cells = np.ascontiguousarray(self.m_cells[field_id], np.uint32).copy()
        cmap = np.empty(self.m_n_cells, np.uint32)
        n_nodes, n_cells = C.c_uint64(), C.c_uint64()
        capi.check(capi.lib.dg_reduce_field(C.byref(self._desc), capi.ptr(nodes, capi.F64P), len(nodes), keep.ctypes.data_as(C.POINTER(C.c_uint8)),
                                            capi.ptr(cells, capi.U32P), len(cells), capi.ptr(cmap, capi.U32P), 0, C.byref(n_nodes), C.byref(n_cells), None))
        self._invalidate()
        self.m_nodes[field_id], self.m_cells[field_id], self.m_cell_map[field_id] = nodes[:n_nodes.value].copy(), cells[:n_cells.value].copy(), cmap

    # ------------------------------------------------------------------ device field cache
    def _device_field(self, field_id):
        if field_id not in self._fields:
            nodes, cells, cmap = self.m_nodes[field_id], self.m_cells[field_id], self.m_cell_map[field_id]
            h = C.c_void_p()
            capi.check(capi.lib.dg_field_create(C.byref(self._desc), capi.ptr(nodes, capi.F64P), len(nodes),
                                                capi.ptr(cells, capi.U32P), len(cells), capi.ptr(cmap, capi.U32P), C.byref(h)))
            self._fields[field_id] = h
        return self._fields[field_id]

    def _invalidate(self):
        for h in self._fields.values():
            capi.lib.dg_field_destroy(h)
        self._fields = {}

    def __del__(self):
        try:
            self._invalidate()
        except Exception:
            pass

    # ------------------------------------------------------------------ interpolate (cubic_lagrange_discrete_grid.cpp:977-1063)
    def interpolate(self, field_id_or_x, x=None, gradient=False):
        """interpolate(x) == interpolate(0, x) (discrete_grid.hpp:38-41).  x: (n, 3).  Returns phi[n], or
        (phi[n], grad[n,3]) when gradient=True.  Out-of-domain / removed cell / missing coefficient -> DBL_MAX, grad 0."""
        if x is None:
            field_id, x = 0, field_id_or_x
        else:
            field_id = field_id_or_x
        x = np.ascontiguousarray(x, np.float64).reshape(-1, 3)
        phi = np.empty(len(x))
        grad = np.empty((len(x), 3)) if gradient else None
        capi.check(capi.lib.dg_interpolate_batch(self._device_field(field_id), capi.ptr(x, capi.F64P), len(x),
                                                 capi.ptr(phi, capi.F64P), capi.ptr(grad, capi.F64P)))
        return (phi, grad) if gradient else phi

    # ------------------------------------------------------------------ save / load (cubic_lagrange_discrete_grid.cpp:678-778)
    def save(self, filename):
        d = self._desc
        with open(filename, "wb") as fh:
            fh.write(struct.pack("<6d", *d.domain_min[:], *d.domain_max[:]))        # AlignedBox3d: min, max
            fh.write(struct.pack("<3I", *d.resolution[:]))
            fh.write(struct.pack("<3d", *d.cell_size[:]))
            fh.write(struct.pack("<3d", *d.inv_cell_size[:]))
            fh.write(struct.pack("<QQ", self.m_n_cells, self.m_n_fields))
            for arrs in (self.m_nodes, self.m_cells, self.m_cell_map):
                fh.write(struct.pack("<Q", len(arrs)))
                for a in arrs:
                    fh.write(struct.pack("<Q", len(a)))
                    fh.write(np.ascontiguousarray(a).tobytes())

    def load(self, filename):
        self._invalidate()
        try:
            fh = open(filename, "rb")
        except OSError:
            # reference: message on std::cerr, object left empty (:725-729)
            import sys
            print("ERROR: Discrete grid can not be loaded. Input file does not exist!", file=sys.stderr)
            return
        with fh:
            head = fh.read(48 + 12 + 24 + 24 + 16)
            mn_mx = struct.unpack_from("<6d", head, 0)
            res = struct.unpack_from("<3I", head, 48)
            cell = struct.unpack_from("<3d", head, 60)
            inv = struct.unpack_from("<3d", head, 84)
            n_cells, n_fields = struct.unpack_from("<QQ", head, 108)
            self._desc = grid_desc(mn_mx[:3], mn_mx[3:], res, cell, inv)     # cell sizes verbatim from the file
            self._sync_members()
            self.m_n_cells, self.m_n_fields = n_cells, n_fields

            def read_nested(dtype, width):
                (outer,) = struct.unpack("<Q", fh.read(8))
                out = []
                for _ in range(outer):
                    (n,) = struct.unpack("<Q", fh.read(8))
                    a = np.frombuffer(fh.read(n * width * np.dtype(dtype).itemsize), dtype=dtype).copy()
                    out.append(a.reshape(n, width) if width > 1 else a)
                return out

            self.m_nodes = read_nested(np.float64, 1)
            self.m_cells = read_nested(np.uint32, 32)
            self.m_cell_map = read_nested(np.uint32, 1)
