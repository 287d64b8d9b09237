"""C-ABI and host-logic tests that need NO GPU: the shared library loads, exports every symbol include/*.h declares,
fails loudly without a device, and the host-side mirror (OBJ reader, .cdf I/O, index helpers, sharding) is correct."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, bits_equal


def test_library_exports_every_declared_symbol(dg):
    hdr = open(os.path.join(ROOT, "include", "discregrid_b200.h")).read()
    declared = sorted(set(re.findall(r"^DG_API[^;(]*?\b(dg_\w+)\s*\(", hdr, re.M)))
    assert len(declared) >= 29
    lib = C.CDLL(dg.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/discregrid_b200.h but not exported"
    from discregrid_b200 import _capi
    assert sorted(_capi.SIGNATURES) == declared          # the Python binding covers the whole header
    assert lib.dg_abi_version() == 1
    out = subprocess.run(["nm", "-D", "--defined-only", dg.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert exported == set(declared), "the library must export exactly the C-ABI (no stray symbols)"


def test_product_does_not_link_or_import_the_oracle(dg):
    out = subprocess.run(["ldd", dg.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "dgref" not in out
    for root, _dirs, files in os.walk(os.path.join(ROOT, "discregrid_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                src = open(os.path.join(root, f)).read()
                assert "oracle_api" not in src and "liboracle" not in src and "libdgref" not in src, f
                assert not re.search(r'#include\s+"[^"]*oracle', src), f


@pytest.mark.skipif(os.path.exists("/dev/nvidiactl"), reason="a GPU is present")
def test_no_cpu_fallback_without_device(dg, box_mesh):
    assert dg.device_count() == 0
    with pytest.raises(dg.DiscregridError) as e:
        dg.TriangleMeshDistance(box_mesh)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)
    g = dg.CubicLagrangeDiscreteGrid(os.path.join(GOLDEN, "box.cdf"))
    with pytest.raises(dg.DiscregridError):
        g.interpolate(0, np.zeros((4, 3)))
    with pytest.raises(dg.DiscregridError):
        g.nodePositions()


def test_grid_helpers_host(dg):
    from discregrid_b200 import _capi as capi
    n = C.c_uint64()
    for res, want in (((16, 16, 16), 32657), ((128, 128, 128), 14926977), ((256, 256, 256), 118425857), ((5, 5, 5), 1296)):
        capi.check(capi.lib.dg_grid_num_nodes((C.c_uint32 * 3)(*res), C.byref(n)))
        assert n.value == want                                  # node counts
    assert capi.lib.dg_grid_num_nodes((C.c_uint32 * 3)(0, 4, 4), C.byref(n)) == capi.DG_ERR_INVALID
    assert capi.lib.dg_grid_num_nodes((C.c_uint32 * 3)(2000, 2000, 2000), C.byref(n)) == capi.DG_ERR_INVALID
    assert b"grid too large" in capi.lib.dg_last_error()
    g = dg.CubicLagrangeDiscreteGrid([0, 0, 0], [1, 2, 3], (4, 5, 6))
    assert g.nCells() == 120 and g.singleToMultiIndex(g.multiToSingleIndex((3, 4, 5))) == (3, 4, 5)
    lo, hi = g.subdomain(g.multiToSingleIndex((1, 2, 3)))
    assert np.allclose(lo, [0.25, 0.8, 1.5]) and np.allclose(hi - lo, g.cellSize())
    with pytest.raises(TypeError):
        g.addFunction(lambda x: 0.0)                            # opaque callables are rejected, not run on the host


def test_generate_sdf_domain_matches_box_cdf(dg, box_mesh):
    gold = dg.CubicLagrangeDiscreteGrid(os.path.join(GOLDEN, "box.cdf"))
    mn, mx = dg.generate_sdf_domain(box_mesh.vertices)
    assert bits_equal(mn, gold.m_domain[0]) and bits_equal(mx, gold.m_domain[1])
    d = dg.grid_desc(mn, mx, (5, 5, 5))
    assert bits_equal(np.array(d.cell_size[:]), gold.m_cell_size) and bits_equal(np.array(d.inv_cell_size[:]), gold.m_inv_cell_size)


def test_cdf_roundtrip_is_byte_identical(dg, tmp_path):
    src = os.path.join(GOLDEN, "box.cdf")
    g = dg.CubicLagrangeDiscreteGrid(src)
    assert g.nFields() == 1 and g.nCells() == 125 and g.m_nodes[0].shape == (1296,) and g.m_cells[0].shape == (125, 32)
    out = tmp_path / "rt.cdf"
    g.save(str(out))
    assert out.read_bytes() == open(src, "rb").read()
    empty = dg.CubicLagrangeDiscreteGrid(str(tmp_path / "missing.cdf"))     # reference: message on stderr, object left empty
    assert empty.nFields() == 0


def test_obj_reader(dg, tmp_path):
    p = tmp_path / "m.obj"
    p.write_text("# c\nv 0 0 0\nv 1.5e0 0 0\nvn 0 0 1\nv 0 1 0\nvt 0 0\nf 1/1/1 2/2/2 3/3/3\nf 3 2 1\n")
    m = dg.TriangleMesh(str(p))
    assert m.nVertices() == 3 and m.nFaces() == 2
    assert np.array_equal(m.faces, [[0, 1, 2], [2, 1, 0]]) and m.vertices[1, 0] == 1.5
    m.exportOBJ(str(tmp_path / "o.obj"))
    m2 = dg.TriangleMesh(str(tmp_path / "o.obj"))
    assert bits_equal(m2.vertices, m.vertices) and np.array_equal(m2.faces, m.faces)
    box = dg.TriangleMesh(os.path.join(GOLDEN, "box.obj"))
    assert box.nVertices() == 8 and box.nFaces() == 12
    t = dg.bumpy_torus()
    assert t.nFaces() == 100000 and t.nVertices() == 50000     # BASELINE.md target mesh
    # closed + consistently oriented: every directed edge appears once, with its reverse
    e = np.concatenate([t.faces[:, [0, 1]], t.faces[:, [1, 2]], t.faces[:, [2, 0]]]).astype(np.int64)
    key = e[:, 0] * t.nVertices() + e[:, 1]
    assert len(np.unique(key)) == len(key) and np.array_equal(np.sort(key), np.sort(e[:, 1] * t.nVertices() + e[:, 0]))


def test_obj_reader_records_and_errors(dg, tmp_path):
    """dg_obj_read accepts exactly what the reference's stream parser does (src/mesh/triangle_mesh.cpp:90-124)"""
    from discregrid_b200 import _capi as capi
    p = tmp_path / "odd.obj"
    p.write_bytes(b"# comment\r\n"
                  b"v 1 2 3\r\n"                       # CRLF
                  b"v   +0.5\t-2.5e-1   1e2 0.75\n"    # blanks, tab, explicit +, a 4th value (ignored)
                  b"v\t9 9 9\n"                        # "v<TAB>" is not "v ": ignored
                  b"vn 0 0 1\nvt 0 1\ng grp\n"
                  b"v 0.1 0.2\n"                        # short record: the missing value stays 0 here (uninitialised in the reference)
                  b"v .5 -.25 1.\n"
                  b"f 1/7/9 2//3 3 4\n"                 # only the first three corners, index before the first '/'
                  b"f  4   1   2")                      # no trailing newline
    m = dg.TriangleMesh(str(p))
    assert np.array_equal(m.vertices, [[1, 2, 3], [0.5, -0.25, 100.0], [0.1, 0.2, 0.0], [0.5, -0.25, 1.0]])
    assert np.array_equal(m.faces, [[0, 1, 2], [3, 0, 1]])
    # correctly rounded like operator>>: 17-digit round trip of awkward values
    vals = np.array([0.1, 1 / 3, 2 ** -1074, 1.7976931348623157e308, 123456789.12345678, 5e-324, 0.30000000000000004], np.float64)
    q = tmp_path / "round.obj"
    q.write_text("".join(f"v {a!r} {-a!r} {a!r}\n" for a in vals.tolist()))
    got = dg.TriangleMesh(str(q)).vertices
    assert bits_equal(got[:, 0], vals) and bits_equal(got[:, 1], -vals)
    bad = tmp_path / "bad.obj"
    bad.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2\n")            # std::stoi("") throws in the reference
    with pytest.raises(dg.DiscregridError) as ei:
        dg.TriangleMesh(str(bad))
    assert ei.value.code == capi.DG_ERR_INVALID and "bad.obj:4" in str(ei.value)
    with pytest.raises(FileNotFoundError):
        dg.TriangleMesh(str(tmp_path / "missing.obj"))
    # malformed exponents: num_get swallows "1e" / "1.5e+" as part of the number and fails the extraction (value 0, the following
    # extractions fail too); a face index with two signs makes std::stoi throw
    mal = tmp_path / "mal.obj"
    mal.write_text("v 1e 2 3\nv 4 1.5e+ 6\nv 7 8 9\nf 1 2 3\n")
    assert np.array_equal(dg.TriangleMesh(str(mal)).vertices, [[0, 0, 0], [4, 0, 0], [7, 8, 9]])
    two = tmp_path / "two.obj"
    two.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf +-1 2 3\n")
    with pytest.raises(dg.DiscregridError):
        dg.TriangleMesh(str(two))


def test_obj_reader_parallel_chunks_and_reference_loader(dg, tmp_path):
    """a file large enough to be parsed by several threads (chunks cut at line starts) gives the same arrays as a line-by-line
    reader; the staged reference meshes equal the reference's own loader bit for bit where that was compiled (oracle/_ref)"""
    import ctypes as C
    from test_oracle_golden import read_obj
    t = dg.bumpy_torus(300, 301, 1.0, 0.4, 0.05, 7, 5)
    p = str(tmp_path / "big.obj")
    t.exportOBJ(p)
    assert os.path.getsize(p) > (1 << 21)
    m = dg.TriangleMesh(p)
    V, F = read_obj(p)
    assert bits_equal(m.vertices, V) and np.array_equal(m.faces, F) and bits_equal(m.vertices, t.vertices) and np.array_equal(m.faces, t.faces)
    from oracle_api import REF_GRID_SO
    from conftest import ref_resource
    if not os.path.exists(REF_GRID_SO) or ref_resource("bunny.obj") is None:
        return
    lib = C.CDLL(REF_GRID_SO)
    lib.refm_open.restype = C.c_void_p; lib.refm_open.argtypes = [C.c_char_p, C.POINTER(C.c_double)]
    lib.refm_sizes.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.refm_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]; lib.refm_close.argtypes = [C.c_void_p]
    for name in ("bunny.obj", "dragon.obj"):
        path = ref_resource(name)
        sec = C.c_double(); h = lib.refm_open(path.encode(), C.byref(sec))
        nv, nf = C.c_uint64(), C.c_uint64(); lib.refm_sizes(h, C.byref(nv), C.byref(nf))
        Vr = np.empty((nv.value, 3)); Fr = np.empty((nf.value, 3), np.uint32)
        lib.refm_copy(h, Vr.ctypes.data, Fr.ctypes.data); lib.refm_close(h)
        mine = dg.TriangleMesh(path)
        assert bits_equal(mine.vertices, Vr) and np.array_equal(mine.faces, Fr)


def test_node_sharding_partition():
    from discregrid_b200.distributed import make_sharding
    for n, world in ((14926977, 8), (32657, 2), (1296, 4), (118425857, 8), (77, 1)):
        sh = make_sharding(n, world)
        seen = np.zeros(n, np.int32)
        for r in range(world):
            for (j, b, e) in sh.chunks_of(r):
                assert 0 <= b <= e <= n and (e - b) <= sh.chunk and b % 1024 == 0 or b == n
                seen[b:e] += 1
                if e > b:
                    assert b == (j * world + r) * sh.chunk           # the all-gather row layout
        assert (seen == 1).all()
        assert sh.padded >= n and sh.padded % (world * sh.rows) == 0


def test_slab_sharding_partition(dg):
    """dg_slab_ranges: whole plane pairs of the four node arrays, every node exactly once, for any world size (host only)"""
    from discregrid_b200.distributed import SlabSharding
    for res, world in (((16, 16, 16), 2), ((128, 128, 128), 8), ((7, 3, 5), 3), ((5, 9, 2), 4), ((1, 1, 1), 2), ((33, 20, 11), 1)):
        desc = dg.grid_desc([0, 0, 0], [1, 1, 1], res)
        sh = SlabSharding(desc, world)
        assert sh.covers_exactly_once()
        nx, ny, nz = res
        nv = (nx + 1) * (ny + 1) * (nz + 1)
        plane = [(ny + 1) * (nx + 1), (ny + 1) * 2 * nx, (nz + 1) * 2 * ny, (nx + 1) * 2 * nz]
        base = [0, nv, nv + 2 * nx * (ny + 1) * (nz + 1), nv + 2 * nx * (ny + 1) * (nz + 1) + 2 * (nx + 1) * ny * (nz + 1)]
        for r in range(world):
            for a, (b, e) in enumerate(sh.ranges[r]):
                assert (b - base[a]) % plane[a] == 0 and (e - base[a]) % plane[a] == 0          # whole planes
                if r + 1 < world:
                    assert ((e - base[a]) // plane[a]) % 2 == 0                                  # interior boundaries on plane pairs


def test_exchanges_gloo_world2(tmp_path):
    """N > 1 host logic on CPU, two gloo ranks, one torchrun for all three shardings' exchange steps: every rank fills what it owns with a
    known function of the node id, runs the sampler's own collective call, and must end up with the whole array --
    node-id chunks (allgather_rows), slabs (allgather_slabs), interleaved plane pairs with 1 and 2 parts per rank (allgather_slots +
    the host statement of the slot layout)."""
    script = tmp_path / "w.py"
    script.write_text(f'''
import os, sys, ctypes as C, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, {ROOT!r})
import discregrid_b200 as dg
from discregrid_b200 import _capi as capi
from discregrid_b200.distributed import make_sharding, allgather_rows, SlabSharding, allgather_slabs, allgather_slots, interleaved_node_slots
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()

# ---- node-id chunks
n = 100003
sh = make_sharding(n, world, rows=3, align=64)
full = torch.full((sh.padded,), -1.0, dtype=torch.float64)
for (j, b, e) in sh.chunks_of(rank):
    full[b:e] = torch.arange(b, e, dtype=torch.float64) * 0.5 + 1.0      # "node value" = f(node id)
allgather_rows(full, sh)
want = torch.arange(n, dtype=torch.float64) * 0.5 + 1.0
assert torch.equal(full[:n], want), ("rows", rank, (full[:n] != want).nonzero()[:5])

# ---- slabs
desc = dg.grid_desc([0, 0, 0], [1, 2, 3], (9, 5, 7))
ss = SlabSharding(desc, world)
full = torch.full((ss.padded,), -1.0, dtype=torch.float64)
for (b, e) in ss.ranges[rank]:
    full[b:e] = torch.arange(b, e, dtype=torch.float64) * 0.25 - 3.0
allgather_slabs(full, ss)
want = torch.arange(ss.n, dtype=torch.float64) * 0.25 - 3.0
assert torch.equal(full[:ss.n], want), ("slabs", rank, (full[:ss.n] != want).nonzero()[:5])

# ---- interleaved plane pairs, `splits` consecutive parts per rank
nn = C.c_uint64(); capi.check(capi.lib.dg_grid_num_nodes(desc.resolution, C.byref(nn))); nn = nn.value
f = np.arange(nn, dtype=np.float64) * 0.25 - 3.0
for splits in (1, 2):
    parts = world * splits
    se = C.c_uint64(); capi.check(capi.lib.dg_interleaved_slot_elems(C.byref(desc), parts, C.byref(se))); se = se.value
    part, pos = interleaved_node_slots(desc, parts, 0, nn)
    slots = torch.full((parts * se,), -1.0, dtype=torch.float64)
    mine = (part // splits) == rank                          # rank r owns parts r*splits .. r*splits + splits - 1
    slots[torch.from_numpy(part[mine].astype(np.int64) * se + pos[mine].astype(np.int64))] = torch.from_numpy(f[mine])
    allgather_slots(slots, se * splits, rank, world)
    got = slots[torch.from_numpy(part.astype(np.int64) * se + pos.astype(np.int64))].numpy()
    assert np.array_equal(got, f), ("slots", splits, rank, np.nonzero(got != f)[0][:5])
dist.barrier()
if rank == 0: print("EXCHANGES_OK")
''')
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29615")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29615", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and "EXCHANGES_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
