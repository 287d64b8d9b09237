"""run by make_golden.py (or alone): ref_aniso_field.cdf -> reference reduceField(0, 0.3 <= v <= 0.9) -> ref_aniso_reduced.cdf"""
import os, struct, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE)))
from oracle_api import Oracle, RefGrid


def write_cdf(path, mn, mx, res, cell, inv, nodes, cells, cmap):
    """one-field file in the reference's layout (cubic_lagrange_discrete_grid.cpp:678-719)"""
    with open(path, "wb") as f:
        f.write(struct.pack("<3d", *mn)); f.write(struct.pack("<3d", *mx)); f.write(struct.pack("<3I", *[int(r) for r in res]))
        f.write(struct.pack("<3d", *cell)); f.write(struct.pack("<3d", *inv)); f.write(struct.pack("<QQ", int(np.prod(res)), 1))
        for arr in (nodes, cells, cmap):
            f.write(struct.pack("<Q", 1)); f.write(struct.pack("<Q", len(arr))); f.write(np.ascontiguousarray(arr).tobytes())


def synthetic_field(orc, mn, mx, res, seed):
    """a blob-shaped field (distance to an off-centre point + a little noise) on the full grid, with the closed-form connectivity"""
    gd, r = orc.grid_desc(mn, mx, res)
    cells = orc.build_cells(r)
    n = int(cells.max()) + 1
    x = orc.node_positions(gd, r, 0, n)
    rng = np.random.default_rng(seed)
    v = np.linalg.norm(x - (np.asarray(mn, float) + np.asarray(mx, float)) / 2 - 0.1, axis=1) + 0.01 * rng.standard_normal(n)
    return gd, r, v, cells


if __name__ == "__main__":
    _orc = Oracle()
    mn_, mx_, res_ = [0.0, 0.0, 0.0], [1.0, 2.0, 4.0], (6, 6, 6)
    gd_, r_, v_, cells_ = synthetic_field(_orc, mn_, mx_, res_, 1)
    src = os.path.join(HERE, "ref_aniso_field.cdf")
    write_cdf(src, mn_, mx_, res_, gd_[6:9], gd_[9:12], v_, cells_, np.arange(len(cells_), dtype=np.uint32))
    g_ = RefGrid(src); g_.reduce_window(0, 0.3, 0.9); g_.save(os.path.join(HERE, "ref_aniso_reduced.cdf")); g_.close()
    print("reduce fixtures written")
