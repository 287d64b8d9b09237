"""K3 parity (GPU): GenerateDensityMap node function vs the CPU oracle on a small grid.  Bar: bit-exact."""
import numpy as np
import pytest

from conftest import bits_equal

pytestmark = pytest.mark.gpu
DBL_MAX = np.finfo(np.float64).max


@pytest.mark.parametrize("no_reduction,h", [(False, 0.03), (True, 0.12)])
def test_density_map_vs_oracle(dg, orc, no_reduction, h):
    s = dg.uv_sphere(10, 16, 0.5)
    mn, mx = dg.generate_sdf_domain(s.vertices)
    g = dg.CubicLagrangeDiscreteGrid(mn, mx, (6, 6, 6))
    g.addFunction(dg.MeshSignedDistance(dg.TriangleMeshDistance(s)))
    fid = g.addFunction(dg.DensityMapFunction(0, h, 1000.0, no_reduction))
    assert fid == 1
    gd, r = orc.grid_desc(g.m_domain[0], g.m_domain[1], g.m_resolution, g.m_cell_size, g.m_inv_cell_size)
    want = orc.density_map(gd, r, g.m_nodes[0], h, 1000.0, no_reduction, 0, g.nNodes())
    got = g.m_nodes[1]
    assert bits_equal(got, want)
    assert (got == 0.0).any() and ((got > 0) & (got < DBL_MAX)).any()
    if not no_reduction:
        assert (got == DBL_MAX).any()
