"""Regenerates the committed golden fixtures.  Run in the build container (needs /root/reference and
oracle/_ref/libdgref.so = the reference's own unmodified TriangleMeshDistance.h, built by `make -C oracle ref`):

    python tests/golden/make_golden.py

Outputs (all small, committed):
  box.obj, box.cdf           verbatim copies of the reference's only golden vector
                             (cmd/generate_sdf/resources/box.{obj,cdf}; box.cdf = GenerateSDF -r "5 5 5" box.obj)
  ref_torus_queries.npz      reference-header results (distance, nearest point, entity, triangle) for 4000 seeded
                             points around a deterministic 4,608-triangle bumpy torus (discregrid_b200.mesh.bumpy_torus)
  ref_torus_tree.npz         the reference's tree (children + internal spheres) and pseudonormals for that torus
  ref_sphere_surface.npz     reference results for points ON / very near a UV sphere's surface (ties, sign near 0)
"""
import os, shutil, sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle_api import RefMesh, Oracle          # noqa: E402
from discregrid_b200.mesh import bumpy_torus, uv_sphere   # noqa: E402 (pure numpy part of the package)

REF_RES = "/root/reference/cmd/generate_sdf/resources"
for f in ("box.obj", "box.cdf"):
    shutil.copyfile(os.path.join(REF_RES, f), os.path.join(HERE, f))
    os.chmod(os.path.join(HERE, f), 0o644)

orc = Oracle()
torus = bumpy_torus(48, 48, 1.0, 0.4, 0.05, 7, 5)            # 4608 triangles
ref = RefMesh(torus.vertices, torus.faces)
mn, mx = orc.generate_sdf_domain(torus.vertices)
rng = np.random.default_rng(20260924)
x = mn + rng.random((4000, 3)) * (mx - mn)
d, near, ent, tri = ref.distance(x, signed=True)
du = ref.distance(x, signed=False)[0]
np.savez_compressed(os.path.join(HERE, "ref_torus_queries.npz"), x=x, distance=d, unsigned=du, nearest=near,
                    entity=ent, triangle=tri, torus_args=np.array([48, 48, 1.0, 0.4, 0.05, 7, 5]))
sph, kids = ref.tree()
pt, pe, pv = ref.pseudonormals()
internal = kids[:, 0] != -1
np.savez_compressed(os.path.join(HERE, "ref_torus_tree.npz"), kids=kids, spheres_internal=sph[internal], pn_tri=pt, pn_edge=pe, pn_vert=pv)

s = uv_sphere(12, 24, 0.75, (0.1, -0.2, 0.05))
refs = RefMesh(s.vertices, s.faces)
V, F = s.vertices, s.faces
pts = [V.copy()]                                             # exactly on vertices
pts.append(0.5 * (V[F[:, 0]] + V[F[:, 1]]))                  # on edge midpoints
pts.append((V[F[:, 0]] + V[F[:, 1]] + V[F[:, 2]]) / 3.0)     # face centroids
c = (V[F[:, 0]] + V[F[:, 1]] + V[F[:, 2]]) / 3.0
pts.append(c * (1 + 1e-9)); pts.append(c * (1 - 1e-9)); pts.append(V * (1 + 1e-12))
xs = np.concatenate(pts, 0)
d, near, ent, tri = refs.distance(xs, signed=True)
np.savez_compressed(os.path.join(HERE, "ref_sphere_surface.npz"), x=xs, distance=d, nearest=near, entity=ent, triangle=tri,
                    sphere_args=np.array([12, 24, 0.75, 0.1, -0.2, 0.05]))
print("golden fixtures written to", HERE)
