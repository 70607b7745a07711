"""OBJ ingestion (crt_obj_load): the stand-in for the reference's assimp import (RayTracer/AssetManager.cpp:67-190)."""
import numpy as np
import pytest

from computational_ray_tracer_b200 import _capi, api

CUBE = """# unit cube, quads, no normals
o cube
v 0 0 0
v 1 0 0
v 1 1 0
v 0 1 0
v 0 0 1
v 1 0 1
v 1 1 1
v 0 1 1
f 1 4 3 2
f 5 6 7 8
f 1 2 6 5
f 2 3 7 6
f 3 4 8 7
f 4 1 5 8
g tri_with_normals
vn 0 0 -1
usemtl red
f -8//1 -5//1 -6//1
f 1/1/1 4/2/1 3/3/1 2/4/1 1/1/1
"""


def test_load_obj_triangulates_and_generates_normals(tmp_path, crt_lib):
    p = tmp_path / "cube.obj"
    p.write_text(CUBE)
    meshes = api.load_obj(p)
    assert [m["name"] for m in meshes] == ["cube", "tri_with_normals"]
    cube, extra = meshes
    assert cube["indices"].shape == (12, 3) and cube["positions"].shape == (36, 3)          # 6 quads -> 12 triangles, no vertex sharing
    assert np.array_equal(cube["indices"].reshape(-1), np.arange(36))
    # flat normals: unit length, perpendicular to the face, equal on the three corners
    P = cube["positions"].reshape(12, 3, 3); N = cube["normals"].reshape(12, 3, 3)
    g = np.cross(P[:, 1] - P[:, 0], P[:, 2] - P[:, 0]); g /= np.linalg.norm(g, axis=1, keepdims=True)
    assert np.allclose(N[:, 0], g, atol=1e-6) and np.allclose(N[:, 0], N[:, 1]) and np.allclose(N[:, 0], N[:, 2])
    # outward orientation of this cube's winding: every normal points away from the centre
    c = P.mean(1) - 0.5
    assert (np.einsum("ij,ij->i", c, N[:, 0]) > 0).all()
    # second mesh: negative indices, v//vn and v/vt/vn forms, pentagon fan -> 1 + 3 triangles, normals from the file
    assert extra["indices"].shape == (4, 3)
    assert np.allclose(extra["normals"], [0, 0, -1])
    assert np.allclose(extra["positions"][:3], [[0, 0, 0], [0, 1, 0], [1, 1, 0]])
    # and it feeds the octree builder
    oc = api.Octtree_Model(api.MeshSet(meshes))
    assert oc.stats()["refs"] == 16
    oc.close()


TEXTURED = """o quad
v 0 0 0
v 2 0 0
v 2 1 0
v 0 1 0
vt 0 0
vt 1 0
vt 1 1
vt 0 1
vn 0 0 1
f 1/1/1 2/2/1 3/3/1 4/4/1
o mirrored
f 1/2/1 2/1/1 3/4/1
o no_uv
f 1//1 2//1 3//1
"""


def test_load_obj_keeps_texcoords_and_builds_the_tangent_space(tmp_path, crt_lib):
    """What the reference's import flags add to MeshCache::Mesh (AssetManager.cpp:104-190): texcoords, CalcTangentSpace tangents, and the
    tangent stored a second time where the bitangent belongs (:153)."""
    p = tmp_path / "quad.obj"
    p.write_text(TEXTURED)
    quad, mirrored, no_uv = api.load_obj(p)
    assert np.array_equal(quad["texcoords"], np.float32([[0, 0], [1, 0], [1, 1], [0, 0], [1, 1], [0, 1]]))
    # u grows along +x on this quad: tangent = +x on every corner, unit length, orthogonal to the normal
    assert np.allclose(quad["tangents"], [1, 0, 0], atol=1e-6)
    assert np.array_equal(quad["bitangents"], quad["tangents"])            # the reference's bug, reproduced
    # mirrored uv (u decreases along +x): the tangent follows du, so it flips
    assert np.allclose(mirrored["tangents"], [-1, 0, 0], atol=1e-6)
    assert "texcoords" not in no_uv and "tangents" not in no_uv             # vertex_available flags stay off without vt
    # the attributes reach the device model: MeshSet carries them through crt_mesh_desc
    ms = api.MeshSet([quad])
    assert bool(ms.descs[0].texcoords) and bool(ms.descs[0].tangents) and bool(ms.descs[0].bitangents)
    assert not bool(api.MeshSet([no_uv]).descs[0].texcoords)


def test_load_obj_errors(tmp_path, crt_lib):
    with pytest.raises(_capi.CrtError, match="cannot open"):
        api.load_obj(tmp_path / "missing.obj")
    bad = tmp_path / "bad.obj"
    bad.write_text("v 0 0 0\nv 1 0 0\nf 1 2 9\n")
    with pytest.raises(_capi.CrtError, match="bad face"):
        api.load_obj(bad)
    empty = tmp_path / "empty.obj"
    empty.write_text("v 0 0 0\n")
    with pytest.raises(_capi.CrtError, match="no faces"):
        api.load_obj(empty)
