"""Invariants of the flattened device layout that the ordered traversal's exactness argument relies on,
checked on the host: the side structures may only ever SKIP triangles a ray cannot hit, so
  * every node's subtree bounds contain every triangle stored beneath it (with the pad),
  * the sub-packets of a leaf (<= 4 triangles each) partition exactly that leaf's references and each box contains its triangles;
    the super-packets of a fat leaf partition its sub-packets and contain their boxes,
  * the non-empty-children mask of an internal node is exact,
  * leaf reference lists are in ascending global-id order (the reference's insertion order)."""
import numpy as np
import pytest

from computational_ray_tracer_b200 import api, scenes

LEAF, PACKETS, SUBPK, COUNT = 0x80000000, 0x40000000, 0x20000000, 0x1FFFFFFF
SUBPACKET, SUPERPACKET = 4, 32                     # crt_host.h

SCENES = {
    "heightfield_fat_leaves": lambda: scenes.heightfield(200),
    "soup": lambda: scenes.random_soup(4000, seed=3),
    "many_lights": lambda: scenes.many_light_scene(96, 200),
}


def _world_triangles(meshes):
    return np.concatenate([m["positions"][m["indices"].reshape(-1)].reshape(-1, 3, 3) for m in meshes]).astype(np.float32)


@pytest.mark.parametrize("name", sorted(SCENES))
def test_flat_layout_invariants(crt_lib, name):
    meshes = SCENES[name]()
    tris = _world_triangles(meshes)                                   # (T, 3, 3), global id order = mesh-major
    oc = api.Octtree_Model(api.MeshSet(meshes))
    f = oc.flat()
    nodes = f["nodes"].reshape(-1, 8)
    a = nodes[:, 3].copy().view(np.uint32); b = nodes[:, 7].copy().view(np.uint32)
    tight = f["node_tight"].reshape(-1, 8)
    refs = f["leaf_refs"]
    n = len(nodes)
    is_leaf = (b & LEAF) != 0
    count = np.where(is_leaf, b & COUNT, 0).astype(np.int64)
    # subtree triangle bounds, recomputed bottom-up from the leaves' reference lists
    lo = np.full((n, 3), np.inf); hi = np.full((n, 3), -np.inf)
    fat = 0
    for i in range(n - 1, -1, -1):
        if is_leaf[i]:
            if count[i] == 0:
                continue
            ids = refs[a[i]:a[i] + count[i]]
            assert (np.diff(ids.astype(np.int64)) > 0).all(), "leaf list must be in ascending global-id order"
            t = tris[ids].reshape(-1, 3)
            lo[i] = t.min(0); hi[i] = t.max(0)
            n_sub = (count[i] + SUBPACKET - 1) // SUBPACKET
            is_fat = n_sub > SUPERPACKET
            assert bool(b[i] & PACKETS) == is_fat and bool(b[i] & SUBPK) == (not is_fat)
            fat += is_fat
            all_boxes = f["pk_boxes"].reshape(-1, 8)

            def box_fields(bx):
                return int(bx[3:4].copy().view(np.uint32)[0]), int(bx[7:8].copy().view(np.uint32)[0])
            h0, hn = int(refs[a[i] - 2]), int(refs[a[i] - 1])
            if is_fat:                       # header -> super-packets, each a run of consecutive sub-packet boxes it must contain
                assert hn == (n_sub + SUPERPACKET - 1) // SUPERPACKET
                sub_ids = []
                for sb in all_boxes[h0:h0 + hn]:
                    first, cnt = box_fields(sb)
                    assert 0 < cnt <= SUPERPACKET
                    inner = all_boxes[first:first + cnt]
                    assert (inner[:, :3] >= sb[:3]).all() and (inner[:, 4:7] <= sb[4:7]).all(), "super-packet box must contain its sub-packet boxes"
                    sub_ids.extend(range(first, first + cnt))
                assert sub_ids == list(range(sub_ids[0], sub_ids[0] + n_sub)), "super-packets must partition the leaf's sub-packets"
                boxes = all_boxes[sub_ids[0]:sub_ids[0] + n_sub]
            else:
                assert hn == n_sub
                boxes = all_boxes[h0:h0 + hn]
            got = []
            for bx in boxes:
                first, cnt = box_fields(bx)
                assert 0 < cnt <= SUBPACKET
                ids_p = f["pk_refs"][first:first + cnt]
                tp = tris[ids_p].reshape(-1, 3)
                assert (tp >= bx[:3]).all() and (tp <= bx[4:7]).all(), "packet box must contain its triangles"
                got.append(ids_p)
            assert np.array_equal(np.sort(np.concatenate(got)), ids), "sub-packets must partition the leaf's references"
        else:
            kids = np.arange(a[i], a[i] + 8)
            lo[i] = lo[kids].min(0); hi[i] = hi[kids].max(0)
            mask = sum(1 << k for k, c in enumerate(kids) if (count[c] > 0 if is_leaf[c] else (b[c] & 0xFF) != 0))
            assert (b[i] & 0xFF) == mask and (b[i] >> 8) == 0, "non-empty-children mask"
    has = np.isfinite(lo[:, 0])
    assert (tight[has, :3] <= lo[has]).all() and (tight[has, 4:7] >= hi[has]).all(), "subtree bounds must contain every triangle beneath the node"
    pad = np.maximum(tight[has, 4:7] - hi[has], lo[has] - tight[has, :3])
    assert (pad > 0).all() and pad.max() < 1.0                        # padded, but not by much (2^-12 of the coordinate magnitude)
    assert (tight[~has, 0] > tight[~has, 4]).all()                    # empty subtrees: inverted box, never hit
    if name == "heightfield_fat_leaves":
        assert fat > 0
    oc.close()
