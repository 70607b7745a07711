"""Shared helpers for the parity tests: build the same scene in the oracle and in the CUDA library."""
import numpy as np

import oracle_lib as O
from computational_ray_tracer_b200 import api, scenes

FLT_MAX = float(np.finfo(np.float32).max)


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a


def camera_1080p_like(width, height, fov=45.0, pos=(0, 0, 0), look=(0, 0, 1), up=(0, 1, 0)):
    return api.camera_matrices(0, 1.0, 1000.0, fov, pos, look, up, width, height)


def pixel_center_rays(width, height, r2c, c2w, step=1):
    """Primary rays through pixel centres, generated on the host with IEEE ops only (no transcendentals): the hit-ID
    parity input the north-star asks for."""
    R = np.asarray(r2c, np.float32).reshape(4, 4).T
    Cw = np.asarray(c2w, np.float32).reshape(4, 4).T
    ys, xs = np.meshgrid(np.arange(1, height + 1, step, dtype=np.float32), np.arange(0, width, step, dtype=np.float32), indexing="ij")
    px = (xs + np.float32(0.5)).reshape(-1); py = (ys + np.float32(0.5)).reshape(-1)
    one = np.ones_like(px); zero = np.zeros_like(px)
    P = np.stack([px, py, zero, one], 0).astype(np.float32)
    c = (R.astype(np.float64) @ P.astype(np.float64)).astype(np.float32)
    d = (c[:3] / c[3:4]).T
    d = d / np.linalg.norm(d, axis=1, keepdims=True)
    dw = (Cw[:3, :3].astype(np.float64) @ d.T.astype(np.float64)).T
    dw = (dw / np.linalg.norm(dw, axis=1, keepdims=True)).astype(np.float32)
    o = np.tile(Cw[:3, 3].astype(np.float32), (len(dw), 1))
    return np.concatenate([o, dw], 1).astype(np.float32)


def random_rays(n, seed, center=(0, 0, 600), spread=300.0, origin_box=50.0):
    rs = np.random.RandomState(seed)
    o = rs.uniform(-origin_box, origin_box, (n, 3)).astype(np.float32)
    tgt = (np.asarray(center) + rs.uniform(-spread, spread, (n, 3))).astype(np.float32)
    d = tgt - o
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    return np.concatenate([o, d], 1).astype(np.float32)


class ScenePair:
    """The same triangle model + octree in the oracle and on the GPU."""

    def __init__(self, ctx, meshes, cull=False, look=(0, 0, 1), rigid=None, precomputed_world=True, materials=None):
        """materials: optional recipe f(scene) -> per-mesh material ids, applied identically to both scenes."""
        self.meshes = meshes
        self.orc = O.OracleScene()
        self.orc.set_model(meshes, rigid=rigid, precomputed_world=precomputed_world, cull_backface=cull, look_dir=look)
        self.orc.build_octree()
        self.ms = api.MeshSet(meshes)
        self.oct = api.Octtree_Model(self.ms, rigid=rigid, precomputed_world=precomputed_world)
        self.gpu = api.Scene(ctx)
        self.cull_bits = self.oct.compute_backface(look) if cull else None
        mm = None
        if materials is not None:
            self.orc.set_mesh_materials(materials(self.orc))
            mm = materials(self.gpu)
        self.gpu.set_model(self.oct, cull_bits=self.cull_bits, mesh_materials=mm)
        self.gpu.commit()

    def close(self):
        self.gpu.close(); self.oct.close(); self.orc.close()
