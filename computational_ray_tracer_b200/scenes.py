"""Procedural scenes of BASELINE.json's configs.

The reference ships no scene files (its models live outside the repository,
Applications/RayTracerTestApp.h:70-73), so every workload is synthetic and is inserted the way
the reference's own code would see it: a `MeshCache::Model` (list of meshes with positions,
normals, indices; RayTracer/AssetManager.h:20-47) whose positions are already in world space
(`precomputed_worldtransform`, RayTracer/Shapes.h:923).  Scene randomness comes from the
reference's PCG32 (`pbrt::RNG(seqIndex)`, ThirdParty/pbrv4/rng.h:24-162).
"""
import math

import numpy as np

_M64 = (1 << 64) - 1
_PCG_MULT = 0x5851F42D4C957F2D


def mix_bits(v):
    """MixBits (ThirdParty/pbrv4/hash.h:67-74)."""
    v &= _M64
    v ^= v >> 31
    v = (v * 0x7FB5D329728EA185) & _M64
    v ^= v >> 27
    v = (v * 0x81DADEF4BC2DD44D) & _M64
    v ^= v >> 33
    return v


class PCG32:
    """pbrt::RNG (ThirdParty/pbrv4/rng.h:24-162): SetSequence(seq) + Uniform<uint32>/<float>."""

    def __init__(self, seq):
        self.state = 0
        self.inc = ((seq << 1) | 1) & _M64
        self.uniform_u32()
        self.state = (self.state + mix_bits(seq)) & _M64
        self.uniform_u32()

    def uniform_u32(self):
        old = self.state
        self.state = (old * _PCG_MULT + self.inc) & _M64
        xs = (((old >> 18) ^ old) >> 27) & 0xFFFFFFFF
        rot = old >> 59
        return ((xs >> rot) | (xs << ((-rot) & 31))) & 0xFFFFFFFF

    def uniform_float(self):
        return float(min(np.float32(1.0), np.float32(self.uniform_u32()) * np.float32(2.0 ** -32)))


def _grid_mesh(nx, ny, x0, x1, y0, y1, zfun, nfun):
    xs = np.linspace(x0, x1, nx + 1, dtype=np.float64)
    ys = np.linspace(y0, y1, ny + 1, dtype=np.float64)
    X, Y = np.meshgrid(xs, ys, indexing="xy")          # (ny+1, nx+1)
    Z = zfun(X, Y)
    pos = np.stack([X, Y, Z], -1).reshape(-1, 3).astype(np.float32)
    nrm = nfun(X, Y).reshape(-1, 3)
    nrm = (nrm / np.linalg.norm(nrm, axis=1, keepdims=True)).astype(np.float32)
    j, i = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
    v00 = (j * (nx + 1) + i).reshape(-1)
    v10 = v00 + 1
    v01 = v00 + (nx + 1)
    v11 = v01 + 1
    # both triangles wind so that cross(p1-p0, p2-p0) points to -z (towards a camera at the origin)
    tris = np.stack([np.stack([v00, v01, v10], -1), np.stack([v10, v01, v11], -1)], 1).reshape(-1, 3)
    return dict(positions=pos, normals=nrm, indices=tris.astype(np.uint32))


def quad_mesh(p00, p10, p01, p11, normal):
    """Two triangles (p00,p10,p11),(p00,p11,p01); caller picks the corner order so the face normal is `normal`."""
    pos = np.array([p00, p10, p01, p11], np.float32)
    n = np.tile(np.asarray(normal, np.float32), (4, 1))
    idx = np.array([[0, 1, 3], [0, 3, 2]], np.uint32)
    # flip winding if the geometric normal disagrees with the requested one
    g = np.cross(pos[1] - pos[0], pos[3] - pos[0])
    if np.dot(g, normal) < 0:
        idx = idx[:, ::-1].copy()
    return dict(positions=pos, normals=n, indices=idx)


def heightfield(n_quads=708, seed=1, extent=400.0, z0=800.0, with_light=True):
    """C2/C5: height-field grid of n_quads^2 quads (708 -> 1 002 528 triangles, 2237 -> 10 008 338).

    z = z0 + sum_k a_k sin(f_k x + phi_k) sin(g_k y + psi_k), four octaves, phases from RNG(seed).
    Mesh 0 = surface (Lambert), mesh 1 = one emissive quad facing the surface.
    """
    rng = PCG32(seed)
    octs = []
    for k in range(4):
        a = 12.0 / (2 ** k)
        f = (2 * math.pi / (2 * extent)) * (1.5 * 2 ** k)
        g = (2 * math.pi / (2 * extent)) * (1.25 * 2 ** k)
        phi = 2 * math.pi * rng.uniform_float()
        psi = 2 * math.pi * rng.uniform_float()
        octs.append((a, f, g, phi, psi))

    def zfun(X, Y):
        Z = np.full_like(X, z0)
        for a, f, g, phi, psi in octs:
            Z += a * np.sin(f * X + phi) * np.sin(g * Y + psi)
        return Z

    def nfun(X, Y):
        hx = np.zeros_like(X)
        hy = np.zeros_like(X)
        for a, f, g, phi, psi in octs:
            hx += a * f * np.cos(f * X + phi) * np.sin(g * Y + psi)
            hy += a * g * np.sin(f * X + phi) * np.cos(g * Y + psi)
        return np.stack([hx, hy, -np.ones_like(X)], -1)

    meshes = [_grid_mesh(n_quads, n_quads, -extent, extent, -extent, extent, zfun, nfun)]
    if with_light:
        zl = z0 - 150.0
        meshes.append(quad_mesh((-150, 300, zl), (150, 300, zl), (-150, 400, zl), (150, 400, zl), (0, 0, 1)))
    return meshes


def cornell_box():
    """C1: box x,y in [-250,250], z in [400,900], open towards the camera; meshes:
    0 white (floor, ceiling, back), 1 red (left), 2 green (right), 3 light (100x100 quad at y=249 facing down)."""
    a, z0, z1 = 250.0, 400.0, 900.0
    floor = quad_mesh((-a, -a, z0), (a, -a, z0), (-a, -a, z1), (a, -a, z1), (0, 1, 0))
    ceil_ = quad_mesh((-a, a, z0), (a, a, z0), (-a, a, z1), (a, a, z1), (0, -1, 0))
    back = quad_mesh((-a, -a, z1), (a, -a, z1), (-a, a, z1), (a, a, z1), (0, 0, -1))
    white = merge_meshes([floor, ceil_, back])
    left = quad_mesh((-a, -a, z0), (-a, -a, z1), (-a, a, z0), (-a, a, z1), (1, 0, 0))
    right = quad_mesh((a, -a, z0), (a, -a, z1), (a, a, z0), (a, a, z1), (-1, 0, 0))
    light = quad_mesh((-50, 249, 600), (50, 249, 600), (-50, 249, 700), (50, 249, 700), (0, -1, 0))
    return [white, left, right, light]


def merge_meshes(ms):
    pos, nrm, idx, base = [], [], [], 0
    for m in ms:
        pos.append(m["positions"]); nrm.append(m["normals"]); idx.append(m["indices"] + base)
        base += len(m["positions"])
    return dict(positions=np.concatenate(pos), normals=np.concatenate(nrm), indices=np.concatenate(idx).astype(np.uint32))


def axis_grid(n=24, z=500.0, extent=240.0, layers=2):
    """Adversarial scene for hit-ID ties: axis-aligned grids whose edges land on pixel-centre rays and on octree
    split planes, stacked in `layers` coincident-in-projection sheets."""
    ms = []
    for l in range(layers):
        zz = z + 40.0 * l
        ms.append(_grid_mesh(n, n, -extent, extent, -extent, extent, lambda X, Y: np.full_like(X, zz),
                             lambda X, Y: np.stack([np.zeros_like(X), np.zeros_like(X), -np.ones_like(X)], -1)))
    return [merge_meshes(ms)]


def random_soup(n_tris=2000, seed=7, center=(0, 0, 600), spread=200.0, size=30.0):
    """Unstructured triangle soup (ragged leaves, overlapping triangles, random orientations)."""
    rs = np.random.RandomState(seed)
    c = rs.uniform(-spread, spread, (n_tris, 3)) + np.asarray(center)
    v = c[:, None, :] + rs.uniform(-size, size, (n_tris, 3, 3))
    pos = v.reshape(-1, 3).astype(np.float32)
    idx = np.arange(3 * n_tris, dtype=np.uint32).reshape(-1, 3)
    p = pos.reshape(-1, 3, 3).astype(np.float64)
    g = np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0])
    g /= np.maximum(np.linalg.norm(g, axis=1, keepdims=True), 1e-30)
    nrm = np.repeat(g, 3, axis=0).astype(np.float32)
    return [dict(positions=pos, normals=nrm, indices=idx)]


def emissive_scatter(n_lights=1000, seed=4, extent=380.0, z_lo=560.0, z_hi=700.0, side=5.0):
    """C4: small emissive triangles scattered in a slab above the height field; returns (mesh, power) where
    power is log-uniform in [1, 100] per triangle (drawn with RNG(seed)) and realised as area (side ~ sqrt(power))."""
    rng = PCG32(seed)
    pos = np.zeros((n_lights, 3, 3), np.float64)
    power = np.zeros(n_lights, np.float32)
    for i in range(n_lights):
        c = np.array([(2 * rng.uniform_float() - 1) * extent, (2 * rng.uniform_float() - 1) * extent,
                      z_lo + (z_hi - z_lo) * rng.uniform_float()])
        ang = 2 * math.pi * rng.uniform_float()
        power[i] = 10.0 ** (2.0 * rng.uniform_float())
        sd = side * math.sqrt(power[i] / 10.0)          # same radiance everywhere: power ~ area
        e1 = np.array([math.cos(ang), math.sin(ang), 0.0]) * sd
        e2 = np.array([-math.sin(ang), math.cos(ang), 0.0]) * sd
        # wind so the geometric normal is +z (facing the height field)
        pos[i, 0], pos[i, 1], pos[i, 2] = c, c + e1, c + e2
    p32 = pos.reshape(-1, 3).astype(np.float32)
    nrm = np.tile(np.array([0, 0, 1], np.float32), (3 * n_lights, 1))
    idx = np.arange(3 * n_lights, dtype=np.uint32).reshape(-1, 3)
    return dict(positions=p32, normals=nrm, indices=idx), power


# ---------------------------------------------------------------------------------------------- materials
# Recipes take any scene object exposing add_spectrum / add_material / add_shape (the CUDA api.Scene and the test
# oracle's scene share those signatures) and return the per-mesh material ids.
SWATCH = {"white": 18, "red": 14, "green": 13, "grey": 20}      # Q_1, M_1, L_1, S_1 (ThirdParty/pbrv4/pixelsensor.cpp:16-237)
COLUMN_MAJOR_IDENTITY = np.eye(4, dtype=np.float32)


def translation(x, y, z):
    """Column-major rigid transform (glm layout): translation in elements 12..14."""
    m = np.eye(4, dtype=np.float32)
    m[3, :3] = (x, y, z)
    return m


def c2_materials(sc, emit_scale=8.0):
    """C2 / C5: Lambert ConstantSpectrum(0.5) surface, one-sided emissive quad carrying the normalised D65 illuminant."""
    grey = sc.add_spectrum(0, c=0.5)
    d65 = sc.add_spectrum(4, n=2)
    surf = sc.add_material(type=0, refl=grey)
    light = sc.add_material(type=0, refl=-1, emit=d65, emit_scale=emit_scale, two_sided=0)
    return [surf, light]


def cornell_materials(sc, with_spheres=True, glass=False):
    """C1: Macbeth-swatch Lambert walls, D65 area light, two spheres (Lambert grey + either Lambert white or BK7 glass)."""
    w = sc.add_spectrum(3, n=SWATCH["white"]); r = sc.add_spectrum(3, n=SWATCH["red"]); g = sc.add_spectrum(3, n=SWATCH["green"])
    gr = sc.add_spectrum(3, n=SWATCH["grey"]); d65 = sc.add_spectrum(4, n=2)
    mw = sc.add_material(type=0, refl=w); mr = sc.add_material(type=0, refl=r); mg = sc.add_material(type=0, refl=g)
    ml = sc.add_material(type=0, refl=-1, emit=d65, emit_scale=20.0)
    mgrey = sc.add_material(type=0, refl=gr)
    if with_spheres:
        if glass:
            bk7 = sc.add_spectrum(2, name="glass_bk7")
            m2 = sc.add_material(type=1, eta=bk7, eta_constant=0)
        else:
            m2 = mw
        sc.add_shape(0, translation(-110, -170, 720), [80.0, -80.0, 80.0, 360.0], material=mgrey)
        sc.add_shape(0, translation(120, -170, 600), [80.0, -80.0, 80.0, 360.0], material=m2)
    return [mw, mr, mg, ml]


def spheres_lattice_materials(sc, n=8, spacing=60.0, radius=22.0, z=650.0):
    """C3: n x n lattice of spheres alternating dispersive glass (BK7 / SF11) and conductors (Cu / Au) over a Lambert floor
    (mesh 0) under an emissive quad (mesh 1)."""
    grey = sc.add_spectrum(3, n=SWATCH["grey"]); d65 = sc.add_spectrum(4, n=2)
    floor = sc.add_material(type=0, refl=grey)
    light = sc.add_material(type=0, refl=-1, emit=d65, emit_scale=12.0, two_sided=0)
    bk7 = sc.add_spectrum(2, name="glass_bk7"); sf11 = sc.add_spectrum(2, name="glass_sf11")
    cu_e = sc.add_spectrum(2, name="cu_eta"); cu_k = sc.add_spectrum(2, name="cu_k")
    au_e = sc.add_spectrum(2, name="au_eta"); au_k = sc.add_spectrum(2, name="au_k")
    mats = [sc.add_material(type=1, eta=bk7, eta_constant=0), sc.add_material(type=2, eta=cu_e, k=cu_k),
            sc.add_material(type=1, eta=sf11, eta_constant=0), sc.add_material(type=2, eta=au_e, k=au_k)]
    x0 = -(n - 1) * spacing / 2
    for j in range(n):
        for i in range(n):
            sc.add_shape(0, translation(x0 + i * spacing, x0 + j * spacing, z), [radius, -radius, radius, 360.0], material=mats[(i + 2 * j) % 4])
    return [floor, light]


def spheres_lattice_meshes(extent=300.0, z_floor=700.0, z_light=350.0):
    """C3 meshes: floor quad behind the lattice (facing the camera) and an emissive quad between camera and lattice, off axis."""
    floor = quad_mesh((-extent, -extent, z_floor), (extent, -extent, z_floor), (-extent, extent, z_floor), (extent, extent, z_floor), (0, 0, -1))
    light = quad_mesh((-120, 260, z_light), (120, 260, z_light), (-120, 260, z_light + 160), (120, 260, z_light + 160), (0, -1, 0))
    return [floor, light]


def many_light_scene(n_quads=354, n_lights=1000):
    """C4: height field (354^2 quads = 250 632 triangles) + n_lights small emissive triangles above it."""
    surface = heightfield(n_quads, seed=4, with_light=False)
    lights, _power = emissive_scatter(n_lights)
    return surface + [lights]


def many_light_materials(sc):
    grey = sc.add_spectrum(0, c=0.5)
    d65 = sc.add_spectrum(4, n=2)
    surf = sc.add_material(type=0, refl=grey)
    light = sc.add_material(type=0, refl=-1, emit=d65, emit_scale=40.0, two_sided=1)
    return [surf, light]


# ---------------------------------------------------------------------------------------------- BASELINE.json configs
# The five workloads of BASELINE.json as concrete synthetic inputs: scene, materials recipe, film, sampler grid,
# integrator limits.  bench.py --config, tools/run_configs.py and the full-size parity tests all read this table.
CONFIGS = {
    "C1": dict(label="C1 Cornell box (12 tris) + 2 spheres + area light, 256x256 @ 16 spp", meshes=cornell_box, materials=cornell_materials,
               width=256, height=256, spp=16, xs=4, ys=4, max_depth=5, rr_depth=0, integrator="path+NEE depth<=5"),
    "C2": dict(label="C2 heightfield 708x708 quads (1002530 tris) + emissive quad, 1920x1080 @ 64 spp", meshes=lambda: heightfield(708),
               materials=c2_materials, width=1920, height=1080, spp=64, xs=8, ys=8, max_depth=5, rr_depth=0, integrator="path+NEE depth<=5"),
    "C3": dict(label="C3 8x8 lattice of dispersive-glass / conductor spheres over a Lambert floor, 1920x1080 @ 256 spp",
               meshes=spheres_lattice_meshes, materials=spheres_lattice_materials, width=1920, height=1080, spp=256, xs=16, ys=16,
               max_depth=16, rr_depth=3, integrator="path+NEE depth<=16, Russian roulette from depth 3"),
    "C4": dict(label="C4 heightfield 354x354 quads (250632 tris) + 1000 emissive triangles (power-CDF light sampling), 1920x1080 @ 64 spp",
               meshes=many_light_scene, materials=many_light_materials, width=1920, height=1080, spp=64, xs=8, ys=8, max_depth=5, rr_depth=0,
               integrator="path+NEE depth<=5"),
    "C5": dict(label="C5 heightfield 2237x2237 quads (10008340 tris) + emissive quad, 3840x2160 @ 1024 spp", meshes=lambda: heightfield(2237, seed=5),
               materials=c2_materials, width=3840, height=2160, spp=1024, xs=32, ys=32, max_depth=5, rr_depth=0, integrator="path+NEE depth<=5"),
}
CAMERA = dict(kind=0, near=1.0, far=1000.0, fov=45.0, pos=(0, 0, 0), look=(0, 0, 1), right=(1, 0, 0), up=(0, 1, 0))     # PerspectiveCamera of every config
