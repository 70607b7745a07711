"""The fixed inputs on which the restated oracle is pinned against the REFERENCE'S OWN compiled code.

`run("ref")` evaluates every case with oracle/_ref/libcrt_ref.so (reference sources compiled unmodified, see
oracle/ref_harness.cpp), `run("oracle")` with the restatement (oracle/_build/liboracle.so).  Both return
{name: ndarray}.  tools/make_ref_golden.py stores run("ref") in tests/golden/ref_pin.npz, so the pin also holds
where /root/reference does not exist; tests/test_cpu_ref_pin.py compares.

Keys listed in TOLERANT are compared with a tolerance (and say why); everything else must be bit-identical.
TEST INFRASTRUCTURE ONLY.
"""
import ctypes as C

import numpy as np

from computational_ray_tracer_b200 import scenes

FLT_MAX = float(np.finfo(np.float32).max)

# Sphere v = (acos(z/r) - thetamin) / ...: the reference calls unqualified `acos` on a float (Shapes.h:382), which is the
# float overload under MSVC but glibc's double acos here; the restatement calls std::acos(float).  <= 1 ulp of theta.
TOLERANT = {"shape0.uv": 2e-6, "shape1.uv": 2e-6,
            # Sphere::calculate_dv (Shapes.h:401-408) calls unqualified cos / sin on floats: the same double-vs-float overload difference
            "shape0.dv": 2e-6, "shape1.dv": 2e-6}


def _rigid(tx, ty, tz, ang=0.0):
    m = scenes.translation(tx, ty, tz)
    c, s = np.float32(np.cos(ang)), np.float32(np.sin(ang))
    m[0, 0], m[0, 2], m[2, 0], m[2, 2] = c, -s, s, c
    return m


def _rays(n, seed, center=(0, 0, 600), spread=300.0, origin_box=50.0):
    rs = np.random.RandomState(seed)
    o = rs.uniform(-origin_box, origin_box, (n, 3)).astype(np.float32)
    tgt = (np.asarray(center) + rs.uniform(-spread, spread, (n, 3))).astype(np.float32)
    d = tgt - o
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    return np.concatenate([o, d], 1).astype(np.float32)


SHAPES = [
    (0, [60.0, -60.0, 60.0, 360.0]), (0, [60.0, -30.0, 45.0, 270.0]),
    (1, [40.0, -50.0, 70.0, 360.0]), (1, [40.0, -50.0, 70.0, 200.0]),
    (2, [10.0, 0.0, 70.0, 360.0]), (2, [10.0, 25.0, 70.0, 300.0]),
    (3, [-60, -40, 0, 70, -30, 5, 0, 65, -10]),
]
CAMERAS = [  # kind, near, far, sensor w/h, fov, pos, look
    (0, 1.0, 1000.0, 0.0, 0.0, 45.0, (0, 0, 0), (0, 0, 1)),
    (0, 0.5, 500.0, 0.0, 0.0, 70.0, (10, 20, -30), (0.3, -0.2, 1)),
    (1, 1.0, 1000.0, 500.0, 400.0, 0.0, (0, 0, 0), (0, 0, 1)),
    (2, 0.0, 10.0, 500.0, 500.0, 0.0, (3, -2, 1), (0.1, 0.1, 1)),
]
MODELS = {
    "soup": lambda: (scenes.random_soup(1200, 7), {}),
    "hf": lambda: (scenes.heightfield(40), {}),
    "cull": lambda: (scenes.random_soup(1500, 9), dict(cull_backface=True, look_dir=(0.2, 0.1, 1))),
    "rigid": lambda: (scenes.random_soup(800, 3, center=(0, 600, 0)), dict(rigid=_rigid(10, -20, 30, 0.3), precomputed_world=False)),
    "rigid_cull": lambda: (scenes.random_soup(900, 4, center=(0, 600, 0)), dict(rigid=_rigid(-15, 5, 20, -0.4), precomputed_world=False,
                                                                                cull_backface=True, look_dir=(-0.1, 0.3, 1))),
}


class _Backend:
    def __init__(self, which):
        self.which = which
        if which == "ref":
            import ref_lib as M
            self.M, self.L, self.p = M, M.lib(), "ref_"
            self.Scene = M.RefScene
        else:
            import oracle_lib as M
            self.M, self.L, self.p = M, M.lib(), "orc_"
            self.Scene = M.OracleScene

    def fn(self, name):
        return getattr(self.L, self.p + name)


def _integers(B, out):
    keys = [b"", b"a", b"hello world!", bytes(range(37)), bytes(range(200, 256)) * 3]
    out["murmur"] = np.array([B.fn("murmur64a")(k, len(k), s) for k in keys for s in (0, 7, 2 ** 63 + 1)], np.uint64)
    vals = [0, 1, 2 ** 63 + 12345, 0xDEADBEEFCAFEBABE, 2 ** 64 - 1]
    out["mixbits"] = np.array([B.fn("mixbits")(v) for v in vals], np.uint64)
    px = [(0, 0, 0), (5, 9, 3), (1919, 1080, -4), (-3, 7, 11), (3839, 2160, 123456)]
    out["hash2"] = np.array([B.fn("hash_pixel_seed")(*p) for p in px], np.uint64)
    out["hash3"] = np.array([B.fn("hash_pixel_dim_seed")(p[0], p[1], d, p[2]) for p in px for d in (0, 1, 5, 1000)], np.uint64)
    rs = np.random.RandomState(0)
    pe = []
    for _ in range(3000):
        l = int(rs.randint(1, 5000)); i = int(rs.randint(0, l)); p = int(rs.randint(0, 2 ** 32))
        pe.append(B.fn("permutation_element")(i, l, p))
    out["permutation"] = np.array(pe, np.int32)
    u32, f32 = [], []
    for mode, seq, off, adv in [(0, 0, 0, 0), (1, 42, 0, 0), (2, 42, 54, 0), (1, 7, 0, 123456789), (1, 7, 0, -1000), (2, 2 ** 63 + 5, 99, 65536 * 7 + 3)]:
        a = np.zeros(32, np.uint32); b = np.zeros(32, np.float32)
        B.fn("pcg32")(mode, seq, off, adv, 32, a.ctypes.data_as(C.POINTER(C.c_uint32)), None)
        B.fn("pcg32")(mode, seq, off, adv, 32, None, B.M.fp(b))
        u32.append(a); f32.append(b)
    out["pcg32_u32"] = np.stack(u32); out["pcg32_float"] = np.stack(f32)


def _sampling(B, out):
    pat = b"1p2211p2"
    seqs = []
    for kind, xs, ys, j in [(0, 4, 4, 1), (1, 4, 4, 1), (1, 8, 8, 1), (1, 3, 5, 0), (1, 16, 16, 1), (1, 32, 32, 1)]:
        for px, py, idx, dim in [(0, 0, 0, 0), (17, 33, 5, 0), (1919, 1080, 14, 3), (5, 5, xs * ys - 1, 7)]:
            a = np.zeros(12, np.float32)
            B.fn("sampler_sequence")(kind, xs, ys, j, 3, px, py, idx, dim, pat, B.M.fp(a))
            seqs.append(a)
    out["sampler"] = np.stack(seqs)
    us = np.concatenate([np.linspace(0, 0.999999, 600, dtype=np.float32), np.float32([0.0, 0.125, 0.5, 0.875, np.nextafter(np.float32(1), np.float32(0))])])
    lam = np.zeros((len(us), 8), np.float32); pdf = np.zeros((len(us), 8), np.float32)
    for i, u in enumerate(us):
        B.fn("sample_visible")(float(u), B.M.fp(lam[i]), B.M.fp(pdf[i]))
    out["visible.lambda"] = lam; out["visible.pdf"] = pdf
    rs = np.random.RandomState(2)
    uv = rs.rand(400, 2).astype(np.float32)
    fs = np.zeros((400, 3), np.float32); cd = np.zeros((400, 2), np.float32)
    for i in range(400):
        B.fn("filter_sample")(0, 0.5, 0.75, float(uv[i, 0]), float(uv[i, 1]), B.M.fp(fs[i]))
        B.fn("concentric_disk")(float(uv[i, 0]), float(uv[i, 1]), B.M.fp(cd[i]))
    out["boxfilter"] = fs; out["concentric"] = cd
    out["gamma"] = np.float32([B.fn("gamma")(n) for n in range(1, 9)])
    # TriangleFilter: SampleTent flips a coin from a global mt19937 (Sampling.h:228-235, not a function of u); what IS deterministic is
    # pinned: the branch each outcome takes, -r + r * SampleLinear(u, 0, 1) and r * SampleLinear(u, 1, 0)
    ul = np.concatenate([np.float32([0.0, 1.0, 0.5]), rs.rand(500).astype(np.float32)])
    out["sample_linear.up"] = np.float32([B.fn("sample_linear")(float(v), 0.0, 1.0) for v in ul])
    out["sample_linear.down"] = np.float32([B.fn("sample_linear")(float(v), 1.0, 0.0) for v in ul])
    ab = rs.randn(1000, 4).astype(np.float32)
    out["dop"] = np.float32([B.fn("difference_of_products")(*[float(v) for v in r]) for r in ab])


def _colour(B, out):
    quiet = getattr(B.M, "_quiet", lambda f, *a: f(*a))
    for w, nm in enumerate(("X", "Y", "Z", "D65")):
        a = np.zeros(471, np.float32)
        quiet(B.fn("dense_table"), w, B.M.fp(a))
        out["dense." + nm] = a
    arrs = [np.zeros(9, np.float32) for _ in range(3)] + [np.zeros(2, np.float32)]
    B.fn("color_constants")(*[B.M.fp(x) for x in arrs])
    for nm, a in zip(("XYZFromSensorRGB", "RGBFromXYZ", "XYZFromRGB", "white"), arrs):
        out["colour." + nm] = a
    rs = np.random.RandomState(4)
    cs = rs.uniform(-3, 3, (300, 3)).astype(np.float32) * np.float32([1e-4, 1e-2, 1.0])
    lam = rs.uniform(360, 830, 300).astype(np.float32)
    out["sigmoid"] = np.float32([B.fn("sigmoid_eval")(float(c[0]), float(c[1]), float(c[2]), float(l)) for c, l in zip(cs, lam)])
    # named spectra (illuminants, glass, metals) at table wavelengths and in between
    lams = np.concatenate([np.arange(355, 836, 5, dtype=np.float32), rs.uniform(360, 830, 64).astype(np.float32)])
    # reference registry name (spectrum.cpp:2691-2712) -> the restatement's table name (oracle_pbrt.cpp named_table)
    table = {"glass-BK7": "glass_bk7", "glass-BAF10": "glass_baf10", "glass-FK51A": "glass_fk51a", "glass-LASF9": "glass_lasf9",
             "glass-F5": "glass_sf5", "glass-F10": "glass_sf10", "glass-F11": "glass_sf11", "metal-Ag-eta": "ag_eta", "metal-Ag-k": "ag_k",
             "metal-Al-eta": "al_eta", "metal-Al-k": "al_k", "metal-Au-eta": "au_eta", "metal-Au-k": "au_k", "metal-Cu-eta": "cu_eta",
             "metal-Cu-k": "cu_k", "metal-CuZn-eta": "cuzn_eta", "metal-CuZn-k": "cuzn_k"}
    names = ["stdillum-A", "stdillum-D50", "stdillum-D65", "stdillum-F1", "stdillum-F2", "stdillum-F11"] + list(table)
    if B.which == "ref":
        for nm in names:
            a = np.zeros(len(lams), np.float32)
            assert B.L.ref_named_spectrum_query(nm.encode(), B.M.fp(lams), len(lams), B.M.fp(a)) == 0, nm
            out["named." + nm] = a
    else:
        sc = B.Scene()
        illum = {"stdillum-A": 0, "stdillum-D50": 1, "stdillum-D65": 2, "stdillum-F1": 3, "stdillum-F2": 4, "stdillum-F11": 5}
        for nm in names:
            sid = sc.add_spectrum(4, n=illum[nm]) if nm in illum else sc.add_spectrum(2, name=table[nm])
            assert sid >= 0, nm
            a = np.zeros(len(lams), np.float32)
            for k in range(0, len(lams), 8):
                chunk = np.zeros(8, np.float32); m = min(8, len(lams) - k); chunk[:m] = lams[k:k + m]
                a[k:k + m] = sc.spectrum_sample(sid, chunk)[:m]
            out["named." + nm] = a
        sc.close()


def _cameras_shapes(B, out):
    for i, (kind, near, far, sw, sh, fov, pos, look) in enumerate(CAMERAS):
        r2c, c2w = B.M.camera_matrices(kind, near, far, sw, sh, fov, pos, look, (1, 0, 0), (0, 1, 0), 640, 480)
        out[f"camera{i}.r2c"] = r2c; out[f"camera{i}.c2w"] = c2w
    for k, (kind, params) in enumerate(SHAPES):
        rigid = _rigid(10, -5, 500, ang=0.4 + 0.1 * k)
        o2r = np.zeros(16, np.float32); r2o = np.zeros(16, np.float32)
        B.fn("shape_matrices")(B.M.fp(B.M.f32(rigid).reshape(-1)), B.M.fp(o2r), B.M.fp(r2o))
        out[f"shape{k}.o2r"] = o2r; out[f"shape{k}.r2o"] = r2o
        sc = B.Scene(); sid = sc.add_shape(kind, rigid, params)
        rays = _rays(4000, 20 + k, center=(10, -5, 500), spread=90.0, origin_box=120.0)
        for tm, tag in ((FLT_MAX, ""), (480.0, ".tmax")):
            r = sc.shape_intersect(sid, rays, tm)
            f = r["found"] > 0
            out[f"shape{k}{tag}.found"] = r["found"]
            if tag == "":
                for key in ("t", "hitp", "n", "uv", "du", "dv", "wo"):
                    out[f"shape{k}.{key}"] = r[key][f]
        sc.close()


def _models(B, out):
    for name, make in MODELS.items():
        meshes, kw = make()
        sc = B.Scene(); sc.set_model(meshes, **kw); sc.build_octree()
        d = sc.octree_dump()
        for k, v in d.items():
            out[f"{name}.octree.{k}"] = v
        out[f"{name}.bounds"] = sc.model_bounds()
        if kw.get("cull_backface"):             # TriModel::ComputeBackFace (Shapes.h:1339-1380): the per-triangle culling table itself
            out[f"{name}.backfacing"] = np.concatenate([sc.backfacing(mi, len(m["indices"])) for mi, m in enumerate(meshes)])
        b = out[f"{name}.bounds"]
        ctr = tuple((b[:3] + b[3:]) / 2)
        rays = np.concatenate([_rays(1500, 11, center=ctr, spread=250.0), _rays(300, 12, center=ctr, spread=200.0, origin_box=5.0)])
        s = sc.traverse_surface(rays)
        out[f"{name}.traverse.found"] = s["found"]
        f = s["found"] > 0
        # tHit of a triangle hit is never assigned by the reference (LocalSurfaceInfo info; Shapes.h:1034) -> not compared
        for key in ("n", "hitp", "uv"):
            out[f"{name}.traverse.{key}"] = s[key][f]
        if B.which == "ref":
            br = sc.brute_force(rays[:400])
        else:
            br = sc.trace(rays[:400], mode=1)
        hit = br["mesh"] >= 0
        out[f"{name}.brute.mesh"] = br["mesh"]; out[f"{name}.brute.tri"] = br["tri"]
        out[f"{name}.brute.t"] = br["t"][hit]; out[f"{name}.brute.bary"] = br["bary"][hit]
        sc.close()


def tier_a_params(B, W, H, cam, sk, xs, ys, j, spp):
    """cam: pos, look, optional lens_radius/focal_distance, kind (0 perspective, 1 orthographic, 2 pinhole), sensor (w, h), near, far
    (pinhole: sensor = box x/y, far = box depth, Cameras.h:313-359)."""
    lens = cam.get("lens_radius", 0.0); foc = cam.get("focal_distance", 0.0)
    kind = cam.get("kind", 0); sw, sh = cam.get("sensor", (0.0, 0.0)); near = cam.get("near", 1.0); far = cam.get("far", 1000.0)
    if B.which == "ref":
        return B.M.make_params(W, H, camera_kind=kind, near=near, far=far, sensor=(sw, sh), pos=cam["pos"], look=cam["look"], lens_radius=lens,
                               focal_distance=foc, sampler_kind=sk, xs=xs, ys=ys, jitter=j, seed=3, albedo=(0.6, 0.6, 0.6), spp_begin=0, spp_end=spp, nthreads=4)
    r2c, c2w = B.M.camera_matrices(kind, near, far, sw, sh, 45.0, cam["pos"], cam["look"], (1, 0, 0), (0, 1, 0), W, H)
    if kind == 2:
        foc = far                       # the restatement's pinhole camera takes the box depth in the focal_distance slot
    return B.M.make_params(W, H, r2c, c2w, lens_radius=lens, focal_distance=foc, camera_kind=kind, sampler_kind=sk, xs=xs, ys=ys, jitter=j, seed=3,
                           mode=0, albedo=(0.6, 0.6, 0.6), spp_begin=0, spp_end=spp, nthreads=4)


TIER_A = [
    ("hf", lambda: (scenes.heightfield(32), {}), dict(pos=(0, 0, 0), look=(0, 0, 1)), (1, 4, 4, 1)),
    ("soup", lambda: (scenes.random_soup(1500, 5), dict(cull_backface=True, look_dir=(0, 0, 1))), dict(pos=(5, -3, 10), look=(0.05, 0.02, 1)), (0, 4, 4, 1)),
    ("lens", lambda: (scenes.heightfield(24), {}), dict(pos=(0, 0, 0), look=(0, 0, 1), lens_radius=20.0, focal_distance=700.0), (1, 3, 3, 0)),
    ("ortho", lambda: (scenes.heightfield(24), {}), dict(pos=(10, -5, 0), look=(0, 0, 1), kind=1, sensor=(700.0, 460.0)), (1, 2, 2, 1)),
    ("pinhole", lambda: (scenes.heightfield(24), {}), dict(pos=(0, 0, 0), look=(0, 0, 1), kind=2, sensor=(40.0, 30.0), far=50.0), (1, 2, 2, 1)),
]


def _tier_a(B, out):
    W, H = 48, 32
    for name, make, cam, (sk, xs, ys, j) in TIER_A:
        meshes, kw = make()
        sc = B.Scene(); sc.set_model(meshes, **kw); sc.build_octree()
        p = tier_a_params(B, W, H, cam, sk, xs, ys, j, 3)
        rs = np.random.RandomState(1)
        pid = rs.randint(0, W * H, 400); idx = rs.randint(0, xs * ys, 400)
        e = sc.eval_samples(p, pid, idx)
        for k, v in e.items():
            out[f"tierA.{name}.sample.{k}"] = v
        film = sc.render(p)["film"] if B.which == "oracle" else sc.render_tier_a(p)
        out[f"tierA.{name}.film"] = film
        rgb8, rgbf = B.M.resolve(film)
        out[f"tierA.{name}.rgb8"] = rgb8; out[f"tierA.{name}.rgbf"] = rgbf
        sc.close()


def pseudo_rgb_table(seed=11):
    """A deterministic stand-in for the contents of `sRGB64binary` (legacy RandomState streams are stable across numpy
    versions): the LOOKUP is what is pinned here, so any table will do -- smooth-ish coefficients of realistic magnitude and
    the generator's own scale nodes."""
    rs = np.random.RandomState(seed)
    k = np.arange(64, dtype=np.float64) / 63.0
    sm = lambda x: x * x * (3.0 - 2.0 * x)
    scale = sm(sm(k)).astype(np.float32)
    data = (rs.uniform(-1, 1, (3, 64, 64, 64, 3)) * np.float32([1e-4, 1e-1, 30.0])).astype(np.float32)
    return scale, np.ascontiguousarray(data)


def rgb_inputs():
    rs = np.random.RandomState(12)
    rgb = rs.rand(500, 3).astype(np.float32)
    special = np.float32([[1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 0], [0.5, 0.5, 0.25], [0.25, 0.5, 0.5], [0.2, 0.2, 0.7], [1, 1, 0.999],
                          [1e-4, 2e-4, 3e-4], [0.999, 1.0, 0.5], [0.3, 0.3, 0.3], [0.0, 0.0, 0.5], [0.7, 0.7, 0.2], [1.0, 0.5, 1.0]])
    return np.concatenate([special, rgb])


def _rgb2spec(B, out):
    """RGBToSpectrumTable::operator() + RGBAlbedo/Illuminant/Unbounded spectra of non-grey RGB on a caller-provided table."""
    scale, data = pseudo_rgb_table()
    B.fn("set_rgb_table")(B.M.fp(scale), B.M.fp(data))
    rgbs = rgb_inputs()
    lams = np.concatenate([np.float32([360, 830, 555.5]), np.random.RandomState(13).uniform(360, 830, 13).astype(np.float32)])
    q = np.zeros((len(rgbs), len(lams)), np.float32)
    samp = np.zeros((3, len(rgbs), 8), np.float32)
    if B.which == "ref":
        lam8 = np.zeros(8, np.float32)
        for i, c in enumerate(rgbs):
            B.L.ref_rgb_albedo_query(B.M.fp(c), B.M.fp(lams), len(lams), B.M.fp(q[i]))
            for kind in range(3):
                B.L.ref_rgb_spectrum_sample(kind, B.M.fp(c), float(0.37 + 0.001 * (i % 100)), B.M.fp(lam8), B.M.fp(samp[kind, i]))
    else:
        sc = B.Scene()
        for i, c in enumerate(rgbs):
            u = float(0.37 + 0.001 * (i % 100))
            lam8 = np.zeros(8, np.float32); pdf8 = np.zeros(8, np.float32)
            B.L.orc_sample_visible(u, B.M.fp(lam8), B.M.fp(pdf8))
            ids = [sc.add_spectrum(7 + kind, interleaved=c) for kind in range(3)]
            assert min(ids) >= 0
            for k in range(0, len(lams), 8):
                chunk = np.zeros(8, np.float32); m = min(8, len(lams) - k); chunk[:m] = lams[k:k + m]
                q[i, k:k + m] = sc.spectrum_sample(ids[0], chunk)[:m]
            for kind in range(3):
                samp[kind, i] = sc.spectrum_sample(ids[kind], lam8)
        sc.close()
        B.L.orc_set_rgb_table(None, None)
    out["albedo.query"] = q
    out["albedo.sample"] = samp[0]; out["illuminant.sample"] = samp[1]; out["unbounded.sample"] = samp[2]


def gaussian_inputs():
    rs = np.random.RandomState(5)
    u = rs.rand(3000, 2).astype(np.float32)
    u[0] = (0, 0); u[1] = (1.0, 0.5); u[2] = (0.5, np.nextafter(np.float32(1), np.float32(0))); u[3] = (1e-7, 0.25)
    return u, [(0.5, 0.5, 0.5), (1.5, 1.0, 0.4), (2.0, 2.0, 1.0)]


def _gaussian_filter(B, out):
    """pbrt::GaussianFilter(radius, sigma).Sample(u) (filters.h:96-163) through its tabulated inverse CDF (Sampling.h:781-848),
    incl. U = 0 ("couldn't find index" -> 0) and U = 1 (zero boundary of the filter: weight 0/0)."""
    u, params = gaussian_inputs()
    quiet = getattr(B.M, "_quiet", lambda f, *a: f(*a))
    for i, (rx, ry, sg) in enumerate(params):
        a = np.zeros((len(u), 3), np.float32)
        quiet(B.fn("gaussian_filter_samples"), rx, ry, sg, B.M.fp(u), len(u), B.M.fp(a))
        out[f"gauss{i}.p"] = a[:, :2].copy()
        w = a[:, 2].copy()
        out[f"gauss{i}.weight_nan"] = np.isnan(w).astype(np.uint8)
        out[f"gauss{i}.weight"] = np.where(np.isnan(w), np.float32(0), w)


SENSORS = [("canon_eos_100d", "stdillum-A"), ("nikon_d810", "stdillum-D65")]


def _sensor(B, out):
    """Measured PixelSensor (pixelsensor.h:37-68; the app's sensor_canon, RayTracerTestApp.h:152-153): XYZFromSensorRGB, ToSensorRGB
    through a Tier A film and its resolve.  The response curves are INPUT data (named spectra of the reference's registry, sampled at
    the integer wavelengths); the reference run exports them and the golden file carries them to where the reference is absent."""
    meshes = scenes.heightfield(24)
    W, H = 40, 24
    cam = dict(pos=(0, 0, 0), look=(0, 0, 1))
    for i, (camname, illum) in enumerate(SENSORS):
        cur = np.zeros((3, 471), np.float32); ill = np.zeros(471, np.float32); mat = np.zeros(9, np.float32)
        ratio = 1.0 / 106.856895
        if B.which == "ref":
            rc = B.M._quiet(B.L.ref_set_sensor, (camname + "_r").encode(), (camname + "_g").encode(), (camname + "_b").encode(), illum.encode(),
                            ratio, B.M.fp(cur), B.M.fp(ill), B.M.fp(mat))
            assert rc == 0, camname
        else:
            cur, ill = _sensor.inputs[i]                      # handed over from the reference run / the golden file
            cur = np.ascontiguousarray(cur); ill = np.ascontiguousarray(ill)
            B.L.orc_set_sensor(B.M.fp(np.ascontiguousarray(cur[0])), B.M.fp(np.ascontiguousarray(cur[1])), B.M.fp(np.ascontiguousarray(cur[2])),
                               B.M.fp(ill), ratio, B.M.fp(mat))
        out[f"sensor{i}.curves"] = cur; out[f"sensor{i}.illum"] = ill; out[f"sensor{i}.matrix"] = mat
        sc = B.Scene(); sc.set_model(meshes); sc.build_octree()
        p = tier_a_params(B, W, H, cam, 1, 2, 2, 1, 2)
        rs = np.random.RandomState(7)
        e = sc.eval_samples(p, rs.randint(0, W * H, 200), rs.randint(0, 4, 200))
        out[f"sensor{i}.sample.rgb"] = e["rgb"]
        film = sc.render(p)["film"] if B.which == "oracle" else sc.render_tier_a(p)
        out[f"sensor{i}.film"] = film
        rgb8, rgbf = B.M.resolve(film)
        out[f"sensor{i}.rgb8"] = rgb8; out[f"sensor{i}.rgbf"] = rgbf
        sc.close()
    if B.which == "ref":
        B.L.ref_set_sensor(None, None, None, None, 0.0, None, None, None)
    else:
        B.L.orc_set_sensor(None, None, None, None, 0.0, None)


_sensor.inputs = None


def _tier_b_parts(B, out):
    """The building blocks of the path integrator that DO exist in the reference (the integrator itself does not, Integrator.h:4-12):
    cosine-hemisphere sampling and pdf (Sampling.h:449-459), TerminateSecondary (spectrum.h:302-310), Shape::Area / Triangle::Area
    (Shapes.h:198,234,455,642,779,949) which weight the light selection."""
    rs = np.random.RandomState(21)
    u = rs.rand(2000, 2).astype(np.float32); u[0] = (0, 0); u[1] = (0.5, 0.5); u[2] = (1.0, 0.0)
    w = np.zeros((len(u), 3), np.float32); pdf = np.zeros(len(u), np.float32)
    B.fn("cosine_hemisphere")(B.M.fp(u), len(u), B.M.fp(w), B.M.fp(pdf))
    out["cosine.w"] = w; out["cosine.pdf"] = pdf
    ts = np.zeros((64, 8), np.float32)
    for i in range(64):
        B.fn("terminate_secondary")(float(i) / 64.0, B.M.fp(ts[i]))
    out["terminate_secondary.pdf"] = ts
    sc = B.Scene()
    out["shape.area"] = np.float32([B.fn("shape_area")(sc.h, sc.add_shape(kind, _rigid(10, -5, 500, 0.4), params)) for kind, params in SHAPES])
    sc.close()
    meshes, kw = MODELS["rigid"]()
    sc = B.Scene(); sc.set_model(meshes, **kw)
    n = 300
    mesh = np.zeros(n, np.int32); tri = rs.randint(0, len(meshes[0]["indices"]), n).astype(np.int32); area = np.zeros(n, np.float32)
    B.fn("triangle_area")(sc.h, B.M.ip(mesh), B.M.ip(tri), n, B.M.fp(area))
    out["triangle.area"] = area
    sc.close()


def sat_inputs(n=6000, seed=31):
    """Box / triangle triples for the Akenine-Moller overlap test: random ones, triangles with a vertex on a box face, axis-aligned and
    coplanar-with-a-face triangles, thin slivers -- the places where the reference's epsilon-free comparisons and its one disabled axis
    test (AABB_triangle_Moller.h:342-344) decide the answer."""
    rs = np.random.RandomState(seed)
    c = rs.uniform(-50, 50, (n, 3)).astype(np.float32); h = rs.uniform(0.5, 20, (n, 3)).astype(np.float32)
    tri = (c[:, None, :] + rs.uniform(-2.0, 2.0, (n, 3, 3)).astype(np.float32) * h[:, None, :]).astype(np.float32)
    k = n // 6
    tri[:k, 0, 0] = c[:k, 0] + h[:k, 0]                                   # a vertex exactly on the +x face
    tri[k:2 * k, :, 2] = (c[k:2 * k, 2] + h[k:2 * k, 2])[:, None]          # triangle coplanar with the +z face
    tri[2 * k:3 * k, :, 1] = tri[2 * k:3 * k, 0:1, 1]                      # axis-aligned (constant y) triangles
    tri[3 * k:4 * k, 2] = tri[3 * k:4 * k, 1] + np.float32(1e-3) * (tri[3 * k:4 * k, 0] - tri[3 * k:4 * k, 1])     # slivers
    return c, h, np.ascontiguousarray(tri.reshape(n, 9))


def _sat(B, out):
    c, h, tri = sat_inputs()
    res = np.zeros(len(c), np.int32)
    B.fn("tri_box_overlap")(B.M.fp(c), B.M.fp(h), B.M.fp(tri), len(c), B.M.ip(res))
    out["overlap"] = res


def attribute_model(n_tris=700, seed=5):
    """A soup whose vertices carry every MeshCache::Mesh attribute (AssetManager.h:20-47): normals, texcoords (some outside [0,1] so the
    clamp of Shapes.h:1046-1047 acts), tangents and bitangents (not unit length: CalculateLocalSurface normalises, :1054,:1063)."""
    m = scenes.random_soup(n_tris, seed)[0]
    rs = np.random.RandomState(seed + 100)
    nv = len(m["positions"])
    m = dict(m, texcoords=rs.uniform(-0.2, 1.2, (nv, 2)).astype(np.float32), tangents=rs.normal(size=(nv, 3)).astype(np.float32),
             bitangents=(3.0 * rs.normal(size=(nv, 3))).astype(np.float32))
    return [m]


def degenerate_triangles():
    """Triangles for Triangle::CalculateLocalSurface called directly (no ray can reach the degenerate branch, Shapes.h:1016-1029, through
    BasicIntersect): ordinary ones, tiny ones whose cross product underflows in float but not in double, collinear and repeated vertices."""
    rs = np.random.RandomState(77)
    tri = rs.uniform(-100, 100, (64, 3, 3)).astype(np.float32)
    tri[16:32] = (tri[16:32, 0:1] + rs.uniform(-1, 1, (16, 3, 3)).astype(np.float32) * np.float32(3e-21)).astype(np.float32)   # |cross| ~ 1e-41
    tri[32:40] = (tri[32:40, 0:1] + rs.uniform(-1, 1, (8, 3, 3)).astype(np.float32) * np.float32(1e-12)).astype(np.float32)    # |cross|^2 underflows
    tri[40:48, 2] = tri[40:48, 0] + np.float32(2.0) * (tri[40:48, 1] - tri[40:48, 0])                                          # collinear
    tri[48:52, 1] = tri[48:52, 0]                                                                                                # repeated vertex
    tri[52:56] = np.float32([[0, 0, 5], [1e-20, 0, 5], [0, 1e-20, 5]])                                                          # axis-aligned, ng = +-z exactly
    bary = rs.dirichlet((1, 1, 1), 64).astype(np.float32)
    rayd = rs.normal(size=(64, 3)); rayd = (rayd / np.linalg.norm(rayd, axis=1, keepdims=True)).astype(np.float32)
    return tri, bary, rayd


def _canon(a):
    """NaN payloads / signs are not part of the contract."""
    a = np.array(a, np.float32)
    a[np.isnan(a)] = np.float32(np.nan)
    return a


def _local_surface(B, out):
    """Triangle::CalculateLocalSurface in full (Shapes.h:982-1083) through Octtree_Model::Traverse: hitp, uv, du, dv, n, wo -- without vertex
    attributes (fixed uv, dpdu / dpdv, geometric normal), with all of them (interpolated uv / tangent / bitangent / normal), and with a
    rigid transform; plus the function called directly on degenerate triangles."""
    cases = {"plain": (scenes.random_soup(600, 8), {}), "normals_only_hf": (scenes.heightfield(32), {}),
             "attributes": (attribute_model(), {}),
             "attributes_rigid": (attribute_model(500, 6), dict(rigid=_rigid(12, -7, 25, 0.35), precomputed_world=False))}
    cases["plain"] = ([dict(cases["plain"][0][0], normals=None)], {})
    for name, (meshes, kw) in cases.items():
        sc = B.Scene(); sc.set_model(meshes, **kw); sc.build_octree()
        b = sc.model_bounds()
        ctr = tuple((b[:3] + b[3:]) / 2)
        s = sc.traverse_local_surface(_rays(1200, 41, center=ctr, spread=220.0))
        out[f"{name}.found"] = s["found"]
        f = s["found"] > 0
        for key in ("hitp", "uv", "du", "dv", "n", "wo"):
            out[f"{name}.{key}"] = s[key][f]
        sc.close()
    tri, bary, rayd = degenerate_triangles()
    pos = tri.reshape(-1, 3)
    sc = B.Scene(); sc.set_model([dict(positions=pos, normals=None, indices=np.arange(len(pos), dtype=np.uint32).reshape(-1, 3))])
    out["direct.info"] = _canon(sc.local_surface_of(np.zeros(len(tri), np.int32), np.arange(len(tri), dtype=np.int32), bary, rayd))
    sc.close()


GROUPS = dict(integers=_integers, sampling=_sampling, colour=_colour, cameras_shapes=_cameras_shapes, models=_models, tier_a=_tier_a,
              rgb2spec=_rgb2spec, gaussian_filter=_gaussian_filter, sensor=_sensor, tier_b_parts=_tier_b_parts, sat=_sat,
              local_surface=_local_surface)


def load_sensor_inputs(golden):
    """Response curves / illuminants of SENSORS as exported by the reference run (a dict like tests/golden/ref_pin.npz)."""
    _sensor.inputs = [(golden[f"sensor/sensor{i}.curves"], golden[f"sensor/sensor{i}.illum"]) for i in range(len(SENSORS))]


def run(which, groups=None):
    B = _Backend(which)
    out = {}
    for g in (groups or GROUPS):
        sub = {}
        GROUPS[g](B, sub)
        out.update({f"{g}/{k}": np.ascontiguousarray(v) for k, v in sub.items()})
    return out


def compare(a, b):
    """Returns the list of keys on which a and b differ (beyond TOLERANT)."""
    bad = []
    assert set(a) == set(b), sorted(set(a) ^ set(b))
    for k in sorted(a):
        x, y = np.asarray(a[k]), np.asarray(b[k])
        if x.shape != y.shape:
            bad.append((k, "shape", x.shape, y.shape)); continue
        tol = TOLERANT.get(k.split("/", 1)[1])
        if tol is not None:
            if not np.allclose(x, y, rtol=0, atol=tol):
                bad.append((k, "tolerance", float(np.abs(x - y).max())))
            continue
        xb = x.view(np.uint32) if x.dtype == np.float32 else x
        yb = y.view(np.uint32) if y.dtype == np.float32 else y
        if not np.array_equal(xb, yb):
            bad.append((k, "bits", int((xb != yb).sum())))
    return bad
