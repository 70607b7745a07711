"""ctypes binding of oracle/_build/liboracle.so -- TEST INFRASTRUCTURE ONLY.

Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference arm.
The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "_build", "liboracle.so")

_f = C.POINTER(C.c_float)
_i = C.POINTER(C.c_int32)
_u = C.POINTER(C.c_uint32)
_ull = C.POINTER(C.c_ulonglong)
_ll = C.POINTER(C.c_longlong)


def build(force=False):
    srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith((".cpp", ".h"))]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs):
        return LIB_PATH
    subprocess.run(["make", "-C", ORACLE_DIR], check=True, capture_output=True)
    return LIB_PATH


class RenderParams(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("r2c", C.c_float * 16), ("c2w", C.c_float * 16),
                ("lens_radius", C.c_float), ("focal_distance", C.c_float), ("camera_kind", C.c_int),
                ("sampler_kind", C.c_int), ("xs", C.c_int), ("ys", C.c_int), ("jitter", C.c_int), ("seed", C.c_int),
                ("filter_kind", C.c_int), ("filter_rx", C.c_float), ("filter_ry", C.c_float),
                ("mode", C.c_int), ("max_depth", C.c_int), ("rr_depth", C.c_int),
                ("ray_eps", C.c_float), ("shadow_eps", C.c_float), ("albedo", C.c_float * 3),
                ("spp_begin", C.c_int), ("spp_end", C.c_int), ("nthreads", C.c_int), ("pixel_stride", C.c_int),
                ("faithful_overheads", C.c_int), ("filter_sigma", C.c_float), ("light_strategy", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(LIB_PATH)
    L.orc_murmur64a.restype = C.c_uint64
    L.orc_murmur64a.argtypes = [C.c_char_p, C.c_uint64, C.c_uint64]
    L.orc_mixbits.restype = C.c_uint64
    L.orc_mixbits.argtypes = [C.c_uint64]
    L.orc_hash_pixel_seed.restype = C.c_uint64
    L.orc_hash_pixel_seed.argtypes = [C.c_int] * 3
    L.orc_hash_pixel_dim_seed.restype = C.c_uint64
    L.orc_hash_pixel_dim_seed.argtypes = [C.c_int] * 4
    L.orc_permutation_element.restype = C.c_int
    L.orc_permutation_element.argtypes = [C.c_uint32] * 3
    L.orc_pcg32.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_int64, C.c_int, _u, _f]
    L.orc_sampler_sequence.argtypes = [C.c_int] * 9 + [C.c_char_p, _f]
    L.orc_sample_visible.argtypes = [C.c_float, _f, _f]
    L.orc_filter_sample.argtypes = [C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, _f]
    L.orc_concentric_disk.argtypes = [C.c_float, C.c_float, _f]
    L.orc_sample_linear.restype = C.c_float
    L.orc_sample_linear.argtypes = [C.c_float] * 3
    L.orc_cosine_hemisphere.argtypes = [_f, C.c_int, _f, _f]
    L.orc_terminate_secondary.argtypes = [C.c_float, _f]
    L.orc_shape_area.restype = C.c_float
    L.orc_shape_area.argtypes = [C.c_void_p, C.c_int]
    L.orc_triangle_area.argtypes = [C.c_void_p, _i, _i, C.c_int, _f]
    L.orc_gaussian_filter_samples.argtypes = [C.c_float, C.c_float, C.c_float, _f, C.c_int, _f]
    L.orc_gamma.restype = C.c_float
    L.orc_gamma.argtypes = [C.c_int]
    L.orc_difference_of_products.restype = C.c_float
    L.orc_difference_of_products.argtypes = [C.c_float] * 4
    L.orc_dense_table.argtypes = [C.c_int, _f]
    L.orc_illuminant_knots.restype = C.c_int
    L.orc_illuminant_knots.argtypes = [C.c_int, _f, _f, C.c_int]
    L.orc_color_constants.argtypes = [_f, _f, _f, _f]
    L.orc_sigmoid_eval.restype = C.c_float
    L.orc_sigmoid_eval.argtypes = [C.c_float] * 4
    L.orc_set_rgb_table.argtypes = [_f, _f]
    L.orc_set_sensor.argtypes = [_f, _f, _f, _f, C.c_float, _f]
    L.orc_rgb_coeffs.restype = C.c_int
    L.orc_rgb_coeffs.argtypes = [_f, _f]
    L.orc_grey_sigmoid.restype = C.c_int
    L.orc_grey_sigmoid.argtypes = [C.c_float, _f]
    L.orc_camera_matrices.argtypes = [C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, _f, _f, _f, _f, C.c_float, C.c_float, _f, _f]
    L.orc_shape_matrices.argtypes = [_f, _f, _f]
    L.orc_scene_create.restype = C.c_void_p
    L.orc_scene_destroy.argtypes = [C.c_void_p]
    L.orc_scene_set_model.restype = C.c_int
    L.orc_scene_set_model.argtypes = [C.c_void_p, C.c_int, _f, _f, _u, _u, _u, _f, C.c_int, C.c_int, _f, _f, _f, _f]
    L.orc_traverse_local_surface.argtypes = [C.c_void_p, _f, C.c_int, _i, _f]
    L.orc_local_surface_of.argtypes = [C.c_void_p, _i, _i, _f, _f, C.c_int, _f]
    L.orc_scene_build_octree.restype = C.c_int
    L.orc_scene_build_octree.argtypes = [C.c_void_p]
    L.orc_octree_stats.argtypes = [C.c_void_p, _i, _f, _ll]
    L.orc_octree_dump.restype = C.c_longlong
    L.orc_octree_dump.argtypes = [C.c_void_p, _f, _i, _i, _ll, _i, C.c_longlong]
    L.orc_model_backfacing.restype = C.c_int
    L.orc_model_backfacing.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_ubyte)]
    L.orc_model_bounds.argtypes = [C.c_void_p, _f]
    L.orc_scene_add_shape.restype = C.c_int
    L.orc_scene_add_shape.argtypes = [C.c_void_p, C.c_int, _f, _f, C.c_int]
    L.orc_scene_add_spectrum.restype = C.c_int
    L.orc_scene_add_spectrum.argtypes = [C.c_void_p, C.c_int, C.c_float, _f, C.c_int, C.c_char_p, C.c_int]
    L.orc_spectrum_sample.argtypes = [C.c_void_p, C.c_int, _f, _f]
    L.orc_scene_add_light.restype = C.c_int
    L.orc_scene_add_light.argtypes = [C.c_void_p, C.c_int, _f, C.c_int, C.c_float]
    L.orc_scene_add_material.restype = C.c_int
    L.orc_scene_add_material.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int]
    L.orc_scene_set_mesh_materials.argtypes = [C.c_void_p, _i, C.c_int]
    L.orc_scene_light_count.restype = C.c_int
    L.orc_scene_light_count.argtypes = [C.c_void_p]
    L.orc_scene_light_cdf.argtypes = [C.c_void_p, _f, _i]
    L.orc_trace.argtypes = [C.c_void_p, C.c_int, _f, C.c_int, _f, _i, _i, _f, _f, C.c_int, _ull]
    L.orc_traverse_surface.argtypes = [C.c_void_p, _f, C.c_int, _i, _f, _f, _f]
    L.orc_scene_closest.argtypes = [C.c_void_p, _f, C.c_int, _i, _i, _i, _f, _f, _f, _f, _i]
    L.orc_tri_box_overlap.argtypes = [_f, _f, _f, C.c_int, _i]
    L.orc_shape_intersect.argtypes = [C.c_void_p, C.c_int, _f, C.c_int, C.c_float, _i, _f, _f, _f, _f]
    L.orc_render.restype = C.c_double
    L.orc_render.argtypes = [C.c_void_p, C.POINTER(RenderParams), _f, _ull]
    L.orc_eval_samples.argtypes = [C.c_void_p, C.POINTER(RenderParams), _i, _i, C.c_int, _f, _f, _f, _f, _f, _f]
    L.orc_resolve.argtypes = [_f, C.c_int, C.POINTER(C.c_ubyte), _f]
    L.orc_hardware_threads.restype = C.c_int
    _lib = L
    return L


def fp(a):
    return a.ctypes.data_as(_f) if a is not None else None


def ip(a):
    return a.ctypes.data_as(_i) if a is not None else None


def up(a):
    return a.ctypes.data_as(_u) if a is not None else None


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


IDENTITY = np.eye(4, dtype=np.float32)


class OracleScene:
    """Thin object wrapper over the orc_scene_* entry points."""

    def __init__(self):
        self.L = lib()
        self.h = C.c_void_p(self.L.orc_scene_create())
        self.n_meshes = 0

    def close(self):
        if self.h:
            self.L.orc_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_model(self, meshes, rigid=None, precomputed_world=True, cull_backface=False, look_dir=(0, 0, 1)):
        """meshes: list of dicts {positions (nv,3), normals (nv,3)|None, indices (nt,3)}; column-major rigid 4x4."""
        pos = f32(np.concatenate([m["positions"] for m in meshes]))
        has_n = all(m.get("normals") is not None for m in meshes)
        nrm = f32(np.concatenate([m["normals"] for m in meshes])) if has_n else None
        idx = np.ascontiguousarray(np.concatenate([m["indices"].reshape(-1) for m in meshes]), dtype=np.uint32)
        nv = np.array([len(m["positions"]) for m in meshes], dtype=np.uint32)
        nt = np.array([len(m["indices"]) for m in meshes], dtype=np.uint32)
        rg = f32(IDENTITY if rigid is None else rigid).reshape(-1)
        look = f32(look_dir)
        attr = [f32(np.concatenate([m[k] for m in meshes])) if all(m.get(k) is not None for m in meshes) else None for k in ("texcoords", "tangents", "bitangents")]
        self.n_meshes = len(meshes)
        self._keep = (pos, nrm, idx, nv, nt, rg, look)
        return self.L.orc_scene_set_model(self.h, len(meshes), fp(pos), fp(nrm), up(nv), up(idx), up(nt), fp(rg), int(precomputed_world), int(cull_backface), fp(look), fp(attr[0]), fp(attr[1]), fp(attr[2]))

    def build_octree(self):
        return self.L.orc_scene_build_octree(self.h)

    def octree_stats(self):
        ints = np.zeros(7, np.int32); avg = C.c_float(); refs = C.c_longlong()
        self.L.orc_octree_stats(self.h, ip(ints), C.byref(avg), C.byref(refs))
        return dict(nodes=int(ints[0]), real_nodes=int(ints[1]), leaves=int(ints[2]), empty_leaves=int(ints[3]), max_leaf=int(ints[4]),
                    depth=int(ints[5]), avg_leaf=float(avg.value), refs=int(refs.value))

    def octree_dump(self):
        n = self.octree_stats()["nodes"]
        total = self.L.orc_octree_dump(self.h, None, None, None, None, None, 0)
        bounds = np.zeros((n, 6), np.float32); leaf = np.zeros(n, np.int32); child = np.zeros((n, 8), np.int32)
        off = np.zeros(n + 1, np.int64); pairs = np.zeros((max(total, 1), 2), np.int32)
        self.L.orc_octree_dump(self.h, fp(bounds), ip(leaf), ip(child), off.ctypes.data_as(_ll), ip(pairs), total)
        return dict(bounds=bounds, leaf=leaf, child=child, list_off=off, pairs=pairs[:total])

    def backfacing(self, mesh, ntris):
        out = np.zeros(ntris, np.uint8)
        n = self.L.orc_model_backfacing(self.h, mesh, out.ctypes.data_as(C.POINTER(C.c_ubyte)))
        return out[:n]

    def model_bounds(self):
        out = np.zeros(6, np.float32)
        self.L.orc_model_bounds(self.h, fp(out))
        return out

    def add_shape(self, kind, rigid, params, material=0):
        rg = f32(rigid).reshape(-1); pr = f32(list(params) + [0] * (9 - len(params)))
        return self.L.orc_scene_add_shape(self.h, kind, fp(rg), fp(pr), material)

    def add_spectrum(self, kind, c=0.0, interleaved=None, n=0, name=None, normalize=False):
        arr = f32(interleaved) if interleaved is not None else None
        if arr is not None and kind == 1:
            n = arr.size
        return self.L.orc_scene_add_spectrum(self.h, kind, float(c), fp(arr), int(n), name.encode() if name else None, int(normalize))

    def spectrum_sample(self, sid, lambdas):
        lam = f32(lambdas); out = np.zeros(8, np.float32)
        self.L.orc_spectrum_sample(self.h, sid, fp(lam), fp(out))
        return out

    def add_material(self, type=0, refl=-1, eta=-1, k=-1, emit=-1, emit_scale=0.0, two_sided=0, eta_constant=1):
        return self.L.orc_scene_add_material(self.h, type, refl, eta, k, emit, float(emit_scale), two_sided, eta_constant)

    def add_light(self, kind, v, spectrum, scale=1.0):
        """Lights.h:5-8: kind 0 point light (v = position), 1 sun (v = direction towards the light)."""
        return self.L.orc_scene_add_light(self.h, int(kind), fp(f32(v)), int(spectrum), C.c_float(scale))

    def set_mesh_materials(self, ids):
        a = np.ascontiguousarray(ids, dtype=np.int32)
        self.L.orc_scene_set_mesh_materials(self.h, ip(a), len(a))

    def lights(self):
        n = self.L.orc_scene_light_count(self.h)
        cdf = np.zeros(max(n, 1), np.float32); mt = np.zeros((max(n, 1), 2), np.int32)
        if n:
            self.L.orc_scene_light_cdf(self.h, fp(cdf), ip(mt))
        return cdf[:n], mt[:n]

    def trace(self, rays, mode=0, tmax=None, nthreads=1, counters=False):
        rays = f32(rays); n = len(rays)
        mesh = np.full(n, -2, np.int32); tri = np.full(n, -2, np.int32); t = np.zeros(n, np.float32); b = np.zeros((n, 3), np.float32)
        cnt = np.zeros(5, np.uint64) if counters else None
        tm = f32(tmax) if tmax is not None else None
        self.L.orc_trace(self.h, mode, fp(rays), n, fp(tm), ip(mesh), ip(tri), fp(t), fp(b), nthreads,
                         cnt.ctypes.data_as(_ull) if counters else None)
        out = dict(mesh=mesh, tri=tri, t=t, bary=b)
        if counters:
            out["counters"] = dict(rays=int(cnt[0]), nodes=int(cnt[1]), tris=int(cnt[2]), leaves=int(cnt[3]), max_queue=int(cnt[4]))
        return out

    def traverse_surface(self, rays):
        rays = f32(rays); n = len(rays)
        found = np.zeros(n, np.int32); nrm = np.zeros((n, 3), np.float32); hp = np.zeros((n, 3), np.float32); uv = np.zeros((n, 2), np.float32)
        self.L.orc_traverse_surface(self.h, fp(rays), n, ip(found), fp(nrm), fp(hp), fp(uv))
        return dict(found=found, n=nrm, hitp=hp, uv=uv)

    def traverse_local_surface(self, rays):
        """Octtree_Model::Traverse -> LocalSurfaceInfo: found, hitp, uv, du, dv, n, wo."""
        rays = f32(rays); n = len(rays)
        found = np.zeros(n, np.int32); info = np.zeros((n, 17), np.float32)
        self.L.orc_traverse_local_surface(self.h, fp(rays), n, ip(found), fp(info))
        return dict(found=found, hitp=info[:, 0:3], uv=info[:, 3:5], du=info[:, 5:8], dv=info[:, 8:11], n=info[:, 11:14], wo=info[:, 14:17])

    def local_surface_of(self, mesh, tri, bary, rayd):
        """Triangle(mesh, tri).CalculateLocalSurface(bary, rayd) -> (n, 17): hitp uv du dv n wo."""
        mesh = np.ascontiguousarray(mesh, np.int32); tri = np.ascontiguousarray(tri, np.int32); bary = f32(bary); rayd = f32(rayd)
        info = np.zeros((len(mesh), 17), np.float32)
        self.L.orc_local_surface_of(self.h, ip(mesh), ip(tri), fp(bary), fp(rayd), len(mesh), fp(info))
        return info

    def scene_closest(self, rays):
        rays = f32(rays); n = len(rays)
        kind = np.zeros(n, np.int32); id0 = np.zeros(n, np.int32); id1 = np.zeros(n, np.int32); t = np.zeros(n, np.float32)
        p = np.zeros((n, 3), np.float32); ns = np.zeros((n, 3), np.float32); ng = np.zeros((n, 3), np.float32); bs = np.zeros(n, np.int32)
        self.L.orc_scene_closest(self.h, fp(rays), n, ip(kind), ip(id0), ip(id1), fp(t), fp(p), fp(ns), fp(ng), ip(bs))
        return dict(kind=kind, id0=id0, id1=id1, t=t, p=p, ns=ns, ng=ng, backside=bs)

    def shape_intersect(self, shape, rays, tmax=np.finfo(np.float32).max):
        rays = f32(rays); n = len(rays)
        found = np.zeros(n, np.int32); t = np.zeros(n, np.float32); hp = np.zeros((n, 3), np.float32); nrm = np.zeros((n, 3), np.float32); uv = np.zeros((n, 2), np.float32)
        self.L.orc_shape_intersect(self.h, shape, fp(rays), n, float(tmax), ip(found), fp(t), fp(hp), fp(nrm), fp(uv))
        fr = np.zeros((n, 9), np.float32)
        self.L.orc_shape_frame.argtypes = [C.c_void_p, C.c_int, _f, C.c_int, C.c_float, _f]
        self.L.orc_shape_frame(self.h, shape, fp(rays), n, float(tmax), fp(fr))
        return dict(found=found, t=t, hitp=hp, n=nrm, uv=uv, du=fr[:, 0:3], dv=fr[:, 3:6], wo=fr[:, 6:9])

    def render(self, params, film=None, counters=False):
        npix = params.width * params.height
        film = np.zeros((npix, 4), np.float32) if film is None else film
        cnt = np.zeros(9, np.uint64) if counters else None
        secs = self.L.orc_render(self.h, C.byref(params), fp(film), cnt.ctypes.data_as(_ull) if counters else None)
        out = dict(film=film, seconds=secs)
        if counters:
            keys = ["rays", "nodes", "tris", "leaves", "max_queue", "paths", "closest_rays", "shadow_rays", "depth_sum"]
            out["counters"] = {k: int(v) for k, v in zip(keys, cnt)}
        return out

    def eval_samples(self, params, pixel_ids, indices):
        pid = np.ascontiguousarray(pixel_ids, np.int32); idx = np.ascontiguousarray(indices, np.int32); n = len(pid)
        ray = np.zeros((n, 6), np.float32); lam = np.zeros((n, 8), np.float32); pdf = np.zeros((n, 8), np.float32)
        L8 = np.zeros((n, 8), np.float32); rgb = np.zeros((n, 3), np.float32); w = np.zeros(n, np.float32)
        self.L.orc_eval_samples(self.h, C.byref(params), ip(pid), ip(idx), n, fp(ray), fp(lam), fp(pdf), fp(L8), fp(rgb), fp(w))
        return dict(ray=ray, lam=lam, pdf=pdf, L=L8, rgb=rgb, weight=w)


def make_params(width, height, r2c, c2w, *, lens_radius=0.0, focal_distance=0.0, camera_kind=0, sampler_kind=1, xs=4, ys=4, jitter=1,
                seed=0, filter_kind=0, filter_r=(0.5, 0.5), mode=0, max_depth=5, rr_depth=0, ray_eps=1e-2, shadow_eps=1e-3,
                albedo=(0.5, 0.5, 0.5), spp_begin=0, spp_end=1, nthreads=1, pixel_stride=1, faithful=0, filter_sigma=0.0, light_strategy=0):
    p = RenderParams()
    p.width, p.height = width, height
    p.r2c[:] = list(f32(r2c).reshape(-1)); p.c2w[:] = list(f32(c2w).reshape(-1))
    p.lens_radius, p.focal_distance, p.camera_kind = lens_radius, focal_distance, camera_kind
    p.sampler_kind, p.xs, p.ys, p.jitter, p.seed = sampler_kind, xs, ys, jitter, seed
    p.filter_kind, p.filter_rx, p.filter_ry = filter_kind, filter_r[0], filter_r[1]
    p.mode, p.max_depth, p.rr_depth, p.ray_eps, p.shadow_eps = mode, max_depth, rr_depth, ray_eps, shadow_eps
    p.albedo[:] = list(albedo)
    p.spp_begin, p.spp_end, p.nthreads, p.pixel_stride, p.faithful_overheads = spp_begin, spp_end, nthreads, pixel_stride, faithful
    p.filter_sigma = filter_sigma
    p.light_strategy = light_strategy
    return p


def camera_matrices(kind, near, far, sw, sh, fov, pos, look, right, up, resx, resy):
    r2c = np.zeros(16, np.float32); c2w = np.zeros(16, np.float32)
    lib().orc_camera_matrices(kind, near, far, sw, sh, fov, fp(f32(pos)), fp(f32(look)), fp(f32(right)), fp(f32(up)), resx, resy, fp(r2c), fp(c2w))
    return r2c, c2w


def resolve(film):
    film = f32(film); n = len(film)
    rgb8 = np.zeros((n, 3), np.uint8); rgbf = np.zeros((n, 3), np.float32)
    lib().orc_resolve(fp(film), n, rgb8.ctypes.data_as(C.POINTER(C.c_ubyte)), fp(rgbf))
    return rgb8, rgbf
