"""ctypes binding of oracle/_ref/libcrt_ref.so -- TEST INFRASTRUCTURE ONLY.

libcrt_ref.so is the REFERENCE'S OWN code (RayTracer/*.h, ThirdParty/pbrv4/*, AABB_triangle_Moller.h) compiled unmodified
from /root/reference by `make -C oracle ref` (oracle/ref_harness.cpp, oracle/refshim/).  It exists only where
/root/reference exists (this container) or where the prebuilt .so travelled (the GPU box); tests that need it skip
otherwise and fall back to tests/golden/ref_pin.npz, which tools/make_ref_golden.py generated from it.
The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "_ref", "libcrt_ref.so")
REFERENCE_ROOT = "/root/reference"

_f = C.POINTER(C.c_float)
_i = C.POINTER(C.c_int32)
_u = C.POINTER(C.c_uint32)
_ll = C.POINTER(C.c_longlong)
_ub = C.POINTER(C.c_ubyte)


def available():
    """True if the compiled reference can be loaded (prebuilt, or buildable because /root/reference is present)."""
    return os.path.exists(LIB_PATH) or os.path.isdir(REFERENCE_ROOT)


def build(force=False):
    if os.path.isdir(REFERENCE_ROOT):
        srcs = [os.path.join(ORACLE_DIR, "ref_harness.cpp"), os.path.join(ORACLE_DIR, "ref_harness_private.cpp"), os.path.join(ORACLE_DIR, "refshim", "glm", "glm.hpp"),
                os.path.join(ORACLE_DIR, "refshim", "ref_prelude.h")]
        stale = not os.path.exists(LIB_PATH) or any(os.path.getmtime(LIB_PATH) < os.path.getmtime(s) for s in srcs)
        if force or stale:
            subprocess.run(["make", "-C", ORACLE_DIR, "ref"], check=True, capture_output=True)
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("oracle/_ref/libcrt_ref.so is absent and /root/reference is not here to build it from")
    return LIB_PATH


class RenderParams(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("camera_kind", C.c_int),
                ("near_", C.c_float), ("far_", C.c_float), ("sensor_w", C.c_float), ("sensor_h", C.c_float), ("fov", C.c_float),
                ("pos", C.c_float * 3), ("look", C.c_float * 3), ("right", C.c_float * 3), ("up", C.c_float * 3),
                ("lens_radius", C.c_float), ("focal_distance", C.c_float),
                ("sampler_kind", C.c_int), ("xs", C.c_int), ("ys", C.c_int), ("jitter", C.c_int), ("seed", C.c_int),
                ("filter_rx", C.c_float), ("filter_ry", C.c_float), ("albedo", C.c_float * 3),
                ("spp_begin", C.c_int), ("spp_end", C.c_int), ("nthreads", C.c_int), ("pixel_stride", C.c_int),
                ("filter_kind", C.c_int), ("filter_sigma", C.c_float)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(LIB_PATH)
    L.ref_describe.restype = C.c_char_p
    L.ref_murmur64a.restype = C.c_uint64
    L.ref_murmur64a.argtypes = [C.c_char_p, C.c_uint64, C.c_uint64]
    for name in ("ref_mixbits", "ref_helper_mixbits"):
        getattr(L, name).restype = C.c_uint64
        getattr(L, name).argtypes = [C.c_uint64]
    L.ref_hash_pixel_seed.restype = C.c_uint64
    L.ref_hash_pixel_seed.argtypes = [C.c_int] * 3
    L.ref_hash_pixel_dim_seed.restype = C.c_uint64
    L.ref_hash_pixel_dim_seed.argtypes = [C.c_int] * 4
    L.ref_permutation_element.restype = C.c_int
    L.ref_permutation_element.argtypes = [C.c_uint32] * 3
    L.ref_pcg32.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_int64, C.c_int, _u, _f]
    L.ref_sampler_sequence.argtypes = [C.c_int] * 9 + [C.c_char_p, _f]
    L.ref_sample_visible.argtypes = [C.c_float, _f, _f]
    L.ref_filter_sample.restype = C.c_int
    L.ref_filter_sample.argtypes = [C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, _f]
    L.ref_concentric_disk.argtypes = [C.c_float, C.c_float, _f]
    L.ref_sample_linear.restype = C.c_float
    L.ref_sample_linear.argtypes = [C.c_float] * 3
    L.ref_cosine_hemisphere.argtypes = [_f, C.c_int, _f, _f]
    L.ref_terminate_secondary.argtypes = [C.c_float, _f]
    L.ref_shape_area.restype = C.c_float
    L.ref_shape_area.argtypes = [C.c_void_p, C.c_int]
    L.ref_triangle_area.argtypes = [C.c_void_p, _i, _i, C.c_int, _f]
    L.ref_gaussian_filter_samples.argtypes = [C.c_float, C.c_float, C.c_float, _f, C.c_int, _f]
    L.ref_gamma.restype = C.c_float
    L.ref_gamma.argtypes = [C.c_int]
    L.ref_difference_of_products.restype = C.c_float
    L.ref_difference_of_products.argtypes = [C.c_float] * 4
    L.ref_dense_table.argtypes = [C.c_int, _f]
    L.ref_named_spectrum_query.restype = C.c_int
    L.ref_named_spectrum_query.argtypes = [C.c_char_p, _f, C.c_int, _f]
    L.ref_interleaved_spectrum_query.argtypes = [_f, C.c_int, C.c_int, _f, C.c_int, _f]
    L.ref_color_constants.argtypes = [_f, _f, _f, _f]
    L.ref_sigmoid_eval.restype = C.c_float
    L.ref_sigmoid_eval.argtypes = [C.c_float] * 4
    L.ref_grey_rgb_spectrum_sample.argtypes = [C.c_int, C.c_float, C.c_float, _f, _f]
    L.ref_to_sensor_rgb.argtypes = [C.c_float, _f, _f]
    L.ref_set_rgb_table.argtypes = [_f, _f]
    L.ref_set_sensor.restype = C.c_int
    L.ref_set_sensor.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_float, _f, _f, _f]
    L.ref_sensor_rgb.argtypes = [C.c_float, _f, _f]
    L.ref_rgb_albedo_query.argtypes = [_f, _f, C.c_int, _f]
    L.ref_rgb_spectrum_sample.argtypes = [C.c_int, _f, C.c_float, _f, _f]
    cam = [C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, _f, _f, _f, _f, C.c_float, C.c_float]
    L.ref_camera_matrices.argtypes = cam + [_f, _f]
    L.ref_camera_rays.argtypes = cam + [C.c_float, C.c_float, _f, C.c_int] + [C.c_int] * 6 + [_f]
    L.ref_shape_matrices.argtypes = [_f, _f, _f]
    L.ref_scene_create.restype = C.c_void_p
    L.ref_scene_destroy.argtypes = [C.c_void_p]
    L.ref_scene_set_model.restype = C.c_int
    L.ref_scene_set_model.argtypes = [C.c_void_p, C.c_int, _f, _f, _u, _u, _u, _f, C.c_int, C.c_int, _f, _f, _f, _f]
    L.ref_traverse_local_surface.argtypes = [C.c_void_p, _f, C.c_int, _i, _f]
    L.ref_local_surface_of.argtypes = [C.c_void_p, _i, _i, _f, _f, C.c_int, _f]
    L.ref_scene_build_octree.restype = C.c_int
    L.ref_scene_build_octree.argtypes = [C.c_void_p]
    L.ref_octree_dump.restype = C.c_longlong
    L.ref_octree_dump.argtypes = [C.c_void_p, _f, _i, _i, _ll, _i, C.c_longlong]
    L.ref_model_bounds.argtypes = [C.c_void_p, _f]
    L.ref_scene_tri_model.restype = C.c_void_p
    L.ref_scene_tri_model.argtypes = [C.c_void_p]
    L.ref_scene_octree.restype = C.c_void_p
    L.ref_scene_octree.argtypes = [C.c_void_p]
    L.ref_private_backfacing.restype = C.c_int
    L.ref_private_backfacing.argtypes = [C.c_void_p, C.c_int, _ub]
    L.ref_private_octree_stats.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 4
    L.ref_scene_add_shape.restype = C.c_int
    L.ref_scene_add_shape.argtypes = [C.c_void_p, C.c_int, _f, _f]
    L.ref_slab_test.argtypes = [_f, _f, _f, C.c_int, _i]
    L.ref_triangle_intersect.argtypes = [C.c_void_p, _i, _i, _f, _f, C.c_int, _i, _f, _f]
    L.ref_brute_force.argtypes = [C.c_void_p, _f, C.c_int, _i, _i, _f, _f]
    L.ref_traverse_surface.argtypes = [C.c_void_p, _f, C.c_int, C.c_int, _i, _f, _f, _f, _f]
    L.ref_surface_of.argtypes = [C.c_void_p, _i, _i, _f, C.c_int, _i, _f, _f, _f, _f]
    L.ref_shape_intersect.argtypes = [C.c_void_p, C.c_int, _f, C.c_int, C.c_float, _i, _f, _f, _f, _f]
    L.ref_tri_box_overlap.argtypes = [_f, _f, _f, C.c_int, _i]
    L.ref_eval_samples.argtypes = [C.c_void_p, C.POINTER(RenderParams), _i, _i, C.c_int, _f, _f, _f, _f, _f, _f]
    L.ref_render_tier_a.argtypes = [C.c_void_p, C.POINTER(RenderParams), _f]
    L.ref_resolve.argtypes = [_f, C.c_int, _ub, _f]
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


def _quiet(fn, *a):
    """The reference prints from CreateOcttree / Init (Octtree_Model.h:48-51, color.cpp:162); keep test output clean."""
    import sys
    sys.stdout.flush()
    saved = os.dup(1)
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 1)
    try:
        return fn(*a)
    finally:
        os.dup2(saved, 1)
        os.close(devnull); os.close(saved)


class RefScene:
    """Same surface as oracle_lib.OracleScene for the probes the compiled reference offers."""

    def __init__(self):
        self.L = lib()
        self.h = C.c_void_p(_quiet(self.L.ref_scene_create))

    def close(self):
        if self.h:
            self.L.ref_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_model(self, meshes, rigid=None, precomputed_world=True, cull_backface=False, look_dir=(0, 0, 1)):
        pos = f32(np.concatenate([m["positions"] for m in meshes]))
        has_n = all(m.get("normals") is not None for m in meshes)
        nrm = f32(np.concatenate([m["normals"] for m in meshes])) if has_n else None
        idx = np.ascontiguousarray(np.concatenate([m["indices"].reshape(-1) for m in meshes]), dtype=np.uint32)
        nv = np.array([len(m["positions"]) for m in meshes], dtype=np.uint32)
        nt = np.array([len(m["indices"]) for m in meshes], dtype=np.uint32)
        rg = f32(IDENTITY if rigid is None else rigid).reshape(-1)
        look = f32(look_dir)
        attr = [f32(np.concatenate([m[k] for m in meshes])) if all(m.get(k) is not None for m in meshes) else None for k in ("texcoords", "tangents", "bitangents")]
        return self.L.ref_scene_set_model(self.h, len(meshes), fp(pos), fp(nrm), up(nv), up(idx), up(nt), fp(rg), int(precomputed_world), int(cull_backface), fp(look), fp(attr[0]), fp(attr[1]), fp(attr[2]))

    def build_octree(self):
        self.n_nodes = _quiet(self.L.ref_scene_build_octree, self.h)
        return self.n_nodes

    def octree_dump(self):
        n = self.n_nodes
        total = self.L.ref_octree_dump(self.h, None, None, None, None, None, 0)
        bounds = np.zeros((n, 6), np.float32); leaf = np.zeros(n, np.int32); child = np.zeros((n, 8), np.int32)
        off = np.zeros(n + 1, np.int64); pairs = np.zeros((max(total, 1), 2), np.int32)
        self.L.ref_octree_dump(self.h, fp(bounds), ip(leaf), ip(child), off.ctypes.data_as(_ll), ip(pairs), total)
        return dict(bounds=bounds, leaf=leaf, child=child, list_off=off, pairs=pairs[:total])

    def model_bounds(self):
        out = np.zeros(6, np.float32)
        self.L.ref_model_bounds(self.h, fp(out))
        return out

    def backfacing(self, mesh, ntris):
        """TriModel::back_facing[mesh] (private; read by oracle/ref_harness_private.cpp)."""
        out = np.zeros(ntris, np.uint8)
        n = self.L.ref_private_backfacing(C.c_void_p(self.L.ref_scene_tri_model(self.h)), mesh, out.ctypes.data_as(_ub))
        return out[:n]

    def octree_stats(self):
        v = [C.c_int() for _ in range(4)]
        self.L.ref_private_octree_stats(C.c_void_p(self.L.ref_scene_octree(self.h)), *[C.byref(x) for x in v])
        return dict(nodes=v[0].value, leaves=v[1].value, empty_leaves=v[2].value, max_leaf=v[3].value)

    def add_shape(self, kind, rigid, params):
        rg = f32(rigid).reshape(-1); pr = f32(list(params) + [0] * (9 - len(params)))
        return self.L.ref_scene_add_shape(self.h, kind, fp(rg), fp(pr))

    def triangle_intersect(self, mesh, tri, rays, tmax):
        rays = f32(rays); n = len(rays)
        mesh = np.ascontiguousarray(mesh, np.int32); tri = np.ascontiguousarray(tri, np.int32); tm = f32(tmax)
        found = np.zeros(n, np.int32); t = np.zeros(n, np.float32); b = np.zeros((n, 3), np.float32)
        self.L.ref_triangle_intersect(self.h, ip(mesh), ip(tri), fp(rays), fp(tm), n, ip(found), fp(t), fp(b))
        return dict(found=found, t=t, bary=b)

    def brute_force(self, rays):
        rays = f32(rays); n = len(rays)
        mesh = np.full(n, -2, np.int32); tri = np.full(n, -2, np.int32); t = np.zeros(n, np.float32); b = np.zeros((n, 3), np.float32)
        self.L.ref_brute_force(self.h, fp(rays), n, ip(mesh), ip(tri), fp(t), fp(b))
        return dict(mesh=mesh, tri=tri, t=t, bary=b)

    def _surface(self, fn, rays, *pre):
        rays = f32(rays); n = len(rays)
        found = np.zeros(n, np.int32); th = np.zeros(n, np.float32); nrm = np.zeros((n, 3), np.float32); hp = np.zeros((n, 3), np.float32); uv = np.zeros((n, 2), np.float32)
        fn(rays, n, found, th, nrm, hp, uv)
        return dict(found=found, t=th, n=nrm, hitp=hp, uv=uv)

    def traverse_surface(self, rays, nthreads=1):
        return self._surface(lambda r, n, f, th, nr, hp, uv: self.L.ref_traverse_surface(self.h, fp(r), n, nthreads, ip(f), fp(th), fp(nr), fp(hp), fp(uv)), rays)

    def surface_of(self, mesh, tri, rays):
        mesh = np.ascontiguousarray(mesh, np.int32); tri = np.ascontiguousarray(tri, np.int32)
        return self._surface(lambda r, n, f, th, nr, hp, uv: self.L.ref_surface_of(self.h, ip(mesh), ip(tri), fp(r), n, ip(f), fp(th), fp(nr), fp(hp), fp(uv)), rays)

    def traverse_local_surface(self, rays):
        """Octtree_Model::Traverse -> LocalSurfaceInfo: found, hitp, uv, du, dv, n, wo."""
        rays = f32(rays); n = len(rays)
        found = np.zeros(n, np.int32); info = np.zeros((n, 17), np.float32)
        self.L.ref_traverse_local_surface(self.h, fp(rays), n, ip(found), fp(info))
        return dict(found=found, hitp=info[:, 0:3], uv=info[:, 3:5], du=info[:, 5:8], dv=info[:, 8:11], n=info[:, 11:14], wo=info[:, 14:17])

    def local_surface_of(self, mesh, tri, bary, rayd):
        """Triangle(mesh, tri).CalculateLocalSurface(bary, rayd) -> (n, 17): hitp uv du dv n wo."""
        mesh = np.ascontiguousarray(mesh, np.int32); tri = np.ascontiguousarray(tri, np.int32); bary = f32(bary); rayd = f32(rayd)
        info = np.zeros((len(mesh), 17), np.float32)
        self.L.ref_local_surface_of(self.h, ip(mesh), ip(tri), fp(bary), fp(rayd), len(mesh), fp(info))
        return info

    def shape_intersect(self, shape, rays, tmax=np.finfo(np.float32).max):
        rays = f32(rays); n = len(rays)
        found = np.zeros(n, np.int32); t = np.zeros(n, np.float32); hp = np.zeros((n, 3), np.float32); nrm = np.zeros((n, 3), np.float32); uv = np.zeros((n, 2), np.float32)
        self.L.ref_shape_intersect(self.h, shape, fp(rays), n, float(tmax), ip(found), fp(t), fp(hp), fp(nrm), fp(uv))
        fr = np.zeros((n, 9), np.float32)
        self.L.ref_shape_frame.argtypes = [C.c_void_p, C.c_int, _f, C.c_int, C.c_float, _f]
        self.L.ref_shape_frame(self.h, shape, fp(rays), n, float(tmax), fp(fr))
        return dict(found=found, t=t, hitp=hp, n=nrm, uv=uv, du=fr[:, 0:3], dv=fr[:, 3:6], wo=fr[:, 6:9])

    def eval_samples(self, params, pixel_ids, indices):
        pid = np.ascontiguousarray(pixel_ids, np.int32); idx = np.ascontiguousarray(indices, np.int32); n = len(pid)
        ray = np.zeros((n, 6), np.float32); lam = np.zeros((n, 8), np.float32); pdf = np.zeros((n, 8), np.float32)
        L8 = np.zeros((n, 8), np.float32); rgb = np.zeros((n, 3), np.float32); w = np.zeros(n, np.float32)
        self.L.ref_eval_samples(self.h, C.byref(params), ip(pid), ip(idx), n, fp(ray), fp(lam), fp(pdf), fp(L8), fp(rgb), fp(w))
        return dict(ray=ray, lam=lam, pdf=pdf, L=L8, rgb=rgb, weight=w)

    def render_tier_a(self, params, film=None):
        npix = params.width * params.height
        film = np.zeros((npix, 4), np.float32) if film is None else film
        self.L.ref_render_tier_a(self.h, C.byref(params), fp(film))
        return film


def make_params(width, height, *, camera_kind=0, near=1.0, far=1000.0, sensor=(0.0, 0.0), fov=45.0, pos=(0, 0, 0), look=(0, 0, 1),
                right=(1, 0, 0), up=(0, 1, 0), lens_radius=0.0, focal_distance=0.0, sampler_kind=1, xs=4, ys=4, jitter=1, seed=0,
                filter_r=(0.5, 0.5), albedo=(0.5, 0.5, 0.5), spp_begin=0, spp_end=1, nthreads=1, pixel_stride=1, filter_kind=0, filter_sigma=0.0):
    p = RenderParams()
    p.width, p.height, p.camera_kind = width, height, camera_kind
    p.near_, p.far_, p.sensor_w, p.sensor_h, p.fov = near, far, sensor[0], sensor[1], fov
    p.pos[:] = list(pos); p.look[:] = list(look); p.right[:] = list(right); p.up[:] = list(up)
    p.lens_radius, p.focal_distance = lens_radius, focal_distance
    p.sampler_kind, p.xs, p.ys, p.jitter, p.seed = sampler_kind, xs, ys, jitter, seed
    p.filter_rx, p.filter_ry = filter_r
    p.albedo[:] = list(albedo)
    p.spp_begin, p.spp_end, p.nthreads, p.pixel_stride = spp_begin, spp_end, nthreads, pixel_stride
    p.filter_kind, p.filter_sigma = filter_kind, filter_sigma
    return p


def camera_matrices(kind, near, far, sw, sh, fov, pos, look, right, up, resx, resy):
    r2c = np.zeros(16, np.float32); c2w = np.zeros(16, np.float32)
    lib().ref_camera_matrices(kind, near, far, sw, sh, fov, fp(f32(pos)), fp(f32(look)), fp(f32(right)), fp(f32(up)), resx, resy, fp(r2c), fp(c2w))
    return r2c, c2w


def camera_rays(kind, near, far, sw, sh, fov, pos, look, right, up, resx, resy, film_xy, lens_radius=0.0, focal_distance=0.0,
                xs=4, ys=4, jitter=1, seed=0, index=0, dim=3):
    xy = f32(film_xy); n = len(xy); out = np.zeros((n, 6), np.float32)
    lib().ref_camera_rays(kind, near, far, sw, sh, fov, fp(f32(pos)), fp(f32(look)), fp(f32(right)), fp(f32(up)), resx, resy,
                          lens_radius, focal_distance, fp(xy), n, xs, ys, jitter, seed, index, dim, fp(out))
    return out


def resolve(film):
    film = f32(film); n = len(film)
    rgb8 = np.zeros((n, 3), np.uint8); rgbf = np.zeros((n, 3), np.float32)
    _quiet(lib().ref_resolve, fp(film), n, rgb8.ctypes.data_as(_ub), fp(rgbf))
    return rgb8, rgbf
