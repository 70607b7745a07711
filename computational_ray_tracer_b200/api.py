"""Python mirror of the reference's object model on top of the C ABI (tests + bench harness).

Names follow the reference: TriModel / Octtree_Model (RayTracer/Shapes.h, Octtree_Model.h), PerspectiveCamera
(Cameras.h), Film (Film.h), the samplers and filters (ThirdParty/pbrv4).  The C++ facade in include/crt/ is the
drop-in for C++ callers; this module is the same thin layer for Python.
"""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import MeshDesc, OctreeStats, RenderConfig, RenderStats, check, f32p, i32p, u8p, u32p

HAS_PATH_INTEGRATOR = True          # crt_render mode 1 (wavefront path integrator with NEE) is built in
DEFAULT_TRACE_MODE = 3          # ordered traversal (one ray per lane) + exact BFS re-trace of order-sensitive rays (identical results)
FLT_MAX = float(np.finfo(np.float32).max)
IDENTITY = np.eye(4, dtype=np.float32)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _fp(a):
    return a.ctypes.data_as(f32p) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(i32p) if a is not None else None


class MeshSet:
    """MeshCache::Model (RayTracer/AssetManager.h:20-47): list of meshes {positions (nv,3), normals (nv,3) | None, indices (nt,3),
    and optionally texcoords (nv,2), tangents (nv,3), bitangents (nv,3)}."""

    def __init__(self, meshes):
        self.meshes = []
        self.attributes = []
        for m in meshes:
            pos = _f32(m["positions"])
            nrm = _f32(m["normals"]) if m.get("normals") is not None else None
            idx = np.ascontiguousarray(m["indices"], dtype=np.uint32).reshape(-1, 3)
            self.meshes.append((pos, nrm, idx))
            self.attributes.append(tuple(_f32(m[k]) if m.get(k) is not None else None for k in ("texcoords", "tangents", "bitangents")))
        self.descs = (MeshDesc * len(self.meshes))()
        for d, (pos, nrm, idx), (uv, tan, bitan) in zip(self.descs, self.meshes, self.attributes):
            d.positions = _fp(pos)
            d.normals = _fp(nrm)
            d.n_vertices = len(pos)
            d.indices = idx.ctypes.data_as(u32p)
            d.n_triangles = len(idx)
            d.texcoords, d.tangents, d.bitangents = _fp(uv), _fp(tan), _fp(bitan)

    def __len__(self):
        return len(self.meshes)

    @property
    def n_triangles(self):
        return sum(len(m[2]) for m in self.meshes)


def load_obj(path):
    """MeshCache::LoadMeshFromFile for Wavefront OBJ (RayTracer/AssetManager.cpp:8-25): list of mesh dicts for MeshSet."""
    L = _capi.load()
    h = C.c_void_p()
    check(L.crt_obj_load(str(path).encode(), C.byref(h)))
    try:
        meshes = []
        for i in range(L.crt_obj_mesh_count(h)):
            nv = C.c_uint32(); nt = C.c_uint32(); name = C.create_string_buffer(256)
            check(L.crt_obj_mesh_info(h, i, C.byref(nv), C.byref(nt), name, 256))
            pos = np.zeros((nv.value, 3), np.float32); nrm = np.zeros((nv.value, 3), np.float32); idx = np.zeros((nt.value, 3), np.uint32)
            check(L.crt_obj_mesh_copy(h, i, _fp(pos), _fp(nrm), idx.ctypes.data_as(u32p)))
            uv = np.zeros((nv.value, 2), np.float32); tan = np.zeros((nv.value, 3), np.float32); bitan = np.zeros((nv.value, 3), np.float32)
            avail = C.c_int()
            check(L.crt_obj_mesh_attributes(h, i, _fp(uv), _fp(tan), _fp(bitan), C.byref(avail)))
            m = dict(name=name.value.decode(errors="replace"), positions=pos, normals=nrm, indices=idx)
            if avail.value:        # Triangle::vertex_available (Shapes.h:917-924): texcoords / tangents / bitangents only when the file has vt
                m.update(texcoords=uv, tangents=tan, bitangents=bitan)
            meshes.append(m)
        return meshes
    finally:
        L.crt_obj_destroy(h)


def shape_matrices(rigid):
    o2r = np.zeros(16, np.float32); r2o = np.zeros(16, np.float32)
    check(_capi.load().crt_shape_matrices(_fp(_f32(rigid).reshape(-1)), _fp(o2r), _fp(r2o)))
    return o2r, r2o


BUILD_INCREMENTAL, BUILD_TOPDOWN, BUILD_GPU = 0, 1, 2


class Octtree_Model:
    """Host octree over a TriModel (RayTracer/Octtree_Model.h): CreateOcttree happens in the constructor.
    algorithm: BUILD_INCREMENTAL = the reference's insertion loop (GetNode ids in reference creation order);
    BUILD_TOPDOWN = the same tree built level by level (same flattened layout, different GetNode numbering);
    BUILD_GPU = that level-by-level construction on the device (needs ctx)."""

    def __init__(self, meshes: MeshSet, rigid=None, precomputed_world=True, algorithm=BUILD_INCREMENTAL, ctx=None):
        self.L = _capi.load()
        if algorithm == BUILD_GPU:
            if ctx is None:
                raise ValueError("BUILD_GPU needs a Context")
            self.meshes = meshes
            self.precomputed_world = bool(precomputed_world)
            self.o2r, _ = shape_matrices(IDENTITY if rigid is None else rigid)
            self.h = C.c_void_p()
            check(self.L.crt_octree_build_gpu(ctx.h, meshes.descs, len(meshes), _fp(self.o2r), int(precomputed_world), C.byref(self.h)))
            return
        check(self.L.crt_octree_set_build_algorithm(algorithm))
        self.meshes = meshes
        self.precomputed_world = bool(precomputed_world)
        self.o2r, _ = shape_matrices(IDENTITY if rigid is None else rigid)
        self.h = C.c_void_p()
        check(self.L.crt_octree_build(meshes.descs, len(meshes), _fp(self.o2r), int(precomputed_world), C.byref(self.h)))

    def close(self):
        if self.h:
            self.L.crt_octree_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def getTreeSize(self):
        return self.L.crt_octree_node_count(self.h)

    def flat(self):
        """The flattened device layout (nothing culled) as numpy arrays."""
        sizes = np.zeros(5, np.uint64)
        check(self.L.crt_octree_flat_sizes(self.h, sizes.ctypes.data_as(_capi.u64p)))
        n = [int(x) for x in sizes]
        out = dict(nodes=np.zeros(n[0], np.float32), leaf_refs=np.zeros(n[1], np.uint32), node_tight=np.zeros(n[2], np.float32),
                   pk_boxes=np.zeros(n[3], np.float32), pk_refs=np.zeros(n[4], np.uint32))
        check(self.L.crt_octree_flat_copy(self.h, _fp(out["nodes"]), out["leaf_refs"].ctypes.data_as(u32p), _fp(out["node_tight"]),
                                          _fp(out["pk_boxes"]), out["pk_refs"].ctypes.data_as(u32p)))
        return out

    def stats(self):
        s = OctreeStats()
        check(self.L.crt_octree_get_stats(self.h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in OctreeStats._fields_}

    def GetNode(self, i, cap=4096):
        b = np.zeros(6, np.float32); leaf = C.c_int32(); child = np.zeros(8, np.int32); pairs = np.zeros((cap, 2), np.int32); n = C.c_int32()
        check(self.L.crt_octree_get_node(self.h, i, _fp(b), C.byref(leaf), _ip(child), _ip(pairs), cap, C.byref(n)))
        return dict(bounds=b, leaf=bool(leaf.value), child=child, pairs=pairs[:min(n.value, cap)], n_pairs=n.value)

    def dump(self):
        n = self.getTreeSize()
        bounds = np.zeros((n, 6), np.float32); leaf = np.zeros(n, np.int32); child = np.zeros((n, 8), np.int32)
        off = np.zeros(n + 1, np.int64); chunks = []
        for i in range(n):
            nd = self.GetNode(i)
            if nd["n_pairs"] > len(nd["pairs"]):
                nd = self.GetNode(i, nd["n_pairs"])
            bounds[i] = nd["bounds"]; leaf[i] = nd["leaf"]; child[i] = nd["child"]
            off[i + 1] = off[i] + nd["n_pairs"]; chunks.append(nd["pairs"].copy())
        pairs = np.concatenate(chunks) if chunks else np.zeros((0, 2), np.int32)
        return dict(bounds=bounds, leaf=leaf, child=child, list_off=off, pairs=pairs)

    def compute_backface(self, look_dir=(0, 0, 1)):
        """TriModel::ComputeBackFace (Shapes.h:1339-1380) -> list of uint8 arrays, one per mesh."""
        out = []
        look = _f32(look_dir)
        for i, (pos, nrm, idx) in enumerate(self.meshes.meshes):
            bits = np.zeros(len(idx), np.uint8)
            check(self.L.crt_model_compute_backface(C.byref(self.meshes.descs[i]), _fp(look), _fp(self.o2r), int(self.precomputed_world), bits.ctypes.data_as(u8p)))
            out.append(bits)
        return out

    def model_bounds(self):
        out = np.zeros(6, np.float32)
        check(self.L.crt_model_bounds(self.meshes.descs, len(self.meshes), _fp(self.o2r), int(self.precomputed_world), _fp(out)))
        return out


class Context:
    def __init__(self, device=0):
        self.L = _capi.load()
        self.h = C.c_void_p()
        check(self.L.crt_context_create(device, C.byref(self.h)))

    def synchronize(self):
        check(self.L.crt_context_synchronize(self.h))

    def set_stream(self, cuda_stream_ptr):
        check(self.L.crt_context_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    # ---- multi-GPU: one NCCL communicator per context (crt_nccl_*), used by Film.reduce
    @staticmethod
    def nccl_unique_id():
        """ncclGetUniqueId: 128 bytes to hand from one rank to all others (any launcher-side broadcast will do)."""
        uid = np.zeros(128, np.uint8)
        check(_capi.load().crt_nccl_unique_id(uid.ctypes.data_as(u8p)))
        return uid

    def nccl_init(self, world, rank, unique_id):
        uid = np.ascontiguousarray(unique_id, np.uint8)
        assert uid.size == 128
        check(self.L.crt_nccl_comm_create(self.h, int(world), int(rank), uid.ctypes.data_as(u8p)))

    def nccl_async_error(self):
        e = C.c_int()
        check(self.L.crt_nccl_async_error(self.h, C.byref(e)))
        return e.value

    @staticmethod
    def nccl_version():
        v = C.c_int(); origin = C.create_string_buffer(512)
        check(_capi.load().crt_nccl_version(C.byref(v), origin, 512))
        return v.value, origin.value.decode(errors="replace")

    def generate_rgb2spec(self):
        """Regenerate the sRGB RGBToSpectrumTable on the GPU and install it (color.h:405-432; the reference's data file is
        absent from its repository).  Returns (scale[64], data[3,64,64,64,3], milliseconds)."""
        scale = np.zeros(RGB2SPEC_RES, np.float32); data = np.zeros(RGB2SPEC_SHAPE, np.float32); ms = np.zeros(1, np.float32)
        check(self.L.crt_rgb2spec_generate(self.h, _fp(scale), _fp(data), _fp(ms)))
        return scale, data, float(ms[0])

    def set_sensor(self, r=None, g=None, b=None, illum=None, imaging_ratio=1.0 / 106.856895):
        """Measured PixelSensor (pixelsensor.h:37-68) from curves sampled at 360..830 nm; no arguments = the XYZ sensor.
        Returns XYZFromSensorRGB (9, column-major).  Re-commit scenes afterwards."""
        out = np.zeros(9, np.float32)
        if r is None:
            check(self.L.crt_context_set_sensor(self.h, None, None, None, None, 0.0, _fp(out)))
        else:
            check(self.L.crt_context_set_sensor(self.h, _fp(_f32(r)), _fp(_f32(g)), _fp(_f32(b)), _fp(_f32(illum)), float(imaging_ratio), _fp(out)))
        return out

    def set_rgb2spec(self, scale, data):
        scale = _f32(scale); data = _f32(data)
        assert scale.size == RGB2SPEC_RES and data.size == int(np.prod(RGB2SPEC_SHAPE))
        check(self.L.crt_rgb2spec_set(self.h, _fp(scale), _fp(data)))

    def close(self):
        if self.h:
            self.L.crt_context_destroy(self.h)
            self.h = C.c_void_p()


RGB2SPEC_RES = 64
RGB2SPEC_SHAPE = (3, 64, 64, 64, 3)


def rgb2spec_lookup(scale, data, rgb):
    """RGBToSpectrumTable::operator() (color.cpp:26-73) on the host: rgb -> sigmoid polynomial (c0, c1, c2)."""
    out = np.zeros(3, np.float32)
    check(_capi.load().crt_rgb2spec_lookup(_fp(_f32(scale)), _fp(_f32(data)), _fp(_f32(rgb)), _fp(out)))
    return out


def rgb2spec_fit(rgb):
    """One cold-start Gauss-Newton fit of a single RGB on the host (the table generator's cell solver)."""
    out = np.zeros(3, np.float32)
    check(_capi.load().crt_rgb2spec_fit(_fp(_f32(rgb)), _fp(out)))
    return out


def rgb2spec_save(path, scale, data):
    check(_capi.load().crt_rgb2spec_save_file(str(path).encode(), _fp(_f32(scale)), _fp(_f32(data))))


def rgb2spec_load(path):
    scale = np.zeros(RGB2SPEC_RES, np.float32); data = np.zeros(RGB2SPEC_SHAPE, np.float32)
    check(_capi.load().crt_rgb2spec_load_file(str(path).encode(), _fp(scale), _fp(data)))
    return scale, data


def kat_local_surface(tri9, bary3, rayd3, on_device=False):
    """Triangle::CalculateLocalSurface on explicit world-space triangles (no vertex attributes): (n, 17) records."""
    tri9 = _f32(tri9).reshape(-1, 9); n = len(tri9)
    out = np.zeros((n, 17), np.float32)
    check(_capi.load().crt_kat_local_surface(_fp(tri9), _fp(_f32(bary3)), _fp(_f32(rayd3)), n, int(on_device), _fp(out)))
    return out


def camera_matrices(kind, near, far, fov, pos, look, worldup, resx, resy, sensor_w=0.0, sensor_h=0.0, right=(1, 0, 0)):
    """PerspectiveCamera (kind 0) / OrthographicCamera (kind 1) matrices (Cameras.h:77-142,213-311)."""
    r2c = np.zeros(16, np.float32); c2w = np.zeros(16, np.float32)
    check(_capi.load().crt_camera_matrices(kind, near, far, sensor_w, sensor_h, fov, _fp(_f32(pos)), _fp(_f32(look)), _fp(_f32(right)),
                                          _fp(_f32(worldup)), float(resx), float(resy), _fp(r2c), _fp(c2w)))
    return r2c, c2w


def make_config(width, height, r2c, c2w, *, lens_radius=0.0, focal_distance=0.0, camera_kind=0, sampler_kind=1, xs=4, ys=4, jitter=1, seed=0,
                filter_kind=0, filter_r=(0.5, 0.5), mode=0, max_depth=5, rr_depth=0, ray_eps=1e-2, shadow_eps=1e-3, albedo=(0.5, 0.5, 0.5),
                spp_begin=0, spp_end=1, rank=0, world=1, partition=0, tile=(32, 32), trace_mode=0, collect_stats=0, time_kernels=0,
                filter_sigma=0.0, light_strategy=0, shade_mode=0):
    c = RenderConfig()
    c.width, c.height = width, height
    c.raster_to_camera[:] = list(_f32(r2c).reshape(-1)); c.camera_to_world[:] = list(_f32(c2w).reshape(-1))
    c.lens_radius, c.focal_distance, c.camera_kind = lens_radius, focal_distance, camera_kind
    c.sampler_kind, c.xs, c.ys, c.jitter, c.seed = sampler_kind, xs, ys, jitter, seed
    c.filter_kind, c.filter_rx, c.filter_ry = filter_kind, filter_r[0], filter_r[1]
    c.mode, c.max_depth, c.rr_depth, c.ray_eps, c.shadow_eps = mode, max_depth, rr_depth, ray_eps, shadow_eps
    c.albedo[:] = list(albedo)
    c.spp_begin, c.spp_end, c.rank, c.world, c.partition = spp_begin, spp_end, rank, world, partition
    c.tile_w, c.tile_h, c.trace_mode = tile[0], tile[1], trace_mode
    c.collect_stats, c.time_kernels = collect_stats, time_kernels
    c.filter_sigma = filter_sigma
    c.light_strategy = light_strategy
    c.shade_mode = shade_mode
    return c


def partition_spp_range(cfg):
    """Sample-index range [begin, end) crt_render gives cfg.rank of cfg.world (partition == 1)."""
    b = C.c_int32(); e = C.c_int32()
    check(_capi.load().crt_partition_spp_range(C.byref(cfg), C.byref(b), C.byref(e)))
    return b.value, e.value


def partition_pixels(cfg):
    """Pixel ids crt_render gives cfg.rank of cfg.world (interleaved tiles when partition == 0)."""
    L = _capi.load()
    n = L.crt_partition_pixel_count(C.byref(cfg))
    if n < 0:
        raise _capi.CrtError(L.crt_last_error().decode())
    out = np.zeros(max(n, 1), np.int32)
    check(L.crt_partition_pixels(C.byref(cfg), _ip(out), n))
    return out[:n]


class Scene:
    def __init__(self, ctx: Context):
        self.L = _capi.load()
        self.ctx = ctx
        self.h = C.c_void_p()
        check(self.L.crt_scene_create(ctx.h, C.byref(self.h)))
        self._keep = []
        self.n_shapes = 0

    def close(self):
        if self.h:
            self.L.crt_scene_destroy(self.h)
            self.h = C.c_void_p()

    def set_model(self, oct: Octtree_Model, cull_bits=None, mesh_materials=None):
        ms = oct.meshes
        cb = None
        if cull_bits is not None:
            cb = (u8p * len(ms))()
            for i, b in enumerate(cull_bits):
                b = np.ascontiguousarray(b, np.uint8)
                self._keep.append(b)
                cb[i] = b.ctypes.data_as(u8p)
        mm = np.ascontiguousarray(mesh_materials, np.int32) if mesh_materials is not None else None
        check(self.L.crt_scene_set_model(self.h, ms.descs, len(ms), _fp(oct.o2r), int(oct.precomputed_world), cb, oct.h, _ip(mm)))

    def add_shape(self, kind, rigid, params, material=0):
        out = C.c_int()
        pr = _f32(list(params) + [0] * (9 - len(params)))
        check(self.L.crt_scene_add_shape(self.h, kind, _fp(_f32(rigid).reshape(-1)), _fp(pr), material, C.byref(out)))
        self.n_shapes += 1
        return out.value

    def add_spectrum(self, kind, c=0.0, interleaved=None, n=0, name=None, normalize=False):
        out = C.c_int()
        arr = _f32(interleaved) if interleaved is not None else None
        if arr is not None and kind == 1:
            n = arr.size
        check(self.L.crt_scene_add_spectrum(self.h, kind, float(c), _fp(arr), int(n), name.encode() if name else None, int(normalize), C.byref(out)))
        return out.value

    def add_material(self, type=0, refl=-1, eta=-1, k=-1, emit=-1, emit_scale=0.0, two_sided=0, eta_constant=1):
        out = C.c_int()
        check(self.L.crt_scene_add_material(self.h, type, refl, eta, k, emit, float(emit_scale), two_sided, eta_constant, C.byref(out)))
        return out.value

    def add_light(self, kind, v, spectrum, scale=1.0):
        """Lights.h:5-8: kind 0 point light (v = position), 1 sun (v = direction towards the light)."""
        out = C.c_int()
        check(self.L.crt_scene_add_light(self.h, int(kind), _fp(_f32(v)), int(spectrum), float(scale), C.byref(out)))
        return out.value

    def commit(self):
        check(self.L.crt_scene_commit(self.h))

    def device_bytes(self):
        return self.L.crt_scene_device_bytes(self.h)

    def lights(self):
        n = self.L.crt_scene_light_count(self.h)
        cdf = np.zeros(max(n, 1), np.float32); mt = np.zeros((max(n, 1), 2), np.int32)
        check(self.L.crt_scene_get_light_cdf(self.h, _fp(cdf), _ip(mt), n))
        return cdf[:n], mt[:n]

    # ---- probes -------------------------------------------------------------------------------------
    def trace_closest(self, rays, mode=0):
        rays = _f32(rays); n = len(rays)
        mesh = np.full(n, -2, np.int32); tri = np.full(n, -2, np.int32); t = np.zeros(n, np.float32); b = np.zeros((n, 3), np.float32)
        check(self.L.crt_trace_closest(self.h, _fp(rays), n, mode, _ip(mesh), _ip(tri), _fp(t), _fp(b)))
        return dict(mesh=mesh, tri=tri, t=t, bary=b)

    def trace_any(self, rays, tmax, mode=0):
        rays = _f32(rays); tm = _f32(tmax); n = len(rays)
        out = np.zeros(n, np.int32)
        check(self.L.crt_trace_any(self.h, _fp(rays), _fp(tm), n, mode, _ip(out)))
        return out

    def traverse_surface(self, rays):
        rays = _f32(rays); n = len(rays)
        found = np.zeros(n, np.int32); nrm = np.zeros((n, 3), np.float32)
        check(self.L.crt_traverse_surface(self.h, _fp(rays), n, _ip(found), _fp(nrm)))
        return dict(found=found, n=nrm)

    LOCAL_SURFACE_FIELDS = (("hitp", 0, 3), ("uv", 3, 5), ("du", 5, 8), ("dv", 8, 11), ("n", 11, 14), ("wo", 14, 17))

    def traverse_local_surface(self, rays, mode=DEFAULT_TRACE_MODE):
        """Octtree_Model::Traverse -> LocalSurfaceInfo (Shapes.h:144-170): found, hitp, uv, du, dv, n, wo."""
        rays = _f32(rays); n = len(rays)
        found = np.zeros(n, np.int32); info = np.zeros((n, 17), np.float32)
        check(self.L.crt_traverse_local_surface(self.h, _fp(rays), n, mode, _ip(found), _fp(info)))
        out = dict(found=found)
        out.update({k: info[:, a:b] for k, a, b in self.LOCAL_SURFACE_FIELDS})
        return out

    def scene_closest(self, rays):
        rays = _f32(rays); n = len(rays)
        kind = np.zeros(n, np.int32); id0 = np.zeros(n, np.int32); id1 = np.zeros(n, np.int32); t = np.zeros(n, np.float32)
        p = np.zeros((n, 3), np.float32); ns = np.zeros((n, 3), np.float32); ng = np.zeros((n, 3), np.float32); bs = np.zeros(n, np.int32)
        check(self.L.crt_scene_closest(self.h, _fp(rays), n, _ip(kind), _ip(id0), _ip(id1), _fp(t), _fp(p), _fp(ns), _fp(ng), _ip(bs)))
        return dict(kind=kind, id0=id0, id1=id1, t=t, p=p, ns=ns, ng=ng, backside=bs)

    def shape_intersect(self, shape, rays, tmax=FLT_MAX):
        rays = _f32(rays); n = len(rays)
        found = np.zeros(n, np.int32); t = np.zeros(n, np.float32); hp = np.zeros((n, 3), np.float32); nrm = np.zeros((n, 3), np.float32); uv = np.zeros((n, 2), np.float32)
        fr = np.zeros((n, 9), np.float32)
        check(self.L.crt_shape_intersect_full(self.h, shape, _fp(rays), n, float(tmax), _ip(found), _fp(t), _fp(hp), _fp(nrm), _fp(uv), _fp(fr)))
        return dict(found=found, t=t, hitp=hp, n=nrm, uv=uv, du=fr[:, 0:3], dv=fr[:, 3:6], wo=fr[:, 6:9])

    def eval_samples(self, cfg, pixel_ids, indices):
        pid = np.ascontiguousarray(pixel_ids, np.int32); idx = np.ascontiguousarray(indices, np.int32); n = len(pid)
        ray = np.zeros((n, 6), np.float32); lam = np.zeros((n, 8), np.float32); pdf = np.zeros((n, 8), np.float32)
        L8 = np.zeros((n, 8), np.float32); rgb = np.zeros((n, 3), np.float32); w = np.zeros(n, np.float32)
        check(self.L.crt_eval_samples(self.h, C.byref(cfg), _ip(pid), _ip(idx), n, _fp(ray), _fp(lam), _fp(pdf), _fp(L8), _fp(rgb), _fp(w)))
        return dict(ray=ray, lam=lam, pdf=pdf, L=L8, rgb=rgb, weight=w)

    def render(self, film, cfg):
        st = RenderStats()
        check(self.L.crt_render(self.h, film.h, C.byref(cfg), C.byref(st)))
        return {k: getattr(st, k) for k, _ in RenderStats._fields_}


class Film:
    """Device-resident Film (RayTracer/Film.h:6-20): per pixel (rgbsum, weightsum)."""

    def __init__(self, ctx: Context, width, height):
        self.L = _capi.load()
        self.ctx = ctx
        self.width, self.height = width, height
        self.h = C.c_void_p()
        check(self.L.crt_film_create(ctx.h, width, height, C.byref(self.h)))

    def close(self):
        if self.h:
            self.L.crt_film_destroy(self.h)
            self.h = C.c_void_p()

    def clear(self):
        check(self.L.crt_film_clear(self.h))

    def attach(self, device_ptr):
        check(self.L.crt_film_attach_device(self.h, C.c_void_p(device_ptr)))

    def reduce(self, root=0):
        """ncclReduce(sum) of the per-GPU films onto `root` on the context's communicator (Context.nccl_init); no-op on one GPU."""
        check(self.L.crt_film_reduce(self.h, int(root)))

    def download(self, out=None):
        out = np.zeros((self.width * self.height, 4), np.float32) if out is None else out
        check(self.L.crt_film_download(self.h, _fp(out)))
        return out

    def upload(self, arr):
        arr = _f32(arr)
        check(self.L.crt_film_upload(self.h, _fp(arr)))

    def resolve(self, want_float=True):
        n = self.width * self.height
        rgb8 = np.zeros((n, 3), np.uint8)
        rgbf = np.zeros((n, 3), np.float32) if want_float else None
        check(self.L.crt_film_resolve(self.h, rgb8.ctypes.data_as(u8p), _fp(rgbf)))
        return rgb8, rgbf
