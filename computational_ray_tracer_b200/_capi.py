"""ctypes binding of libcrt_b200.so (include/crt_b200.h).  No torch types cross this boundary.

The library is the product: if it is missing the import fails -- there is no Python/CPU fallback.
"""
import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
# CRT_B200_LIB selects another build of the same library (tuning experiments: tools/build_variants.py); default = the in-tree one
LIB_PATH = os.environ.get("CRT_B200_LIB") or os.path.join(PKG, "libcrt_b200.so")

f32p = C.POINTER(C.c_float)
i32p = C.POINTER(C.c_int32)
u32p = C.POINTER(C.c_uint32)
u8p = C.POINTER(C.c_uint8)
u64p = C.POINTER(C.c_uint64)


class MeshDesc(C.Structure):
    _fields_ = [("positions", f32p), ("normals", f32p), ("n_vertices", C.c_uint32), ("indices", u32p), ("n_triangles", C.c_uint32),
                ("texcoords", f32p), ("tangents", f32p), ("bitangents", f32p)]


class OctreeStats(C.Structure):
    _fields_ = [("nodes", C.c_int32), ("real_nodes", C.c_int32), ("leaves", C.c_int32), ("empty_leaves", C.c_int32),
                ("max_leaf", C.c_int32), ("depth", C.c_int32), ("avg_leaf", C.c_float), ("refs", C.c_int64)]


class RenderConfig(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("raster_to_camera", C.c_float * 16), ("camera_to_world", C.c_float * 16),
                ("lens_radius", C.c_float), ("focal_distance", C.c_float), ("camera_kind", C.c_int32),
                ("sampler_kind", C.c_int32), ("xs", C.c_int32), ("ys", C.c_int32), ("jitter", C.c_int32), ("seed", C.c_int32),
                ("filter_kind", C.c_int32), ("filter_rx", C.c_float), ("filter_ry", C.c_float),
                ("mode", C.c_int32), ("max_depth", C.c_int32), ("rr_depth", C.c_int32),
                ("ray_eps", C.c_float), ("shadow_eps", C.c_float), ("albedo", C.c_float * 3),
                ("spp_begin", C.c_int32), ("spp_end", C.c_int32), ("rank", C.c_int32), ("world", C.c_int32),
                ("partition", C.c_int32), ("tile_w", C.c_int32), ("tile_h", C.c_int32), ("trace_mode", C.c_int32),
                ("collect_stats", C.c_int32), ("time_kernels", C.c_int32), ("filter_sigma", C.c_float), ("light_strategy", C.c_int32), ("shade_mode", C.c_int32)]


class RenderStats(C.Structure):
    _fields_ = [("paths", C.c_uint64), ("closest_rays", C.c_uint64), ("shadow_rays", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("depth_sum", C.c_uint64), ("exact_retraced_rays", C.c_uint64), ("queue_overflow_rays", C.c_uint64),
                ("nodes_visited", C.c_uint64), ("tris_tested", C.c_uint64), ("leaves_visited", C.c_uint64), ("max_queue", C.c_uint64),
                ("trace_launches", C.c_uint64), ("graph_launches", C.c_uint64), ("trace_ms", C.c_float), ("total_ms", C.c_float)]


EXPORTS = {
    # name: (restype, argtypes)
    "crt_last_error": (C.c_char_p, []),
    "crt_version": (C.c_int, []),
    "crt_context_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "crt_context_destroy": (None, [C.c_void_p]),
    "crt_context_synchronize": (C.c_int, [C.c_void_p]),
    "crt_context_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "crt_obj_load": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "crt_obj_destroy": (None, [C.c_void_p]),
    "crt_obj_mesh_count": (C.c_int, [C.c_void_p]),
    "crt_obj_mesh_info": (C.c_int, [C.c_void_p, C.c_int, u32p, u32p, C.c_char_p, C.c_int]),
    "crt_obj_mesh_copy": (C.c_int, [C.c_void_p, C.c_int, f32p, f32p, u32p]),
    "crt_obj_mesh_attributes": (C.c_int, [C.c_void_p, C.c_int, f32p, f32p, f32p, C.POINTER(C.c_int)]),
    "crt_octree_build": (C.c_int, [C.POINTER(MeshDesc), C.c_uint32, f32p, C.c_int, C.POINTER(C.c_void_p)]),
    "crt_octree_build_gpu": (C.c_int, [C.c_void_p, C.POINTER(MeshDesc), C.c_uint32, f32p, C.c_int, C.POINTER(C.c_void_p)]),
    "crt_octree_destroy": (None, [C.c_void_p]),
    "crt_octree_set_build_algorithm": (C.c_int, [C.c_int]),
    "crt_octree_flat_sizes": (C.c_int, [C.c_void_p, u64p]),
    "crt_octree_flat_copy": (C.c_int, [C.c_void_p, f32p, u32p, f32p, f32p, u32p]),
    "crt_model_compute_backface": (C.c_int, [C.POINTER(MeshDesc), f32p, f32p, C.c_int, u8p]),
    "crt_model_bounds": (C.c_int, [C.POINTER(MeshDesc), C.c_uint32, f32p, C.c_int, f32p]),
    "crt_octree_get_stats": (C.c_int, [C.c_void_p, C.POINTER(OctreeStats)]),
    "crt_octree_node_count": (C.c_int, [C.c_void_p]),
    "crt_octree_get_node": (C.c_int, [C.c_void_p, C.c_int, f32p, i32p, i32p, i32p, C.c_int32, i32p]),
    "crt_scene_create": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "crt_scene_destroy": (None, [C.c_void_p]),
    "crt_scene_set_model": (C.c_int, [C.c_void_p, C.POINTER(MeshDesc), C.c_uint32, f32p, C.c_int, C.POINTER(u8p), C.c_void_p, i32p]),
    "crt_scene_add_shape": (C.c_int, [C.c_void_p, C.c_int, f32p, f32p, C.c_int, C.POINTER(C.c_int)]),
    "crt_scene_add_spectrum": (C.c_int, [C.c_void_p, C.c_int, C.c_float, f32p, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int)]),
    "crt_scene_add_material": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "crt_scene_add_light": (C.c_int, [C.c_void_p, C.c_int, f32p, C.c_int, C.c_float, C.POINTER(C.c_int)]),
    "crt_scene_commit": (C.c_int, [C.c_void_p]),
    "crt_scene_light_count": (C.c_int, [C.c_void_p]),
    "crt_scene_get_light_cdf": (C.c_int, [C.c_void_p, f32p, i32p, C.c_int]),
    "crt_scene_device_bytes": (C.c_size_t, [C.c_void_p]),
    "crt_trace_closest": (C.c_int, [C.c_void_p, f32p, C.c_int, C.c_int, i32p, i32p, f32p, f32p]),
    "crt_scene_closest": (C.c_int, [C.c_void_p, f32p, C.c_int, i32p, i32p, i32p, f32p, f32p, f32p, f32p, i32p]),
    "crt_trace_any": (C.c_int, [C.c_void_p, f32p, f32p, C.c_int, C.c_int, i32p]),
    "crt_traverse_surface": (C.c_int, [C.c_void_p, f32p, C.c_int, i32p, f32p]),
    "crt_traverse_local_surface": (C.c_int, [C.c_void_p, f32p, C.c_int, C.c_int, i32p, f32p]),
    "crt_kat_local_surface": (C.c_int, [f32p, f32p, f32p, C.c_int, C.c_int, f32p]),
    "crt_shape_intersect": (C.c_int, [C.c_void_p, C.c_int, f32p, C.c_int, C.c_float, i32p, f32p, f32p, f32p, f32p]),
    "crt_shape_intersect_full": (C.c_int, [C.c_void_p, C.c_int, f32p, C.c_int, C.c_float, i32p, f32p, f32p, f32p, f32p, f32p]),
    "crt_film_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "crt_film_destroy": (None, [C.c_void_p]),
    "crt_film_clear": (C.c_int, [C.c_void_p]),
    "crt_film_attach_device": (C.c_int, [C.c_void_p, C.c_void_p]),
    "crt_film_device_ptr": (C.c_void_p, [C.c_void_p]),
    "crt_film_download": (C.c_int, [C.c_void_p, f32p]),
    "crt_film_upload": (C.c_int, [C.c_void_p, f32p]),
    "crt_film_resolve": (C.c_int, [C.c_void_p, u8p, f32p]),
    "crt_film_reduce_nccl": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "crt_film_reduce": (C.c_int, [C.c_void_p, C.c_int]),
    "crt_nccl_unique_id": (C.c_int, [u8p]),
    "crt_nccl_comm_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, u8p]),
    "crt_nccl_comm_destroy": (C.c_int, [C.c_void_p]),
    "crt_nccl_async_error": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "crt_nccl_version": (C.c_int, [C.POINTER(C.c_int), C.c_char_p, C.c_int]),
    "crt_render": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(RenderConfig), C.POINTER(RenderStats)]),
    "crt_partition_spp_range": (C.c_int, [C.POINTER(RenderConfig), i32p, i32p]),
    "crt_partition_pixel_count": (C.c_int, [C.POINTER(RenderConfig)]),
    "crt_partition_pixels": (C.c_int, [C.POINTER(RenderConfig), i32p, C.c_int32]),
    "crt_eval_samples": (C.c_int, [C.c_void_p, C.POINTER(RenderConfig), i32p, i32p, C.c_int, f32p, f32p, f32p, f32p, f32p, f32p]),
    "crt_dense_table": (C.c_int, [C.c_int, f32p]),
    "crt_color_constants": (C.c_int, [f32p, f32p, f32p, f32p]),
    "crt_camera_matrices": (C.c_int, [C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, f32p, f32p, f32p, f32p, C.c_float, C.c_float, f32p, f32p]),
    "crt_shape_matrices": (C.c_int, [f32p, f32p, f32p]),
    "crt_shape_area": (C.c_int, [C.c_int, f32p, f32p]),
    "crt_shape_bounds": (C.c_int, [C.c_int, f32p, f32p, f32p]),
    "crt_camera_generate_rays": (C.c_int, [C.c_int, f32p, f32p, C.c_float, C.c_float, f32p, f32p, C.c_int, C.c_int, f32p]),
    "crt_context_set_sensor": (C.c_int, [C.c_void_p, f32p, f32p, f32p, f32p, C.c_float, f32p]),
    "crt_measured_sensor_matrix": (C.c_int, [f32p, f32p, f32p, f32p, f32p]),
    "crt_rgb2spec_generate": (C.c_int, [C.c_void_p, f32p, f32p, f32p]),
    "crt_rgb2spec_set": (C.c_int, [C.c_void_p, f32p, f32p]),
    "crt_rgb2spec_load_file": (C.c_int, [C.c_char_p, f32p, f32p]),
    "crt_rgb2spec_save_file": (C.c_int, [C.c_char_p, f32p, f32p]),
    "crt_rgb2spec_lookup": (C.c_int, [f32p, f32p, f32p, f32p]),
    "crt_rgb2spec_fit": (C.c_int, [f32p, f32p]),
    "crt_kat_hash": (C.c_int, [C.c_char_p, C.c_uint64, C.c_uint64, C.c_int, u64p]),
    "crt_kat_permutation": (C.c_int, [u32p, u32p, u32p, C.c_int, C.c_int, i32p]),
    "crt_kat_pcg32": (C.c_int, [C.c_int, C.c_uint64, C.c_uint64, C.c_int64, C.c_int, C.c_int, u32p, f32p]),
    "crt_kat_gaussian_filter": (C.c_int, [C.c_float, C.c_float, C.c_float, f32p, C.c_int, C.c_int, f32p]),
    "crt_kat_sampler": (C.c_int, [C.c_int] * 9 + [C.c_char_p, C.c_int, f32p]),
}

_lib = None


def load():
    """Load libcrt_b200.so; raises if it has not been built (python -m computational_ray_tracer_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m computational_ray_tracer_b200.build` "
                           "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(L, name)       # AttributeError if the header declares something the library lacks
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


class CrtError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise CrtError(load().crt_last_error().decode(errors="replace"))
