/* crt_b200.h -- C ABI of libcrt_b200.so, the B200 (sm_100a) implementation of the reference's
 * render hot path (camera ray -> octree closest/any hit -> shade -> film).
 *
 * The reference (GiboDidact/Computational_ray_tracer) has no FFI/plugin layer: its boundary is the
 * in-process C++ object model.  Each entry point below names the reference interface it stands in
 * for (paths relative to the reference root).  A C++ facade with the reference's class shapes
 * (include/crt/facade.hpp) and a ctypes binding (computational_ray_tracer_b200/_capi.py) sit on top.
 *
 * Conventions: every function returns 0 on success, non-zero on error (crt_last_error() gives the
 * text; the reference instead prints to std::cout and returns {} -- Shapes.h:1103-1107).  Handles are
 * opaque and owned by the caller until the matching *_destroy.  Host pointers stay owned by the
 * caller and may be freed as soon as the call returns.  Matrices are 16 floats, column-major (glm).
 * One context = one GPU; calls on one context must be externally serialised (the reference's
 * shapes/octree are likewise read-only-shared and its samplers per-thread, RayTracerTestApp.h:361-392).
 * There is NO CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef CRT_B200_H
#define CRT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct crt_context crt_context;   /* device + stream + scratch                                   */
typedef struct crt_octree crt_octree;     /* host-side Octtree_Model (RayTracer/Octtree_Model.h)          */
typedef struct crt_scene crt_scene;       /* device-resident scene: flattened octree, triangles, shapes   */
typedef struct crt_film crt_film;         /* device-resident Film (RayTracer/Film.h:6-20)                 */

const char* crt_last_error(void);
int crt_version(void);

/* ---- context ------------------------------------------------------------------------------- */
int crt_context_create(int device, crt_context** out);
void crt_context_destroy(crt_context* ctx);
int crt_context_synchronize(crt_context* ctx);
/* Run the library's work on a caller-owned stream (cudaStream_t as void*); NULL = the context's own. */
int crt_context_set_stream(crt_context* ctx, void* cuda_stream);

/* ---- mesh data model: MeshCache::Mesh / Model (RayTracer/AssetManager.h:20-47) ---------------- */
typedef struct crt_mesh_desc {
    const float* positions;    /* 3 floats per vertex                                                */
    const float* normals;      /* 3 floats per vertex, or NULL (geometric normal is used)            */
    uint32_t n_vertices;
    const uint32_t* indices;   /* 3 per triangle, local to this mesh                                 */
    uint32_t n_triangles;
    /* the remaining MeshCache::Mesh arrays, read by Triangle::CalculateLocalSurface when the matching
     * Triangle::vertex_available flag is set (Shapes.h:917-924,1036-1063); NULL = not available          */
    const float* texcoords;    /* 2 floats per vertex, or NULL: (u,v) = p0(0,0) p1(1,0) p2(0,1) (Shapes.h:999) */
    const float* tangents;     /* 3 floats per vertex, or NULL: du = dpdu                            */
    const float* bitangents;   /* 3 floats per vertex, or NULL: dv = dpdv                            */
} crt_mesh_desc;

/* ---- asset ingestion: MeshCache::LoadMeshFromFile / ASSIMPLoader (RayTracer/AssetManager.cpp:8-25,67-190) ------
 * Wavefront OBJ only (assimp itself is not part of the reference repository): triangulated, one vertex per face
 * corner, flat normals generated where the file has none, texture coordinates kept, tangent space computed -- what the
 * reference's import flags (aiProcess_Triangulate | CalcTangentSpace | GenNormals) produce.                          */
typedef struct crt_obj crt_obj;
int crt_obj_load(const char* path, crt_obj** out);
void crt_obj_destroy(crt_obj* obj);
int crt_obj_mesh_count(const crt_obj* obj);
int crt_obj_mesh_info(const crt_obj* obj, int mesh, uint32_t* n_vertices, uint32_t* n_triangles, char* name, int name_cap);
int crt_obj_mesh_copy(const crt_obj* obj, int mesh, float* positions, float* normals, uint32_t* indices);
/* The remaining MeshCache::Mesh arrays as ASSIMPLoader::Process_Mesh fills them (AssetManager.cpp:104-190): texcoords (2 per vertex) from
 * `vt`, tangents (3) by assimp's CalcTangentSpace rule, and bitangents == tangents -- the reference stores the tangent twice (:153).
 * *available = 1 when the mesh has texture coordinates on every corner; otherwise the arrays are zero-filled (:139,:156-157).        */
int crt_obj_mesh_attributes(const crt_obj* obj, int mesh, float* texcoords, float* tangents, float* bitangents, int* available);

/* ---- Octtree_Model (RayTracer/Octtree_Model.h:29-63,180-366; ThirdParty/AABB_triangle_Moller.h) -
 * Host build, exactly the reference's incremental insert / lazy 8-way split at 40 triangles.
 * `object_to_render` is TriModel's ObjectToRender (Shapes.h:175-182); with precomputed_world != 0 the
 * positions are used as they are (Shapes.h:1117-1128, Octtree_Model.h:192).
 * cull_bits: optional per-mesh arrays (one byte per triangle) = TriModel::back_facing
 * (Shapes.h:1339-1380); pass NULL to disable culling.  They only matter at flatten time.            */
int crt_octree_build(const crt_mesh_desc* meshes, uint32_t n_meshes, const float* object_to_render,
                     int precomputed_world, crt_octree** out);
void crt_octree_destroy(crt_octree* oct);
/* Which builder crt_octree_build uses (process-wide): 0 = the reference's incremental insertion (node ids in reference
 * creation order: GetNode(i) matches Octtree_Model::GetNode(i)); 1 = top-down construction of the same tree (same cells,
 * same per-leaf triangle order, same flattened device layout; node ids in breadth-first creation order).             */
int crt_octree_set_build_algorithm(int algorithm);
/* CreateOcttree on the GPU (level-synchronous, csrc/crt_build.cuh): the same tree and the same flattened layout as
 * crt_octree_build, node ids breadth-first.  1 M triangles in tens of milliseconds instead of seconds.                  */
int crt_octree_build_gpu(crt_context* ctx, const crt_mesh_desc* meshes, uint32_t n_meshes, const float* object_to_render,
                         int precomputed_world, crt_octree** out);
/* The flattened device layout of an octree with nothing culled: element counts of
 * {nodes (floats), leaf_refs, node_tight (floats), pk_boxes (floats), pk_refs}, then the arrays themselves.           */
int crt_octree_flat_sizes(const crt_octree* oct, uint64_t* sizes5);
int crt_octree_flat_copy(const crt_octree* oct, float* nodes, uint32_t* leaf_refs, float* node_tight, float* pk_boxes, uint32_t* pk_refs);
/* TriModel::ComputeBackFace (Shapes.h:1339-1380): per-triangle "averaged vertex normal faces look_dir". */
int crt_model_compute_backface(const crt_mesh_desc* mesh, const float* look_dir3, const float* object_to_render,
                               int precomputed_world, uint8_t* out_bits);
/* TriModel::Bounds (Shapes.h:1282-1300,1390-1397), including the FLT_MIN max-initialiser quirk.      */
int crt_model_bounds(const crt_mesh_desc* meshes, uint32_t n_meshes, const float* object_to_render,
                     int precomputed_world, float* out_min3_max3);

typedef struct crt_octree_stats {          /* Octtree_Model::PrintInfo (Octtree_Model.h:134-176)     */
    int32_t nodes, real_nodes, leaves, empty_leaves, max_leaf, depth;
    float avg_leaf;
    int64_t refs;
} crt_octree_stats;
int crt_octree_get_stats(const crt_octree* oct, crt_octree_stats* out);
int crt_octree_node_count(const crt_octree* oct);                       /* getTreeSize(), :129        */
/* GetNode(i) (:178), reference (creation) order: bounds[6], leaf flag, child ids[8] (-1 for a leaf),
 * number of (mesh,tri) pairs copied into pairs (up to cap).                                           */
int crt_octree_get_node(const crt_octree* oct, int i, float* bounds6, int32_t* leaf, int32_t* child8,
                        int32_t* pairs, int32_t cap, int32_t* n_pairs);

/* ---- scene ------------------------------------------------------------------------------------ */
int crt_scene_create(crt_context* ctx, crt_scene** out);
void crt_scene_destroy(crt_scene* scene);
/* Attach the triangle model + its octree (flattened to BFS-ordered 32-byte nodes, leaf reference lists and
 * 48-byte world-space triangles; back-face-culled and degenerate triangles are dropped from the leaf
 * lists because Octtree_Model::Traverse skips / Triangle::BasicIntersect rejects them unconditionally,
 * Octtree_Model.h:91-96, Shapes.h:1131-1134).  mesh_materials: one material id per mesh (may be NULL). */
int crt_scene_set_model(crt_scene* scene, const crt_mesh_desc* meshes, uint32_t n_meshes, const float* object_to_render,
                        int precomputed_world, const uint8_t* const* cull_bits, const crt_octree* oct,
                        const int32_t* mesh_materials);
/* Analytic shapes (Shapes.h:209-905).  kind: 0 Sphere(r,zmin,zmax,phimax_deg) 1 Cylinder(r,zmin,zmax,phimax_deg)
 * 2 Disk(h,inner_r,outer_r,phimax_deg) 3 TriangleSimple(p1,p2,p3).  rigid = the constructor's rigidtransform.
 * At most 65535 shapes per scene. */
int crt_scene_add_shape(crt_scene* scene, int kind, const float* rigid16, const float* params9, int material, int* out_id);
/* Spectra (ThirdParty/pbrv4/spectrum.h:355-638).  kind: 0 constant(c); 1 piecewise-linear from n interleaved
 * (lambda,value) floats [FromInterleaved, spectrum.cpp:134-165]; 2 named table (see crt_named_table_count);
 * 3 Macbeth swatch i=n (pixelsensor.cpp:16-237); 4 normalised std illuminant n (0 A,1 D50,2 D65,3 F1,4 F2,5 F11);
 * 5 grey RGBAlbedoSpectrum(c,c,c); 6 grey RGBIlluminantSpectrum(c,c,c) (x sRGB's D65); 7 RGBAlbedoSpectrum(rgb),
 * 8 RGBIlluminantSpectrum(rgb), 9 RGBUnboundedSpectrum(rgb) with rgb = interleaved[0..2] (spectrum.cpp:249-270;
 * non-grey rgb needs the context's RGB -> spectrum table, see crt_rgb2spec_*).                              */
int crt_scene_add_spectrum(crt_scene* scene, int kind, float c, const float* interleaved, int n, const char* name,
                           int normalize, int* out_id);
/* Materials (Tier B; intent notes Shading.h:1-20).  type: 0 Lambert, 1 smooth dielectric, 2 smooth conductor. */
int crt_scene_add_material(crt_scene* scene, int type, int refl, int eta, int k, int emit, float emit_scale,
                           int two_sided, int eta_constant, int* out_id);
/* Lights (RayTracer/Lights.h:5-8, intent notes): kind 0 point light -- v = position, radiant intensity scale * spectrum, "r^2 falloff";
 * kind 1 sun -- v = direction towards the light (normalised here), irradiance scale * spectrum.  Sampled once each at every Lambert hit
 * of the path integrator, after the emissive-triangle sample(s).                                                                       */
int crt_scene_add_light(crt_scene* scene, int kind, const float* v3, int spectrum, float scale, int* out_id);
int crt_scene_commit(crt_scene* scene);     /* upload everything; build the emissive-triangle CDF          */
int crt_scene_light_count(const crt_scene* scene);
int crt_scene_get_light_cdf(const crt_scene* scene, float* cdf, int32_t* mesh_tri_pairs, int cap);
size_t crt_scene_device_bytes(const crt_scene* scene);

/* ---- probes: the parity surface ------------------------------------------------------------------ */
/* Octtree_Model::Traverse up to the hit record (Octtree_Model.h:66-122): closest hit per ray, BFS order,
 * shrinking tMax, strict '<'.  rays = 6 floats (o, d) each, HOST memory.  Outputs (HOST, any may be NULL):
 * mesh_id/tri_id (-1 = miss), t, bary (b0,b1,b2).  mode 0 = exact BFS emulation; 3 = ordered traversal (one ray per
 * lane) with exact-BFS re-trace of order-sensitive rays; identical results.                                    */
int crt_trace_closest(crt_scene* scene, const float* rays, int n, int mode, int32_t* mesh_id, int32_t* tri_id,
                      float* t, float* bary3);
/* Scene-level closest hit (mesh via the octree, then analytic shapes in list order, strict '<') and the
 * Tier-B surface record: kind (-1 miss, 0 triangle, 1 shape), id0 (mesh or shape), id1 (tri), t, p, ns, ng, backside. */
int crt_scene_closest(crt_scene* scene, const float* rays, int n, int32_t* kind, int32_t* id0, int32_t* id1, float* t,
                      float* p3, float* ns3, float* ng3, int32_t* backside);
/* Occlusion with a fixed per-ray tMax (order independent): out[i] = 1 if anything is hit in (0, tmax[i]).
 * mode as in crt_trace_closest (0 BFS order, 1 ordered depth-first with early exit).                         */
int crt_trace_any(crt_scene* scene, const float* rays, const float* tmax, int n, int mode, int32_t* out);
/* Octtree_Model::Traverse incl. Triangle::CalculateLocalSurface (Shapes.h:982-1083): normal n as Li uses it. */
int crt_traverse_surface(crt_scene* scene, const float* rays, int n, int32_t* found, float* nrm3);
/* Octtree_Model::Traverse's full return value (Octtree_Model.h:66-127 -> Triangle::CalculateLocalSurface, Shapes.h:982-1083): the
 * LocalSurfaceInfo record (Shapes.h:144-170) as 17 floats per ray -- hitp(3) u v du(3) dv(3) n(3) wo(3) -- computed on the device
 * (uv / tangent / bitangent / normal interpolation where the model carries those attributes, else the fixed uv, dpdu / dpdv incl. the
 * degenerate-frame fallback :1016-1029, and the geometric normal; n faces the ray).  tHit is not part of it: the reference never assigns
 * it for triangle hits (:1034).  mode as in crt_trace_closest.                                                                          */
int crt_traverse_local_surface(crt_scene* scene, const float* rays, int n, int mode, int32_t* found, float* info17);
/* Triangle::CalculateLocalSurface on explicit world-space triangles without vertex attributes (tri9), for given barycentrics and
 * normalised ray directions; on the device when on_device != 0.  Exists to pin the degenerate-frame branch, which no ray can reach
 * through Traverse (BasicIntersect rejects zero-area triangles first, Shapes.h:1131-1134).                                           */
int crt_kat_local_surface(const float* tri9, const float* bary3, const float* rayd3, int n, int on_device, float* info17);
/* Shape::Intersect for one analytic shape (Shapes.h:244-270 and siblings).                                    */
int crt_shape_intersect(crt_scene* scene, int shape, const float* rays, int n, float tmax, int32_t* found, float* t,
                        float* hitp3, float* nrm3, float* uv2);
/* The same with the rest of the LocalSurfaceInfo record: du, dv (calculate_du / calculate_dv of each shape, Shapes.h:393-408,589-600,
 * 731-739,884-892) and wo, after LocalSurfaceInfo::Transform(ObjectToRender) (Shapes.h:147-160); 9 floats per ray.                     */
int crt_shape_intersect_full(crt_scene* scene, int shape, const float* rays, int n, float tmax, int32_t* found, float* t,
                             float* hitp3, float* nrm3, float* uv2, float* du_dv_wo9);

/* ---- film -------------------------------------------------------------------------------------- */
int crt_film_create(crt_context* ctx, int width, int height, crt_film** out);
void crt_film_destroy(crt_film* film);
int crt_film_clear(crt_film* film);
/* Use caller-owned device memory (width*height*4 floats: rgbsum, weightsum) as the film, e.g. a torch tensor
 * that torch.distributed will reduce.  NULL detaches.                                                       */
int crt_film_attach_device(crt_film* film, void* device_ptr);
void* crt_film_device_ptr(crt_film* film);
int crt_film_download(crt_film* film, float* host_rgbw);                 /* width*height*4 floats          */
int crt_film_upload(crt_film* film, const float* host_rgbw);             /* resume (checkpoint)  */
/* Resolve (RayTracerTestApp.h:425-452): rgbsum/weightsum -> XYZFromSensorRGB -> sRGB -> clamp -> *255.       */
int crt_film_resolve(crt_film* film, uint8_t* host_rgb8, float* host_rgbf);
/* Multi-GPU (one process and one context per GPU): the per-GPU films are combined with ONE ncclReduce(sum) onto `root` over
 * NVLink -- the reference's threads write one shared Film (RayTracerTestApp.h:361-366); here that is the only exchange step.
 * libnccl is resolved with dlopen at first use: the copy already loaded in the process (e.g. torch's), else $CRT_NCCL_LIB,
 * else libnccl.so.2 from the library search path.
 *   crt_nccl_unique_id    : ncclGetUniqueId (128 bytes) -- call on one rank, hand the bytes to the others through the launcher
 *   crt_nccl_comm_create  : ncclCommInitRank; the communicator belongs to the context (destroyed with it)
 *   crt_film_reduce       : in-place ncclReduce(sum) of the film on the context's communicator and stream; no-op without one
 *   crt_film_reduce_nccl  : the same on a caller-owned ncclComm_t
 *   crt_nccl_async_error  : ncclCommGetAsyncError of the context's communicator (0 = healthy)
 *   crt_nccl_version      : version and origin of the NCCL the library bound to                                                  */
int crt_nccl_unique_id(uint8_t* id128);
int crt_nccl_comm_create(crt_context* ctx, int world, int rank, const uint8_t* id128);
int crt_nccl_comm_destroy(crt_context* ctx);
int crt_nccl_async_error(crt_context* ctx, int* async_error);
int crt_nccl_version(int* version, char* origin, int origin_cap);
int crt_film_reduce(crt_film* film, int root);
int crt_film_reduce_nccl(crt_film* film, void* nccl_comm, int root);

/* ---- render: evaluate_pixel + Li + thread pool (Applications/RayTracerTestApp.h:218-409) ---------- */
typedef struct crt_render_config {
    int32_t width, height;
    float raster_to_camera[16];    /* CameraBase::M_RastertoCamera (Cameras.h:303-310)                  */
    float camera_to_world[16];     /* CameraBase::M_CameratoWorld  (Cameras.h:130-142)                  */
    float lens_radius, focal_distance;
    int32_t camera_kind;           /* 0 PerspectiveCamera, 1 OrthographicCamera, 2 PinholeCamera (raster_to_camera = M_RastertoScreen,
                                      focal_distance = box depth; Cameras.h:313-359)                       */
    int32_t sampler_kind;          /* 0 IndependentSampler(xs*ys spp), 1 StratifiedSampler(xs,ys,jitter) */
    int32_t xs, ys, jitter, seed;
    int32_t filter_kind;           /* 0 BoxFilter, 1 TriangleFilter (deterministic tent), 2 GaussianFilter
                                      (filters.h:96-163; tabulated inverse CDF, sigma = filter_sigma below) */
    float filter_rx, filter_ry;
    int32_t mode;                  /* 0 reference Li (RayTracerTestApp.h:218-284), 1 path integrator      */
    int32_t max_depth, rr_depth;
    float ray_eps, shadow_eps;
    float albedo[3];               /* Tier A `colors` (RayTracerTestApp.h:207); grey only                */
    int32_t spp_begin, spp_end;    /* sample indices [begin, end) -- the reference's pixel_index loop    */
    int32_t rank, world;           /* multi-GPU partition of the image; world<=1 = everything            */
    int32_t partition;             /* 0 interleaved tiles (tile_w x tile_h, tile_id % world == rank), 1 spp range */
    int32_t tile_w, tile_h;
    int32_t trace_mode;            /* 0 exact BFS kernel; 3 ordered traversal, one ray per lane, + exact re-trace of order-sensitive rays
                                      (production; identical results)                                                       */
    int32_t collect_stats;         /* count nodes/triangles visited (instrumented kernels; not for timing) */
    int32_t time_kernels;          /* bracket every traversal launch with CUDA events -> stats.trace_ms    */
    float filter_sigma;            /* GaussianFilter sigma; <= 0 selects the class default 0.5 (filters.h:100) */
    int32_t light_strategy;        /* path integrator, emissive triangles at a Lambert hit: 0 one sample, light picked by the power CDF;
                                      1 "1 sample from each light source" (Shading.h:4).  Point / sun lights are always sampled once each. */
    int32_t shade_mode;            /* path integrator, how a bounce is shaded (identical films): 0 automatic (staged when the scene has analytic
                                      shapes), 1 fused (surface record + all materials in one kernel), 2 staged (surface-record kernel, then
                                      one kernel per material type over that type's queue) */
} crt_render_config;

typedef struct crt_render_stats {
    uint64_t paths, closest_rays, shadow_rays, kernel_launches;
    uint64_t depth_sum;            /* path integrator: sum over paths of the realised depth                */
    uint64_t exact_retraced_rays, queue_overflow_rays;
    uint64_t nodes_visited, tris_tested, leaves_visited, max_queue;   /* collect_stats only                 */
    uint64_t trace_launches;       /* traversal launches covered by trace_ms                              */
    uint64_t graph_launches;       /* waves replayed from a CUDA graph (small frames; their kernels are counted in kernel_launches) */
    float trace_ms, total_ms;      /* CUDA-event times on the library's stream                            */
} crt_render_stats;

int crt_render(crt_scene* scene, crt_film* film, const crt_render_config* cfg, crt_render_stats* stats);
/* The partition crt_render applies for (cfg->rank, cfg->world), exposed as host-only helpers (no GPU needed):
 * the static split of RayTracerTestApp.h:375-394 re-cut for GPUs.  partition 1: contiguous sample-index range of
 * [spp_begin, spp_end) for this rank, every pixel; partition 0: all sample indices, the pixels of the interleaved
 * tile_w x tile_h tiles with tile_id % world == rank (ascending pixel ids).                                     */
int crt_partition_spp_range(const crt_render_config* cfg, int32_t* begin, int32_t* end);
int crt_partition_pixel_count(const crt_render_config* cfg);
int crt_partition_pixels(const crt_render_config* cfg, int32_t* pixel_ids, int32_t cap);
/* Per-sample probe (parity with evaluate_pixel): for each (pixel_id, sample index) the camera ray (6),
 * wavelengths (8), pdf (8), radiance L (8), clamped sensor RGB (3) and filter weight.  HOST pointers.        */
int crt_eval_samples(crt_scene* scene, const crt_render_config* cfg, const int32_t* pixel_ids, const int32_t* indices,
                     int n, float* ray6, float* lambda8, float* pdf8, float* L8, float* rgb3, float* weight);

/* ---- host-side constants the oracle must agree with bit for bit ----------------------------------- */
/* which: 0 X, 1 Y, 2 Z matching curves, 3 sRGB illuminant (D65) -- DenselySampledSpectrum, 471 values.     */
int crt_dense_table(int which, float* out471);
/* XYZFromSensorRGB (pixelsensor.h:70-79), RGBFromXYZ, XYZFromRGB (colorspace.cpp:13-28): 9 floats each,
 * column-major; white = sRGB white point xy.                                                              */
int crt_color_constants(float* sensor9, float* rgb_from_xyz9, float* xyz_from_rgb9, float* white2);
/* CameraBase / PerspectiveCamera (kind 0) / OrthographicCamera (1) matrices (Cameras.h:77-142,213-311); kind 2 =
 * PinholeCamera: sensor_w/h = box.xy, raster_to_camera16 receives M_RastertoScreen (Cameras.h:313-359).        */
int crt_camera_matrices(int kind, float near_, float far_, float sensor_w, float sensor_h, float fov_deg,
                        const float* pos3, const float* look3, const float* right3, const float* worldup3,
                        float res_x, float res_y, float* raster_to_camera16, float* camera_to_world16);
/* Shape transform convention (Shapes.h:175-182): rigid -> ObjectToRender, RenderToObject.                   */
int crt_shape_matrices(const float* rigid16, float* object_to_render16, float* render_to_object16);
/* Shape::Area (Shapes.h:234-237, 455-458, 642-645, 779-782) and Shape::Bounds = TransformBounds(object box, ObjectToRender)
 * (Shapes.h:239-242, 459-462, 647-650, 784-792; Bounds3::Transform :60-98).  kind / params9 as in crt_scene_add_shape.   */
int crt_shape_area(int kind, const float* params9, float* out);
int crt_shape_bounds(int kind, const float* rigid16, const float* params9, float* out_min3_max3);
/* CameraBase::generateRay(pixel, sampler) (Cameras.h:179; :231-242, :273-297, :340-352) for n film positions (x, y) and, for a thin
 * lens, the n Get2D() draws it consumes (NULL = no lens).  kind and matrices as in crt_render_config; ray6 = (o, d) per position.
 * on_device != 0 runs the device function the renderer itself uses.                                                            */
int crt_camera_generate_rays(int kind, const float* raster_to_camera16, const float* camera_to_world16, float lens_radius, float focal_distance,
                             const float* film_xy2, const float* lens_u2, int n, int on_device, float* ray6);

/* ---- film sensor (Film::pixel_sensor, RayTracer/Film.h:18; PixelSensor, ThirdParty/pbrv4/pixelsensor.h:28-101) -----------
 * Default: the app's sensor_xyz = PixelSensor(sRGB, stdillum-D65, 1/CIE_Y_integral) (RayTracerTestApp.h:149).
 * crt_context_set_sensor installs the measured-sensor constructor (pixelsensor.h:37-68, the app's sensor_canon :152-153): r/g/b
 * response curves and the sensor illuminant sampled at the 471 integer wavelengths 360..830 nm (what its DenselySampledSpectrum
 * members and 1 nm sums see), XYZFromSensorRGB by LinearLeastSquares over the 24 Macbeth swatches (helpers.h:257-274, including
 * its [col][row] indexing, which is why the reference's author notes "doesnt work", :151).  r471 == NULL restores the default.
 * Scenes capture the sensor at crt_scene_commit; films resolve with the context's current matrix.                          */
int crt_context_set_sensor(crt_context* ctx, const float* r471, const float* g471, const float* b471, const float* illum471,
                           float imaging_ratio, float* xyz_from_sensor_rgb9_out);
/* The same matrix without a context (host only).                                                                           */
int crt_measured_sensor_matrix(const float* r471, const float* g471, const float* b471, const float* illum471,
                               float* xyz_from_sensor_rgb9_out);

/* ---- RGB -> spectrum table (RGBToSpectrumTable, ThirdParty/pbrv4/color.h:405-432, color.cpp:26-166) ---------------
 * The reference loads `float scale[64]` + `float data[3][64][64][64][3]` from ../rgb2spec/sRGB64binary, a file its
 * repository does not contain.  A table is a property of the context; non-grey RGB spectra (crt_scene_add_spectrum kinds
 * 7-9, crt_render_config.albedo) need one and fail with an explicit error otherwise.
 *   crt_rgb2spec_generate : regenerate the sRGB table on the GPU (Jakob-Hanika optimiser, csrc/crt_rgb2spec.cuh),
 *                           install it in the context and optionally copy it out (either pointer may be NULL).
 *   crt_rgb2spec_set      : install a caller-provided table (e.g. read from the reference's file).
 *   crt_rgb2spec_load_file / save_file : the reference's binary layout (color.cpp:107-158): 4 bytes, 64 floats, data.
 *   crt_rgb2spec_lookup   : RGBToSpectrumTable::operator() on the host (no GPU needed): rgb3 -> (c0, c1, c2).
 *   crt_rgb2spec_fit      : one cold-start fit of a single RGB on the host (no GPU needed), coefficients in nm.        */
#define CRT_RGB2SPEC_RES 64
#define CRT_RGB2SPEC_DATA_FLOATS (3 * 64 * 64 * 64 * 3)
int crt_rgb2spec_generate(crt_context* ctx, float* scale64_out, float* data_out, float* milliseconds_out);
int crt_rgb2spec_set(crt_context* ctx, const float* scale64, const float* data);
int crt_rgb2spec_load_file(const char* path, float* scale64_out, float* data_out);
int crt_rgb2spec_save_file(const char* path, const float* scale64, const float* data);
int crt_rgb2spec_lookup(const float* scale64, const float* data, const float* rgb3, float* coeffs3_out);
int crt_rgb2spec_fit(const float* rgb3, float* coeffs3_out);

/* Integer primitives of the sampler stack, exported for known-answer tests (run on the device when
 * on_device != 0): MurmurHash64A (hash.h:18-63), MixBits (:67-74), PermutationElement
 * (HelperFunctions.h:175-203), PCG32 (rng.h:24-162), sampler sequences (samplers.h:38-136).              */
int crt_kat_hash(const uint8_t* key, uint64_t len, uint64_t seed, int on_device, uint64_t* out);
int crt_kat_permutation(const uint32_t* i, const uint32_t* l, const uint32_t* p, int n, int on_device, int32_t* out);
int crt_kat_pcg32(int mode, uint64_t seq, uint64_t offset, int64_t advance, int n, int on_device, uint32_t* out_u32, float* out_f);
/* pbrt::GaussianFilter(radius, sigma).Sample(u) (filters.h:96-163, RayTracer/Sampling.h:781-848): (p.x, p.y, weight) per u */
int crt_kat_gaussian_filter(float rx, float ry, float sigma, const float* u2, int n, int on_device, float* out3);
int crt_kat_sampler(int kind, int xs, int ys, int jitter, int seed, int px, int py, int index, int dim, const char* pattern,
                    int on_device, float* out);

#ifdef __cplusplus
}
#endif
#endif /* CRT_B200_H */
