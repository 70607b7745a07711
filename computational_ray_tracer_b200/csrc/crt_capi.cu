// crt_capi.cu -- device half of the C ABI (include/crt_b200.h): context, scene upload, film, render loop, probes.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>            // types and enums only: libnccl itself is resolved with dlopen (see the NCCL section)

#include <algorithm>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "crt_host.h"
#include "crt_build.cuh"
#include "crt_path.cuh"
#include "crt_rgb2spec.cuh"

using namespace crt;

#define CRT_CUDA(call)                                                                                     \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess) {                                                                           \
            set_error(std::string(#call) + ": " + cudaGetErrorString(e_));                                 \
            return 2;                                                                                      \
        }                                                                                                  \
    } while (0)

namespace {

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t resize(size_t count) {
        if (count <= n && p) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; n = 0;
        if (count == 0) return cudaSuccess;
        cudaError_t e = cudaMalloc(&p, count * sizeof(T));
        if (e == cudaSuccess) n = count;
        return e;
    }
    cudaError_t upload(const T* src, size_t count, cudaStream_t s) {
        cudaError_t e = resize(count);
        if (e != cudaSuccess || count == 0) return e;
        return cudaMemcpyAsync(p, src, count * sizeof(T), cudaMemcpyHostToDevice, s);
    }
    size_t bytes() const { return n * sizeof(T); }
};

inline int cdiv(int a, int b) { return (a + b - 1) / b; }

}  // namespace

struct WaveGraphKey {           // everything a captured wave bakes into its kernel arguments
    RenderConst rc;
    const void* scene; unsigned scene_gen; const void* film; int n, per_wave, max_depth, trace_mode, light_strategy, shade_mode; const void* pixel_list; unsigned wave_gen; const void* stream;
};

struct crt_context {
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    int sm_count = 148;
    long long max_wave_slots = 1ll << 24;       // min(kMaxWaveSlots, what a third of the free device memory holds), crt_context_create
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    std::vector<cudaEvent_t> wave_events;       // pairs bracketing traversal launches when cfg->time_kernels
    DevBuf<float> gauss_cdf;                    // GaussianFilter CDF tables (x then y), rebuilt when (rx, ry, sigma) change
    float gauss_key[3] = {0, 0, 0}, gauss_exp[2] = {0, 0};
    // Film::pixel_sensor: empty = the XYZ sensor (RayTracerTestApp.h:149); else r/g/b response curves (3 x 471), XYZFromSensorRGB, ratio
    std::vector<float> sensor_curves; float sensor_matrix[9] = {0}; float sensor_ratio = 1.0f / 106.856895f; unsigned sensor_gen = 0;
    std::vector<float> rgb_scale, rgb_data;     // RGBToSpectrumTable (zNodes[64], coeffs[3][64][64][64][3]); empty until set/generated
    // wave scratch (grow-only)
    DevBuf<float4> ray_o, ray_d, ray_k, ray_s, hit_tb, lambda, pdf, beta, L;
    DevBuf<int> hit_ref, pixel, flags, occluded, pixel_list, index_list, overflow_list, retrace_list;
    DevBuf<float> weight;
    DevBuf<SamplerState> sampler;
    DevBuf<int> counters;                // [0] work cursor, [1] overflow count, [2] work cursor (overflow pass), ...
    DevBuf<uint32_t> gqueue;
    DevBuf<unsigned long long> stats;    // [0..4] traversal statistics, [8] closest rays, [9] shadow rays, [10] depth sum
    // Tier B wavefront queues
    DevBuf<int> active_a, active_b, sh_path, qcount;      // qcount: [2*b] active count entering bounce b, [2*b+1] shadow count of bounce b
    DevBuf<float4> sh_o, sh_d, sh_k, sh_s, sh_contrib;
    // additional next-event slots (one per point / sun light, or per emissive triangle under light_strategy 1): `xq_slots` shadow queues of
    // wave capacity each, laid out slot-major, + their counters per bounce
    DevBuf<float4> xq_o, xq_d, xq_k, xq_s, xq_contrib;
    DevBuf<int> xq_path, xq_count;
    int xq_slots = 0; size_t xq_capacity = 0;
    int ensure_nee_slots(int slots, size_t n);
    // staged shading (crt_path.cuh): surface records of the active paths and one path-id queue per material type; mq_count[4*b + type]
    DevBuf<float4> hit_a, hit_b, hit_c;
    DevBuf<int> mq_ids, mq_count;
    size_t staged_capacity = 0;
    int ensure_staged(size_t n);
    int event_cursor = 0;
    size_t wave_capacity = 0;
    // CUDA graph of one full path-integrator wave (crt_render, small frames), and what it was captured for
    cudaGraphExec_t wave_graph = nullptr;
    WaveGraphKey wave_key;
    crt_render_stats wave_rs;
    DevBuf<int> d_cursor;
    unsigned wave_gen = 0;               // bumped whenever the wave buffers are reallocated
    void* nccl_comm = nullptr;           // ncclComm_t of crt_nccl_comm_create (one communicator per context = per GPU)
    int nccl_world = 1, nccl_rank = 0;
    int ensure_wave(size_t n, bool tier_b);
    PathBuffers path_buffers() {
        PathBuffers pb;
        pb.depth_sum = nullptr;
        pb.ray_o = ray_o.p; pb.ray_d = ray_d.p; pb.ray_k = ray_k.p; pb.ray_s = ray_s.p; pb.hit_ref = hit_ref.p; pb.hit_tb = hit_tb.p; pb.lambda = lambda.p; pb.pdf = pdf.p;
        pb.weight = weight.p; pb.pixel = pixel.p; pb.beta = beta.p; pb.L = L.p; pb.sampler = sampler.p; pb.flags = flags.p;
        return pb;
    }
};

static const int kGlobalQueueCap = 1 << 16;
static const int kMaxDepth = 64;
static const int kMaxSamplesPerWave = 64;                // small frames put more sample indices into a wave (fewer, fuller launches); 1080p: 8
static const int kRootLeafMaxTris = 16;                  // a one-leaf octree of at most this many triangles is traversed inside the shading kernels
static const int kMaxShapes = 65535;                     // closest_over_shapes_warp / occluded_by_shapes_warp: (lane << 16 | shape index) pool entries
static const int kMaxNeeSlots = 16;                      // point / sun lights + (light_strategy 1) emissive triangles sampled one each
// Path slots of a path-integrator wave.  A bounce ends with a tail in which a few long rays keep the GPU nearly idle; the more paths a
// wave holds, the smaller the share of those tails (C2 at 2^23 / 2^24 / 2^25 / 2^26 slots: 492 / 515 / 527 / 535 Mpaths/s, C3: 749 / 804 /
// 835 / 852).  2^26 slots are ~30 GB of wave state (kWaveBytesPerSlot), a sixth of a B200's HBM; a context on a smaller or busier device
// gets as many slots as fit a third of its free memory (crt_context_create).
#ifndef CRT_MAX_WAVE_SLOTS_LOG2
#define CRT_MAX_WAVE_SLOTS_LOG2 26
#endif
static const long long kMaxWaveSlots = 1ll << CRT_MAX_WAVE_SLOTS_LOG2;
static const long long kWaveBytesPerSlot = 448;
static const long long kGraphMaxSlots = 1ll << 21;      // waves of at most 2 M path slots are replayed from a CUDA graph (launch-bound regime)


int crt_context::ensure_nee_slots(int slots, size_t n) {
    if (slots <= xq_slots && n <= xq_capacity) return 0;
    slots = std::max(slots, xq_slots); n = std::max(n, xq_capacity);
    const size_t total = (size_t)slots * n;
    CRT_CUDA(xq_o.resize(total)); CRT_CUDA(xq_d.resize(total)); CRT_CUDA(xq_k.resize(total)); CRT_CUDA(xq_s.resize(total));
    CRT_CUDA(xq_contrib.resize(2 * total)); CRT_CUDA(xq_path.resize(total));
    CRT_CUDA(xq_count.resize((size_t)(kMaxDepth + 2) * kMaxNeeSlots));
    xq_slots = slots; xq_capacity = n;
    ++wave_gen;
    return 0;
}
int crt_context::ensure_staged(size_t n) {
    if (n <= staged_capacity) return 0;
    CRT_CUDA(hit_a.resize(n)); CRT_CUDA(hit_b.resize(n)); CRT_CUDA(hit_c.resize(n)); CRT_CUDA(mq_ids.resize(3 * n));
    CRT_CUDA(mq_count.resize((size_t)4 * (kMaxDepth + 2)));
    staged_capacity = n;
    ++wave_gen;
    return 0;
}
int crt_context::ensure_wave(size_t n, bool tier_b) {
    if (n > wave_capacity) {
        CRT_CUDA(ray_o.resize(n)); CRT_CUDA(ray_d.resize(n)); CRT_CUDA(ray_k.resize(n)); CRT_CUDA(ray_s.resize(n)); CRT_CUDA(hit_tb.resize(n)); CRT_CUDA(hit_ref.resize(n));
        CRT_CUDA(lambda.resize(2 * n)); CRT_CUDA(pdf.resize(2 * n)); CRT_CUDA(weight.resize(n)); CRT_CUDA(pixel.resize(n));
        CRT_CUDA(occluded.resize(n)); CRT_CUDA(overflow_list.resize(n)); CRT_CUDA(retrace_list.resize(n));
        wave_capacity = n;
        ++wave_gen;
    }
    if (tier_b && beta.n < 2 * n) {
        CRT_CUDA(beta.resize(2 * n)); CRT_CUDA(L.resize(2 * n)); CRT_CUDA(sampler.resize(n)); CRT_CUDA(flags.resize(n));
        CRT_CUDA(active_a.resize(n)); CRT_CUDA(active_b.resize(n)); CRT_CUDA(sh_path.resize(n));
        CRT_CUDA(sh_o.resize(n)); CRT_CUDA(sh_d.resize(n)); CRT_CUDA(sh_k.resize(n)); CRT_CUDA(sh_s.resize(n)); CRT_CUDA(sh_contrib.resize(2 * n));
        ++wave_gen;
    }
    if (!counters.p) { CRT_CUDA(counters.resize(8 * (2 + kMaxNeeSlots) * (kMaxDepth + 2) + 8)); CRT_CUDA(cudaMemset(counters.p, 0, counters.bytes())); CRT_CUDA(stats.resize(16)); CRT_CUDA(qcount.resize(2 * (kMaxDepth + 2))); }
    return 0;
}

struct crt_scene {
    crt_context* ctx = nullptr;
    // host staging
    std::vector<float> h_nodes;
    std::vector<uint32_t> h_leaf_refs, h_pk_refs;
    std::vector<float> h_pk_boxes, h_node_tight;
    std::vector<float> h_tris;        // 12 floats per triangle
    std::vector<float> h_tri_nrm;     // 12 floats per triangle or empty
    std::vector<float> h_tri_uv, h_tri_tan, h_tri_bitan;   // 8 / 12 / 12 floats per triangle or empty (MeshCache::Mesh texcoords, tangents, bitangents)
    std::vector<uint32_t> mesh_first;
    std::vector<int32_t> mesh_material;
    std::vector<DevShape> h_shapes;
    std::vector<DevShapeBox> h_shape_boxes;
    std::vector<DevMaterial> h_materials;
    std::vector<DevSpectrum> h_spectra;
    std::vector<float> h_pool;
    std::vector<DevLight> h_lights;
    std::vector<DevDeltaLight> h_delta;
    DevBuf<DevDeltaLight> d_delta;
    std::vector<float> h_light_cdf;
    std::vector<int32_t> h_light_pairs;
    float light_total = 0;
    bool has_model = false, committed = false;
    int root_leaf_tris = 0;          // > 0: the octree is its root leaf with this many (at most kRootLeafMaxTris) listed triangles
    unsigned commit_gen = 0;            // bumped by every crt_scene_commit (captured CUDA graphs hold the scene's device pointers)
    int retransform = 0;
    float model_o2r[16];
    int octree_depth = 0;
    // device
    DevBuf<float4> d_nodes, d_tris, d_tri_nrm, d_tri_uv, d_tri_tan, d_tri_bitan;
    DevBuf<uint32_t> d_leaf_refs, d_pk_refs;
    DevBuf<float4> d_pk_boxes, d_node_tight;
    DevBuf<DevShape> d_shapes;
    DevBuf<DevShapeBox> d_shape_boxes;
    DevBuf<float4> d_shape_bvh;
    std::vector<float> h_shape_bvh;         // 8 floats per node, depth-first (crt_device_scene.h)
    DevBuf<DevMaterial> d_materials;
    DevBuf<DevSpectrum> d_spectra;
    DevBuf<float> d_pool, d_light_cdf, d_tables, d_color;
    unsigned sensor_gen = 0;            // context sensor generation captured by the last commit
    DevBuf<DevLight> d_lights;
    DeviceScene view;
    // the big staging vectors are page-locked so that crt_scene_commit's uploads run at PCIe speed
    std::vector<void*> pinned;
    void unpin() { for (void* p : pinned) cudaHostUnregister(p); pinned.clear(); }
    template <typename T> void pin(std::vector<T>& v) {
        if (v.empty()) return;
        if (cudaHostRegister(v.data(), v.size() * sizeof(T), cudaHostRegisterDefault) == cudaSuccess) pinned.push_back(v.data());
        else cudaGetLastError();          // pageable uploads still work
    }
};

struct crt_film {
    crt_context* ctx = nullptr;
    int width = 0, height = 0;
    DevBuf<float4> own;
    float4* data = nullptr;
    DevBuf<unsigned char> rgb8;
    DevBuf<float> rgbf;
};

extern "C" {

// ================================================================ context =================================
int crt_context_create(int device, crt_context** out) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        set_error(std::string("no CUDA device: ") + cudaGetErrorString(e) + " (libcrt_b200 has no CPU fallback)");
        return 2;
    }
    if (device < 0 || device >= count) { set_error("context_create: bad device index"); return 1; }
    CRT_CUDA(cudaSetDevice(device));
    std::unique_ptr<crt_context, void (*)(crt_context*)> c(new crt_context, crt_context_destroy);      // released on every error path below
    c->device = device;
    cudaDeviceProp prop;
    CRT_CUDA(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    size_t free_b = 0, total_b = 0;
    CRT_CUDA(cudaMemGetInfo(&free_b, &total_b));
    c->max_wave_slots = std::max<long long>(1ll << 20, std::min<long long>(kMaxWaveSlots, (long long)(free_b / 3) / kWaveBytesPerSlot));
    CRT_CUDA(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    for (auto& ev : c->ev) CRT_CUDA(cudaEventCreate(&ev));
    *out = c.release();
    return 0;
}
void crt_context_destroy(crt_context* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (auto& ev : c->ev) if (ev) cudaEventDestroy(ev);
    for (auto& ev : c->wave_events) cudaEventDestroy(ev);
    if (c->wave_graph) cudaGraphExecDestroy(c->wave_graph);
    crt_nccl_comm_destroy(c);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}
int crt_context_synchronize(crt_context* c) { CRT_CUDA(cudaSetDevice(c->device)); CRT_CUDA(cudaStreamSynchronize(c->stream)); return 0; }
int crt_context_set_stream(crt_context* c, void* s) { c->stream = s ? (cudaStream_t)s : c->own_stream; return 0; }

// ================================================================ GPU octree build ========================
// Octtree_Model::CreateOcttree on the device (crt_build.cuh): same tree, same flattened layout as crt_octree_build;
// node ids are breadth-first (like algorithm 1 of crt_octree_set_build_algorithm).
static int octree_build_gpu_impl(crt_context* c, const crt_mesh_desc* meshes, uint32_t n_meshes, const float* o2r, int precomputed_world, crt_octree** out);
int crt_octree_build_gpu(crt_context* c, const crt_mesh_desc* meshes, uint32_t n_meshes, const float* o2r, int precomputed_world, crt_octree** out) {
    try {           // thrust reports CUDA failures by throwing: nothing may unwind through the C ABI
        return octree_build_gpu_impl(c, meshes, n_meshes, o2r, precomputed_world, out);
    } catch (const std::exception& e) {
        set_error(std::string("octree_build_gpu: ") + e.what());
        return 2;
    }
}
static int octree_build_gpu_impl(crt_context* c, const crt_mesh_desc* meshes, uint32_t n_meshes, const float* o2r, int precomputed_world, crt_octree** out) {
    if (!c || !out) { set_error("octree_build_gpu: bad arguments"); return 1; }
    CRT_CUDA(cudaSetDevice(c->device));
    std::unique_ptr<crt_octree> oct(crt::octree_prepare(meshes, n_meshes, o2r, precomputed_world));
    if (!oct) return 1;
    cudaStream_t st = c->stream;
    const unsigned n_tris = oct->mesh_first[n_meshes];
    DevBuf<float> d_pos;
    CRT_CUDA(d_pos.upload(reinterpret_cast<const float*>(oct->world_pos.data()), 9 * (size_t)n_tris, st));
    BuildNode root;
    for (int a = 0; a < 3; ++a) { root.lo[a] = oct->nodes[0].bmin[a]; root.hi[a] = oct->nodes[0].bmax[a]; }
    root.tau = -1; root.start = 0; root.count = 0;
    DevBuf<unsigned> refs[2], ref_node[2], flags, offs, split_flag, split_rank, child_cnt, child_start;
    DevBuf<long long> split_gid;
    DevBuf<unsigned char> masks;
    DevBuf<BuildNode> nodes[2];
    auto policy = thrust::cuda::par.on(st);
    // root: the triangles that overlap the model's bounds at all (AddTriangle's first test, Octtree_Model.h:206-212)
    CRT_CUDA(flags.resize(n_tris + 1)); CRT_CUDA(offs.resize(n_tris + 1));
    CRT_CUDA(cudaMemsetAsync(flags.p, 0, (n_tris + 1) * sizeof(unsigned), st));
    if (n_tris) k_build_root_filter<<<cdiv((int)n_tris, 256), 256, 0, st>>>(d_pos.p, n_tris, root, flags.p);
    thrust::exclusive_scan(policy, flags.p, flags.p + n_tris + 1, offs.p);
    unsigned n_refs = 0;
    CRT_CUDA(cudaMemcpyAsync(&n_refs, offs.p + n_tris, sizeof n_refs, cudaMemcpyDeviceToHost, st));
    CRT_CUDA(cudaStreamSynchronize(st));
    int cur = 0;
    CRT_CUDA(refs[0].resize(std::max(n_refs, 1u))); CRT_CUDA(ref_node[0].resize(std::max(n_refs, 1u)));
    if (n_tris) k_build_root_compact<<<cdiv((int)n_tris, 256), 256, 0, st>>>(flags.p, offs.p, n_tris, refs[0].p, ref_node[0].p);
    root.count = n_refs;
    CRT_CUDA(nodes[0].upload(&root, 1, st));
    unsigned n_nodes = 1;
    std::vector<BuildNode> h_nodes;
    std::vector<unsigned> h_flag, h_rank, h_refs;
    std::vector<int> parent_of_cur, parent_of_next;
    oct->nodes.clear();
    size_t level_base = 0;
    const int kMaxBuildLevels = 64;
    for (int level = 0; level < kMaxBuildLevels; ++level) {
        CRT_CUDA(masks.resize(std::max(n_refs, 1u)));
        CRT_CUDA(split_flag.resize(n_nodes + 1)); CRT_CUDA(split_rank.resize(n_nodes + 1)); CRT_CUDA(split_gid.resize(n_nodes));
        CRT_CUDA(child_cnt.resize(8 * (size_t)n_nodes + 1)); CRT_CUDA(child_start.resize(8 * (size_t)n_nodes + 1));
        CRT_CUDA(cudaMemsetAsync(split_flag.p + n_nodes, 0, sizeof(unsigned), st));
        CRT_CUDA(cudaMemsetAsync(child_cnt.p + 8 * (size_t)n_nodes, 0, sizeof(unsigned), st));
        if (n_refs) k_build_masks<<<cdiv((int)n_refs, 256), 256, 0, st>>>(nodes[cur].p, refs[cur].p, ref_node[cur].p, n_refs, d_pos.p, masks.p);
        k_build_decide<<<cdiv((int)n_nodes, 8), 256, 0, st>>>(nodes[cur].p, n_nodes, refs[cur].p, masks.p, split_flag.p, split_gid.p, child_cnt.p);
        CRT_CUDA(cudaGetLastError());
        thrust::exclusive_scan(policy, split_flag.p, split_flag.p + n_nodes + 1, split_rank.p);
        thrust::exclusive_scan(policy, child_cnt.p, child_cnt.p + 8 * (size_t)n_nodes + 1, child_start.p);
        unsigned n_split = 0, next_refs = 0;
        CRT_CUDA(cudaMemcpyAsync(&n_split, split_rank.p + n_nodes, sizeof n_split, cudaMemcpyDeviceToHost, st));
        CRT_CUDA(cudaMemcpyAsync(&next_refs, child_start.p + 8 * (size_t)n_nodes, sizeof next_refs, cudaMemcpyDeviceToHost, st));
        // this level goes back to the host octree (breadth-first numbering)
        h_nodes.resize(n_nodes); h_flag.resize(n_nodes); h_rank.resize(n_nodes); h_refs.resize(std::max(n_refs, 1u));
        CRT_CUDA(cudaMemcpyAsync(h_nodes.data(), nodes[cur].p, n_nodes * sizeof(BuildNode), cudaMemcpyDeviceToHost, st));
        CRT_CUDA(cudaMemcpyAsync(h_flag.data(), split_flag.p, n_nodes * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
        CRT_CUDA(cudaMemcpyAsync(h_rank.data(), split_rank.p, n_nodes * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
        if (n_refs) CRT_CUDA(cudaMemcpyAsync(h_refs.data(), refs[cur].p, n_refs * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
        CRT_CUDA(cudaStreamSynchronize(st));
        if (n_split) {
            const int nx = cur ^ 1;
            CRT_CUDA(nodes[nx].resize(8 * (size_t)n_split)); CRT_CUDA(refs[nx].resize(std::max(next_refs, 1u))); CRT_CUDA(ref_node[nx].resize(std::max(next_refs, 1u)));
            k_build_scatter<<<cdiv((int)n_nodes, 8), 256, 0, st>>>(nodes[cur].p, n_nodes, refs[cur].p, masks.p, split_flag.p, split_rank.p, split_gid.p,
                                                                    child_cnt.p, child_start.p, nodes[nx].p, refs[nx].p, ref_node[nx].p);
            CRT_CUDA(cudaGetLastError());
        }
        // (host work below overlaps the scatter kernel)
        const size_t next_base = level_base + n_nodes;
        oct->nodes.resize(next_base);
        for (unsigned j = 0; j < n_nodes; ++j) {
            HostOctreeNode& hn = oct->nodes[level_base + j];
            for (int a = 0; a < 3; ++a) { hn.bmin[a] = h_nodes[j].lo[a]; hn.bmax[a] = h_nodes[j].hi[a]; }
            if (h_flag[j]) {
                hn.leaf = false;
                for (int k = 0; k < 8; ++k) hn.child[k] = (int)(next_base + 8 * (size_t)h_rank[j] + k);
            } else {
                hn.leaf = true;
                hn.tris.assign(h_refs.begin() + h_nodes[j].start, h_refs.begin() + h_nodes[j].start + h_nodes[j].count);
            }
        }
        for (unsigned j = 0; j < n_nodes; ++j)
            if (h_flag[j]) for (int k = 0; k < 8; ++k) parent_of_next.push_back((int)(level_base + j));
        if (level > 0) for (unsigned j = 0; j < n_nodes; ++j) oct->nodes[level_base + j].parent = parent_of_cur[j];
        else oct->nodes[0].parent = 0;
        parent_of_cur.swap(parent_of_next); parent_of_next.clear();
        if (!n_split) break;
        if (level == kMaxBuildLevels - 1) { set_error("octree_build_gpu: the tree is deeper than 64 levels"); return 1; }
        level_base = next_base;
        cur ^= 1; n_nodes = 8 * n_split; n_refs = next_refs;
    }
    CRT_CUDA(cudaStreamSynchronize(st));
    *out = oct.release();
    return 0;
}

// ================================================================ scene ===================================
int crt_scene_create(crt_context* ctx, crt_scene** out) {
    if (!ctx) { set_error("scene_create: null context"); return 1; }
    auto* s = new crt_scene;
    s->ctx = ctx;
    m4_identity(s->model_o2r);
    std::memset(&s->view, 0, sizeof s->view);
    *out = s;
    return 0;
}
void crt_scene_destroy(crt_scene* s) {
    if (!s) return;
    cudaSetDevice(s->ctx->device);
    cudaStreamSynchronize(s->ctx->stream);
    s->unpin();
    delete s;
}

int crt_scene_set_model(crt_scene* s, const crt_mesh_desc* meshes, uint32_t n_meshes, const float* o2r, int precomputed_world,
                        const uint8_t* const* cull_bits, const crt_octree* oct, const int32_t* mesh_materials) {
    if (!s || !meshes || !oct) { set_error("scene_set_model: bad arguments"); return 1; }
    if (oct->mesh_first.size() != n_meshes + 1) { set_error("scene_set_model: octree was built for a different model"); return 1; }
    CRT_CUDA(cudaSetDevice(s->ctx->device));
    CRT_CUDA(cudaStreamSynchronize(s->ctx->stream));
    s->unpin();
    const uint32_t total = oct->mesh_first[n_meshes];
    s->mesh_first = oct->mesh_first;
    s->mesh_material.assign(n_meshes, 0);
    if (mesh_materials) s->mesh_material.assign(mesh_materials, mesh_materials + n_meshes);
    std::memcpy(s->model_o2r, o2r, 64);
    s->retransform = precomputed_world ? 1 : 0;
    // an attribute is available (Triangle::vertex_available, Shapes.h:917-924: one flag set per TriModel) when every mesh carries it
    bool all_normals = true, all_uv = true, all_tan = true, all_bitan = true;
    for (uint32_t m = 0; m < n_meshes; ++m) {
        all_normals = all_normals && meshes[m].normals != nullptr; all_uv = all_uv && meshes[m].texcoords != nullptr;
        all_tan = all_tan && meshes[m].tangents != nullptr; all_bitan = all_bitan && meshes[m].bitangents != nullptr;
    }
    s->h_tri_uv.clear(); s->h_tri_tan.clear(); s->h_tri_bitan.clear();
    if (all_uv) s->h_tri_uv.assign(8 * (size_t)total, 0.0f);
    if (all_tan) s->h_tri_tan.assign(12 * (size_t)total, 0.0f);
    if (all_bitan) s->h_tri_bitan.assign(12 * (size_t)total, 0.0f);
    s->h_tris.resize(12 * (size_t)total);
    s->h_tri_nrm.clear();
    if (all_normals) s->h_tri_nrm.resize(12 * (size_t)total);
    std::vector<uint8_t> skip(total, 0);
    for (uint32_t m = 0; m < n_meshes; ++m)
        for (uint32_t t = 0; t < meshes[m].n_triangles; ++t) {
            const size_t gid = (size_t)oct->mesh_first[m] + t;
            const f3* w = &oct->world_pos[3 * gid];
            float* d = &s->h_tris[12 * gid];
            int32_t tags[3] = {s->mesh_material[m], (int32_t)m, (int32_t)t};
            for (int k = 0; k < 3; ++k) {
                d[4 * k] = w[k].x; d[4 * k + 1] = w[k].y; d[4 * k + 2] = w[k].z;
                std::memcpy(&d[4 * k + 3], &tags[k], 4);
            }
            if (all_normals) {
                float* nn = &s->h_tri_nrm[12 * gid];
                for (int k = 0; k < 3; ++k) {
                    const float* src = &meshes[m].normals[3 * (size_t)meshes[m].indices[3 * (size_t)t + k]];
                    nn[4 * k] = src[0]; nn[4 * k + 1] = src[1]; nn[4 * k + 2] = src[2]; nn[4 * k + 3] = 0;
                }
            }
            for (int k = 0; k < 3; ++k) {
                const size_t vi = meshes[m].indices[3 * (size_t)t + k];
                if (all_uv) { s->h_tri_uv[8 * gid + 2 * k] = meshes[m].texcoords[2 * vi]; s->h_tri_uv[8 * gid + 2 * k + 1] = meshes[m].texcoords[2 * vi + 1]; }
                if (all_tan) for (int a = 0; a < 3; ++a) s->h_tri_tan[12 * gid + 4 * k + a] = meshes[m].tangents[3 * vi + a];
                if (all_bitan) for (int a = 0; a < 3; ++a) s->h_tri_bitan[12 * gid + 4 * k + a] = meshes[m].bitangents[3 * vi + a];
            }
            // back-face-culled triangles are skipped by the traversal loop (Octtree_Model.h:91-96) ...
            if (cull_bits && cull_bits[m] && cull_bits[m][t]) skip[gid] = 1;
            // ... and degenerate ones always miss (Shapes.h:1131-1134): length(cross(p2-p0, p1-p0)) == 0
            f3 c = cross3(w[2] - w[0], w[1] - w[0]);
            if (sqrtf(dot3(c, c)) == 0) skip[gid] = 1;
        }
    FlatOctree flat;
    oct->flatten(skip, &flat);
    s->h_nodes.swap(flat.nodes);
    s->h_leaf_refs.swap(flat.leaf_refs);
    s->h_pk_boxes.swap(flat.pk_boxes);
    s->h_node_tight.swap(flat.node_tight);
    s->h_pk_refs.swap(flat.pk_refs);
    s->octree_depth = flat.depth;
    s->pin(s->h_nodes); s->pin(s->h_leaf_refs); s->pin(s->h_pk_boxes); s->pin(s->h_node_tight); s->pin(s->h_pk_refs); s->pin(s->h_tris); s->pin(s->h_tri_nrm); s->pin(s->h_tri_uv); s->pin(s->h_tri_tan); s->pin(s->h_tri_bitan);
    s->has_model = true;
    s->committed = false;
    return 0;
}

// The shape constructors (Shapes.h:220-231 Sphere, :459-466 Cylinder, :648-655 Disk, :771-777 TriangleSimple) and Shape's transform convention (:175-182)
static int shape_from_params(int kind, const float* rigid16, const float* p, int material, DevShape& sh) {
    if (kind < 0 || kind > 3) { set_error("shape: unknown kind (0 Sphere, 1 Cylinder, 2 Disk, 3 TriangleSimple)"); return 1; }
    if (!rigid16 || !p) { set_error("shape: null argument"); return 1; }
    std::memset(&sh, 0, sizeof sh);
    sh.kind = kind; sh.material = material;
    shape_matrices(rigid16, sh.o2r, sh.r2o);
    normal_matrix(sh.o2r, sh.nmat);
    const float deg2rad = 0.01745329251994329576923690768489f;
    if (kind == SHAPE_SPHERE) {            // Sphere ctor, Shapes.h:220-231
        float r = p[0];
        float zmin = gclamp(p[1], -r, r), zmax = gclamp(p[2], -r, r);
        sh.p[0] = r; sh.p[1] = zmin; sh.p[2] = zmax;
        sh.p[3] = acosf(gclamp(zmin / r, -1.f, 1.f));
        sh.p[4] = acosf(gclamp(zmax / r, -1.f, 1.f));
        sh.p[5] = gclamp(p[3], 0.0f, 360.f) * deg2rad;
    } else if (kind == SHAPE_CYLINDER) {   // Shapes.h:459-466
        sh.p[0] = p[0]; sh.p[1] = p[1]; sh.p[2] = p[2]; sh.p[3] = p[3] * deg2rad;
    } else if (kind == SHAPE_DISK) {       // Shapes.h:648-655
        sh.p[0] = p[0]; sh.p[1] = p[1]; sh.p[2] = p[2]; sh.p[3] = p[3] * deg2rad;
    } else {
        for (int i = 0; i < 9; ++i) sh.p[i] = p[i];
    }
    return 0;
}
// Shape::Area (Shapes.h:234-237 Sphere, :455-458 Cylinder, :642-645 Disk, :779-782 TriangleSimple), host arithmetic
int crt_shape_area(int kind, const float* params9, float* out) {
    DevShape sh;
    float id[16];
    m4_identity(id);
    if (!out) { set_error("shape_area: null output"); return 1; }
    if (int e = shape_from_params(kind, id, params9, 0, sh)) return e;
    if (kind == SHAPE_SPHERE) *out = sh.p[5] * sh.p[0] * (sh.p[2] - sh.p[1]);
    else if (kind == SHAPE_CYLINDER) *out = (sh.p[2] - sh.p[1]) * sh.p[0] * sh.p[3];
    else if (kind == SHAPE_DISK) *out = sh.p[3] * .5f * (sh.p[2] * sh.p[2] - sh.p[1] * sh.p[1]);
    else {
        f3 p1 = mk3(sh.p[0], sh.p[1], sh.p[2]), p2 = mk3(sh.p[3], sh.p[4], sh.p[5]), p3 = mk3(sh.p[6], sh.p[7], sh.p[8]);
        *out = 0.5f * length3(cross3(p2 - p1, p3 - p1));
    }
    return 0;
}
// Shape::Bounds = TransformBounds(object-space box, ObjectToRender) (Shapes.h:239-242, :459-462, :647-650, :784-792; Bounds3::Transform :60-98,
// whose max side starts at FLT_MIN)
int crt_shape_bounds(int kind, const float* rigid16, const float* params9, float* out_min3_max3) {
    DevShape sh;
    if (!out_min3_max3) { set_error("shape_bounds: null output"); return 1; }
    if (int e = shape_from_params(kind, rigid16, params9, 0, sh)) return e;
    float lo[3], hi[3];
    if (kind == SHAPE_SPHERE || kind == SHAPE_CYLINDER) { lo[0] = lo[1] = -sh.p[0]; hi[0] = hi[1] = sh.p[0]; lo[2] = sh.p[1]; hi[2] = sh.p[2]; }
    else if (kind == SHAPE_DISK) { lo[0] = lo[1] = -sh.p[2]; hi[0] = hi[1] = sh.p[2]; lo[2] = hi[2] = sh.p[0]; }
    else for (int a = 0; a < 3; ++a) { lo[a] = gmin(gmin(sh.p[a], sh.p[3 + a]), sh.p[6 + a]); hi[a] = gmax(gmax(sh.p[a], sh.p[3 + a]), sh.p[6 + a]); }
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {FLT_MIN, FLT_MIN, FLT_MIN};
    const float cx[2] = {lo[0], hi[0]}, cy[2] = {lo[1], hi[1]}, cz[2] = {lo[2], hi[2]};
    for (int i = 0; i < 8; ++i) {
        const f3 q = xform_point(sh.o2r, mk3(cx[i & 1], cy[(i >> 1) & 1], cz[(i >> 2) & 1]));
        for (int a = 0; a < 3; ++a) { mn[a] = std::min(mn[a], comp(q, a)); mx[a] = std::max(mx[a], comp(q, a)); }
    }
    for (int a = 0; a < 3; ++a) { out_min3_max3[a] = mn[a]; out_min3_max3[3 + a] = mx[a]; }
    return 0;
}

int crt_scene_add_shape(crt_scene* s, int kind, const float* rigid16, const float* p, int material, int* out_id) {
    DevShape sh;
    if (int e = shape_from_params(kind, rigid16, p, material, sh)) return e;
    if (s->h_shapes.size() >= (size_t)kMaxShapes) {
        set_error("scene_add_shape: at most " + std::to_string(kMaxShapes) + " analytic shapes per scene (the pooled shape tests pack the shape index into 16 bits)");
        return 1;
    }
    s->h_shapes.push_back(sh);
    {   // padded world bounds: the 8 corners of the object-space box of the FULL shape (clipping only removes surface)
        float olo[3], ohi[3];
        if (kind == SHAPE_SPHERE) { for (int a = 0; a < 3; ++a) { olo[a] = -sh.p[0]; ohi[a] = sh.p[0]; } }
        else if (kind == SHAPE_CYLINDER) { olo[0] = olo[1] = -sh.p[0]; ohi[0] = ohi[1] = sh.p[0]; olo[2] = std::min(sh.p[1], sh.p[2]); ohi[2] = std::max(sh.p[1], sh.p[2]); }
        else if (kind == SHAPE_DISK) { olo[0] = olo[1] = -sh.p[2]; ohi[0] = ohi[1] = sh.p[2]; olo[2] = ohi[2] = sh.p[0]; }
        else {
            for (int a = 0; a < 3; ++a) { olo[a] = std::min(sh.p[a], std::min(sh.p[3 + a], sh.p[6 + a])); ohi[a] = std::max(sh.p[a], std::max(sh.p[3 + a], sh.p[6 + a])); }
        }
        float wlo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, whi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX}, mag = 0;
        for (int i = 0; i < 8; ++i) {
            f3 w = xform_point(sh.o2r, mk3(i & 1 ? ohi[0] : olo[0], i & 2 ? ohi[1] : olo[1], i & 4 ? ohi[2] : olo[2]));
            for (int a = 0; a < 3; ++a) { wlo[a] = std::min(wlo[a], comp(w, a)); whi[a] = std::max(whi[a], comp(w, a)); mag = std::max(mag, std::fabs(comp(w, a))); }
        }
        const float ext = std::max(whi[0] - wlo[0], std::max(whi[1] - wlo[1], whi[2] - wlo[2]));
        const float pad = std::max(mag, ext) * 0x1p-10f + 1e-4f;          // ~1e-3 relative: far beyond the intersection routines' rounding
        DevShapeBox bx;
        bx.lo = make_float4(wlo[0] - pad, wlo[1] - pad, wlo[2] - pad, 0);
        bx.hi = make_float4(whi[0] + pad, whi[1] + pad, whi[2] + pad, 0);
        s->h_shape_boxes.push_back(bx);
    }
    s->committed = false;
    if (out_id) *out_id = (int)s->h_shapes.size() - 1;
    return 0;
}

static int add_piecewise(crt_scene* s, const PiecewiseLinear& pl) {
    DevSpectrum sp;
    std::memset(&sp, 0, sizeof sp);
    sp.kind = SPEC_PIECEWISE;
    sp.offset = (int)s->h_pool.size();
    sp.n = (int)pl.lambdas.size();
    s->h_pool.insert(s->h_pool.end(), pl.lambdas.begin(), pl.lambdas.end());
    s->h_pool.insert(s->h_pool.end(), pl.values.begin(), pl.values.end());
    // per-nanometre interval index (crt_device_scene.h): entry b = the last knot at or below (lut_base + b) nm, 0 if there is none
    if (sp.n >= 2) {
        const int base = (int)std::floor(pl.lambdas.front()), count = (int)std::floor(pl.lambdas.back()) - base + 1;
        if (count > 0 && count <= 8192) {
            sp.lut_base = base; sp.lut_n = count;
            int o = 0;
            for (int b = 0; b < count; ++b) {
                const float edge = (float)(base + b);
                while (o + 1 < sp.n && pl.lambdas[o + 1] <= edge) ++o;
                float bits;
                std::memcpy(&bits, &o, 4);
                s->h_pool.push_back(bits);
            }
        }
    }
    s->h_spectra.push_back(sp);
    return (int)s->h_spectra.size() - 1;
}
// RGBToSpectrumTable::operator() for uniform rgb (color.cpp:35-37); the 64^3 table file is not part of the reference repo
static bool grey_sigmoid(float g, float* c) {
    c[0] = 0; c[1] = 0;
    c[2] = (g - .5f) / sqrtf(g * (1 - g));
    return true;
}

// ---------------------------------------------------------------- film sensor ------------------------------------------
int crt_context_set_sensor(crt_context* c, const float* r471, const float* g471, const float* b471, const float* illum471, float imaging_ratio,
                           float* matrix9_out) {
    if (!c) { set_error("set_sensor: null context"); return 1; }
    if (!r471) {                                    // back to PixelSensor(sRGB, stdillum-D65, 1 / CIE_Y_integral)
        c->sensor_curves.clear(); c->sensor_ratio = 1.0f / 106.856895f; ++c->sensor_gen;
        if (matrix9_out) std::memcpy(matrix9_out, host_spectra().XYZFromSensorRGB, 36);
        return 0;
    }
    if (!g471 || !b471 || !illum471) { set_error("set_sensor: need r, g, b response curves and the sensor illuminant (471 floats each, 360..830 nm)"); return 1; }
    c->sensor_curves.assign(r471, r471 + 471);
    c->sensor_curves.insert(c->sensor_curves.end(), g471, g471 + 471);
    c->sensor_curves.insert(c->sensor_curves.end(), b471, b471 + 471);
    measured_sensor_matrix(r471, g471, b471, illum471, c->sensor_matrix);
    c->sensor_ratio = imaging_ratio;
    ++c->sensor_gen;
    if (matrix9_out) std::memcpy(matrix9_out, c->sensor_matrix, 36);
    return 0;
}

int crt_measured_sensor_matrix(const float* r471, const float* g471, const float* b471, const float* illum471, float* matrix9_out) {
    if (!r471 || !g471 || !b471 || !illum471 || !matrix9_out) { set_error("measured_sensor_matrix: null argument"); return 1; }
    measured_sensor_matrix(r471, g471, b471, illum471, matrix9_out);
    return 0;
}

// ---------------------------------------------------------------- RGB -> spectrum table --------------------------------
// RGBColorSpace::ToRGBCoeffs (colorspace.cpp:38-43): ClampZero, then RGBToSpectrumTable::operator()
static int rgb_coeffs(crt_context* c, const float* rgb_in, float* cc) {
    float rgb[3] = {std::max(0.0f, rgb_in[0]), std::max(0.0f, rgb_in[1]), std::max(0.0f, rgb_in[2])};
    if (rgb[0] == rgb[1] && rgb[1] == rgb[2]) { grey_sigmoid(rgb[0], cc); return 0; }
    if (!c || c->rgb_data.empty()) {
        set_error("non-grey RGB needs the sRGB spectrum table, which the reference repository does not contain (color.cpp:114): "
                  "call crt_rgb2spec_generate(ctx, ...) or crt_rgb2spec_set(ctx, ...) first");
        return 1;
    }
    rgb2spec_lookup(c->rgb_scale.data(), c->rgb_data.data(), rgb, cc);
    return 0;
}

static void rgb2spec_model(Rgb2SpecModel& M) {
    const HostSpectra& h = host_spectra();
    for (int k = 0; k < 3; ++k) M.white[k] = 0;
    for (int i = 0; i < kRgb2SpecSamples; ++i) {
        const double xyz[3] = {h.X[i], h.Y[i], h.Z[i]}, I = h.D65dense[i];
        for (int k = 0; k < 3; ++k) {
            double v = 0;
            for (int j = 0; j < 3; ++j) v += (double)h.RGBFromXYZ[3 * j + k] * xyz[j];
            M.rgb_tbl[k][i] = v * I;
            M.white[k] += xyz[k] * I;
        }
    }
    // the illuminant is normalised to unit luminance (sum of ybar * I = 1): a perfect reflector has Y = 1, RGB = (1, 1, 1)
    const double yn = M.white[1];
    for (int k = 0; k < 3; ++k) {
        M.white[k] /= yn;
        for (int i = 0; i < kRgb2SpecSamples; ++i) M.rgb_tbl[k][i] /= yn;
    }
    for (int i = 0; i < 9; ++i) M.rgb_to_xyz[i] = h.XYZFromRGB[i];
}

__global__ void __launch_bounds__(256) k_rgb2spec(const Rgb2SpecModel* __restrict__ M, float* __restrict__ data) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int res = kRgb2SpecRes;
    if (warp >= 3 * res * res) return;
    r2s_chain<32>(*M, warp / (res * res), (warp / res) % res, warp % res, lane, data);
}

int crt_rgb2spec_generate(crt_context* c, float* scale_out, float* data_out, float* ms_out) {
    if (!c) { set_error("rgb2spec_generate: needs a context (the table is built on the GPU)"); return 1; }
    CRT_CUDA(cudaSetDevice(c->device));
    Rgb2SpecModel model;
    rgb2spec_model(model);
    DevBuf<Rgb2SpecModel> d_model;
    DevBuf<float> d_data;
    const size_t n = (size_t)3 * 64 * 64 * 64 * 3;
    CRT_CUDA(d_model.upload(&model, 1, c->stream));
    CRT_CUDA(d_data.resize(n));
    cudaEvent_t e0 = c->ev[2], e1 = c->ev[3];              // the context's own events: nothing to release on an error path
    CRT_CUDA(cudaEventRecord(e0, c->stream));
    const int warps = 3 * 64 * 64;
    k_rgb2spec<<<(warps * 32 + 255) / 256, 256, 0, c->stream>>>(d_model.p, d_data.p);
    CRT_CUDA(cudaGetLastError());
    CRT_CUDA(cudaEventRecord(e1, c->stream));
    c->rgb_data.resize(n); c->rgb_scale.resize(64);
    CRT_CUDA(cudaMemcpyAsync(c->rgb_data.data(), d_data.p, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CRT_CUDA(cudaStreamSynchronize(c->stream));
    float ms = 0;
    CRT_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    for (int k = 0; k < 64; ++k) c->rgb_scale[k] = (float)r2s_scale(k);
    if (scale_out) std::memcpy(scale_out, c->rgb_scale.data(), 64 * sizeof(float));
    if (data_out) std::memcpy(data_out, c->rgb_data.data(), n * sizeof(float));
    if (ms_out) *ms_out = ms;
    return 0;
}
int crt_rgb2spec_set(crt_context* c, const float* scale, const float* data) {
    if (!c || !scale || !data) { set_error("rgb2spec_set: null argument"); return 1; }
    for (int k = 1; k < 64; ++k) if (!(scale[k] > scale[k - 1])) { set_error("rgb2spec_set: scale must increase strictly"); return 1; }
    c->rgb_scale.assign(scale, scale + 64);
    c->rgb_data.assign(data, data + (size_t)3 * 64 * 64 * 64 * 3);
    return 0;
}
int crt_rgb2spec_lookup(const float* scale, const float* data, const float* rgb, float* cc) {
    if (!scale || !data || !rgb || !cc) { set_error("rgb2spec_lookup: null argument"); return 1; }
    rgb2spec_lookup(scale, data, rgb, cc);
    return 0;
}
int crt_rgb2spec_fit(const float* rgb, float* cc) {
    if (!rgb || !cc) { set_error("rgb2spec_fit: null argument"); return 1; }
    static Rgb2SpecModel model;
    static bool ready = false;
    if (!ready) { rgb2spec_model(model); ready = true; }
    double c[3] = {0, 0, 0}, t[3] = {rgb[0], rgb[1], rgb[2]};
    r2s_gauss_newton<1>(model, t, c, 0);
    r2s_to_nm(c, cc);
    return 0;
}
// color.cpp:107-158: 4 bytes (printed, otherwise unused), 64 floats of scale, then float[3][64][64][64][3], host byte order
int crt_rgb2spec_load_file(const char* path, float* scale, float* data) {
    FILE* f = path ? std::fopen(path, "rb") : nullptr;
    if (!f) { set_error(std::string("rgb2spec_load_file: cannot open ") + (path ? path : "(null)")); return 1; }
    unsigned char head[4];
    const size_t n = (size_t)3 * 64 * 64 * 64 * 3;
    bool ok = std::fread(head, 1, 4, f) == 4 && std::fread(scale, sizeof(float), 64, f) == 64 && std::fread(data, sizeof(float), n, f) == n;
    std::fclose(f);
    if (!ok) { set_error("rgb2spec_load_file: short file"); return 1; }
    return 0;
}
int crt_rgb2spec_save_file(const char* path, const float* scale, const float* data) {
    FILE* f = path ? std::fopen(path, "wb") : nullptr;
    if (!f) { set_error(std::string("rgb2spec_save_file: cannot open ") + (path ? path : "(null)")); return 1; }
    const unsigned char head[4] = {0, 0, 0, 64};        // read back by UtoInt (color.cpp:101-104) as 64
    const size_t n = (size_t)3 * 64 * 64 * 64 * 3;
    bool ok = std::fwrite(head, 1, 4, f) == 4 && std::fwrite(scale, sizeof(float), 64, f) == 64 && std::fwrite(data, sizeof(float), n, f) == n;
    std::fclose(f);
    if (!ok) { set_error("rgb2spec_save_file: write failed"); return 1; }
    return 0;
}

int crt_scene_add_spectrum(crt_scene* s, int kind, float c, const float* interleaved, int n, const char* name, int normalize, int* out_id) {
    DevSpectrum sp;
    std::memset(&sp, 0, sizeof sp);
    int id = -1;
    switch (kind) {
        case 0: sp.kind = SPEC_CONSTANT; sp.c0 = c; s->h_spectra.push_back(sp); id = (int)s->h_spectra.size() - 1; break;
        case 1:
            if (!interleaved || n < 4 || (n & 1)) { set_error("add_spectrum: need >= 2 (lambda,value) pairs"); return 1; }
            id = add_piecewise(s, PiecewiseLinear::from_interleaved(interleaved, n, normalize != 0));
            break;
        case 2: {
            int cnt = 0;
            const float* t = name ? named_table(name, &cnt) : nullptr;
            if (!t) { set_error(std::string("add_spectrum: unknown named table ") + (name ? name : "(null)")); return 1; }
            id = add_piecewise(s, PiecewiseLinear::from_interleaved(t, cnt, normalize != 0));
            break;
        }
        case 3: {
            int cnt = 0;
            const float* t = swatch_table(n, &cnt);
            if (!t) { set_error("add_spectrum: swatch index out of range"); return 1; }
            id = add_piecewise(s, PiecewiseLinear::from_interleaved(t, cnt, false));
            break;
        }
        case 4:
            if (n < 0 || n > 5) { set_error("add_spectrum: illuminant index out of range"); return 1; }
            id = add_piecewise(s, host_spectra().illum[n]);
            break;
        case 5: {   // RGBAlbedoSpectrum(grey), spectrum.cpp:249-254 + ClampZero
            float g = std::max(0.0f, c), cc[3];
            grey_sigmoid(g, cc);
            sp.kind = SPEC_SIGMOID; sp.c0 = cc[0]; sp.c1 = cc[1]; sp.c2 = cc[2]; sp.scale = 1;
            s->h_spectra.push_back(sp); id = (int)s->h_spectra.size() - 1;
            break;
        }
        case 6: {   // RGBIlluminantSpectrum(grey), spectrum.cpp:264-270
            float m = c, scale = 2 * m, g = scale ? c / scale : 0, cc[3];
            g = std::max(0.0f, g);
            grey_sigmoid(g, cc);
            sp.kind = SPEC_SIGMOID_ILLUM; sp.c0 = cc[0]; sp.c1 = cc[1]; sp.c2 = cc[2]; sp.scale = scale;
            s->h_spectra.push_back(sp); id = (int)s->h_spectra.size() - 1;
            break;
        }
        case 7: case 8: case 9: {   // RGBAlbedoSpectrum / RGBIlluminantSpectrum / RGBUnboundedSpectrum (spectrum.cpp:249-270)
            if (!interleaved) { set_error("add_spectrum: kinds 7-9 take rgb in interleaved[0..2]"); return 1; }
            float rgb[3] = {interleaved[0], interleaved[1], interleaved[2]}, scale = 1, cc[3];
            if (kind != 7) {
                float m = std::max(rgb[0], rgb[1]);
                m = std::max(m, rgb[2]);
                scale = 2 * m;
                for (float& v : rgb) v = scale ? v / scale : 0.0f;
            }
            if (rgb_coeffs(s->ctx, rgb, cc)) return 1;
            sp.kind = kind == 7 ? SPEC_SIGMOID : kind == 8 ? SPEC_SIGMOID_ILLUM : SPEC_SIGMOID_UNBOUNDED;
            sp.c0 = cc[0]; sp.c1 = cc[1]; sp.c2 = cc[2]; sp.scale = scale;
            s->h_spectra.push_back(sp); id = (int)s->h_spectra.size() - 1;
            break;
        }
        default: set_error("add_spectrum: unknown kind"); return 1;
    }
    s->committed = false;
    if (out_id) *out_id = id;
    return 0;
}

int crt_scene_add_material(crt_scene* s, int type, int refl, int eta, int k, int emit, float emit_scale, int two_sided, int eta_constant, int* out_id) {
    DevMaterial m;
    m.type = type; m.refl = refl; m.eta = eta; m.k = k; m.emit = emit; m.emit_scale = emit_scale; m.two_sided = two_sided; m.eta_constant = eta_constant;
    int ns = (int)s->h_spectra.size();
    if (refl >= ns || eta >= ns || k >= ns || emit >= ns) { set_error("add_material: spectrum id out of range"); return 1; }
    s->h_materials.push_back(m);
    s->committed = false;
    if (out_id) *out_id = (int)s->h_materials.size() - 1;
    return 0;
}

// Threaded BVH over the padded shape boxes (crt_device_scene.h): median split of the box centres along the widest axis, nodes emitted
// depth-first, each carrying the index to continue at when its box is missed.
static void shape_bvh_emit(const std::vector<DevShapeBox>& boxes, int* ids, int n, std::vector<float>& out) {
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX}, clo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, chi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int k = 0; k < n; ++k) {
        const DevShapeBox& b = boxes[ids[k]];
        const float l[3] = {b.lo.x, b.lo.y, b.lo.z}, h[3] = {b.hi.x, b.hi.y, b.hi.z};
        for (int a = 0; a < 3; ++a) {
            lo[a] = std::min(lo[a], l[a]); hi[a] = std::max(hi[a], h[a]);
            const float c = 0.5f * (l[a] + h[a]);
            clo[a] = std::min(clo[a], c); chi[a] = std::max(chi[a], c);
        }
    }
    const size_t at = out.size();
    out.resize(at + 8);
    int leaf = n == 1 ? ids[0] : -1;
    if (n > 1) {
        int ax = 0;
        if (chi[1] - clo[1] > chi[ax] - clo[ax]) ax = 1;
        if (chi[2] - clo[2] > chi[ax] - clo[ax]) ax = 2;
        auto centre = [&](int id) { const DevShapeBox& b = boxes[id]; return ax == 0 ? b.lo.x + b.hi.x : ax == 1 ? b.lo.y + b.hi.y : b.lo.z + b.hi.z; };
        std::stable_sort(ids, ids + n, [&](int x, int y) { return centre(x) < centre(y); });
        shape_bvh_emit(boxes, ids, n / 2, out);
        shape_bvh_emit(boxes, ids + n / 2, n - n / 2, out);
    }
    const int skip = (int)(out.size() / 8);
    float* d = &out[at];
    d[0] = lo[0]; d[1] = lo[1]; d[2] = lo[2]; std::memcpy(&d[3], &skip, 4);
    d[4] = hi[0]; d[5] = hi[1]; d[6] = hi[2]; std::memcpy(&d[7], &leaf, 4);
}
static void build_shape_bvh(const std::vector<DevShapeBox>& boxes, std::vector<float>& out) {
    out.clear();
    if (boxes.empty()) return;
    std::vector<int> ids(boxes.size());
    for (size_t i = 0; i < ids.size(); ++i) ids[i] = (int)i;
    shape_bvh_emit(boxes, ids.data(), (int)ids.size(), out);
}

// Lights.h:5-8: point light (kind 0, v = position) / sun (kind 1, v = direction towards the light)
int crt_scene_add_light(crt_scene* s, int kind, const float* v3, int spectrum, float scale, int* out_id) {
    if (!s || !v3 || (kind != 0 && kind != 1)) { set_error("add_light: kind must be 0 (point light) or 1 (sun)"); return 1; }
    if (spectrum < 0 || spectrum >= (int)s->h_spectra.size()) { set_error("add_light: spectrum id out of range"); return 1; }
    DevDeltaLight dl;
    dl.kind = kind; dl.spectrum = spectrum; dl.scale = scale;
    f3 v = mk3(v3[0], v3[1], v3[2]);
    if (kind == 1) {
        if (!(dot3(v, v) > 0)) { set_error("add_light: a sun needs a direction"); return 1; }
        v = normalize3(v);
    }
    dl.v[0] = v.x; dl.v[1] = v.y; dl.v[2] = v.z;
    s->h_delta.push_back(dl);
    s->committed = false;
    if (out_id) *out_id = (int)s->h_delta.size() - 1;
    return 0;
}

int crt_scene_commit(crt_scene* s) {
    crt_context* c = s->ctx;
    CRT_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    DeviceScene& v = s->view;
    std::memset(&v, 0, sizeof v);
    if (s->has_model) {
        CRT_CUDA(s->d_nodes.upload((const float4*)s->h_nodes.data(), s->h_nodes.size() / 4, st));
        CRT_CUDA(s->d_leaf_refs.upload(s->h_leaf_refs.data(), s->h_leaf_refs.size(), st));
        CRT_CUDA(s->d_tris.upload((const float4*)s->h_tris.data(), s->h_tris.size() / 4, st));
        if (!s->h_tri_nrm.empty()) CRT_CUDA(s->d_tri_nrm.upload((const float4*)s->h_tri_nrm.data(), s->h_tri_nrm.size() / 4, st));
        CRT_CUDA(s->d_pk_boxes.upload((const float4*)s->h_pk_boxes.data(), s->h_pk_boxes.size() / 4, st));
        CRT_CUDA(s->d_pk_refs.upload(s->h_pk_refs.data(), s->h_pk_refs.size(), st));
        v.nodes = s->d_nodes.p; v.leaf_refs = s->d_leaf_refs.p; v.tris = s->d_tris.p;
        CRT_CUDA(s->d_node_tight.upload((const float4*)s->h_node_tight.data(), s->h_node_tight.size() / 4, st));
        v.pk_boxes = s->d_pk_boxes.p; v.pk_refs = s->d_pk_refs.p; v.node_tight = s->d_node_tight.p;
        v.tri_nrm = s->h_tri_nrm.empty() ? nullptr : s->d_tri_nrm.p;
        if (!s->h_tri_uv.empty()) CRT_CUDA(s->d_tri_uv.upload((const float4*)s->h_tri_uv.data(), s->h_tri_uv.size() / 4, st));
        if (!s->h_tri_tan.empty()) CRT_CUDA(s->d_tri_tan.upload((const float4*)s->h_tri_tan.data(), s->h_tri_tan.size() / 4, st));
        if (!s->h_tri_bitan.empty()) CRT_CUDA(s->d_tri_bitan.upload((const float4*)s->h_tri_bitan.data(), s->h_tri_bitan.size() / 4, st));
        v.tri_uv = s->h_tri_uv.empty() ? nullptr : s->d_tri_uv.p;
        v.tri_tan = s->h_tri_tan.empty() ? nullptr : s->d_tri_tan.p;
        v.tri_bitan = s->h_tri_bitan.empty() ? nullptr : s->d_tri_bitan.p;
        v.n_nodes = (int)(s->h_nodes.size() / 8); v.n_tris = (int)(s->h_tris.size() / 12);
        v.has_model = 1; v.retransform_surface = s->retransform;
        s->root_leaf_tris = 0;
        if (v.n_nodes == 1) {
            uint32_t bbits;
            std::memcpy(&bbits, &s->h_nodes[7], 4);
            const int count = (int)(bbits & CRT_LEAF_COUNT_MASK);
            if ((bbits & CRT_LEAF_FLAG) && count > 0 && count <= kRootLeafMaxTris) s->root_leaf_tris = count;
        }
        std::memcpy(v.model_o2r, s->model_o2r, 64);
        // materials may have been assigned after set_model: refresh the per-triangle tag
    }
    // emissive triangles and their selection CDF (weight = area * emit_scale), mesh-major order
    s->h_lights.clear(); s->h_light_cdf.clear(); s->h_light_pairs.clear(); s->light_total = 0;
    if (s->has_model && !s->h_materials.empty()) {
        for (size_t m = 0; m + 1 < s->mesh_first.size(); ++m) {
            int mat = s->mesh_material[m];
            if (mat < 0 || mat >= (int)s->h_materials.size()) { set_error("commit: mesh material id out of range"); return 1; }
            if (s->h_materials[mat].emit < 0) continue;
            for (uint32_t gid = s->mesh_first[m]; gid < s->mesh_first[m + 1]; ++gid) {
                const float* d = &s->h_tris[12 * (size_t)gid];
                f3 p0 = mk3(d[0], d[1], d[2]), p1 = mk3(d[4], d[5], d[6]), p2 = mk3(d[8], d[9], d[10]);
                f3 cr = cross3(p1 - p0, p2 - p0);
                float len = length3(cr);
                DevLight L;
                L.area = 0.5f * len;
                f3 nn = cr * (1.0f / len);
                if (!(L.area > 0)) continue;
                L.p0[0] = p0.x; L.p0[1] = p0.y; L.p0[2] = p0.z; L.p1[0] = p1.x; L.p1[1] = p1.y; L.p1[2] = p1.z;
                L.p2[0] = p2.x; L.p2[1] = p2.y; L.p2[2] = p2.z; L.n[0] = nn.x; L.n[1] = nn.y; L.n[2] = nn.z;
                L.material = mat;
                s->h_lights.push_back(L);
                s->light_total += L.area * s->h_materials[mat].emit_scale;
                s->h_light_cdf.push_back(s->light_total);
                s->h_light_pairs.push_back((int32_t)m);
                s->h_light_pairs.push_back((int32_t)(gid - s->mesh_first[m]));
            }
        }
    }
    CRT_CUDA(s->d_shapes.upload(s->h_shapes.data(), s->h_shapes.size(), st));
    CRT_CUDA(s->d_shape_boxes.upload(s->h_shape_boxes.data(), s->h_shape_boxes.size(), st));
    build_shape_bvh(s->h_shape_boxes, s->h_shape_bvh);
    CRT_CUDA(s->d_shape_bvh.upload((const float4*)s->h_shape_bvh.data(), s->h_shape_bvh.size() / 4, st));
    CRT_CUDA(s->d_materials.upload(s->h_materials.data(), s->h_materials.size(), st));
    CRT_CUDA(s->d_spectra.upload(s->h_spectra.data(), s->h_spectra.size(), st));
    CRT_CUDA(s->d_pool.upload(s->h_pool.data(), s->h_pool.size(), st));
    CRT_CUDA(s->d_lights.upload(s->h_lights.data(), s->h_lights.size(), st));
    CRT_CUDA(s->d_light_cdf.upload(s->h_light_cdf.data(), s->h_light_cdf.size(), st));
    CRT_CUDA(s->d_delta.upload(s->h_delta.data(), s->h_delta.size(), st));
    v.shapes = s->d_shapes.p; v.shape_boxes = s->d_shape_boxes.p; v.n_shapes = (int)s->h_shapes.size();
    v.shape_bvh = s->d_shape_bvh.p; v.n_shape_nodes = (int)(s->h_shape_bvh.size() / 8);
    v.materials = s->d_materials.p; v.n_materials = (int)s->h_materials.size();
    v.spectra = s->d_spectra.p; v.n_spectra = (int)s->h_spectra.size();
    v.pool = s->d_pool.p;
    v.lights = s->d_lights.p; v.light_cdf = s->d_light_cdf.p; v.n_lights = (int)s->h_lights.size(); v.light_total = s->light_total;
    v.delta_lights = s->d_delta.p; v.n_delta = (int)s->h_delta.size();
    // global tables: X, Y, Z, D65dense, F1 knots
    const HostSpectra& hs = host_spectra();
    std::vector<float> tab;
    if (c->sensor_curves.empty()) { tab.insert(tab.end(), hs.X, hs.X + 471); tab.insert(tab.end(), hs.Y, hs.Y + 471); tab.insert(tab.end(), hs.Z, hs.Z + 471); }
    else tab.insert(tab.end(), c->sensor_curves.begin(), c->sensor_curves.end());
    v.imaging_ratio = c->sensor_ratio;
    s->sensor_gen = c->sensor_gen;
    tab.insert(tab.end(), hs.D65dense, hs.D65dense + 471);
    const PiecewiseLinear& f1 = hs.illum[3];
    tab.insert(tab.end(), f1.lambdas.begin(), f1.lambdas.end());
    tab.insert(tab.end(), f1.values.begin(), f1.values.end());
    CRT_CUDA(s->d_tables.upload(tab.data(), tab.size(), st));
    v.cieX = s->d_tables.p; v.cieY = v.cieX + 471; v.cieZ = v.cieY + 471; v.d65dense = v.cieZ + 471;
    v.f1_lambdas = v.d65dense + 471; v.f1_n = (int)f1.lambdas.size(); v.f1_values = v.f1_lambdas + v.f1_n;
    const float* sm = c->sensor_curves.empty() ? hs.XYZFromSensorRGB : c->sensor_matrix;
    std::vector<float> col(sm, sm + 9);
    col.insert(col.end(), hs.RGBFromXYZ, hs.RGBFromXYZ + 9);
    CRT_CUDA(s->d_color.upload(col.data(), col.size(), st));
    CRT_CUDA(cudaStreamSynchronize(st));     // host staging vectors may be reused after return
    s->committed = true;
    ++s->commit_gen;
    return 0;
}
int crt_scene_light_count(const crt_scene* s) { return (int)s->h_lights.size(); }
int crt_scene_get_light_cdf(const crt_scene* s, float* cdf, int32_t* pairs, int cap) {
    int n = std::min(cap, (int)s->h_lights.size());
    for (int i = 0; i < n; ++i) { cdf[i] = s->h_light_cdf[i]; pairs[2 * i] = s->h_light_pairs[2 * i]; pairs[2 * i + 1] = s->h_light_pairs[2 * i + 1]; }
    return 0;
}
size_t crt_scene_device_bytes(const crt_scene* s) {
    return s->d_nodes.bytes() + s->d_node_tight.bytes() + s->d_leaf_refs.bytes() + s->d_pk_boxes.bytes() + s->d_pk_refs.bytes() + s->d_tris.bytes() + s->d_tri_nrm.bytes() + s->d_tri_uv.bytes() + s->d_tri_tan.bytes() + s->d_tri_bitan.bytes() + s->d_shapes.bytes() + s->d_pool.bytes() +
           s->d_lights.bytes() + s->d_light_cdf.bytes() + s->d_tables.bytes();
}

}  // extern "C"

// ================================================================ traversal launch helpers =================
namespace {

// closest-hit (or any-hit) pass over the rays A describes (ray arrays, optional indirection, host or device count, outputs).
//   trace_mode 3: k_trace_wide, then ONE k_trace launch over the hand-over list of order-sensitive rays (32 CTAs; a ray whose
//                 shared-memory FIFO overflows continues at once in its warp's global-memory ring);
//   trace_mode 0: k_trace over everything, then k_trace over the rays whose shared-memory FIFO overflowed, with global-memory rings.
// `slot` selects the launch's 8 counters in ctx->counters (zeroed once per wave / probe by zero_trace_counters):
//   [0] cursor of the first pass, [1] size of its hand-over / overflow list, [2] cursor of the second pass;
// counters[8 * kTraceSlots] accumulates the rays lost to a ring overflow (lost_rays() reads and clears it)
// With time_it the launches are bracketed by a pair of events from ctx->wave_events.
static const int kTraceSlots = (2 + kMaxNeeSlots) * (kMaxDepth + 2);
int zero_trace_counters(crt_context* c) {
    CRT_CUDA(cudaMemsetAsync(c->counters.p, 0, 8 * kTraceSlots * sizeof(int), c->stream));
    return 0;
}
int ensure_trace_rings(crt_context* c) {            // 32 CTAs x 8 warps x 64 Ki entries: allocated once, never inside a stream capture
    if (!c->gqueue.p) CRT_CUDA(c->gqueue.resize((size_t)32 * CRT_TRACE_WARPS * kGlobalQueueCap));
    return 0;
}
// rays the exact pass had to give up on (reported as misses): read and cleared; the caller turns a non-zero count into an error
int lost_rays(crt_context* c, int* out) {
    CRT_CUDA(cudaMemcpyAsync(out, c->counters.p + 8 * kTraceSlots, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CRT_CUDA(cudaMemsetAsync(c->counters.p + 8 * kTraceSlots, 0, sizeof(int), c->stream));
    CRT_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}
int fail_on_lost_rays(crt_context* c) {
    int lost = 0;
    if (int e = lost_rays(c, &lost)) return e;
    if (lost) { set_error(std::to_string(lost) + " ray(s) overflowed the exact traversal's global FIFO ring and were reported as misses"); return 3; }
    return 0;
}
template <bool ANY>
int launch_trace(crt_scene* s, TraceArgs A, bool stats, bool time_it = false, int trace_mode = 0, int slot = 0) {
    crt_context* c = s->ctx;
    cudaStream_t st = c->stream;
    int* cnt = c->counters.p + 8 * slot;
    A.gqueue = nullptr; A.gqcap = 0;
    A.stats = c->stats.p;
    if (time_it) {
        while ((int)c->wave_events.size() < c->event_cursor + 2) { cudaEvent_t e; CRT_CUDA(cudaEventCreate(&e)); c->wave_events.push_back(e); }
        CRT_CUDA(cudaEventRecord(c->wave_events[c->event_cursor], st));
    }
    const int threads = CRT_TRACE_WARPS * 32, ring_grid = 32;
    TraceArgs E = A;                       // the second, exact pass
    if (trace_mode == 3) {
        TraceArgs F = A;
        F.work_counter = cnt; F.overflow_count = cnt + 1; F.overflow_list = c->retrace_list.p;
        int wgrid = std::min(c->sm_count * CRT_WIDE_MINBLOCKS, std::max(1, cdiv(A.n, 32 * CRT_TRACE_WARPS)));
        if (stats) k_trace_wide<ANY, true><<<wgrid, threads, 0, st>>>(s->view, F);
        else k_trace_wide<ANY, false><<<wgrid, threads, 0, st>>>(s->view, F);
        CRT_CUDA(cudaGetLastError());
    } else {
        TraceArgs F = A;
        F.work_counter = cnt; F.overflow_count = cnt + 1; F.overflow_list = c->retrace_list.p;
        int grid = std::min(c->sm_count * 4, std::max(1, cdiv(A.n, CRT_TRACE_CHUNK * CRT_TRACE_WARPS)));
        if (stats) k_trace<ANY, true><<<grid, threads, 0, st>>>(s->view, F);
        else k_trace<ANY, false><<<grid, threads, 0, st>>>(s->view, F);
        CRT_CUDA(cudaGetLastError());
    }
    // second pass: always launched (it exits at once when the list is empty), so no host round trip is needed
    E.ray_index = c->retrace_list.p; E.n_ptr = cnt + 1; E.n = 0;
    E.work_counter = cnt + 2; E.gqueue = c->gqueue.p; E.gqcap = kGlobalQueueCap;
    E.overflow_count = c->counters.p + 8 * kTraceSlots; E.overflow_list = nullptr;       // rays lost to a ring overflow: accumulated until read
    if (stats) k_trace<ANY, true><<<ring_grid, threads, 0, st>>>(s->view, E);
    else k_trace<ANY, false><<<ring_grid, threads, 0, st>>>(s->view, E);
    CRT_CUDA(cudaGetLastError());
    if (time_it) { CRT_CUDA(cudaEventRecord(c->wave_events[c->event_cursor + 1], st)); c->event_cursor += 2; }
    return 0;
}
// the common case: n rays in ctx->ray_o/ray_d, identity indexing, results in hit_ref/hit_tb (or occluded)
TraceArgs wave_trace_args(crt_context* c, int n) {
    TraceArgs A;
    std::memset(&A, 0, sizeof A);
    A.ray_o = c->ray_o.p; A.ray_d = c->ray_d.p; A.ray_k = c->ray_k.p; A.ray_s = c->ray_s.p; A.n = n;
    A.hit_ref = c->hit_ref.p; A.hit_tb = c->hit_tb.p; A.occluded = c->occluded.p;
    return A;
}

int upload_rays(crt_scene* s, const float* rays, const float* tmax, int n) {
    crt_context* c = s->ctx;
    if (c->ensure_wave((size_t)n, false)) return 2;
    DevBuf<float> d_rays, d_tmax;
    CRT_CUDA(d_rays.upload(rays, 6 * (size_t)n, c->stream));
    if (tmax) CRT_CUDA(d_tmax.upload(tmax, (size_t)n, c->stream));
    if (int e = ensure_trace_rings(c)) return e;
    if (int e = zero_trace_counters(c)) return e;          // every probe's traversal uses counter slot 0
    k_pack_rays<<<cdiv(n, 256), 256, 0, c->stream>>>(d_rays.p, tmax ? d_tmax.p : nullptr, n, c->ray_o.p, c->ray_d.p, c->ray_k.p, c->ray_s.p);
    CRT_CUDA(cudaGetLastError());
    CRT_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}
template <typename T>
int download(const T* dev, T* host, size_t n, cudaStream_t st) {
    if (!host) return 0;
    CRT_CUDA(cudaMemcpyAsync(host, dev, n * sizeof(T), cudaMemcpyDeviceToHost, st));
    return 0;
}
int check_scene(crt_scene* s, bool need_model) {
    if (!s) { set_error("null scene"); return 1; }
    if (!s->committed) { set_error("scene not committed (call crt_scene_commit)"); return 1; }
    if (need_model && !s->has_model) { set_error("scene has no triangle model"); return 1; }
    if (s->sensor_gen != s->ctx->sensor_gen) { set_error("the context's film sensor changed after crt_scene_commit: commit the scene again"); return 1; }
    CRT_CUDA(cudaSetDevice(s->ctx->device));
    return 0;
}

}  // namespace

extern "C" {

// ================================================================ probes ==================================
int crt_trace_closest(crt_scene* s, const float* rays, int n, int mode, int32_t* mesh_id, int32_t* tri_id, float* t, float* bary3) {
    if (int e = check_scene(s, true)) return e;
    if (n <= 0) return 0;
    if (mode != 0 && mode != 3) { set_error("trace_closest: mode must be 0 (exact BFS) or 3 (ordered traversal, one ray per lane, + exact re-trace of order-sensitive rays)"); return 1; }
    crt_context* c = s->ctx;
    if (int e = upload_rays(s, rays, nullptr, n)) return e;
    if (int e = launch_trace<false>(s, wave_trace_args(c, n), false, false, mode)) return e;
    if (int e = fail_on_lost_rays(c)) return e;
    DevBuf<int> d_mesh, d_tri;
    DevBuf<float> d_t, d_b;
    CRT_CUDA(d_mesh.resize(n)); CRT_CUDA(d_tri.resize(n)); CRT_CUDA(d_t.resize(n)); CRT_CUDA(d_b.resize(3 * (size_t)n));
    k_unpack_hits<<<cdiv(n, 256), 256, 0, c->stream>>>(s->view, c->hit_ref.p, c->hit_tb.p, n, d_mesh.p, d_tri.p, d_t.p, d_b.p);
    CRT_CUDA(cudaGetLastError());
    if (download(d_mesh.p, mesh_id, n, c->stream) || download(d_tri.p, tri_id, n, c->stream) || download(d_t.p, t, n, c->stream) ||
        download(d_b.p, bary3, 3 * (size_t)n, c->stream)) return 2;
    CRT_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}
int crt_trace_any(crt_scene* s, const float* rays, const float* tmax, int n, int mode, int32_t* out) {
    if (int e = check_scene(s, true)) return e;
    if (n <= 0) return 0;
    if (mode != 0 && mode != 3) { set_error("trace_any: mode must be 0 or 3"); return 1; }
    crt_context* c = s->ctx;
    if (int e = upload_rays(s, rays, tmax, n)) return e;
    if (int e = launch_trace<true>(s, wave_trace_args(c, n), false, false, mode)) return e;
    if (int e = fail_on_lost_rays(c)) return e;
    if (download(c->occluded.p, out, n, c->stream)) return 2;
    CRT_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}
int crt_traverse_surface(crt_scene* s, const float* rays, int n, int32_t* found, float* nrm3) {
    if (int e = check_scene(s, true)) return e;
    if (n <= 0) return 0;
    crt_context* c = s->ctx;
    if (int e = upload_rays(s, rays, nullptr, n)) return e;
    if (int e = launch_trace<false>(s, wave_trace_args(c, n), false)) return e;
    DevBuf<int> d_found; DevBuf<float> d_n;
    CRT_CUDA(d_found.resize(n)); CRT_CUDA(d_n.resize(3 * (size_t)n));
    CRT_CUDA(cudaMemsetAsync(d_n.p, 0, 3 * (size_t)n * sizeof(float), c->stream));
    k_traverse_surface<<<cdiv(n, 256), 256, 0, c->stream>>>(s->view, c->ray_d.p, c->hit_ref.p, c->hit_tb.p, n, d_found.p, d_n.p);
    CRT_CUDA(cudaGetLastError());
    if (download(d_found.p, found, n, c->stream) || download(d_n.p, nrm3, 3 * (size_t)n, c->stream)) return 2;
    CRT_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}
int crt_traverse_local_surface(crt_scene* s, const float* rays, int n, int mode, int32_t* found, float* info17) {
    if (int e = check_scene(s, true)) return e;
    if (n <= 0) return 0;
    if (mode != 0 && mode != 3) { set_error("traverse_local_surface: mode must be 0 or 3"); return 1; }
    crt_context* c = s->ctx;
    if (int e = upload_rays(s, rays, nullptr, n)) return e;
    if (int e = launch_trace<false>(s, wave_trace_args(c, n), false, false, mode)) return e;
    DevBuf<int> d_found; DevBuf<float> d_info;
    CRT_CUDA(d_found.resize(n)); CRT_CUDA(d_info.resize(17 * (size_t)n));
    CRT_CUDA(cudaMemsetAsync(d_info.p, 0, 17 * (size_t)n * sizeof(float), c->stream));
    k_traverse_local_surface<<<cdiv(n, 128), 128, 0, c->stream>>>(s->view, c->ray_d.p, c->hit_ref.p, c->hit_tb.p, n, d_found.p, d_info.p);
    CRT_CUDA(cudaGetLastError());
    if (download(d_found.p, found, n, c->stream) || download(d_info.p, info17, 17 * (size_t)n, c->stream)) return 2;
    CRT_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}
int crt_kat_local_surface(const float* tri9, const float* bary3, const float* rayd3, int n, int on_device, float* info17) {
    if (!tri9 || !bary3 || !rayd3 || !info17 || n < 0) { set_error("kat_local_surface: bad argument"); return 1; }
    if (!on_device) {
        for (int i = 0; i < n; ++i) {
            const float* t = tri9 + 9 * (size_t)i;
            LocalSurfaceDev o;
            local_surface_core(mk3(t[0], t[1], t[2]), mk3(t[3], t[4], t[5]), mk3(t[6], t[7], t[8]), bary3[3 * i], bary3[3 * i + 1], bary3[3 * i + 2],
                               mk3(rayd3[3 * i], rayd3[3 * i + 1], rayd3[3 * i + 2]), nullptr, nullptr, nullptr, nullptr, o);
            float* w = info17 + 17 * (size_t)i;
            w[0] = o.hitp.x; w[1] = o.hitp.y; w[2] = o.hitp.z; w[3] = o.u; w[4] = o.v;
            w[5] = o.du.x; w[6] = o.du.y; w[7] = o.du.z; w[8] = o.dv.x; w[9] = o.dv.y; w[10] = o.dv.z;
            w[11] = o.n.x; w[12] = o.n.y; w[13] = o.n.z; w[14] = o.wo.x; w[15] = o.wo.y; w[16] = o.wo.z;
        }
        return 0;
    }
    DevBuf<float> d_tri, d_b, d_d, d_out;
    CRT_CUDA(d_tri.upload(tri9, 9 * (size_t)n, nullptr)); CRT_CUDA(d_b.upload(bary3, 3 * (size_t)n, nullptr)); CRT_CUDA(d_d.upload(rayd3, 3 * (size_t)n, nullptr));
    CRT_CUDA(d_out.resize(17 * (size_t)std::max(n, 1)));
    if (n) k_kat_local_surface<<<cdiv(n, 128), 128>>>(d_tri.p, d_b.p, d_d.p, n, d_out.p);
    CRT_CUDA(cudaGetLastError());
    CRT_CUDA(cudaMemcpy(info17, d_out.p, 17 * (size_t)n * sizeof(float), cudaMemcpyDeviceToHost));
    return 0;
}
static int shape_intersect_impl(crt_scene* s, int shape, const float* rays, int n, float tmax, int32_t* found, float* t, float* hitp3, float* nrm3, float* uv2,
                                float* frame9) {
    if (int e = check_scene(s, false)) return e;
    if (shape < 0 || shape >= (int)s->h_shapes.size()) { set_error("shape_intersect: bad shape id"); return 1; }
    if (n <= 0) return 0;
    crt_context* c = s->ctx;
    if (int e = upload_rays(s, rays, nullptr, n)) return e;
    DevBuf<int> d_found; DevBuf<float> d_t, d_p, d_n, d_uv, d_f;
    CRT_CUDA(d_found.resize(n)); CRT_CUDA(d_t.resize(n)); CRT_CUDA(d_p.resize(3 * (size_t)n)); CRT_CUDA(d_n.resize(3 * (size_t)n)); CRT_CUDA(d_uv.resize(2 * (size_t)n));
    CRT_CUDA(cudaMemsetAsync(d_t.p, 0, n * sizeof(float), c->stream));
    CRT_CUDA(cudaMemsetAsync(d_p.p, 0, 3 * (size_t)n * sizeof(float), c->stream));
    CRT_CUDA(cudaMemsetAsync(d_n.p, 0, 3 * (size_t)n * sizeof(float), c->stream));
    CRT_CUDA(cudaMemsetAsync(d_uv.p, 0, 2 * (size_t)n * sizeof(float), c->stream));
    if (frame9) { CRT_CUDA(d_f.resize(9 * (size_t)n)); CRT_CUDA(cudaMemsetAsync(d_f.p, 0, 9 * (size_t)n * sizeof(float), c->stream)); }
    k_shape_intersect<<<cdiv(n, 128), 128, 0, c->stream>>>(s->view, shape, c->ray_o.p, c->ray_d.p, n, tmax, d_found.p, d_t.p, d_p.p, d_n.p, d_uv.p, frame9 ? d_f.p : nullptr);
    CRT_CUDA(cudaGetLastError());
    if (download(d_found.p, found, n, c->stream) || download(d_t.p, t, n, c->stream) || download(d_p.p, hitp3, 3 * (size_t)n, c->stream) ||
        download(d_n.p, nrm3, 3 * (size_t)n, c->stream) || download(d_uv.p, uv2, 2 * (size_t)n, c->stream)) return 2;
    if (frame9 && download(d_f.p, frame9, 9 * (size_t)n, c->stream)) return 2;
    CRT_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}
int crt_shape_intersect(crt_scene* s, int shape, const float* rays, int n, float tmax, int32_t* found, float* t, float* hitp3, float* nrm3, float* uv2) {
    return shape_intersect_impl(s, shape, rays, n, tmax, found, t, hitp3, nrm3, uv2, nullptr);
}
int crt_shape_intersect_full(crt_scene* s, int shape, const float* rays, int n, float tmax, int32_t* found, float* t, float* hitp3, float* nrm3, float* uv2,
                             float* du_dv_wo9) {
    if (!du_dv_wo9) { set_error("shape_intersect_full: null output"); return 1; }
    return shape_intersect_impl(s, shape, rays, n, tmax, found, t, hitp3, nrm3, uv2, du_dv_wo9);
}

// ================================================================ film ====================================
int crt_film_create(crt_context* ctx, int width, int height, crt_film** out) {
    if (!ctx || width <= 0 || height <= 0) { set_error("film_create: bad arguments"); return 1; }
    CRT_CUDA(cudaSetDevice(ctx->device));
    auto* f = new crt_film;
    f->ctx = ctx; f->width = width; f->height = height;
    if (f->own.resize((size_t)width * height) != cudaSuccess) { set_error("film_create: out of device memory"); delete f; return 2; }
    f->data = f->own.p;
    cudaMemsetAsync(f->data, 0, (size_t)width * height * sizeof(float4), ctx->stream);
    *out = f;
    return 0;
}
void crt_film_destroy(crt_film* f) {
    if (!f) return;
    cudaSetDevice(f->ctx->device);
    cudaStreamSynchronize(f->ctx->stream);
    delete f;
}
int crt_film_clear(crt_film* f) {
    CRT_CUDA(cudaSetDevice(f->ctx->device));
    CRT_CUDA(cudaMemsetAsync(f->data, 0, (size_t)f->width * f->height * sizeof(float4), f->ctx->stream));
    return 0;
}
int crt_film_attach_device(crt_film* f, void* p) { f->data = p ? (float4*)p : f->own.p; return 0; }
void* crt_film_device_ptr(crt_film* f) { return f->data; }
int crt_film_download(crt_film* f, float* host) {
    CRT_CUDA(cudaSetDevice(f->ctx->device));
    CRT_CUDA(cudaMemcpyAsync(host, f->data, (size_t)f->width * f->height * sizeof(float4), cudaMemcpyDeviceToHost, f->ctx->stream));
    CRT_CUDA(cudaStreamSynchronize(f->ctx->stream));
    return 0;
}
int crt_film_upload(crt_film* f, const float* host) {
    CRT_CUDA(cudaSetDevice(f->ctx->device));
    CRT_CUDA(cudaMemcpyAsync(f->data, host, (size_t)f->width * f->height * sizeof(float4), cudaMemcpyHostToDevice, f->ctx->stream));
    CRT_CUDA(cudaStreamSynchronize(f->ctx->stream));
    return 0;
}
int crt_film_resolve(crt_film* f, uint8_t* host_rgb8, float* host_rgbf) {
    crt_context* c = f->ctx;
    CRT_CUDA(cudaSetDevice(c->device));
    const int npix = f->width * f->height;
    const HostSpectra& hs = host_spectra();
    DevBuf<float> col;
    const float* sm = c->sensor_curves.empty() ? hs.XYZFromSensorRGB : c->sensor_matrix;
    std::vector<float> h(sm, sm + 9);
    h.insert(h.end(), hs.RGBFromXYZ, hs.RGBFromXYZ + 9);
    CRT_CUDA(col.upload(h.data(), h.size(), c->stream));
    if (host_rgb8) CRT_CUDA(f->rgb8.resize(3 * (size_t)npix));
    if (host_rgbf) CRT_CUDA(f->rgbf.resize(3 * (size_t)npix));
    k_film_resolve<<<cdiv(npix, 256), 256, 0, c->stream>>>(f->data, npix, col.p, col.p + 9, host_rgb8 ? f->rgb8.p : nullptr, host_rgbf ? f->rgbf.p : nullptr);
    CRT_CUDA(cudaGetLastError());
    if (host_rgb8) CRT_CUDA(cudaMemcpyAsync(host_rgb8, f->rgb8.p, 3 * (size_t)npix, cudaMemcpyDeviceToHost, c->stream));
    if (host_rgbf) CRT_CUDA(cudaMemcpyAsync(host_rgbf, f->rgbf.p, 3 * (size_t)npix * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CRT_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}
// ---------------------------------------------------------------- NCCL (multi-GPU film reduce) --------------------------
// The only exchange step of a render: one ncclReduce(sum) of the per-GPU films onto the root over NVLink.
// libnccl is resolved at first use with dlopen -- the copy already in the process if there is one (torch's bundled libnccl.so.2),
// else $CRT_NCCL_LIB, else the system's -- so the library carries no link-time dependency on it; types and enums come from <nccl.h>.
namespace {
struct NcclApi {
    void* handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclReduce) Reduce = nullptr;
    decltype(&ncclCommGetAsyncError) CommGetAsyncError = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
    std::string origin;
};
NcclApi* nccl_api() {
    static NcclApi api;
    if (api.handle) return &api;            // a failed lookup is NOT cached: NCCL may be loaded later
    const char* env = std::getenv("CRT_NCCL_LIB");
    void* h = nullptr;
    if (env && *env) { h = dlopen(env, RTLD_NOW | RTLD_LOCAL); api.origin = env; }
    if (!h) { h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_LOCAL); api.origin = "libnccl.so.2 (already loaded in the process)"; }
    if (!h) { h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL); api.origin = "libnccl.so.2 (library search path)"; }
    if (!h) { set_error(std::string("NCCL not found (set CRT_NCCL_LIB): ") + dlerror()); return nullptr; }
    NcclApi a;
    a.handle = h; a.origin = api.origin;
    a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))dlsym(h, "ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy))dlsym(h, "ncclCommDestroy");
    a.Reduce = (decltype(a.Reduce))dlsym(h, "ncclReduce");
    a.CommGetAsyncError = (decltype(a.CommGetAsyncError))dlsym(h, "ncclCommGetAsyncError");
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(h, "ncclGetErrorString");
    a.GetVersion = (decltype(a.GetVersion))dlsym(h, "ncclGetVersion");
    if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.Reduce || !a.CommGetAsyncError || !a.GetErrorString) {
        set_error("NCCL library lacks a required entry point: " + a.origin);
        dlclose(h);
        return nullptr;
    }
    api = a;
    return &api;
}
int nccl_fail(NcclApi* n, const char* what, ncclResult_t r) {
    set_error(std::string(what) + ": " + n->GetErrorString(r));
    return 2;
}
}  // namespace

int crt_nccl_unique_id(uint8_t* id128) {
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    NcclApi* n = nccl_api();
    if (!n) return 1;
    if (!id128) { set_error("nccl_unique_id: null output"); return 1; }
    ncclUniqueId id;
    ncclResult_t r = n->GetUniqueId(&id);
    if (r != ncclSuccess) return nccl_fail(n, "ncclGetUniqueId", r);
    std::memcpy(id128, &id, 128);
    return 0;
}
int crt_nccl_comm_create(crt_context* c, int world, int rank, const uint8_t* id128) {
    NcclApi* n = nccl_api();
    if (!n) return 1;
    if (!c || !id128 || world < 1 || rank < 0 || rank >= world) { set_error("nccl_comm_create: bad arguments"); return 1; }
    if (c->nccl_comm) { set_error("nccl_comm_create: the context already has a communicator"); return 1; }
    CRT_CUDA(cudaSetDevice(c->device));
    ncclUniqueId id;
    std::memcpy(&id, id128, 128);
    ncclComm_t comm = nullptr;
    ncclResult_t r = n->CommInitRank(&comm, world, id, rank);
    if (r != ncclSuccess) return nccl_fail(n, "ncclCommInitRank", r);
    c->nccl_comm = comm; c->nccl_world = world; c->nccl_rank = rank;
    return 0;
}
int crt_nccl_comm_destroy(crt_context* c) {
    if (!c || !c->nccl_comm) return 0;
    NcclApi* n = nccl_api();
    if (!n) return 1;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    n->CommDestroy((ncclComm_t)c->nccl_comm);
    c->nccl_comm = nullptr;
    return 0;
}
// SURVEY 5 (failure detection): ncclCommGetAsyncError of the context's communicator; *async_error = 0 when healthy
int crt_nccl_async_error(crt_context* c, int* async_error) {
    if (!c || !c->nccl_comm || !async_error) { set_error("nccl_async_error: no communicator"); return 1; }
    NcclApi* n = nccl_api();
    if (!n) return 1;
    ncclResult_t st = ncclSuccess;
    ncclResult_t r = n->CommGetAsyncError((ncclComm_t)c->nccl_comm, &st);
    if (r != ncclSuccess) return nccl_fail(n, "ncclCommGetAsyncError", r);
    *async_error = (int)st;
    if (st != ncclSuccess && st != ncclInProgress) set_error(std::string("NCCL asynchronous error: ") + n->GetErrorString(st));
    return 0;
}
int crt_nccl_version(int* version, char* origin, int origin_cap) {
    NcclApi* n = nccl_api();
    if (!n) return 1;
    if (version) { *version = 0; if (n->GetVersion) n->GetVersion(version); }
    if (origin && origin_cap > 0) { std::strncpy(origin, n->origin.c_str(), origin_cap - 1); origin[origin_cap - 1] = 0; }
    return 0;
}
// ncclReduce(sum) of the film onto `root`, in place, on the context's stream; comm: a caller-owned ncclComm_t
int crt_film_reduce_nccl(crt_film* f, void* comm, int root) {
    NcclApi* n = nccl_api();
    if (!n) return 1;
    if (!f || !comm) { set_error("film_reduce_nccl: null film or communicator"); return 1; }
    CRT_CUDA(cudaSetDevice(f->ctx->device));
    ncclResult_t r = n->Reduce(f->data, f->data, (size_t)f->width * f->height * 4, ncclFloat32, ncclSum, root, (ncclComm_t)comm, f->ctx->stream);
    if (r != ncclSuccess) return nccl_fail(n, "ncclReduce", r);
    return 0;
}
// the same on the context's own communicator (crt_nccl_comm_create); world == 1 (no communicator) is a no-op
int crt_film_reduce(crt_film* f, int root) {
    if (!f) { set_error("film_reduce: null film"); return 1; }
    if (!f->ctx->nccl_comm) return 0;
    return crt_film_reduce_nccl(f, f->ctx->nccl_comm, root);
}

// ================================================================ render ==================================
// Continuous_Inversion_Sampler's constructor for pdf(x) = max(0, Gaussian(x, 0, sigma) - Gaussian(radius, 0, sigma)) on [-radius, radius],
// N = 10000 (RayTracer/Sampling.h:784-806, filters.h:102-105).  Host libm, so the table equals a CPU build of the reference.
static void gaussian_filter_build(float radius, float sigma, float* cdf, float* exp_out) {
    const int N = CRT_GAUSS_N;
    const float a = -radius, b = radius, e = gaussian_pbrt(radius, sigma);
    float delta_x = (b - a) / (float)N;
    float sum = 0;
    cdf[0] = 0.0f;
    for (int n = 1; n < N + 1; n++) {
        float current_x = std::min(std::max(a + delta_x * n, a), b);
        sum += delta_x * std::max<float>(0, gaussian_pbrt(current_x, sigma) - e);
        cdf[n] = sum;
    }
    float scaling_term = 1.0f / cdf[N];
    for (int n = 1; n < N; n++) cdf[n] *= scaling_term;
    cdf[N] = 1.0f;
    *exp_out = e;
}
// GaussianFilter(radius, sigma).Sample(u) on the host or on the device (parity probe): out = (p.x, p.y, weight) per sample
__global__ void k_gauss_probe(GaussFilter g, float rx, float ry, const float* u2, int n, float* out3) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    f2 u; u.x = u2[2 * i]; u.y = u2[2 * i + 1];
    FilterSample fs = filter_sample(2, rx, ry, u, g);
    out3[3 * i] = fs.px; out3[3 * i + 1] = fs.py; out3[3 * i + 2] = fs.weight;
}
int crt_kat_gaussian_filter(float rx, float ry, float sigma, const float* u2, int n, int on_device, float* out3) {
    if (!u2 || !out3 || n < 0 || !(rx > 0 && ry > 0 && sigma > 0)) { set_error("kat_gaussian_filter: bad argument"); return 1; }
    std::vector<float> cdf(2 * (CRT_GAUSS_N + 1));
    float ex[2];
    gaussian_filter_build(rx, sigma, cdf.data(), &ex[0]);
    gaussian_filter_build(ry, sigma, cdf.data() + CRT_GAUSS_N + 1, &ex[1]);
    if (!on_device) {
        GaussFilter g{cdf.data(), cdf.data() + CRT_GAUSS_N + 1, sigma, ex[0], ex[1]};
        for (int i = 0; i < n; ++i) {
            f2 u; u.x = u2[2 * i]; u.y = u2[2 * i + 1];
            FilterSample fs = filter_sample(2, rx, ry, u, g);
            out3[3 * i] = fs.px; out3[3 * i + 1] = fs.py; out3[3 * i + 2] = fs.weight;
        }
        return 0;
    }
    DevBuf<float> d_cdf, d_u, d_out;
    CRT_CUDA(d_cdf.upload(cdf.data(), cdf.size(), nullptr));
    CRT_CUDA(d_u.upload(u2, (size_t)2 * n, nullptr));
    CRT_CUDA(d_out.resize((size_t)3 * n));
    GaussFilter g{d_cdf.p, d_cdf.p + CRT_GAUSS_N + 1, sigma, ex[0], ex[1]};
    if (n) k_gauss_probe<<<(n + 255) / 256, 256>>>(g, rx, ry, d_u.p, n, d_out.p);
    CRT_CUDA(cudaGetLastError());
    CRT_CUDA(cudaMemcpy(out3, d_out.p, (size_t)3 * n * sizeof(float), cudaMemcpyDeviceToHost));
    return 0;
}

static int build_render_const(crt_context* ctx, const crt_render_config* cfg, RenderConst& rc) {
    std::memset(&rc, 0, sizeof rc);
    rc.width = cfg->width; rc.height = cfg->height;
    std::memcpy(rc.cam.r2c, cfg->raster_to_camera, 64);
    std::memcpy(rc.cam.c2w, cfg->camera_to_world, 64);
    rc.cam.lens_radius = cfg->lens_radius; rc.cam.focal_distance = cfg->focal_distance; rc.cam.kind = cfg->camera_kind;
    rc.sampler.kind = cfg->sampler_kind; rc.sampler.xs = cfg->xs; rc.sampler.ys = cfg->ys; rc.sampler.jitter = cfg->jitter; rc.sampler.seed = cfg->seed;
    rc.filter_kind = cfg->filter_kind; rc.filter_rx = cfg->filter_rx; rc.filter_ry = cfg->filter_ry;
    rc.gauss = GaussFilter{nullptr, nullptr, 0, 0, 0};
    if (cfg->filter_kind < 0 || cfg->filter_kind > 2) { set_error("render: filter_kind must be 0 (box), 1 (triangle) or 2 (gaussian)"); return 1; }
    if (cfg->filter_kind == 2) {                 // GaussianFilter(radius, sigma), filters.h:99-105
        const float sigma = cfg->filter_sigma > 0 ? cfg->filter_sigma : 0.5f;
        if (!(cfg->filter_rx > 0 && cfg->filter_ry > 0)) { set_error("render: GaussianFilter needs a positive radius"); return 1; }
        if (!ctx->gauss_cdf.p || ctx->gauss_key[0] != cfg->filter_rx || ctx->gauss_key[1] != cfg->filter_ry || ctx->gauss_key[2] != sigma) {
            std::vector<float> cdf(2 * (CRT_GAUSS_N + 1));
            gaussian_filter_build(cfg->filter_rx, sigma, cdf.data(), &ctx->gauss_exp[0]);
            gaussian_filter_build(cfg->filter_ry, sigma, cdf.data() + CRT_GAUSS_N + 1, &ctx->gauss_exp[1]);
            CRT_CUDA(ctx->gauss_cdf.upload(cdf.data(), cdf.size(), ctx->stream));
            CRT_CUDA(cudaStreamSynchronize(ctx->stream));        // cdf is a local
            ctx->gauss_key[0] = cfg->filter_rx; ctx->gauss_key[1] = cfg->filter_ry; ctx->gauss_key[2] = sigma;
        }
        rc.gauss = GaussFilter{ctx->gauss_cdf.p, ctx->gauss_cdf.p + CRT_GAUSS_N + 1, sigma, ctx->gauss_exp[0], ctx->gauss_exp[1]};
    }
    rc.max_depth = cfg->max_depth; rc.rr_depth = cfg->rr_depth; rc.ray_eps = cfg->ray_eps; rc.shadow_eps = cfg->shadow_eps;
    rc.light_strategy = cfg->light_strategy;
    if (cfg->xs <= 0 || cfg->ys <= 0) { set_error("render: sampler grid must be positive"); return 1; }
    if (cfg->sampler_kind == 1 && !cfg->jitter && cfg->spp_end > cfg->xs * cfg->ys) {
        set_error("render: StratifiedSampler without jitter refuses sample indices >= xs*ys (samplers.h:83-87)");
        return 1;
    }
    // RGBIlluminantSpectrum(sRGB, (1,1,1)): scale = 2, rsp = table(0.5 grey); RGBAlbedoSpectrum(sRGB, colors)
    grey_sigmoid(0.5f, rc.light_c);
    rc.light_scale = 2.0f;
    if (rgb_coeffs(ctx, cfg->albedo, rc.albedo_c)) return 1;       // non-grey `colors` go through the context's RGB -> spectrum table
    return 0;
}

// pixels owned by (rank, world) under interleaved tiles; world <= 1 or partition == 1 (spp split): all pixels
static void owned_pixels(const crt_render_config* cfg, std::vector<int>& out) {
    out.clear();
    if (cfg->world <= 1 || cfg->partition == 1) return;
    const int tw = std::max(1, cfg->tile_w), th = std::max(1, cfg->tile_h);
    const int tiles_x = (cfg->width + tw - 1) / tw;
    for (int p = 0; p < cfg->width * cfg->height; ++p) {
        int x = p % cfg->width, row = p / cfg->width;
        int tile = (row / th) * tiles_x + (x / tw);
        if (tile % cfg->world == cfg->rank) out.push_back(p);
    }
}

static void spp_range(const crt_render_config* cfg, int& s_begin, int& s_end) {
    s_begin = cfg->spp_begin; s_end = cfg->spp_end;
    if (cfg->world > 1 && cfg->partition == 1) {       // contiguous spp ranges
        int total = cfg->spp_end - cfg->spp_begin, per = total / cfg->world, rem = total % cfg->world;
        s_begin = cfg->spp_begin + cfg->rank * per + std::min(cfg->rank, rem);
        s_end = s_begin + per + (cfg->rank < rem ? 1 : 0);
    }
}
static int check_partition(const crt_render_config* cfg) {
    if (!cfg) { set_error("partition: null config"); return 1; }
    if (cfg->world > 1 && (cfg->rank < 0 || cfg->rank >= cfg->world)) { set_error("partition: rank outside [0, world)"); return 1; }
    if (cfg->width <= 0 || cfg->height <= 0) { set_error("partition: empty image"); return 1; }
    return 0;
}
int crt_partition_spp_range(const crt_render_config* cfg, int32_t* begin, int32_t* end) {
    if (int e = check_partition(cfg)) return e;
    int b, e2;
    spp_range(cfg, b, e2);
    *begin = b; *end = e2;
    return 0;
}
int crt_partition_pixel_count(const crt_render_config* cfg) {
    if (check_partition(cfg)) return -1;
    if (cfg->world <= 1 || cfg->partition == 1) return cfg->width * cfg->height;
    std::vector<int> owned;
    owned_pixels(cfg, owned);
    return (int)owned.size();
}
int crt_partition_pixels(const crt_render_config* cfg, int32_t* pixel_ids, int32_t cap) {
    if (int e = check_partition(cfg)) return e;
    std::vector<int> owned;
    if (cfg->world <= 1 || cfg->partition == 1) { owned.resize((size_t)cfg->width * cfg->height); for (size_t i = 0; i < owned.size(); ++i) owned[i] = (int)i; }
    else owned_pixels(cfg, owned);
    if ((int)owned.size() > cap) { set_error("partition_pixels: output buffer too small"); return 1; }
    std::copy(owned.begin(), owned.end(), pixel_ids);
    return 0;
}

// One wave: n path slots, slot i = (pixel_list ? pixel_list[i] : i, index_list ? index_list[i] : sample_index).
// mode 0: raygen -> closest hit -> reference Li + splat.  mode 1: raygen -> bounce loop (closest, shade + NEE,
// any-hit, resolve) -> splat.  Nothing here synchronises with the host: queue sizes stay on the device.
// nsamp > 1 (path integrator only): the wave holds nsamp consecutive sample indices of each of n_pix pixel slots
// (n = nsamp * n_pix), which keeps the deep-bounce launches full; the film is then updated per pixel in index order.
// staged (per-material) or fused shading of a bounce, crt_render_config.shade_mode
// Automatic: staged when the scene has analytic shapes (their intersection and surface code is what makes the fused kernel too large,
// C3: 426 -> 646 Mpaths/s) unless the frame is so small that launches dominate (C1: 327 fused vs 310 staged).
static bool use_staged_shading(const crt_scene* s, const crt_render_config* cfg) {
    if (cfg->mode != 1 || cfg->shade_mode == 1) return false;
    if (cfg->shade_mode == 2) return true;
    return s->view.n_shapes > 0 && (long long)cfg->width * cfg->height >= (1ll << 18);
}
static int run_wave(crt_scene* s, const crt_render_config* cfg, const RenderConst& rc, const int* pixel_list, const int* index_list,
                    int sample_index, int n, float4* film, const SampleDebugOut& dbg, crt_render_stats& rs, int nsamp = 1, const int* sample_cursor = nullptr) {
    crt_context* c = s->ctx;
    cudaStream_t st = c->stream;
    const bool stats = cfg->collect_stats != 0, time_it = cfg->time_kernels != 0;
    PathBuffers pb = c->path_buffers();
    if (cfg->mode == 0) pb.sampler = nullptr;
    const int n_pix = nsamp > 1 ? n / nsamp : 0;
    if (int e = ensure_trace_rings(c)) return e;          // no-op after the first call (crt_render makes it before any stream capture)
    if (int e = zero_trace_counters(c)) return e;
    k_raygen<<<cdiv(n, 256), 256, 0, st>>>(rc, pb, pixel_list, index_list, sample_index, n, n_pix, sample_cursor);
    rs.kernel_launches += 1;
    rs.paths += (uint64_t)n;
    if (cfg->mode == 0) {
        if (int e = launch_trace<false>(s, wave_trace_args(c, n), stats, time_it, cfg->trace_mode, 0)) return e;
        k_shade_li<<<cdiv(n, 256), 256, 0, st>>>(s->view, rc, pb, film, dbg, n);
        rs.kernel_launches += 3; rs.trace_launches += 1;
        rs.closest_rays += (uint64_t)n;
        CRT_CUDA(cudaGetLastError());
        return 0;
    }
    // ---- Tier B
    if (dbg.ray6) k_dump_rays<<<cdiv(n, 256), 256, 0, st>>>(pb, dbg.ray6, n);
    CRT_CUDA(cudaMemsetAsync(c->qcount.p, 0, c->qcount.bytes(), st));
    PathDebugOut nodbg;
    std::memset(&nodbg, 0, sizeof nodbg);
    int* lists[2] = {c->active_a.p, c->active_b.p};
    const int n_each = cfg->light_strategy == 1 ? s->view.n_lights : 0, n_slots = n_each + s->view.n_delta;
    const int slots_per_bounce = 2 + kMaxNeeSlots;          // trace-counter slots: closest, CDF shadow, one per additional next-event slot
    if (n_slots > 0) CRT_CUDA(cudaMemsetAsync(c->xq_count.p, 0, c->xq_count.bytes(), st));
    const bool staged = use_staged_shading(s, cfg);
    // A model that is one root leaf of a few triangles is traversed inside the shading kernels (trace_root_leaf): no traversal launches.
    // (Not with the exact-BFS trace mode or the instrumented kernels, which keep the launches -- and so cross-check this path.)
    const bool root_leaf = s->root_leaf_tris > 0 && cfg->trace_mode == 3 && !stats;
    DeviceScene V = s->view;
    V.root_leaf = root_leaf ? 1 : 0;
    // with time_kernels the kernels that then contain the traversal are bracketed like traversal launches (stats.trace_ms)
    auto tic = [&]() -> int {
        if (!(time_it && root_leaf)) return 0;
        while ((int)c->wave_events.size() < c->event_cursor + 2) { cudaEvent_t e; CRT_CUDA(cudaEventCreate(&e)); c->wave_events.push_back(e); }
        CRT_CUDA(cudaEventRecord(c->wave_events[c->event_cursor], st));
        return 0;
    };
    auto toc = [&]() -> int {
        if (!(time_it && root_leaf)) return 0;
        CRT_CUDA(cudaEventRecord(c->wave_events[c->event_cursor + 1], st));
        c->event_cursor += 2; rs.trace_launches += 1;
        return 0;
    };
    HitRecords H = {nullptr, nullptr, nullptr};
    if (staged) {
        H.a = c->hit_a.p; H.b = c->hit_b.p; H.c = c->hit_c.p;
        CRT_CUDA(cudaMemsetAsync(c->mq_count.p, 0, c->mq_count.bytes(), st));
    }
    const int staged_grid = std::min(cdiv(n, CRT_STAGED_THREADS), c->sm_count * CRT_STAGED_GRID);      // persistent: a warp strides over its queue
    for (int b = 0; b <= cfg->max_depth; ++b) {
        PathQueues Q;
        Q.count_active = 1;
        Q.active = b == 0 ? nullptr : lists[b & 1];
        Q.n_active = b == 0 ? nullptr : c->qcount.p + 2 * b;
        Q.n = n;
        Q.next_active = lists[(b + 1) & 1]; Q.n_next = c->qcount.p + 2 * (b + 1);
        Q.sh_o = c->sh_o.p; Q.sh_d = c->sh_d.p; Q.sh_k = c->sh_k.p; Q.sh_s = c->sh_s.p; Q.sh_contrib = c->sh_contrib.p; Q.sh_path = c->sh_path.p;
        Q.n_shadow = c->qcount.p + 2 * b + 1;
        Q.ray_counters = c->stats.p + 8;
        if (s->has_model && !root_leaf) {
            TraceArgs A = wave_trace_args(c, n);
            A.ray_index = Q.active; A.n_ptr = Q.n_active;
            if (int e = launch_trace<false>(s, A, stats, time_it, cfg->trace_mode, slots_per_bounce * b)) return e;
            rs.kernel_launches += 2; rs.trace_launches += 1;
        }
        // additional next-event slots (one sample from each emissive triangle under light_strategy 1, then the point / sun lights): their
        // shadow rays are generated from the path state of this hit, i.e. before the shade kernel advances it
        const size_t wave_n = c->xq_capacity;
        auto slot_queues = [&](int j) {
            PathQueues X = Q;
            X.sh_o = c->xq_o.p + (size_t)j * wave_n; X.sh_d = c->xq_d.p + (size_t)j * wave_n; X.sh_k = c->xq_k.p + (size_t)j * wave_n; X.sh_s = c->xq_s.p + (size_t)j * wave_n;
            X.sh_contrib = c->xq_contrib.p + 2 * (size_t)j * wave_n; X.sh_path = c->xq_path.p + (size_t)j * wave_n;
            X.n_shadow = c->xq_count.p + (size_t)b * kMaxNeeSlots + j;
            X.count_active = 0;
            return X;
        };
        MaterialQueues M = {c->mq_ids.p, c->mq_count.p + 4 * b, (int)c->staged_capacity};
        if (staged) {
            if (int e = tic()) return e;
            k_path_hit<<<staged_grid, CRT_STAGED_THREADS, 0, st>>>(V, pb, Q, H, M);
            if (int e = toc()) return e;
            rs.kernel_launches += 1;
        }
        for (int j = 0; j < n_slots; ++j) {
            const bool tri = j < n_each;
            k_path_nee_slot<<<cdiv(n, 128), 128, 0, st>>>(V, rc, pb, slot_queues(j), H, tri ? 0 : 1, tri ? j : j - n_each);
            rs.kernel_launches += 1;
        }
        if (staged) {
            k_path_shade_mat<MAT_LAMBERT><<<staged_grid, CRT_STAGED_THREADS, 0, st>>>(V, rc, pb, Q, H, M);
            k_path_shade_mat<MAT_DIELECTRIC><<<staged_grid, CRT_STAGED_THREADS, 0, st>>>(V, rc, pb, Q, H, M);
            k_path_shade_mat<MAT_CONDUCTOR><<<staged_grid, CRT_STAGED_THREADS, 0, st>>>(V, rc, pb, Q, H, M);
            rs.kernel_launches += 3;
        } else {
            if (int e = tic()) return e;
            k_path_shade<<<cdiv(n, 128), 128, 0, st>>>(V, rc, pb, Q, nodbg);
            if (int e = toc()) return e;
            rs.kernel_launches += 1;
        }
        // shadow queues: the CDF sample's (filled by the shade kernel), then the slots, each traced and added to L in the oracle's order
        bool counted = false;
        for (int j = -1; j < n_slots; ++j) {
            if (j < 0 && !(s->view.n_lights > 0 && cfg->light_strategy == 0)) continue;
            PathQueues X = j < 0 ? Q : slot_queues(j);
            X.count_active = counted ? 0 : 1;
            counted = true;
            if (s->has_model && !root_leaf) {
                TraceArgs A;
                std::memset(&A, 0, sizeof A);
                A.ray_o = X.sh_o; A.ray_d = X.sh_d; A.ray_k = X.sh_k; A.ray_s = X.sh_s; A.n = n; A.n_ptr = X.n_shadow; A.occluded = c->occluded.p;
                if (int e = launch_trace<true>(s, A, stats, time_it, cfg->trace_mode, slots_per_bounce * b + 2 + j)) return e;
                rs.kernel_launches += 2; rs.trace_launches += 1;
            }
            if (int e = tic()) return e;
            k_shadow_resolve<<<cdiv(n, 256), 256, 0, st>>>(V, pb, X, c->occluded.p);       // also adds this bounce's ray counts
            if (int e = toc()) return e;
            rs.kernel_launches += 1;
        }
        if (!counted) {
            k_path_count<<<1, 1, 0, st>>>(Q, b);
            rs.kernel_launches += 1;
        }
    }
    pb.depth_sum = c->stats.p + 10;
    if (nsamp > 1) {
        k_path_splat_multi<<<cdiv(n_pix, 256), 256, 0, st>>>(s->view, pb, film, n_pix, nsamp);
        rs.kernel_launches += 1;
    } else {
        k_path_splat<<<cdiv(n, 256), 256, 0, st>>>(s->view, pb, film, dbg, n);
        k_path_depth_sum<<<cdiv(n, 256), 256, 0, st>>>(pb, n);
        rs.kernel_launches += 2;
    }
    CRT_CUDA(cudaGetLastError());
    return 0;
}

static int check_render_mode(crt_scene* s, const crt_render_config* cfg) {
    if (cfg->mode != 0 && cfg->mode != 1) { set_error("render: unknown integrator mode"); return 1; }
    if (cfg->trace_mode != 0 && cfg->trace_mode != 3) { set_error("render: unknown trace_mode (0 = exact BFS kernel, 3 = ordered traversal + exact re-trace)"); return 1; }
    if (cfg->mode == 1) {
        if (cfg->max_depth < 0 || cfg->max_depth > kMaxDepth) { set_error("render: max_depth outside [0, 64]"); return 1; }
        if (s->h_materials.empty()) { set_error("render: the path integrator needs materials (crt_scene_add_material)"); return 1; }
        if (s->has_model && !s->retransform) { set_error("render: the path integrator needs world-space meshes (precomputed_world != 0)"); return 1; }
        for (const DevShape& sh : s->h_shapes)
            if (sh.material < 0 || sh.material >= (int)s->h_materials.size()) { set_error("render: shape material id out of range"); return 1; }
        if (cfg->shade_mode < 0 || cfg->shade_mode > 2) { set_error("render: shade_mode must be 0 (automatic), 1 (fused) or 2 (staged per material type)"); return 1; }
        if (cfg->light_strategy != 0 && cfg->light_strategy != 1) { set_error("render: light_strategy must be 0 (one sample by the power CDF) or 1 (one sample from each light)"); return 1; }
        const int slots = (cfg->light_strategy == 1 ? (int)s->h_lights.size() : 0) + (int)s->h_delta.size();
        if (slots > kMaxNeeSlots) {
            set_error("render: " + std::to_string(slots) + " next-event slots (point / sun lights + emissive triangles under light_strategy 1) exceed the limit of " +
                      std::to_string(kMaxNeeSlots));
            return 1;
        }
    }
    return 0;
}

int crt_render(crt_scene* s, crt_film* film, const crt_render_config* cfg, crt_render_stats* stats) {
    if (!cfg) { set_error("render: null config"); return 1; }
    if (int e = check_scene(s, cfg->mode == 0)) return e;
    if (int e = check_partition(cfg)) return e;
    if (int e = check_render_mode(s, cfg)) return e;
    if (!film || film->width != cfg->width || film->height != cfg->height) { set_error("render: film size does not match the config"); return 1; }
    crt_context* c = s->ctx;
    cudaStream_t st = c->stream;
    RenderConst rc;
    if (int e = build_render_const(c, cfg, rc)) return e;
    std::vector<int> owned;
    owned_pixels(cfg, owned);
    const bool use_list = cfg->world > 1 && cfg->partition == 0;
    const int n = use_list ? (int)owned.size() : cfg->width * cfg->height;
    int s_begin, s_end;
    spp_range(cfg, s_begin, s_end);
    // sample indices per wave: as many as keep a wave within ctx->max_wave_slots path slots (the path integrator's deep bounces
    // have few live paths per index; several indices per wave keep those launches busy).  Tier A splats inside its
    // shading kernel and keeps one index per wave.
    int per_wave = 1;
    if (cfg->mode == 1 && n > 0) per_wave = (int)std::max<long long>(1, std::min<long long>(kMaxSamplesPerWave, c->max_wave_slots / n));
    if (c->ensure_wave((size_t)std::max(n, 1) * per_wave, cfg->mode == 1)) return 2;
    if (cfg->mode == 1) {
        const int slots = (cfg->light_strategy == 1 ? s->view.n_lights : 0) + s->view.n_delta;
        if (slots > 0 && c->ensure_nee_slots(slots, c->wave_capacity)) return 2;
        if (use_staged_shading(s, cfg) && c->ensure_staged(c->wave_capacity)) return 2;
    }
    if (use_list) CRT_CUDA(c->pixel_list.upload(owned.data(), owned.size(), st));
    SampleDebugOut nodbg;
    std::memset(&nodbg, 0, sizeof nodbg);
    crt_render_stats rs;
    std::memset(&rs, 0, sizeof rs);
    c->event_cursor = 0;
    if (int e = ensure_trace_rings(c)) return e;
    CRT_CUDA(cudaMemsetAsync(c->stats.p, 0, c->stats.bytes(), st));
    CRT_CUDA(cudaEventRecord(c->ev[0], st));
    int idx = s_begin;
    // Small frames are launch bound (C1: ~100 launches of a few microseconds each per wave): their full waves are captured ONCE into a
    // CUDA graph -- the wave's first sample index lives in device memory and the graph's last node advances it -- and replayed.
    const int full_waves = n > 0 ? (s_end - s_begin) / per_wave : 0;
    const bool graphable = cfg->mode == 1 && !cfg->time_kernels && !cfg->collect_stats && full_waves >= 2 && (long long)n * per_wave <= kGraphMaxSlots;
    if (graphable) {
        WaveGraphKey key;
        std::memset(&key, 0, sizeof key);
        key.rc = rc; key.scene = s; key.scene_gen = s->commit_gen; key.film = film->data; key.n = n; key.per_wave = per_wave; key.max_depth = cfg->max_depth;
        key.trace_mode = cfg->trace_mode; key.light_strategy = cfg->light_strategy; key.shade_mode = cfg->shade_mode; key.pixel_list = use_list ? c->pixel_list.p : nullptr; key.wave_gen = c->wave_gen; key.stream = st;
        if (!c->d_cursor.p) CRT_CUDA(c->d_cursor.resize(1));
        if (!c->wave_graph || std::memcmp(&key, &c->wave_key, sizeof key) != 0) {
            if (c->wave_graph) { cudaGraphExecDestroy(c->wave_graph); c->wave_graph = nullptr; }
            cudaGraph_t g = nullptr;
            std::memset(&c->wave_rs, 0, sizeof c->wave_rs);
            CRT_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            int rc_wave = run_wave(s, cfg, rc, use_list ? c->pixel_list.p : nullptr, nullptr, 0, n * per_wave, film->data, nodbg, c->wave_rs, per_wave, c->d_cursor.p);
            k_add_int<<<1, 1, 0, st>>>(c->d_cursor.p, per_wave);
            cudaError_t ce = cudaStreamEndCapture(st, &g);
            if (rc_wave) { if (g) cudaGraphDestroy(g); return rc_wave; }
            CRT_CUDA(ce);
            ce = cudaGraphInstantiate(&c->wave_graph, g, 0);
            cudaGraphDestroy(g);
            CRT_CUDA(ce);
            c->wave_key = key;
        }
        k_set_int<<<1, 1, 0, st>>>(c->d_cursor.p, s_begin);
        for (int w = 0; w < full_waves; ++w) CRT_CUDA(cudaGraphLaunch(c->wave_graph, st));
        rs.paths += c->wave_rs.paths * full_waves; rs.kernel_launches += (c->wave_rs.kernel_launches + 1) * full_waves + 1;
        rs.trace_launches += c->wave_rs.trace_launches * full_waves; rs.graph_launches += full_waves;
        idx += full_waves * per_wave;
    }
    for (; idx < s_end && n > 0; idx += per_wave) {
        const int ns = std::min(per_wave, s_end - idx);
        if (int e = run_wave(s, cfg, rc, use_list ? c->pixel_list.p : nullptr, nullptr, idx, n * ns, film->data, nodbg, rs, ns)) return e;
    }
    CRT_CUDA(cudaGetLastError());
    CRT_CUDA(cudaEventRecord(c->ev[1], st));
    CRT_CUDA(cudaStreamSynchronize(st));
    CRT_CUDA(cudaEventElapsedTime(&rs.total_ms, c->ev[0], c->ev[1]));
    for (int e = 0; e + 1 < c->event_cursor; e += 2) {
        float ms = 0;
        CRT_CUDA(cudaEventElapsedTime(&ms, c->wave_events[e], c->wave_events[e + 1]));
        rs.trace_ms += ms;
    }
    {
        unsigned long long h[16];
        CRT_CUDA(cudaMemcpy(h, c->stats.p, sizeof h, cudaMemcpyDeviceToHost));
        if (cfg->collect_stats) { rs.nodes_visited = h[0]; rs.tris_tested = h[1]; rs.leaves_visited = h[2]; rs.max_queue = h[3]; }
        if (cfg->mode == 1) { rs.closest_rays = h[8]; rs.shadow_rays = h[9]; rs.depth_sum = h[10]; }
        int lost = 0;
        if (int e = lost_rays(c, &lost)) return e;
        rs.queue_overflow_rays = (uint64_t)lost;       // rays the exact pass gave up on (global FIFO ring overflow): an error, below
        rs.exact_retraced_rays = h[11];
    }
    if (stats) *stats = rs;
    if (rs.queue_overflow_rays) {
        set_error(std::to_string(rs.queue_overflow_rays) + " ray(s) overflowed the exact traversal's global FIFO ring and were rendered as misses");
        return 3;
    }
    return 0;
}

int crt_eval_samples(crt_scene* s, const crt_render_config* cfg, const int32_t* pixel_ids, const int32_t* indices, int n, float* ray6,
                     float* lambda8, float* pdf8, float* L8, float* rgb3, float* weight) {
    if (!cfg) { set_error("eval_samples: null config"); return 1; }
    if (int e = check_scene(s, cfg->mode == 0)) return e;
    if (int e = check_render_mode(s, cfg)) return e;
    if (n <= 0) return 0;
    crt_context* c = s->ctx;
    cudaStream_t st = c->stream;
    RenderConst rc;
    if (int e = build_render_const(c, cfg, rc)) return e;
    if (c->ensure_wave((size_t)n, cfg->mode == 1)) return 2;
    if (cfg->mode == 1) {
        const int slots = (cfg->light_strategy == 1 ? s->view.n_lights : 0) + s->view.n_delta;
        if (slots > 0 && c->ensure_nee_slots(slots, c->wave_capacity)) return 2;
        if (use_staged_shading(s, cfg) && c->ensure_staged(c->wave_capacity)) return 2;
    }
    CRT_CUDA(c->pixel_list.upload(pixel_ids, n, st));
    CRT_CUDA(c->index_list.upload(indices, n, st));
    DevBuf<float> d_ray, d_lam, d_pdf, d_L, d_rgb, d_w;
    CRT_CUDA(d_ray.resize(6 * (size_t)n)); CRT_CUDA(d_lam.resize(8 * (size_t)n)); CRT_CUDA(d_pdf.resize(8 * (size_t)n));
    CRT_CUDA(d_L.resize(8 * (size_t)n)); CRT_CUDA(d_rgb.resize(3 * (size_t)n)); CRT_CUDA(d_w.resize(n));
    SampleDebugOut dbg = {d_ray.p, d_lam.p, d_pdf.p, d_L.p, d_rgb.p, d_w.p};
    crt_render_stats rs;
    std::memset(&rs, 0, sizeof rs);
    crt_render_config cf = *cfg;
    cf.collect_stats = 0; cf.time_kernels = 0;
    if (int e = run_wave(s, &cf, rc, c->pixel_list.p, c->index_list.p, 0, n, nullptr, dbg, rs)) return e;
    if (download(d_ray.p, ray6, 6 * (size_t)n, st) || download(d_lam.p, lambda8, 8 * (size_t)n, st) || download(d_pdf.p, pdf8, 8 * (size_t)n, st) ||
        download(d_L.p, L8, 8 * (size_t)n, st) || download(d_rgb.p, rgb3, 3 * (size_t)n, st) || download(d_w.p, weight, n, st)) return 2;
    CRT_CUDA(cudaStreamSynchronize(st));
    return 0;
}

// Scene::Closest probe (oracle_render.cpp:38-93): mesh via the octree, then analytic shapes, and the surface record
int crt_scene_closest(crt_scene* s, const float* rays, int n, int32_t* kind, int32_t* id0, int32_t* id1, float* t, float* p3, float* ns3,
                      float* ng3, int32_t* backside) {
    if (int e = check_scene(s, false)) return e;
    if (n <= 0) return 0;
    crt_context* c = s->ctx;
    cudaStream_t st = c->stream;
    if (int e = upload_rays(s, rays, nullptr, n)) return e;
    if (c->ensure_wave((size_t)n, true)) return 2;
    if (s->has_model) { if (int e = launch_trace<false>(s, wave_trace_args(c, n), false)) return e; }
    DevBuf<int> d_kind, d_id0, d_id1, d_bs; DevBuf<float> d_t, d_p, d_ns, d_ng;
    CRT_CUDA(d_kind.resize(n)); CRT_CUDA(d_id0.resize(n)); CRT_CUDA(d_id1.resize(n)); CRT_CUDA(d_bs.resize(n)); CRT_CUDA(d_t.resize(n));
    CRT_CUDA(d_p.resize(3 * (size_t)n)); CRT_CUDA(d_ns.resize(3 * (size_t)n)); CRT_CUDA(d_ng.resize(3 * (size_t)n));
    PathDebugOut dbg = {d_kind.p, d_id0.p, d_id1.p, d_t.p, d_p.p, d_ns.p, d_ng.p, d_bs.p};
    CRT_CUDA(cudaMemsetAsync(c->qcount.p, 0, c->qcount.bytes(), st));
    CRT_CUDA(cudaMemsetAsync(c->flags.p, 0, n * sizeof(int), st));
    PathQueues Q;
    std::memset(&Q, 0, sizeof Q);
    Q.n = n; Q.next_active = c->active_a.p; Q.n_next = c->qcount.p + 2; Q.n_shadow = c->qcount.p + 1;
    Q.sh_o = c->sh_o.p; Q.sh_d = c->sh_d.p; Q.sh_k = c->sh_k.p; Q.sh_s = c->sh_s.p; Q.sh_contrib = c->sh_contrib.p; Q.sh_path = c->sh_path.p;
    RenderConst rc;
    std::memset(&rc, 0, sizeof rc);
    k_path_shade<<<cdiv(n, 128), 128, 0, st>>>(s->view, rc, c->path_buffers(), Q, dbg);
    CRT_CUDA(cudaGetLastError());
    if (download(d_kind.p, kind, n, st) || download(d_id0.p, id0, n, st) || download(d_id1.p, id1, n, st) || download(d_bs.p, backside, n, st) ||
        download(d_t.p, t, n, st) || download(d_p.p, p3, 3 * (size_t)n, st) || download(d_ns.p, ns3, 3 * (size_t)n, st) ||
        download(d_ng.p, ng3, 3 * (size_t)n, st)) return 2;
    CRT_CUDA(cudaStreamSynchronize(st));
    return 0;
}

// ================================================================ known-answer entry points ================
int crt_kat_hash(const uint8_t* key, uint64_t len, uint64_t seed, int on_device, uint64_t* out) {
    if (!on_device) { *out = murmur64a(key, len, seed); return 0; }
    DevBuf<unsigned char> d_key; DevBuf<uint64_t> d_out;
    CRT_CUDA(d_key.upload(key, std::max<uint64_t>(len, 1), 0));
    CRT_CUDA(d_out.resize(1));
    k_kat_hash<<<1, 1>>>(d_key.p, len, seed, d_out.p);
    CRT_CUDA(cudaMemcpy(out, d_out.p, 8, cudaMemcpyDeviceToHost));
    return 0;
}
int crt_kat_permutation(const uint32_t* i, const uint32_t* l, const uint32_t* p, int n, int on_device, int32_t* out) {
    if (!on_device) { for (int k = 0; k < n; ++k) out[k] = permutation_element(i[k], l[k], p[k]); return 0; }
    DevBuf<uint32_t> di, dl, dp; DevBuf<int> d_out;
    CRT_CUDA(di.upload(i, n, 0)); CRT_CUDA(dl.upload(l, n, 0)); CRT_CUDA(dp.upload(p, n, 0)); CRT_CUDA(d_out.resize(n));
    k_kat_permutation<<<cdiv(n, 128), 128>>>(di.p, dl.p, dp.p, n, d_out.p);
    CRT_CUDA(cudaMemcpy(out, d_out.p, n * sizeof(int), cudaMemcpyDeviceToHost));
    return 0;
}
int crt_kat_pcg32(int mode, uint64_t seq, uint64_t offset, int64_t adv, int n, int on_device, uint32_t* out_u32, float* out_f) {
    if (!on_device) {
        Pcg32 r;
        r.state = 0x853c49e6748fea9bULL; r.inc = 0xda3e39cb94b95bdbULL;
        if (mode == 1) pcg_set_sequence(r, seq, mix_bits(seq));
        if (mode == 2) pcg_set_sequence(r, seq, offset);
        if (adv) pcg_advance(r, adv);
        for (int i = 0; i < n; ++i) { if (out_u32) out_u32[i] = pcg_next_u32(r); else out_f[i] = pcg_next_float(r); }
        return 0;
    }
    DevBuf<uint32_t> du; DevBuf<float> df;
    if (out_u32) CRT_CUDA(du.resize(n)); else CRT_CUDA(df.resize(n));
    k_kat_pcg32<<<1, 1>>>(mode, seq, offset, adv, n, out_u32 ? du.p : nullptr, df.p);
    if (out_u32) CRT_CUDA(cudaMemcpy(out_u32, du.p, n * 4, cudaMemcpyDeviceToHost));
    else CRT_CUDA(cudaMemcpy(out_f, df.p, n * 4, cudaMemcpyDeviceToHost));
    return 0;
}
int crt_kat_sampler(int kind, int xs, int ys, int jitter, int seed, int px, int py, int index, int dim, const char* pattern, int on_device, float* out) {
    SamplerCfg c; c.kind = kind; c.xs = xs; c.ys = ys; c.jitter = jitter; c.seed = seed;
    size_t nout = 0;
    for (const char* ch = pattern; *ch; ++ch) nout += (*ch == '1') ? 1 : 2;
    if (!on_device) {
        SamplerState s;
        sampler_start(c, s, px, py, index, dim);
        for (const char* ch = pattern; *ch; ++ch) {
            if (*ch == '1') *out++ = sampler_get1d(c, s);
            else { f2 v = sampler_get2d(c, s); *out++ = v.x; *out++ = v.y; }
        }
        return 0;
    }
    DevBuf<char> d_pat; DevBuf<float> d_out;
    CRT_CUDA(d_pat.upload(pattern, std::strlen(pattern) + 1, 0));
    CRT_CUDA(d_out.resize(nout));
    k_kat_sampler<<<1, 1>>>(c, px, py, index, dim, d_pat.p, d_out.p);
    CRT_CUDA(cudaMemcpy(out, d_out.p, nout * 4, cudaMemcpyDeviceToHost));
    return 0;
}

// CameraBase::generateRay (Cameras.h:231-242 Orthographic, :273-297 Perspective, :340-352 Pinhole) for n film positions; lens_u2 = the Get2D()
// draws the thin lens consumes (NULL or lens_radius <= 0: no lens).  Host arithmetic, or the very device function k_raygen runs.
__global__ void k_camera_rays(DevCamera cam, const float* film_xy2, const float* lens_u2, int n, float* ray6) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    f2 u; u.x = lens_u2 ? lens_u2[2 * i] : 0.0f; u.y = lens_u2 ? lens_u2[2 * i + 1] : 0.0f;
    f3 o, d;
    camera_ray_core(cam, film_xy2[2 * i], film_xy2[2 * i + 1], u, o, d);
    ray6[6 * i] = o.x; ray6[6 * i + 1] = o.y; ray6[6 * i + 2] = o.z; ray6[6 * i + 3] = d.x; ray6[6 * i + 4] = d.y; ray6[6 * i + 5] = d.z;
}
int crt_camera_generate_rays(int kind, const float* raster_to_camera16, const float* camera_to_world16, float lens_radius, float focal_distance,
                             const float* film_xy2, const float* lens_u2, int n, int on_device, float* ray6) {
    if (!raster_to_camera16 || !camera_to_world16 || !film_xy2 || !ray6 || n < 0 || kind < 0 || kind > 2) { set_error("camera_generate_rays: bad argument"); return 1; }
    DevCamera cam;
    std::memcpy(cam.r2c, raster_to_camera16, 64); std::memcpy(cam.c2w, camera_to_world16, 64);
    cam.lens_radius = lens_u2 ? lens_radius : 0.0f; cam.focal_distance = focal_distance; cam.kind = kind;
    if (!on_device) {
        for (int i = 0; i < n; ++i) {
            f2 u; u.x = lens_u2 ? lens_u2[2 * i] : 0.0f; u.y = lens_u2 ? lens_u2[2 * i + 1] : 0.0f;
            f3 o, d;
            camera_ray_core(cam, film_xy2[2 * i], film_xy2[2 * i + 1], u, o, d);
            ray6[6 * i] = o.x; ray6[6 * i + 1] = o.y; ray6[6 * i + 2] = o.z; ray6[6 * i + 3] = d.x; ray6[6 * i + 4] = d.y; ray6[6 * i + 5] = d.z;
        }
        return 0;
    }
    DevBuf<float> d_xy, d_u, d_out;
    CRT_CUDA(d_xy.upload(film_xy2, 2 * (size_t)n, nullptr));
    if (lens_u2) CRT_CUDA(d_u.upload(lens_u2, 2 * (size_t)n, nullptr));
    CRT_CUDA(d_out.resize(6 * (size_t)std::max(n, 1)));
    if (n) k_camera_rays<<<cdiv(n, 128), 128>>>(cam, d_xy.p, lens_u2 ? d_u.p : nullptr, n, d_out.p);
    CRT_CUDA(cudaGetLastError());
    CRT_CUDA(cudaMemcpy(ray6, d_out.p, 6 * (size_t)n * sizeof(float), cudaMemcpyDeviceToHost));
    return 0;
}

}  // extern "C"
