// crt_kernels.cuh -- the wavefront stages as __global__ kernels for sm_100a.
//
//   k_raygen        evaluate_pixel up to the camera ray (RayTracerTestApp.h:287-323): sampler start, wavelengths,
//                   filter offset, Cameras.h generateRay
//   k_trace         Octtree_Model::Traverse closest hit / fixed-tMax occlusion, exact BFS order, one warp per ray (crt_trace.cuh)
//   k_trace_wide    the same answers by ordered traversal, one ray per lane (production; order-sensitive rays go to k_trace)
//   k_shade_li      the reference's Li + ToSensorRGB + film accumulation (RayTracerTestApp.h:218-284, :326-337)
//   k_film_resolve  RayTracerTestApp.h:425-452
// plus thread-per-ray probe kernels used by the parity tests.
#pragma once
#include "crt_shapes.cuh"
#include "crt_spectrum.cuh"
#include "crt_trace.cuh"
#include "crt_sat.h"

namespace crt {

// SoA path state for one wave (one sample index over the owned pixels)
struct PathBuffers {
    float4* ray_o;      // o.xyz, tMax
    float4* ray_d;      // d.xyz, -
    float4* ray_k;      // traversal constants formed once by the producer of the ray (store_ray): 1/d.xyz, bits(kz | octant order << 2)
    float4* ray_s;      // shear Sx, Sy, Sz (Shapes.h:1156-1158), -
    int* hit_ref;       // global triangle id or -1
    float4* hit_tb;     // t, b0, b1, b2
    float4* lambda;     // 2 per path
    float4* pdf;        // 2 per path
    float* weight;      // filter weight
    int* pixel;         // film pixel id
    // Tier B
    float4* beta;       // 2 per path
    float4* L;          // 2 per path
    SamplerState* sampler;
    int* flags;         // bit0 specularBounce, bits 8.. depth
    unsigned long long* depth_sum;   // optional: sum of realised path depths
};

struct RenderConst {
    int width, height;
    DevCamera cam;
    SamplerCfg sampler;
    int filter_kind;
    float filter_rx, filter_ry;
    GaussFilter gauss;          // filter_kind 2 only
    // Tier A shading constants (RGBIlluminantSpectrum(1,1,1), RGBAlbedoSpectrum(colors), RayTracerTestApp.h:246-255)
    float light_c[3], light_scale, albedo_c[3];
    // Tier B
    int max_depth, rr_depth;
    float ray_eps, shadow_eps;
    int light_strategy;         // emissive triangles: 0 one sample by the power CDF, 1 one sample from each (Shading.h:4)
};

// A ray as the traversal kernels read it: origin + tMax, direction, and the per-ray constants of Bounds3::IntersectP (1/d, Shapes.h:109)
// and Triangle::BasicIntersect (permutation + shear, Shapes.h:1142-1158), formed here -- by ray_setup, the same code the exact kernel runs --
// so that k_trace_wide's refill, which runs at a few lanes per warp, only loads them.
CRT_D void store_ray(float4* ray_o, float4* ray_d, float4* ray_k, float4* ray_s, size_t i, f3 o, f3 d, float tMax) {
    RayConst rc;
    ray_setup(rc, o, d);
    const int order = (d.x < 0 ? 1 : 0) | (d.z < 0 ? 2 : 0) | (d.y > 0 ? 4 : 0);      // child bits: x: bit0 = +x, z: bit1 = +z, y: bit2 = -y (crt_host.cpp split())
    ray_o[i] = make_float4(o.x, o.y, o.z, tMax);
    ray_d[i] = make_float4(d.x, d.y, d.z, 0.0f);
    ray_k[i] = make_float4(rc.inv_d.x, rc.inv_d.y, rc.inv_d.z, __int_as_float(rc.kz | (order << 2)));
    ray_s[i] = make_float4(rc.Sx, rc.Sy, rc.Sz, 0.0f);
}

CRT_D void store8(float4* dst, size_t i, const Spec8& s) {
    dst[2 * i] = make_float4(s.v[0], s.v[1], s.v[2], s.v[3]);
    dst[2 * i + 1] = make_float4(s.v[4], s.v[5], s.v[6], s.v[7]);
}
CRT_D void load8(const float4* src, size_t i, Spec8& s) {
    float4 a = src[2 * i], b = src[2 * i + 1];
    s.v[0] = a.x; s.v[1] = a.y; s.v[2] = a.z; s.v[3] = a.w; s.v[4] = b.x; s.v[5] = b.y; s.v[6] = b.z; s.v[7] = b.w;
}

// SampleUniformDiskConcentric (RayTracer/Sampling.h:383-403)
CRT_HD f2 sample_disk_concentric(f2 u) {
    const float PiOver4 = 0.78539816339744830961f, PiOver2 = 1.57079632679489661923f;
    f2 r;
    float ox = 2 * u.x - 1, oy = 2 * u.y - 1;
    if (ox == 0 && oy == 0) { r.x = 0; r.y = 0; return r; }
    float theta, rad;
    if (fabsf(ox) > fabsf(oy)) { rad = ox; theta = PiOver4 * (oy / ox); }
    else { rad = oy; theta = PiOver2 - PiOver4 * (ox / oy); }
    r.x = rad * cosf(theta); r.y = rad * sinf(theta);
    return r;
}

// CameraBase::generateRay for Perspective (Cameras.h:273-297), Orthographic (:231-242) and Pinhole (:340-352) cameras; lens_u = the sampler's
// Get2D() the thin lens consumes (read only when lens_radius > 0)
CRT_HD void camera_ray_core(const DevCamera& cam, float px, float py, f2 lens_u, f3& o, f3& d) {
    f4 c = mul_m4_v4(cam.r2c, px, py, 0.0f, 1.0f);
    if (cam.kind == 1) {
        o = mk3(c.x, c.y, c.z);
        d = mk3(0, 0, 1);
    } else if (cam.kind == 2) {      // PinholeCamera::generateRay (Cameras.h:340-352): r2c carries M_RastertoScreen, focal_distance the box depth
        o = mk3(c.x, c.y, c.z);
        f3 pinhole = mk3(0.0f, 0.0f, cam.focal_distance);       // 0 * hole_radius * cos/sin(0) = 0 exactly
        d = normalize3(pinhole - o);
    } else {
        f3 near_pos = mk3(c.x / c.w, c.y / c.w, c.z / c.w);
        o = mk3(0, 0, 0);
        d = normalize3(near_pos);
        if (cam.lens_radius > 0) {
            f2 dk = sample_disk_concentric(lens_u);
            float lx = cam.lens_radius * dk.x, ly = cam.lens_radius * dk.y;
            float ft = cam.focal_distance / d.z;
            f3 pfocus = o + d * ft;
            o = mk3(lx, ly, 0);
            d = normalize3(pfocus - o);
        }
    }
    // Ray::Transform (Shapes.h:37-41)
    o = xform_point(cam.c2w, o);
    d = xform_dir_normalized(cam.c2w, d);
}
CRT_D void camera_generate_ray(const DevCamera& cam, const SamplerCfg& sc, SamplerState& ss, float px, float py, f3& o, f3& d) {
    f2 u; u.x = 0; u.y = 0;
    if (cam.kind == 0 && cam.lens_radius > 0) u = sampler_get2d(sc, ss);
    camera_ray_core(cam, px, py, u, o, d);
}

// pixel_list == nullptr: path slot i renders pixel i.  index_list != nullptr: per-slot sample index (probe mode).
// n_pix > 0: the wave holds several sample indices, slot i = (sample_index + i / n_pix, pixel slot i % n_pix).
// sample_cursor != nullptr: the first sample index of the wave is read from device memory (waves replayed from one CUDA graph).
__global__ void __launch_bounds__(256) k_raygen(RenderConst rc, PathBuffers pb, const int* pixel_list, const int* index_list, int sample_index, int n, int n_pix,
                                                const int* sample_cursor) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (sample_cursor) sample_index = *sample_cursor;
    int slot = i, index = index_list ? index_list[i] : sample_index;
    if (n_pix > 0) { slot = i % n_pix; index = sample_index + i / n_pix; }
    int pixel_id = pixel_list ? pixel_list[slot] : slot;
    int x_pix = pixel_id % rc.width;
    int y_pix = (int)((float)rc.height - floorf((float)pixel_id / (float)rc.width));     // RayTracerTestApp.h:289-291
    SamplerState ss;
    sampler_start(rc.sampler, ss, x_pix, y_pix, index, 0);
    Spec8 lambda, pdf;
    sample_visible(sampler_get1d(rc.sampler, ss), lambda, pdf);
    f2 u = sampler_get2d(rc.sampler, ss);                                                    // GetPixel2D
    FilterSample fs = filter_sample(rc.filter_kind, rc.filter_rx, rc.filter_ry, u, rc.gauss);
    float fx = ((float)x_pix + .5f) + fs.px, fy = ((float)y_pix + .5f) + fs.py;
    f3 o, d;
    camera_generate_ray(rc.cam, rc.sampler, ss, fx, fy, o, d);
    store_ray(pb.ray_o, pb.ray_d, pb.ray_k, pb.ray_s, i, o, d, FLT_MAX);
    store8(pb.lambda, i, lambda);
    store8(pb.pdf, i, pdf);
    pb.weight[i] = fs.weight;
    pb.pixel[i] = pixel_id;
    if (pb.sampler) {            // path integrator: sampler state + fresh path state (beta = 1, L = 0, specularBounce = true, depth 0)
        pb.sampler[i] = ss;
        const float4 one = make_float4(1, 1, 1, 1), zero = make_float4(0, 0, 0, 0);
        pb.beta[2 * (size_t)i] = one; pb.beta[2 * (size_t)i + 1] = one;
        pb.L[2 * (size_t)i] = zero; pb.L[2 * (size_t)i + 1] = zero;
        pb.flags[i] = 1;
    }
}

__global__ void k_set_int(int* p, int v) { *p = v; }
__global__ void k_add_int(int* p, int v) { *p += v; }

// ---- traversal kernel ---------------------------------------------------------------------------------
#ifndef CRT_TRACE_WARPS
#define CRT_TRACE_WARPS 8
#endif
#define CRT_TRACE_QCAP 512          // FIFO entries per warp in shared memory (one entry = 8 child nodes)
#define CRT_TRACE_CHUNK 8           // rays fetched per warp per atomic

struct TraceArgs {
    const float4* ray_o; const float4* ray_d;
    const float4* ray_k; const float4* ray_s;       // store_ray's traversal constants (k_trace_wide); the exact kernel re-derives them
    const int* ray_index;       // optional indirection (re-trace lists); nullptr = identity
    const int* n_ptr;           // optional device-side count (overrides n when non-null)
    int n;
    int* hit_ref; float4* hit_tb;       // closest
    int* occluded;                       // any-hit
    int* work_counter;          // atomic ray cursor, zeroed before launch
    uint32_t* gqueue; int gqcap;         // global-memory FIFO for the overflow pass (nullptr = shared memory FIFO)
    int* overflow_count; int* overflow_list;
    unsigned long long* stats;  // nodes, tris, leaves, max_queue, rays
};

// Exact BFS pass.  The FIFO of a ray lives in shared memory (CRT_TRACE_QCAP entries per warp).  A ray that overflows it
//   * is appended to A.overflow_list for a later pass over a global-memory FIFO (A.gqueue == nullptr: the first pass of trace_mode 0), or
//   * is re-traced at once by the same warp with its private global-memory ring (A.gqueue != nullptr: the hand-over pass of trace_mode 3
//     and the overflow pass of trace_mode 0).  A ray that overflows even that ring (A.gqcap entries, i.e. 8 * A.gqcap queued child
//     groups) is reported as a miss and counted in A.overflow_count: crt_render returns an error instead of a silently wrong film.
template <bool ANY, bool STATS>
__global__ void __launch_bounds__(CRT_TRACE_WARPS * 32) k_trace(DeviceScene S, TraceArgs A) {
    __shared__ uint32_t s_queue[CRT_TRACE_WARPS * CRT_TRACE_QCAP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* q = s_queue + warp * CRT_TRACE_QCAP;
    uint32_t* gq = A.gqueue ? A.gqueue + (size_t)(blockIdx.x * CRT_TRACE_WARPS + warp) * A.gqcap : nullptr;
    const int n = A.n_ptr ? *A.n_ptr : A.n;
    // rays per atomic: 8 for full launches, down to 1 for the short hand-over lists of trace_mode 3 (a handful of
    // order-sensitive rays should spread over the warps, not queue behind each other in one)
    const int chunk = min(CRT_TRACE_CHUNK, max(1, n / (int)(gridDim.x * CRT_TRACE_WARPS * 2)));
    TraceStats st = {0, 0, 0, 0};
    unsigned nrays = 0;
    while (true) {
        int base = 0;
        if (lane == 0) base = atomicAdd(A.work_counter, chunk);
        base = __shfl_sync(CRT_FULL, base, 0);
        if (base >= n) break;
        // stage the chunk: lanes 0..7 fetch origins, 8..15 directions (two coalesced 128-byte requests)
        float4 stage = make_float4(0, 0, 0, 0);
        int my = base + (lane & 7);
        int ridx = -1;
        if (lane < 16 && (lane & 7) < chunk && my < n) {
            ridx = A.ray_index ? A.ray_index[my] : my;
            stage = (lane < 8) ? A.ray_o[ridx] : A.ray_d[ridx];
        }
        const int cnt = min(chunk, n - base);
        for (int r = 0; r < cnt; ++r) {
            float4 o4, d4;
            o4.x = __shfl_sync(CRT_FULL, stage.x, r); o4.y = __shfl_sync(CRT_FULL, stage.y, r);
            o4.z = __shfl_sync(CRT_FULL, stage.z, r); o4.w = __shfl_sync(CRT_FULL, stage.w, r);
            d4.x = __shfl_sync(CRT_FULL, stage.x, 8 + r); d4.y = __shfl_sync(CRT_FULL, stage.y, 8 + r);
            d4.z = __shfl_sync(CRT_FULL, stage.z, 8 + r);
            const int out_idx = __shfl_sync(CRT_FULL, ridx, r);
            RayConst rcst;
            ray_setup(rcst, mk3(o4.x, o4.y, o4.z), mk3(d4.x, d4.y, d4.z));
            WarpHit hit;
            bool ok = trace_bfs_warp<ANY, STATS>(S, rcst, o4.w, q, CRT_TRACE_QCAP, hit, &st);
            bool lost = false;
            if (!ok && gq) {
                ok = trace_bfs_warp<ANY, STATS>(S, rcst, o4.w, gq, A.gqcap, hit, &st);
                if (!ok) { lost = true; ok = true; hit.ref = -1; hit.t = hit.b0 = hit.b1 = hit.b2 = 0; }
            }
            if (STATS) nrays++;
            if (lane == 0) {
                if (lost) atomicAdd(A.overflow_count, 1);
                if (!ok) {
                    int slot = atomicAdd(A.overflow_count, 1); A.overflow_list[slot] = out_idx;
                } else if (ANY) {
                    A.occluded[out_idx] = hit.ref >= 0 ? 1 : 0;
                } else {
                    A.hit_ref[out_idx] = hit.ref;
                    A.hit_tb[out_idx] = make_float4(hit.t, hit.b0, hit.b1, hit.b2);
                }
            }
        }
    }
    if (STATS && lane == 0 && A.stats) {
        atomicAdd(&A.stats[0], (unsigned long long)st.nodes);
        atomicAdd(&A.stats[1], (unsigned long long)st.tris);
        atomicAdd(&A.stats[2], (unsigned long long)st.leaves);
        atomicMax(&A.stats[3], (unsigned long long)st.max_queue);
        atomicAdd(&A.stats[4], (unsigned long long)nrays);
    }
}

// Ordered traversal, one ray per lane for the descent (crt_trace.cuh "One ray per LANE"), trace_mode 3: the production kernel.
#ifndef CRT_WIDE_MINBLOCKS
#define CRT_WIDE_MINBLOCKS 4        // 62-64 registers, no spills; measured 2/3/4/5 CTAs per SM: 218 / 269 / 294 / 285 Mpaths/s on C2
#endif
#ifndef CRT_WIDE_LEAF_WAIT
#define CRT_WIDE_LEAF_WAIT 8        // parked leaves that trigger a leaf phase (measured 6 / 9 / 12 / 16: 508.6 / 509.8 / 507.1 / 499.5 Mpaths/s on C2)
#endif
#ifndef CRT_WIDE_CHUNK
#define CRT_WIDE_CHUNK 64           // rays a warp reserves per atomic
#endif
#ifndef CRT_WIDE_REFILL_MIN
#define CRT_WIDE_REFILL_MIN 8       // idle lanes that trigger a refill from the ray queue (also refilled when nobody can work)
#endif
template <bool ANY, bool STATS>
__global__ void __launch_bounds__(CRT_TRACE_WARPS * 32, CRT_WIDE_MINBLOCKS) k_trace_wide(DeviceScene S, TraceArgs A) {
    __shared__ uint2 s_stack[CRT_TRACE_WARPS * CRT_WIDE_STACK * 32];
    __shared__ uint2 s_pkq[CRT_TRACE_WARPS * CRT_PKQ_CAP], s_dq[CRT_TRACE_WARPS * CRT_PKQ_CAP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint2* stk = s_stack + warp * CRT_WIDE_STACK * 32 + lane;       // entry e of this lane: stk[e * 32]
    uint2* pkq = s_pkq + warp * CRT_PKQ_CAP;      // surviving sub-packets awaiting their triangle batch (crt_trace.cuh, stage 2)
    uint2* dq = s_dq + warp * CRT_PKQ_CAP;        // surviving super-packets of fat leaves awaiting stage 1
    const int n = A.n_ptr ? *A.n_ptr : A.n;
    const unsigned lt_mask = (1u << lane) - 1u;
    TraceStats st = {0, 0, 0, 0};
    unsigned nrays = 0;
    // rays are reserved from the launch's queue a chunk at a time (one atomic per chunk, not per refill); small launches (deep bounces)
    // take smaller chunks so that their rays spread over all warps
    const int chunk = min(CRT_WIDE_CHUNK, max(4, n / (int)(gridDim.x * CRT_TRACE_WARPS * 2)));
    int pool_next = 0, pool_end = 0;
    bool more = true;
    LaneRay r;
    r.ctl = 0; r.leaf = 0; r.href = -1; r.out_idx = -1;
    r.o = mk3(0, 0, 0); r.inv_d = r.o; r.Sx = r.Sy = r.Sz = 0;
    r.tMax0 = r.tbest = r.bound = 0; r.t2 = INFINITY;
    while (true) {
        // ---- retire, refill
        if (r.status() >= 2) {
            if (r.status() == 3) {
                int slot = atomicAdd(A.overflow_count, 1);
                A.overflow_list[slot] = r.out_idx;
                atomicAdd(&A.stats[11], 1ull);
            } else if (ANY) {
                A.occluded[r.out_idx] = r.href >= 0 ? 1 : 0;
            } else {
                A.hit_ref[r.out_idx] = r.href;
                if (r.href < 0) A.hit_tb[r.out_idx] = make_float4(0, 0, 0, 0);      // a hit's record was stored when it was accepted
            }
            r.set_status(0);
        }
        const unsigned idle = __ballot_sync(CRT_FULL, r.status() == 0);
        if ((more || pool_next < pool_end) && (__popc(idle) >= CRT_WIDE_REFILL_MIN || idle == CRT_FULL)) {
            if (pool_next == pool_end) {
                int base = 0;
                if (lane == 0) base = atomicAdd(A.work_counter, chunk);
                base = __shfl_sync(CRT_FULL, base, 0);
                pool_next = min(base, n); pool_end = min(base + chunk, n);
                if (base + chunk >= n) more = false;
            }
            const int my = pool_next + __popc(idle & lt_mask);
            pool_next = min(pool_end, pool_next + __popc(idle));
            if (r.status() == 0 && my < pool_end) {
                // the ray's traversal constants (1/d, shear, kz, octant order) were formed by the kernel that produced the ray
                // (store_ray: the same ray_setup code, hence the same bits), at full lane occupancy instead of here at 3 of 32
                const int ridx = A.ray_index ? A.ray_index[my] : my;
                const float4 o4 = A.ray_o[ridx], k4 = A.ray_k[ridx], s4 = A.ray_s[ridx];
                r.o = mk3(o4.x, o4.y, o4.z); r.inv_d = mk3(k4.x, k4.y, k4.z); r.Sx = s4.x; r.Sy = s4.y; r.Sz = s4.z;
                r.ctl = (uint32_t)__float_as_int(k4.w) | (1u << 5);          // kz | octant order << 2 (store_ray), status 1, empty stack
                r.tMax0 = o4.w; r.tbest = o4.w; r.bound = ANY ? o4.w : fast_bound(o4.w); r.t2 = INFINITY;
                r.href = -1;
                r.out_idx = ridx; r.leaf = 0;
                if (STATS) { nrays++; st.nodes++; }
                float m;
                const bool pinf = slab_unbounded_oi(r.o, r.inv_d, __ldg(&S.nodes[0]), __ldg(&S.nodes[1]), m);
                if (pinf && !(m > r.bound)) { stk[0] = make_uint2(0u, __float_as_uint(m)); r.set_sp(1); }
                else r.set_status(2);
            }
        }
        if (!__ballot_sync(CRT_FULL, r.status() != 0)) { if (more || pool_next < pool_end) continue; break; }
        // ---- node step: every lane that can descend pops its stack and tests the non-empty child cells of that node, far to near
        const bool want = r.traversing() && r.leaf == 0 && r.sp() > 0;
        if (want) {
            uint2 e;
            bool live;
            int sp = r.sp();
            // subtree bounds are tested when an entry is popped (one test per visited node) instead of for every child pushed
            do {
                e = stk[(--sp) * 32];
                live = !(__uint_as_float(e.y) > r.bound);
                if (live) {
                    float mt;
                    live = slab_unbounded_oi(r.o, r.inv_d, __ldg(&S.node_tight[2 * (size_t)e.x]), __ldg(&S.node_tight[2 * (size_t)e.x + 1]), mt) && !(mt > r.bound);
                }
            } while (!live && sp > 0);
            r.set_sp(sp);
            if (live) {
                const float4 plo = __ldg(&S.nodes[2 * (size_t)e.x]), phi = __ldg(&S.nodes[2 * (size_t)e.x + 1]);
                const uint32_t a = __float_as_uint(plo.w), b = __float_as_uint(phi.w);
                if (b & CRT_LEAF_FLAG) r.leaf = a | ((b & CRT_LEAF_PACKETS) ? 0x80000000u : 0u);      // park it (a >= 2: the list follows a header)
                else {
                    if (STATS) st.nodes += 8;
                    const int flip = r.flip();
                    // the 8 child cells are octants of this node's box: derive them with the host's own arithmetic (crt_sat.h
                    // child_cell, Octtree_Model.h:282-300) instead of loading 8 x 32 bytes; the parent's b word says which octants are empty
                    // (child_cell: hd = (max - min) / 2; C = min + hd; hd += 0.01; a child spans [C - hd, C] or [C, C + hd] per axis)
                    f3 hd = mk3(phi.x - plo.x, phi.y - plo.y, phi.z - plo.z) / 2.0f;
                    const f3 C = mk3(plo.x, plo.y, plo.z) + hd;
                    hd = hd + mk3(0.01f, 0.01f, 0.01f);
                    // the 9 plane distances shared by the 8 children, each formed exactly as the slab test forms it: (plane - o) * inv
                    const f3 tl = mk3(((C.x + -hd.x) - r.o.x) * r.inv_d.x, ((C.y + -hd.y) - r.o.y) * r.inv_d.y, ((C.z + -hd.z) - r.o.z) * r.inv_d.z);
                    const f3 tc = mk3(((C.x + 0.0f) - r.o.x) * r.inv_d.x, ((C.y + 0.0f) - r.o.y) * r.inv_d.y, ((C.z + 0.0f) - r.o.z) * r.inv_d.z);
                    const f3 th = mk3(((C.x + hd.x) - r.o.x) * r.inv_d.x, ((C.y + hd.y) - r.o.y) * r.inv_d.y, ((C.z + hd.z) - r.o.z) * r.inv_d.z);
                    // Along each axis a child spans the low half [l, c] or the high half [c, h] of the parent: the slab test's swap
                    // and its widening of the far plane (Shapes.h:108-113) are formed once per half (6x) instead of once per child
                    // (24x), with the very same operations -- (tNear > tFar) ? swap, tFar *= 1 + 2 gamma(3) -- so every value is
                    // bitwise what slab_unbounded_t forms.  min_t / max_t only grow / shrink, so "min_t > max_t at some axis" ==
                    // "final min_t > final max_t".
                    const float K = 1 + 2 * gamma_n(3);
                    const bool sxl = tl.x > tc.x, sxh = tc.x > th.x, syl = tl.y > tc.y, syh = tc.y > th.y, szl = tl.z > tc.z, szh = tc.z > th.z;
                    const float nXl = sxl ? tc.x : tl.x, fXl = (sxl ? tl.x : tc.x) * K, nXh = sxh ? th.x : tc.x, fXh = (sxh ? tc.x : th.x) * K;
                    const float nYl = syl ? tc.y : tl.y, fYl = (syl ? tl.y : tc.y) * K, nYh = syh ? th.y : tc.y, fYh = (syh ? tc.y : th.y) * K;
                    const float nZl = szl ? tc.z : tl.z, fZl = (szl ? tl.z : tc.z) * K, nZh = szh ? th.z : tc.z, fZh = (szh ? tc.z : th.z) * K;
#ifdef CRT_WIDE_ALL_OCTANTS
#pragma unroll 1
                    for (int kk = 7; kk >= 0; --kk) {
                        const int k = kk ^ flip;
                        if (!((b >> k) & 1u)) continue;                            // empty octant (mask kept in the parent's b word)
#else
                    // visit only the non-empty octants (mask kept in the parent's b word), far to near in this ray's order: bit kk of
                    // pm <-> child kk ^ flip (XOR-ing the bit index = swapping bit pairs / nibble pairs / nibbles)
                    uint32_t pm = b & 0xffu;
                    if (flip & 1) pm = ((pm & 0x55u) << 1) | ((pm & 0xaau) >> 1);
                    if (flip & 2) pm = ((pm & 0x33u) << 2) | ((pm & 0xccu) >> 2);
                    if (flip & 4) pm = ((pm & 0x0fu) << 4) | ((pm & 0xf0u) >> 4);
#pragma unroll 1
                    while (pm) {
                        const int kk = 31 - __clz(pm);
                        pm ^= 1u << kk;
                        const int k = kk ^ flip;
#endif
                        const bool xh = k & 1, zh = k & 2, yh = !(k & 4);          // child on the high side of the centre plane (bit 2 set = -y)
                        const float m = fmaxf(zh ? nZh : nZl, fmaxf(yh ? nYh : nYl, fmaxf(xh ? nXh : nXl, 0.0f)));
                        const float mx = fminf(zh ? fZh : fZl, fminf(yh ? fYh : fYl, fminf(xh ? fXh : fXl, INFINITY)));
                        if (m > mx || m > r.bound) continue;
                        if (r.sp() >= CRT_WIDE_STACK) { r.set_status(3); break; }                                        // overflow: exact kernel
                        stk[r.sp() * 32] = make_uint2(a + (uint32_t)k, __float_as_uint(m));
                        r.ctl += 1u << 8;
                    }
                    if (STATS) st.max_queue = max(st.max_queue, (unsigned)r.sp());
                }
            }
        }
        __syncwarp();
        // ---- leaf phase
        const unsigned parked = __ballot_sync(CRT_FULL, r.traversing() && r.leaf != 0);
        const unsigned can_descend = __ballot_sync(CRT_FULL, r.traversing() && r.leaf == 0 && r.sp() > 0);
        if (parked && (__popc(parked) >= CRT_WIDE_LEAF_WAIT || !can_descend))
            wide_leaf_phase<ANY, STATS>(S, r, pkq, dq, &st, ANY ? nullptr : A.hit_tb);
        if (r.traversing() && r.sp() == 0 && r.leaf == 0)
            r.set_status((!ANY && r.href >= 0 && !(r.t2 > r.bound)) ? 3 : 2);
    }
    if (STATS && A.stats) {
        unsigned nodes = st.nodes, tris = st.tris, leaves = st.leaves, mq = st.max_queue;
        for (int o = 16; o > 0; o >>= 1) {
            nodes += __shfl_xor_sync(CRT_FULL, nodes, o); tris += __shfl_xor_sync(CRT_FULL, tris, o); leaves += __shfl_xor_sync(CRT_FULL, leaves, o);
            mq = max(mq, __shfl_xor_sync(CRT_FULL, mq, o)); nrays += __shfl_xor_sync(CRT_FULL, nrays, o);
        }
        if (lane == 0) {
            atomicAdd(&A.stats[0], (unsigned long long)nodes);
            atomicAdd(&A.stats[1], (unsigned long long)tris);
            atomicAdd(&A.stats[2], (unsigned long long)leaves);
            atomicMax(&A.stats[3], (unsigned long long)mq);
            atomicAdd(&A.stats[4], (unsigned long long)nrays);
        }
    }
}

// ---- Tier A shading -------------------------------------------------------------------------------------
// Triangle::CalculateLocalSurface restricted to what Li reads: the normal (Shapes.h:1066-1075)
CRT_D f3 triangle_li_normal(const DeviceScene& S, int ref, float b0, float b1, float b2, f3 ray_d) {
    f3 n;
    if (S.tri_nrm) {
        float4 n0 = S.tri_nrm[3 * (size_t)ref], n1 = S.tri_nrm[3 * (size_t)ref + 1], n2 = S.tri_nrm[3 * (size_t)ref + 2];
        n = normalize3((mk3(n0.x, n0.y, n0.z) * b0 + mk3(n1.x, n1.y, n1.z) * b1) + mk3(n2.x, n2.y, n2.z) * b2);
    } else {
        float4 v0 = S.tris[3 * (size_t)ref], v1 = S.tris[3 * (size_t)ref + 1], v2 = S.tris[3 * (size_t)ref + 2];
        f3 p0 = mk3(v0.x, v0.y, v0.z), p1 = mk3(v1.x, v1.y, v1.z), p2 = mk3(v2.x, v2.y, v2.z);
        if (S.retransform_surface) {      // the reference re-applies ObjectToRender here even to world-space meshes (:995-997)
            p0 = xform_point(S.model_o2r, p0); p1 = xform_point(S.model_o2r, p1); p2 = xform_point(S.model_o2r, p2);
        }
        n = normalize3(cross3(p0 - p2, p1 - p2));
    }
    f3 rayd = normalize3(ray_d);          // TriangleIntersect::rayd = glm::normalize(ray.d), Shapes.h:1259
    if (dot3(n, rayd) > 0) n = -n;
    return n;
}

// Triangle::CalculateLocalSurface in full (Shapes.h:982-1083): the LocalSurfaceInfo record (Shapes.h:144-170) of a triangle hit.
// tHit is never assigned by the reference for triangles (Shapes.h:1034) and is therefore not part of the record here.
struct LocalSurfaceDev { f3 hitp; float u, v; f3 du, dv, n, wo; };
// p0_w, p1_w, p2_w: the vertices after ObjectToRender (:995-997); uv / tan / bitan / nrm: the three vertices' attributes, or nullptr where
// the model's vertex_available flag is off (:917-924); rayd: TriangleIntersect::rayd = normalize(ray.d) (:1259)
CRT_HD void local_surface_core(f3 p0_w, f3 p1_w, f3 p2_w, float b0, float b1, float b2, f3 rayd, const f2* uv, const f3* tan, const f3* bitan, const f3* nrm,
                              LocalSurfaceDev& o) {
    // "for now (u,v) p0(0,0), p1(1,0), p2(0,1)" (:999-1000)
    const float uv0x = 0, uv0y = 0, uv1x = 1, uv1y = 0, uv2x = 0, uv2y = 1;
    const float duv02x = uv0x - uv2x, duv02y = uv0y - uv2y, duv12x = uv1x - uv2x, duv12y = uv1y - uv2y;
    const f3 dp02 = p0_w - p2_w, dp12 = p1_w - p2_w;
    const float determinant = duv02x * duv12y - duv02y * duv12x;
    f3 dpdu = mk3(0, 0, 0), dpdv = mk3(0, 0, 0);
    const bool degenerateUV = fabsf(determinant) < 1e-9f;
    if (!degenerateUV) {
        const float invdet = 1 / determinant;
        dpdu = (duv12y * dp02 - duv02y * dp12) * invdet;
        dpdv = (duv02x * dp12 - duv12x * dp02) * invdet;
    }
    // std::pow(glm::length(cross), 2) == 0 in double <=> the float length is 0
    if (degenerateUV || length3(cross3(dpdu, dpdv)) == 0) {
        f3 ng = cross3(p2_w - p0_w, p1_w - p0_w);
        if (length3(ng) == 0) {                                    // glm::cross(dvec3, dvec3), :1018
            const f3 a = p2_w - p0_w, b = p1_w - p0_w;
            const double ax = a.x, ay = a.y, az = a.z, bx = b.x, by = b.y, bz = b.z;
            ng = mk3((float)(ay * bz - by * az), (float)(az * bx - bz * ax), (float)(ax * by - bx * ay));
        }
        ng = normalize3(ng);
        // Frisvad/Duff frame (:1024-1029); std::pow(float, int) promotes those two components to double
        const float sign = copysignf(1.0f, ng.z);
        const float a = -1 / (sign + ng.z);
        const float b = ng.x * ng.y * a;
        dpdu = mk3((float)(1 + (double)sign * ((double)ng.x * (double)ng.x) * (double)a), sign * b, -sign * ng.x);
        dpdv = mk3(b, (float)((double)sign + ((double)ng.y * (double)ng.y) * (double)a), -ng.y);
    }
    o.hitp = (p0_w * b0 + p1_w * b1) + p2_w * b2;
    float hu, hv;
    if (uv) { hu = (uv[0].x * b0 + uv[1].x * b1) + uv[2].x * b2; hv = (uv[0].y * b0 + uv[1].y * b1) + uv[2].y * b2; }
    else { hu = (uv0x * b0 + uv1x * b1) + uv2x * b2; hv = (uv0y * b0 + uv1y * b1) + uv2y * b2; }
    o.u = gclamp(hu, 0.0f, 1.0f); o.v = gclamp(hv, 0.0f, 1.0f);
    o.du = tan ? normalize3((tan[0] * b0 + tan[1] * b1) + tan[2] * b2) : dpdu;
    o.dv = bitan ? normalize3((bitan[0] * b0 + bitan[1] * b1) + bitan[2] * b2) : dpdv;
    o.n = nrm ? normalize3((nrm[0] * b0 + nrm[1] * b1) + nrm[2] * b2) : normalize3(cross3(dp02, dp12));
    if (dot3(o.n, rayd) > 0) o.n = -o.n;
    o.wo = mk3(0, 0, 1);
}
CRT_D f3 ld3(const float4* p, size_t i) { const float4 v = p[i]; return mk3(v.x, v.y, v.z); }
CRT_D void triangle_local_surface(const DeviceScene& S, int ref, float b0, float b1, float b2, f3 ray_d, LocalSurfaceDev& o) {
    f3 p0 = ld3(S.tris, 3 * (size_t)ref), p1 = ld3(S.tris, 3 * (size_t)ref + 1), p2 = ld3(S.tris, 3 * (size_t)ref + 2);
    if (S.retransform_surface) {      // the reference re-applies ObjectToRender here even to world-space meshes (:995-997)
        p0 = xform_point(S.model_o2r, p0); p1 = xform_point(S.model_o2r, p1); p2 = xform_point(S.model_o2r, p2);
    }
    f2 uv[3]; f3 tn[3], bt[3], nr[3];
    if (S.tri_uv) {
        const float4 a = S.tri_uv[2 * (size_t)ref], b = S.tri_uv[2 * (size_t)ref + 1];
        uv[0].x = a.x; uv[0].y = a.y; uv[1].x = a.z; uv[1].y = a.w; uv[2].x = b.x; uv[2].y = b.y;
    }
    if (S.tri_tan) { tn[0] = ld3(S.tri_tan, 3 * (size_t)ref); tn[1] = ld3(S.tri_tan, 3 * (size_t)ref + 1); tn[2] = ld3(S.tri_tan, 3 * (size_t)ref + 2); }
    if (S.tri_bitan) { bt[0] = ld3(S.tri_bitan, 3 * (size_t)ref); bt[1] = ld3(S.tri_bitan, 3 * (size_t)ref + 1); bt[2] = ld3(S.tri_bitan, 3 * (size_t)ref + 2); }
    if (S.tri_nrm) { nr[0] = ld3(S.tri_nrm, 3 * (size_t)ref); nr[1] = ld3(S.tri_nrm, 3 * (size_t)ref + 1); nr[2] = ld3(S.tri_nrm, 3 * (size_t)ref + 2); }
    local_surface_core(p0, p1, p2, b0, b1, b2, normalize3(ray_d), S.tri_uv ? uv : nullptr, S.tri_tan ? tn : nullptr, S.tri_bitan ? bt : nullptr,
                       S.tri_nrm ? nr : nullptr, o);
}

struct SampleDebugOut { float* ray6; float* lambda8; float* pdf8; float* L8; float* rgb3; float* weight; };

CRT_D void splat_or_debug(const DeviceScene& S, const PathBuffers& pb, int i, const Spec8& L, const Spec8& lambda, const Spec8& pdf,
                          float4* film, const SampleDebugOut& dbg) {
    f3 cam = to_sensor_rgb(S, L, lambda, pdf);
    cam.x = gclamp(cam.x, 0.0f, 1.0f); cam.y = gclamp(cam.y, 0.0f, 1.0f); cam.z = gclamp(cam.z, 0.0f, 1.0f);   // RayTracerTestApp.h:332-334
    float w = pb.weight[i];
    if (film) {
        int p = pb.pixel[i];
        float4 f = film[p];
        f.x += w * cam.x; f.y += w * cam.y; f.z += w * cam.z; f.w += w;
        film[p] = f;
    }
    if (dbg.L8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { dbg.L8[8 * i + k] = L.v[k]; dbg.lambda8[8 * i + k] = lambda.v[k]; dbg.pdf8[8 * i + k] = pdf.v[k]; }
        dbg.rgb3[3 * i] = cam.x; dbg.rgb3[3 * i + 1] = cam.y; dbg.rgb3[3 * i + 2] = cam.z;
        dbg.weight[i] = w;
    }
}

__global__ void __launch_bounds__(256) k_shade_li(DeviceScene S, RenderConst rc, PathBuffers pb, float4* film, SampleDebugOut dbg, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Spec8 lambda, pdf, L;
    load8(pb.lambda, i, lambda);
    load8(pb.pdf, i, pdf);
    int ref = pb.hit_ref[i];
    if (ref >= 0) {
        float4 tb = pb.hit_tb[i];
        float4 d4 = pb.ray_d[i];
        f3 n = triangle_li_normal(S, ref, tb.y, tb.z, tb.w, mk3(d4.x, d4.y, d4.z));
        float cosv = gclamp(dot3(n, mk3(0, 0, -1)), 0.0f, 1.0f);
#pragma unroll
        for (int k = 0; k < CRT_NLAMBDA; ++k) {
            float lam = lambda.v[k];
            float light = (rc.light_scale * sigmoid_eval(rc.light_c[0], rc.light_c[1], rc.light_c[2], lam)) * dense_lookup(S.d65dense, lam);
            float ambient = piecewise_query(S.f1_lambdas, S.f1_values, S.f1_n, lam) * 0.3f;
            float mat = sigmoid_eval(rc.albedo_c[0], rc.albedo_c[1], rc.albedo_c[2], lam);
            float r = 0.0f + ambient;
            r = r + (light * mat) * cosv;
            L.v[k] = r;
        }
    } else {
#pragma unroll
        for (int k = 0; k < CRT_NLAMBDA; ++k) L.v[k] = 0.0f;
    }
    if (dbg.ray6) {
        float4 o4 = pb.ray_o[i], d4 = pb.ray_d[i];
        dbg.ray6[6 * i] = o4.x; dbg.ray6[6 * i + 1] = o4.y; dbg.ray6[6 * i + 2] = o4.z;
        dbg.ray6[6 * i + 3] = d4.x; dbg.ray6[6 * i + 4] = d4.y; dbg.ray6[6 * i + 5] = d4.z;
    }
    splat_or_debug(S, pb, i, L, lambda, pdf, film, dbg);
}

// RayTracerTestApp.h:425-452
__global__ void __launch_bounds__(256) k_film_resolve(const float4* film, int npix, const float* sensor9, const float* rgbfromxyz9,
                                                        unsigned char* rgb8, float* rgbf) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    float4 f = film[i];
    // weightsum == 0: a pixel nothing was splatted to -- e.g. a rank's film under the tile partition before the reduce.  The reference's
    // 0 / 0 would be NaN and its cast to unsigned char undefined; such pixels resolve to black.
    f3 sensor_rgb = f.w != 0 ? mk3(f.x, f.y, f.z) / f.w : mk3(0, 0, 0);
    float m[9], q[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) { m[k] = sensor9[k]; q[k] = rgbfromxyz9[k]; }
    f3 xyz = mul_m3_v3(m, sensor_rgb);
    f3 o = mul_m3_v3(q, xyz);
    o.x = gclamp(o.x, 0.0f, 1.0f); o.y = gclamp(o.y, 0.0f, 1.0f); o.z = gclamp(o.z, 0.0f, 1.0f);
    if (rgbf) { rgbf[3 * i] = o.x; rgbf[3 * i + 1] = o.y; rgbf[3 * i + 2] = o.z; }
    if (rgb8) {
        rgb8[3 * i] = (unsigned char)(255.0f * o.x);
        rgb8[3 * i + 1] = (unsigned char)(255.0f * o.y);
        rgb8[3 * i + 2] = (unsigned char)(255.0f * o.z);
    }
}

// ---- probes ------------------------------------------------------------------------------------------------
__global__ void k_pack_rays(const float* rays6, const float* tmax, int n, float4* ray_o, float4* ray_d, float4* ray_k, float4* ray_s) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    store_ray(ray_o, ray_d, ray_k, ray_s, i, mk3(rays6[6 * i], rays6[6 * i + 1], rays6[6 * i + 2]), mk3(rays6[6 * i + 3], rays6[6 * i + 4], rays6[6 * i + 5]),
              tmax ? tmax[i] : FLT_MAX);
}
__global__ void k_unpack_hits(DeviceScene S, const int* hit_ref, const float4* hit_tb, int n, int* mesh_id, int* tri_id, float* t, float* bary3) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int ref = hit_ref[i];
    float4 tb = hit_tb[i];
    if (ref >= 0) {
        mesh_id[i] = __float_as_int(S.tris[3 * (size_t)ref + 1].w);
        tri_id[i] = __float_as_int(S.tris[3 * (size_t)ref + 2].w);
    } else { mesh_id[i] = -1; tri_id[i] = -1; tb = make_float4(0, 0, 0, 0); }
    if (t) t[i] = tb.x;
    if (bary3) { bary3[3 * i] = tb.y; bary3[3 * i + 1] = tb.z; bary3[3 * i + 2] = tb.w; }
}
__global__ void k_traverse_surface(DeviceScene S, const float4* ray_d, const int* hit_ref, const float4* hit_tb, int n, int* found, float* nrm3) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int ref = hit_ref[i];
    found[i] = ref >= 0;
    if (ref < 0) return;
    float4 tb = hit_tb[i], d4 = ray_d[i];
    f3 nn = triangle_li_normal(S, ref, tb.y, tb.z, tb.w, mk3(d4.x, d4.y, d4.z));
    nrm3[3 * i] = nn.x; nrm3[3 * i + 1] = nn.y; nrm3[3 * i + 2] = nn.z;
}
// Octtree_Model::Traverse's full return value, 17 floats per ray: hitp(3) u v du(3) dv(3) n(3) wo(3)
__global__ void k_traverse_local_surface(DeviceScene S, const float4* ray_d, const int* hit_ref, const float4* hit_tb, int n, int* found, float* out17) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int ref = hit_ref[i];
    found[i] = ref >= 0;
    if (ref < 0) return;
    float4 tb = hit_tb[i], d4 = ray_d[i];
    LocalSurfaceDev o;
    triangle_local_surface(S, ref, tb.y, tb.z, tb.w, mk3(d4.x, d4.y, d4.z), o);
    float* w = out17 + 17 * (size_t)i;
    w[0] = o.hitp.x; w[1] = o.hitp.y; w[2] = o.hitp.z; w[3] = o.u; w[4] = o.v;
    w[5] = o.du.x; w[6] = o.du.y; w[7] = o.du.z; w[8] = o.dv.x; w[9] = o.dv.y; w[10] = o.dv.z;
    w[11] = o.n.x; w[12] = o.n.y; w[13] = o.n.z; w[14] = o.wo.x; w[15] = o.wo.y; w[16] = o.wo.z;
}
// CalculateLocalSurface for explicit triangles (no attributes): tri9 = world-space vertices, bary3, rayd3 (already normalised) -> out17
__global__ void k_kat_local_surface(const float* tri9, const float* bary3, const float* rayd3, int n, float* out17) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* t = tri9 + 9 * (size_t)i;
    LocalSurfaceDev o;
    local_surface_core(mk3(t[0], t[1], t[2]), mk3(t[3], t[4], t[5]), mk3(t[6], t[7], t[8]), bary3[3 * i], bary3[3 * i + 1], bary3[3 * i + 2],
                       mk3(rayd3[3 * i], rayd3[3 * i + 1], rayd3[3 * i + 2]), nullptr, nullptr, nullptr, nullptr, o);
    float* w = out17 + 17 * (size_t)i;
    w[0] = o.hitp.x; w[1] = o.hitp.y; w[2] = o.hitp.z; w[3] = o.u; w[4] = o.v;
    w[5] = o.du.x; w[6] = o.du.y; w[7] = o.du.z; w[8] = o.dv.x; w[9] = o.dv.y; w[10] = o.dv.z;
    w[11] = o.n.x; w[12] = o.n.y; w[13] = o.n.z; w[14] = o.wo.x; w[15] = o.wo.y; w[16] = o.wo.z;
}
__global__ void k_shape_intersect(DeviceScene S, int shape, const float4* ray_o, const float4* ray_d, int n, float tmax, int* found, float* t,
                                  float* hitp3, float* nrm3, float* uv2, float* frame9) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 o4 = ray_o[i], d4 = ray_d[i];
    ShapeIsect is;
    bool hit = shape_basic(S.shapes[shape], mk3(o4.x, o4.y, o4.z), mk3(d4.x, d4.y, d4.z), tmax, is);
    found[i] = hit;
    if (!hit) return;
    SurfaceInfo si;
    shape_surface(S.shapes[shape], is, si);
    t[i] = si.tHit;
    hitp3[3 * i] = si.hitp.x; hitp3[3 * i + 1] = si.hitp.y; hitp3[3 * i + 2] = si.hitp.z;
    nrm3[3 * i] = si.n.x; nrm3[3 * i + 1] = si.n.y; nrm3[3 * i + 2] = si.n.z;
    uv2[2 * i] = si.u; uv2[2 * i + 1] = si.v;
    if (frame9) {
        float* w = frame9 + 9 * (size_t)i;
        w[0] = si.du.x; w[1] = si.du.y; w[2] = si.du.z; w[3] = si.dv.x; w[4] = si.dv.y; w[5] = si.dv.z; w[6] = si.wo.x; w[7] = si.wo.y; w[8] = si.wo.z;
    }
}

// known-answer kernels for the integer stack
__global__ void k_kat_permutation(const uint32_t* i, const uint32_t* l, const uint32_t* p, int n, int* out) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[k] = permutation_element(i[k], l[k], p[k]);
}
__global__ void k_kat_hash(const unsigned char* key, uint64_t len, uint64_t seed, uint64_t* out) { *out = murmur64a(key, len, seed); }
__global__ void k_kat_pcg32(int mode, uint64_t seq, uint64_t offset, int64_t adv, int n, uint32_t* out_u32, float* out_f) {
    Pcg32 r;
    r.state = 0x853c49e6748fea9bULL; r.inc = 0xda3e39cb94b95bdbULL;
    if (mode == 1) pcg_set_sequence(r, seq, mix_bits(seq));
    if (mode == 2) pcg_set_sequence(r, seq, offset);
    if (adv) pcg_advance(r, adv);
    for (int i = 0; i < n; ++i) {
        if (out_u32) out_u32[i] = pcg_next_u32(r);
        else out_f[i] = pcg_next_float(r);
    }
}
__global__ void k_kat_sampler(SamplerCfg c, int px, int py, int index, int dim, const char* pattern, float* out) {
    SamplerState s;
    sampler_start(c, s, px, py, index, dim);
    for (const char* ch = pattern; *ch; ++ch) {
        if (*ch == '1') *out++ = sampler_get1d(c, s);
        else { f2 v = sampler_get2d(c, s); *out++ = v.x; *out++ = v.y; }
    }
}

}  // namespace crt
