// crt_trace.cuh -- octree traversal for sm_100a: one WARP per ray.
//
// What is reproduced: Octtree_Model::Traverse (RayTracer/Octtree_Model.h:66-122) -- breadth-first visit
// with a FIFO frontier, Bounds3::IntersectP slab test against the *current* tMax (Shapes.h:100-124),
// leaf triangles tested in list order with pbrt's watertight test (Shapes.h:1101-1260) against the
// current tMax, strict '<' update.
//
// How: the tree is renumbered breadth-first at flatten time, so the 8 children of a node are 256
// contiguous bytes and the FIFO only stores "first child" indices (one entry per internal node that
// passed).  All the floating-point work that does not depend on tMax is done lane-parallel -- up to 32 child
// boxes or 32 leaf triangles per step, each lane executing exactly the reference's operation sequence for
// its element -- and the few tMax-dependent comparisons are then replayed serially, in the reference's visit
// order, from lane results exchanged with warp shuffles:
//   * IntersectP(ray, tMax) == P_inf && !(m > tMax), where P_inf is the test's outcome with no upper bound
//     and m its final min_t (min_t never depends on tMax and only grows axis by axis);
//   * BasicIntersect(ray, tMax) == ok_inf && !(det<0 ? tScaled < tMax*det : tScaled > tMax*det), followed by
//     the caller's `t < tMax`.
// The visit order, every operand and every rounding are therefore the reference's: hit ids are bit-exact
// by construction, not by tolerance.
#pragma once
#include "crt_device_scene.h"
#include "crt_host.h"          // CRT_SUBPACKET / CRT_SUPERPACKET: the packet sizes the flattener used

namespace crt {

#define CRT_FULL 0xffffffffu

struct RayConst {
    f3 o, d, inv_d;
    int kx, ky, kz;
    float Sx, Sy, Sz;
};

CRT_D f3 permute3(f3 v, int kx, int ky, int kz) { return mk3(comp(v, kx), comp(v, ky), comp(v, kz)); }

CRT_D void ray_setup(RayConst& rc, f3 o, f3 d) {
    rc.o = o; rc.d = d;
    rc.inv_d = mk3(1 / d.x, 1 / d.y, 1 / d.z);                     // Shapes.h:109
    f3 ad = mk3(fabsf(d.x), fabsf(d.y), fabsf(d.z));
    int kz = (ad.x > ad.y) ? ((ad.x > ad.z) ? 0 : 2) : ((ad.y > ad.z) ? 1 : 2);   // MaxComponentIndex, helpers.h:64-66
    int kx = kz + 1; if (kx == 3) kx = 0;
    int ky = kx + 1; if (ky == 3) ky = 0;
    rc.kx = kx; rc.ky = ky; rc.kz = kz;
    f3 dp = permute3(d, kx, ky, kz);
    rc.Sx = -dp.x / dp.z; rc.Sy = -dp.y / dp.z; rc.Sz = 1 / dp.z;     // Shapes.h:1156-1158
}

// Bounds3::IntersectP without an upper bound: returns P_inf and the final min_t (Shapes.h:100-124)
CRT_D bool slab_unbounded(const RayConst& rc, float4 lo, float4 hi, float& min_t_out) {
    const float K = 1 + 2 * gamma_n(3);
    float min_t = 0, max_t = INFINITY;
    bool pass = true;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        float pmin = i == 0 ? lo.x : (i == 1 ? lo.y : lo.z);
        float pmax = i == 0 ? hi.x : (i == 1 ? hi.y : hi.z);
        float o = comp(rc.o, i), inv = comp(rc.inv_d, i);
        float tNear = (pmin - o) * inv;
        float tFar = (pmax - o) * inv;
        if (tNear > tFar) { float t = tNear; tNear = tFar; tFar = t; }
        tFar *= K;
        // == "tNear > min_t ? tNear : min_t" / "tFar < max_t ? tFar : max_t": min_t and max_t are never NaN, and a NaN
        // tNear/tFar (0 * inf) leaves them unchanged under both forms
        min_t = fmaxf(tNear, min_t);
        max_t = fminf(tFar, max_t);
        if (min_t > max_t) pass = false;
    }
    min_t_out = min_t;
    return pass;
}

// The same test from the six plane distances (pmin - o) * inv, (pmax - o) * inv already formed by the caller: the 8 child
// cells of a node share 9 planes, so k_trace_wide forms each product once per node instead of once per child.
CRT_D bool slab_unbounded_t(float nx, float fx, float ny, float fy, float nz, float fz, float& min_t_out) {
    const float K = 1 + 2 * gamma_n(3);
    float min_t = 0, max_t = INFINITY;
    bool pass = true;
    { float tNear = nx, tFar = fx; if (tNear > tFar) { float t = tNear; tNear = tFar; tFar = t; } tFar *= K; min_t = fmaxf(tNear, min_t); max_t = fminf(tFar, max_t); if (min_t > max_t) pass = false; }
    { float tNear = ny, tFar = fy; if (tNear > tFar) { float t = tNear; tNear = tFar; tFar = t; } tFar *= K; min_t = fmaxf(tNear, min_t); max_t = fminf(tFar, max_t); if (min_t > max_t) pass = false; }
    { float tNear = nz, tFar = fz; if (tNear > tFar) { float t = tNear; tNear = tFar; tFar = t; } tFar *= K; min_t = fmaxf(tNear, min_t); max_t = fminf(tFar, max_t); if (min_t > max_t) pass = false; }
    min_t_out = min_t;
    return pass;
}

struct TriCand {
    float det, tScaled, t, b0, b1, b2;
};
template <int K> CRT_D float compk(f3 v) { return K == 0 ? v.x : (K == 1 ? v.y : v.z); }
// Triangle::BasicIntersect minus its two tMax comparisons (Shapes.h:1136-1259).  Degenerate triangles never
// reach this point: they are removed from the leaf lists at flatten time (Shapes.h:1131-1134).
// KZ = MaxComponentIndex(|d|) as a template parameter: the permutation (helpers.h:64-66, Shapes.h:1142-1148) costs
// nothing; one ray per warp makes the dispatch on kz a uniform branch.
template <int KZ>
CRT_D bool tri_test_unbounded_k(const RayConst& rc, f3 p0, f3 p1, f3 p2, TriCand& c) {
    constexpr int KX = (KZ + 1) % 3, KY = (KX + 1) % 3;
    f3 q0 = p0 - rc.o, q1 = p1 - rc.o, q2 = p2 - rc.o;
    f3 p0t = mk3(compk<KX>(q0), compk<KY>(q0), compk<KZ>(q0));
    f3 p1t = mk3(compk<KX>(q1), compk<KY>(q1), compk<KZ>(q1));
    f3 p2t = mk3(compk<KX>(q2), compk<KY>(q2), compk<KZ>(q2));
    p0t.x += rc.Sx * p0t.z; p0t.y += rc.Sy * p0t.z;
    p1t.x += rc.Sx * p1t.z; p1t.y += rc.Sy * p1t.z;
    p2t.x += rc.Sx * p2t.z; p2t.y += rc.Sy * p2t.z;
    float e0 = diff_of_products(p1t.x, p2t.y, p1t.y, p2t.x);
    float e1 = diff_of_products(p2t.x, p0t.y, p2t.y, p0t.x);
    float e2 = diff_of_products(p0t.x, p1t.y, p0t.y, p1t.x);
    if (e0 == 0.0f || e1 == 0.0f || e2 == 0.0f) {                     // double-precision edge fallback, :1174-1184
        double p2txp1ty = (double)p2t.x * (double)p1t.y, p2typ1tx = (double)p2t.y * (double)p1t.x;
        e0 = (float)(p2typ1tx - p2txp1ty);
        double p0txp2ty = (double)p0t.x * (double)p2t.y, p0typ2tx = (double)p0t.y * (double)p2t.x;
        e1 = (float)(p0typ2tx - p0txp2ty);
        double p1txp0ty = (double)p1t.x * (double)p0t.y, p1typ0tx = (double)p1t.y * (double)p0t.x;
        e2 = (float)(p1typ0tx - p1txp0ty);
    }
    if ((e0 < 0 || e1 < 0 || e2 < 0) && (e0 > 0 || e1 > 0 || e2 > 0)) return false;
    float det = e0 + e1 + e2;
    if (det == 0) return false;
    p0t.z *= rc.Sz; p1t.z *= rc.Sz; p2t.z *= rc.Sz;
    float tScaled = e0 * p0t.z + e1 * p1t.z + e2 * p2t.z;
    if (det < 0 && tScaled >= 0) return false;
    if (det > 0 && tScaled <= 0) return false;
    float invDet = 1 / det;
    float t = tScaled * invDet;
    if (isnan(t)) return false;
    float maxZt = max3_std(fabsf(p0t.z), fabsf(p1t.z), fabsf(p2t.z));
    float deltaZ = gamma_n(3) * maxZt;
    float maxXt = max3_std(fabsf(p0t.x), fabsf(p1t.x), fabsf(p2t.x));
    float maxYt = max3_std(fabsf(p0t.y), fabsf(p1t.y), fabsf(p2t.y));
    float deltaX = gamma_n(5) * (maxXt + maxZt);
    float deltaY = gamma_n(5) * (maxYt + maxZt);
    float deltaE = 2 * (gamma_n(2) * maxXt * maxYt + deltaY * maxXt + deltaX * maxYt);
    float maxE = max3_std(fabsf(e0), fabsf(e1), fabsf(e2));
    float deltaT = 3 * (gamma_n(3) * maxE * maxZt + deltaE * maxZt + deltaZ * maxE) * fabsf(invDet);
    if (t <= deltaT) return false;
    c.det = det; c.tScaled = tScaled; c.t = t;
    c.b0 = e0 * invDet; c.b1 = e1 * invDet; c.b2 = e2 * invDet;
    return true;
}
// Per-lane kz (lanes of one batch belong to different rays): the permutation becomes two selects per component.
CRT_D bool tri_test_unbounded_dyn(f3 o, float Sx, float Sy, float Sz, int kz, f3 p0, f3 p1, f3 p2, TriCand& c) {
    f3 q0 = p0 - o, q1 = p1 - o, q2 = p2 - o;
    const bool z0 = kz == 0, z1 = kz == 1;
    // kz = 0: (kx,ky,kz) = (y,z,x); kz = 1: (z,x,y); kz = 2: (x,y,z)
    f3 p0t = mk3(z0 ? q0.y : (z1 ? q0.z : q0.x), z0 ? q0.z : (z1 ? q0.x : q0.y), z0 ? q0.x : (z1 ? q0.y : q0.z));
    f3 p1t = mk3(z0 ? q1.y : (z1 ? q1.z : q1.x), z0 ? q1.z : (z1 ? q1.x : q1.y), z0 ? q1.x : (z1 ? q1.y : q1.z));
    f3 p2t = mk3(z0 ? q2.y : (z1 ? q2.z : q2.x), z0 ? q2.z : (z1 ? q2.x : q2.y), z0 ? q2.x : (z1 ? q2.y : q2.z));
    p0t.x += Sx * p0t.z; p0t.y += Sy * p0t.z;
    p1t.x += Sx * p1t.z; p1t.y += Sy * p1t.z;
    p2t.x += Sx * p2t.z; p2t.y += Sy * p2t.z;
    float e0 = diff_of_products(p1t.x, p2t.y, p1t.y, p2t.x);
    float e1 = diff_of_products(p2t.x, p0t.y, p2t.y, p0t.x);
    float e2 = diff_of_products(p0t.x, p1t.y, p0t.y, p1t.x);
    if (e0 == 0.0f || e1 == 0.0f || e2 == 0.0f) {                     // double-precision edge fallback, :1174-1184
        double p2txp1ty = (double)p2t.x * (double)p1t.y, p2typ1tx = (double)p2t.y * (double)p1t.x;
        e0 = (float)(p2typ1tx - p2txp1ty);
        double p0txp2ty = (double)p0t.x * (double)p2t.y, p0typ2tx = (double)p0t.y * (double)p2t.x;
        e1 = (float)(p0typ2tx - p0txp2ty);
        double p1txp0ty = (double)p1t.x * (double)p0t.y, p1typ0tx = (double)p1t.y * (double)p0t.x;
        e2 = (float)(p1typ0tx - p1txp0ty);
    }
    if ((e0 < 0 || e1 < 0 || e2 < 0) && (e0 > 0 || e1 > 0 || e2 > 0)) return false;
    float det = e0 + e1 + e2;
    if (det == 0) return false;
    p0t.z *= Sz; p1t.z *= Sz; p2t.z *= Sz;
    float tScaled = e0 * p0t.z + e1 * p1t.z + e2 * p2t.z;
    if (det < 0 && tScaled >= 0) return false;
    if (det > 0 && tScaled <= 0) return false;
    float invDet = 1 / det;
    float t = tScaled * invDet;
    if (isnan(t)) return false;
    float maxZt = max3_std(fabsf(p0t.z), fabsf(p1t.z), fabsf(p2t.z));
    float deltaZ = gamma_n(3) * maxZt;
    float maxXt = max3_std(fabsf(p0t.x), fabsf(p1t.x), fabsf(p2t.x));
    float maxYt = max3_std(fabsf(p0t.y), fabsf(p1t.y), fabsf(p2t.y));
    float deltaX = gamma_n(5) * (maxXt + maxZt);
    float deltaY = gamma_n(5) * (maxYt + maxZt);
    float deltaE = 2 * (gamma_n(2) * maxXt * maxYt + deltaY * maxXt + deltaX * maxYt);
    float maxE = max3_std(fabsf(e0), fabsf(e1), fabsf(e2));
    float deltaT = 3 * (gamma_n(3) * maxE * maxZt + deltaE * maxZt + deltaZ * maxE) * fabsf(invDet);
    if (t <= deltaT) return false;
    c.det = det; c.tScaled = tScaled; c.t = t;
    c.b0 = e0 * invDet; c.b1 = e1 * invDet; c.b2 = e2 * invDet;
    return true;
}
CRT_D bool tri_test_unbounded(const RayConst& rc, f3 p0, f3 p1, f3 p2, TriCand& c) {
    if (rc.kz == 0) return tri_test_unbounded_k<0>(rc, p0, p1, p2, c);
    if (rc.kz == 1) return tri_test_unbounded_k<1>(rc, p0, p1, p2, c);
    return tri_test_unbounded_k<2>(rc, p0, p1, p2, c);
}
// the tMax-dependent half of BasicIntersect (Shapes.h:1201-1209)
CRT_D bool tri_rejected_by_tmax(float det, float tScaled, float tMax) {
    return det < 0 ? (tScaled < tMax * det) : (tScaled > tMax * det);
}

struct WarpHit {
    int ref;            // global triangle id, -1 = miss
    float t, b0, b1, b2;
};
struct TraceStats { unsigned nodes, tris, leaves, max_queue; };

// Breadth-first closest hit (ANY == false) or fixed-tMax occlusion (ANY == true) for one ray per warp.
// q: this warp's FIFO ring (shared or global memory), qcap a power of two.  Returns false on FIFO overflow.
template <bool ANY, bool STATS>
CRT_D bool trace_bfs_warp(const DeviceScene& S, const RayConst& rc, float tMax0, uint32_t* q, int qcap, WarpHit& hit, TraceStats* st) {
    const int lane = threadIdx.x & 31;
    float tMax = tMax0;
    hit.ref = -1; hit.t = 0; hit.b0 = hit.b1 = hit.b2 = 0;
    int head = 0, tail = 0;
    bool first = true;
    while (first || head < tail) {
        bool valid;
        uint32_t node_idx = 0;
        if (first) { valid = (lane == 0); first = false; }
        else {
            int ngroups = min(4, tail - head);
            int gi = lane >> 3;
            valid = gi < ngroups;
            if (valid) node_idx = q[(head + gi) & (qcap - 1)] + (lane & 7);
            head += ngroups;
        }
        float4 lo = make_float4(0, 0, 0, 0), hi = lo;
        float m = 0;
        bool pinf = false;
        if (valid) {
            lo = __ldg(&S.nodes[2 * (size_t)node_idx]);
            hi = __ldg(&S.nodes[2 * (size_t)node_idx + 1]);
            pinf = slab_unbounded(rc, lo, hi, m);
        }
        if (STATS) { unsigned vm = __ballot_sync(CRT_FULL, valid); if (lane == 0) st->nodes += __popc(vm); }
        unsigned mask = __ballot_sync(CRT_FULL, valid && pinf);
        while (mask) {
            int l = __ffs(mask) - 1;
            mask &= mask - 1;
            float ml = __shfl_sync(CRT_FULL, m, l);
            if (ml > tMax) continue;                                  // IntersectP(ray, tMax) fails on the current bound
            uint32_t a = __float_as_uint(__shfl_sync(CRT_FULL, lo.w, l));
            uint32_t b = __float_as_uint(__shfl_sync(CRT_FULL, hi.w, l));
            if (b & CRT_LEAF_FLAG) {
                int count = (int)(b & CRT_LEAF_COUNT_MASK);
                if (STATS && lane == 0) { st->leaves++; st->tris += count; }
                for (int base = 0; base < count; base += 32) {
                    int i = base + lane;
                    TriCand c;
                    c.det = c.tScaled = c.t = c.b0 = c.b1 = c.b2 = 0;
                    bool ok = false;
                    uint32_t ref = 0;
                    if (i < count) {
                        ref = __ldg(&S.leaf_refs[a + i]);
                        float4 v0 = __ldg(&S.tris[3 * (size_t)ref]);
                        float4 v1 = __ldg(&S.tris[3 * (size_t)ref + 1]);
                        float4 v2 = __ldg(&S.tris[3 * (size_t)ref + 2]);
                        ok = tri_test_unbounded(rc, mk3(v0.x, v0.y, v0.z), mk3(v1.x, v1.y, v1.z), mk3(v2.x, v2.y, v2.z), c);
                    }
                    unsigned cm = __ballot_sync(CRT_FULL, ok);
                    while (cm) {
                        int cl = __ffs(cm) - 1;
                        cm &= cm - 1;
                        float det = __shfl_sync(CRT_FULL, c.det, cl);
                        float ts = __shfl_sync(CRT_FULL, c.tScaled, cl);
                        float t = __shfl_sync(CRT_FULL, c.t, cl);
                        if (tri_rejected_by_tmax(det, ts, tMax)) continue;
                        if (t < tMax) {
                            if (ANY) { hit.ref = 1; return true; }
                            tMax = t;
                            hit.ref = (int)__shfl_sync(CRT_FULL, ref, cl);
                            hit.t = t;
                            hit.b0 = __shfl_sync(CRT_FULL, c.b0, cl);
                            hit.b1 = __shfl_sync(CRT_FULL, c.b1, cl);
                            hit.b2 = __shfl_sync(CRT_FULL, c.b2, cl);
                        }
                    }
                }
            } else {
                if (tail - head >= qcap) return false;               // FIFO overflow: caller re-traces with a bigger ring
                if (lane == 0) q[tail & (qcap - 1)] = a;
                ++tail;
                if (STATS && lane == 0) st->max_queue = max(st->max_queue, (unsigned)(tail - head));
            }
        }
        __syncwarp();
    }
    return true;
}

// Octtree_Model::Traverse / IntersectP for a tree that consists of its root leaf (fewer than 40 triangles never split,
// Octtree_Model.h:204-214: the Cornell box, the floor under a field of analytic shapes).  The loop above then degenerates to the root's
// cell test and the root's list in order with the shrinking tMax; one lane does that for its own ray, with the same per-triangle
// arithmetic and the same tMax decisions, so the path integrator needs no traversal launch for such scenes.
template <bool ANY>
CRT_D bool trace_root_leaf(const DeviceScene& S, f3 o, f3 d, float tMax, int& ref_out, float4& tb_out) {
    RayConst rc;
    ray_setup(rc, o, d);
    ref_out = -1; tb_out = make_float4(0, 0, 0, 0);
    const float4 lo = __ldg(&S.nodes[0]), hi = __ldg(&S.nodes[1]);
    float m;
    if (!slab_unbounded(rc, lo, hi, m) || m > tMax) return false;
    const uint32_t a = __float_as_uint(lo.w);
    const int count = (int)(__float_as_uint(hi.w) & CRT_LEAF_COUNT_MASK);
    for (int i = 0; i < count; ++i) {
        const uint32_t ref = __ldg(&S.leaf_refs[a + i]);
        const float4 v0 = __ldg(&S.tris[3 * (size_t)ref]), v1 = __ldg(&S.tris[3 * (size_t)ref + 1]), v2 = __ldg(&S.tris[3 * (size_t)ref + 2]);
        TriCand c;
        if (!tri_test_unbounded_dyn(rc.o, rc.Sx, rc.Sy, rc.Sz, rc.kz, mk3(v0.x, v0.y, v0.z), mk3(v1.x, v1.y, v1.z), mk3(v2.x, v2.y, v2.z), c)) continue;
        if (tri_rejected_by_tmax(c.det, c.tScaled, tMax)) continue;
        if (c.t < tMax) {
            if (ANY) return true;
            tMax = c.t;
            ref_out = (int)ref;
            tb_out = make_float4(c.t, c.b0, c.b1, c.b2);
        }
    }
    return ref_out >= 0;
}

// ------------------------------------------------------------------------------------------------------------
// Ordered traversal (trace_mode 3, the production path).  Same tree, same triangle lists and the same triangle
// arithmetic as above -- only the visit order changes: depth-first, children nearest-octant first, so a hit found
// early culls everything behind it.  The reference's answer is order dependent only among NEAR-TIES (candidates
// whose t differ by rounding), so this pass
//   * culls with a slack bound = tbest * (1 + 2^-12), far wider than any rounding in the tMax tests,
//   * tracks the smallest t among candidates of a DIFFERENT triangle than the current best (t2),
//   * and reports the ray as order-sensitive when t2 <= bound at the end (or when its small stack overflowed).
// Order-sensitive rays are re-traced by the exact BFS kernel; for all others the unique in-band candidate is what
// the BFS loop accepts last (DESIGN.md section 5 gives the argument), with identical t and barycentrics.
// Occlusion queries (fixed tMax) are order independent outright.
#define CRT_FAST_EPS 0x1p-12f
CRT_D float fast_bound(float tbest) { return fminf(tbest * (1.0f + CRT_FAST_EPS), FLT_MAX); }

// ------------------------------------------------------------------------------------------------------------
// One ray per LANE for the descent.  Each lane walks the octree for its own ray with a small stack in shared memory; a lane that
// pops a non-empty leaf parks it, and parked leaves of all lanes are then tested by the whole warp in three pooled stages:
//   stage 0 (fat leaves only): the super-packet boxes of all parked fat leaves are concatenated and dealt to the 32 lanes, one box
//            test per lane against the owning ray; survivors become work descriptors (owner, first sub-packet, sub-packet count);
//   stage 1: the sub-packet boxes (<= CRT_SUBPACKET triangles each, crt_host.h) of all work descriptors -- the parked ordinary leaves
//            themselves, then the surviving super-packets -- are dealt to the lanes the same way; survivors are queued in a ring;
//   stage 2: the ring is consumed 32 / CRT_SUBPACKET sub-packets at a time: one watertight triangle test per lane with the owning
//            ray's constants fetched by shuffle, candidates folded into the owner's state.
#ifndef CRT_WIDE_STACK
#define CRT_WIDE_STACK 16          // entries per lane (8 B each): 32 KB per CTA; deeper stacks cost L1 (shared carve-out) -- overflow goes to the exact kernel
#endif

struct LaneRay {            // per-lane ray constants + traversal state (small fields packed: the kernel lives at 64 registers)
    f3 o, inv_d;
    float Sx, Sy, Sz;
    float tMax0, tbest, bound, t2;
    int href;
    int out_idx;
    uint32_t leaf;             // parked leaf: offset of its reference list in leaf_refs | 0x80000000 for a fat leaf; 0 = none
    uint32_t ctl;              // kz (bits 0-1) | octant order (2-4) | status (5-6) | stack depth (8..)
    // status: 0 idle, 1 traversing, 2 finished (result ready), 3 finished, needs the exact pass
    CRT_D int kz() const { return (int)(ctl & 3u); }
    CRT_D int flip() const { return (int)((ctl >> 2) & 7u); }
    CRT_D int status() const { return (int)((ctl >> 5) & 3u); }
    CRT_D void set_status(int v) { ctl = (ctl & ~0x60u) | ((uint32_t)v << 5); }
    CRT_D int sp() const { return (int)(ctl >> 8); }
    CRT_D void set_sp(int v) { ctl = (ctl & 0xffu) | ((uint32_t)v << 8); }
    CRT_D bool traversing() const { return (ctl & 0x60u) == 0x20u; }
};

CRT_D bool slab_unbounded_oi(f3 o, f3 inv_d, float4 lo, float4 hi, float& min_t_out) {
    RayConst rc;
    rc.o = o; rc.inv_d = inv_d;
    return slab_unbounded(rc, lo, hi, min_t_out);
}

// inclusive prefix sum over the warp
CRT_D int warp_incl_scan(int v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(CRT_FULL, v, o); if (lane >= o) v += u; }
    return v;
}
// owner of item j = number of lanes whose inclusive count is <= j (lanes with count 0 are skipped automatically)
CRT_D int warp_owner_of(int incl, int j) {
    int owner = 0;
#pragma unroll
    for (int step = 16; step > 0; step >>= 1) {
        const int v = __shfl_sync(CRT_FULL, incl, owner + step - 1);
        if (v <= j) owner += step;
    }
    return min(owner, 31);
}

// Stage 2.  Sub-packets that survive their box test are queued in
// a small per-warp ring in shared memory -- (first reference, count | owner lane << 8) -- and consumed 32 / CRT_SUBPACKET at a
// time, so every triangle batch but the last of a phase is full no matter how the survivors are spread over the box batches.
#define CRT_PKQ_CAP 64                                     // ring entries per warp: < 32 / CRT_SUBPACKET left over + 32 new
#define CRT_PKQ_BATCH (32 / CRT_SUBPACKET)                 // entries one triangle batch consumes
static_assert(CRT_SUBPACKET >= 1 && CRT_SUBPACKET <= 32 && (CRT_SUBPACKET & (CRT_SUBPACKET - 1)) == 0, "CRT_SUBPACKET must be a power of two <= 32");

template <bool ANY>
CRT_D void wide_triangle_batch(const DeviceScene& S, LaneRay& r, const uint2* pkq, int head, int navail, float4* hit_tb) {
    const int lane = threadIdx.x & 31;
    const int ei = lane / CRT_SUBPACKET, k = lane % CRT_SUBPACKET;
    uint2 e = make_uint2(0u, 0u);
    if (ei < navail) e = pkq[(head + ei) & (CRT_PKQ_CAP - 1)];
    const int own = (int)(e.y >> 8);
    const bool valid = k < (int)(e.y & 0xffu);
    // the owning ray's constants (valid lanes only use them; all lanes take part in the shuffles)
    const f3 o = mk3(__shfl_sync(CRT_FULL, r.o.x, own), __shfl_sync(CRT_FULL, r.o.y, own), __shfl_sync(CRT_FULL, r.o.z, own));
    const float Sx = __shfl_sync(CRT_FULL, r.Sx, own), Sy = __shfl_sync(CRT_FULL, r.Sy, own), Sz = __shfl_sync(CRT_FULL, r.Sz, own);
    const int kz = __shfl_sync(CRT_FULL, r.kz(), own);
    const float tMax0 = __shfl_sync(CRT_FULL, r.tMax0, own), bound = __shfl_sync(CRT_FULL, r.bound, own);
    TriCand tc;
    tc.det = tc.tScaled = tc.t = tc.b0 = tc.b1 = tc.b2 = 0;
    bool ok = false;
    uint32_t ref = 0;
    if (valid) {
        ref = __ldg(&S.pk_refs[e.x + (uint32_t)k]);
        const float4 v0 = __ldg(&S.tris[3 * (size_t)ref]);
        const float4 v1 = __ldg(&S.tris[3 * (size_t)ref + 1]);
        const float4 v2 = __ldg(&S.tris[3 * (size_t)ref + 2]);
        ok = tri_test_unbounded_dyn(o, Sx, Sy, Sz, kz, mk3(v0.x, v0.y, v0.z), mk3(v1.x, v1.y, v1.z), mk3(v2.x, v2.y, v2.z), tc);
        // candidates are exactly the triangles the reference loop could ever accept (its tests at the initial tMax)
        ok = ok && !tri_rejected_by_tmax(tc.det, tc.tScaled, tMax0) && tc.t < tMax0;
        if (!ANY) ok = ok && !(tc.t > bound);
    }
    unsigned cm = __ballot_sync(CRT_FULL, ok);
    while (cm) {
        const int cl = __ffs(cm) - 1;
        cm &= cm - 1;
        const int sl = __shfl_sync(CRT_FULL, own, cl);
        const float t = __shfl_sync(CRT_FULL, tc.t, cl);
        const int rr = (int)__shfl_sync(CRT_FULL, ref, cl);
        const float b0 = __shfl_sync(CRT_FULL, tc.b0, cl), b1 = __shfl_sync(CRT_FULL, tc.b1, cl), b2 = __shfl_sync(CRT_FULL, tc.b2, cl);
        if (lane != sl) continue;
        if (ANY) { r.href = 1; r.set_status(2); continue; }
        if (rr == r.href) continue;                        // the same triangle met again in another leaf
        if (t < r.tbest) {
            if (r.href >= 0) r.t2 = fminf(r.t2, r.tbest);
            r.tbest = t; r.bound = fast_bound(t);
            r.href = rr;
            hit_tb[r.out_idx] = make_float4(t, b0, b1, b2);
        } else if (!(t > r.bound)) {
            r.t2 = fminf(r.t2, t);
        }
    }
}

// One pooled box stage step: every lane holds a work descriptor (D_FIRST = first box, D_N = box count, 0 = nothing, D_OWNER = the lane
// whose ray it belongs to); box J of the concatenated list is tested against its owner's ray.  A surviving lane gets in E what it must
// enqueue: (the box's first item, its item count | owner << 8).
#define CRT_WIDE_BOX_TEST(INCL, D_FIRST, D_N, D_OWNER, J, TOTAL, E, PASS)                                                                   \
    {                                                                                                                                        \
        const int desc_ = warp_owner_of(INCL, J);                                                                                            \
        const int first_ = __shfl_sync(CRT_FULL, (INCL) - (D_N), desc_);                                                                     \
        const uint32_t box0_ = __shfl_sync(CRT_FULL, D_FIRST, desc_);                                                                        \
        const int own_ = __shfl_sync(CRT_FULL, D_OWNER, desc_);                                                                              \
        const f3 o_ = mk3(__shfl_sync(CRT_FULL, r.o.x, own_), __shfl_sync(CRT_FULL, r.o.y, own_), __shfl_sync(CRT_FULL, r.o.z, own_));       \
        const f3 inv_ = mk3(__shfl_sync(CRT_FULL, r.inv_d.x, own_), __shfl_sync(CRT_FULL, r.inv_d.y, own_), __shfl_sync(CRT_FULL, r.inv_d.z, own_)); \
        const float bound_ = __shfl_sync(CRT_FULL, r.bound, own_);                                                                           \
        const int ostat_ = __shfl_sync(CRT_FULL, r.status(), own_);                                                                            \
        PASS = false;                                                                                                                        \
        E = make_uint2(0u, 0u);                                                                                                              \
        if ((J) < (TOTAL) && ostat_ == 1) {                                                                                                  \
            const size_t pi_ = (size_t)box0_ + (size_t)((J) - first_);                                                                       \
            const float4 lo_ = __ldg(&S.pk_boxes[2 * pi_]), hi_ = __ldg(&S.pk_boxes[2 * pi_ + 1]);                                           \
            float m_;                                                                                                                        \
            PASS = slab_unbounded_oi(o_, inv_, lo_, hi_, m_) && !(m_ > bound_);                                                              \
            E = make_uint2(__float_as_uint(lo_.w), __float_as_uint(hi_.w) | ((uint32_t)own_ << 8));                                          \
        }                                                                                                                                    \
    }

template <bool ANY, bool STATS>
CRT_D void wide_leaf_phase(const DeviceScene& S, LaneRay& r, uint2* pkq, uint2* dq, TraceStats* st, float4* hit_tb) {
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const bool parked = r.traversing() && r.leaf != 0;
    const bool fat = parked && (r.leaf & 0x80000000u);
    uint32_t h0 = 0;
    int hn = 0;
    if (parked) { const uint32_t a = r.leaf & 0x7fffffffu; h0 = __ldg(&S.leaf_refs[a - 2]); hn = (int)__ldg(&S.leaf_refs[a - 1]); }
    if (STATS) { st->nodes += hn; st->leaves += parked ? 1 : 0; }
    // stage 0 input: the super-packets of the parked fat leaves
    const int f_n = fat ? hn : 0;
    const int f_incl = warp_incl_scan(f_n);
    const int f_total = __shfl_sync(CRT_FULL, f_incl, 31);
    int f_base = 0;
    // stage 1 input of the first round: the parked ordinary leaves themselves
    uint32_t d_first = fat ? 0u : h0;
    int d_n = fat ? 0 : hn, d_owner = lane;
    int thead = 0, ttail = 0, dhead = 0, dtail = 0;
    while (true) {
        const bool more_rounds = f_base < f_total || dtail > dhead;
        // ---- stage 1 over the current descriptors, stage 2 whenever the ring holds a full batch (and, in the last round, the rest)
        const int incl = warp_incl_scan(d_n);
        const int total = __shfl_sync(CRT_FULL, incl, 31);
        for (int base = 0; base < total || (base == 0 && !more_rounds && ttail > thead); base += 32) {
            uint2 e;
            bool pass;
            CRT_WIDE_BOX_TEST(incl, d_first, d_n, d_owner, base + lane, total, e, pass);
            const unsigned pm = __ballot_sync(CRT_FULL, pass);
            if (pass) pkq[(ttail + __popc(pm & lt_mask)) & (CRT_PKQ_CAP - 1)] = e;
            if (STATS && pass) st->tris += e.y & 0xffu;
            ttail += __popc(pm);
            __syncwarp();
            const bool last = base + 32 >= total && !more_rounds;
            while (ttail - thead >= CRT_PKQ_BATCH || (last && ttail > thead)) {
                wide_triangle_batch<ANY>(S, r, pkq, thead, min(CRT_PKQ_BATCH, ttail - thead), hit_tb);
                thead += CRT_PKQ_BATCH;
            }
            __syncwarp();
        }
        if (!more_rounds) break;
        // ---- stage 0: super-packet boxes of the fat leaves, until 32 descriptors are queued or the boxes are exhausted
        while (dtail - dhead < 32 && f_base < f_total) {
            uint2 e;
            bool pass;
            CRT_WIDE_BOX_TEST(f_incl, h0, f_n, lane, f_base + lane, f_total, e, pass);
            const unsigned pm = __ballot_sync(CRT_FULL, pass);
            if (pass) dq[(dtail + __popc(pm & lt_mask)) & (CRT_PKQ_CAP - 1)] = e;
            if (STATS && pass) st->nodes += e.y & 0xffu;
            dtail += __popc(pm);
            f_base += 32;
            __syncwarp();
        }
        // the next round's descriptors: one surviving super-packet per lane
        const int navail = min(32, dtail - dhead);
        uint2 e = make_uint2(0u, 0u);
        if (lane < navail) e = dq[(dhead + lane) & (CRT_PKQ_CAP - 1)];
        d_first = e.x; d_n = (int)(e.y & 0xffu); d_owner = (int)(e.y >> 8);
        dhead += navail;
        __syncwarp();
    }
    if (parked) r.leaf = 0;
}

}  // namespace crt
