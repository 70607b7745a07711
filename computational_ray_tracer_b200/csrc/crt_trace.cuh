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

// ------------------------------------------------------------------------------------------------------------
// Ordered traversal (trace_mode 1).  Same tree, same slab test, same triangle lists and the same triangle
// arithmetic as above -- only the visit order changes: depth-first, children nearest-octant first, so a hit found
// early culls everything behind it.  The reference's answer is order dependent only among NEAR-TIES (candidates
// whose t differ by rounding), so this pass
//   * culls with a slack bound = tbest * (1 + 2^-12), far wider than any rounding in the tMax tests,
//   * tracks the smallest t among candidates of a DIFFERENT triangle than the current best (t2),
//   * and reports the ray as order-sensitive when t2 <= bound at the end (or when its small stack overflowed).
// Order-sensitive rays are re-traced by the exact BFS kernel; for all others the unique in-band candidate is what
// the BFS loop accepts last (DESIGN.md section 5 gives the argument), with identical t and barycentrics.
// Occlusion queries (fixed tMax) are order independent outright.
#define CRT_FAST_STACK 64            // uint4 entries per warp: (first child | leaf start, count|flag, entry t, -)
#define CRT_FAST_EPS 0x1p-12f
CRT_D float fast_bound(float tbest) { return fminf(tbest * (1.0f + CRT_FAST_EPS), FLT_MAX); }

struct OrderedState { float tbest, bound, t2; };

// One list of triangle references (a whole small leaf, or one packet of a fat leaf), 32 at a time.
// Returns true only for ANY when an occluder was found.
template <bool ANY>
CRT_D bool ordered_test_refs(const DeviceScene& S, const RayConst& rc, float tMax0, const uint32_t* refs, int count, OrderedState& os, WarpHit& hit) {
    const int lane = threadIdx.x & 31;
    for (int base = 0; base < count; base += 32) {
        const int i = base + lane;
        TriCand tc;
        tc.det = tc.tScaled = tc.t = tc.b0 = tc.b1 = tc.b2 = 0;
        bool ok = false;
        uint32_t ref = 0;
        if (i < count) {
            ref = __ldg(&refs[i]);
            float4 v0 = __ldg(&S.tris[3 * (size_t)ref]);
            float4 v1 = __ldg(&S.tris[3 * (size_t)ref + 1]);
            float4 v2 = __ldg(&S.tris[3 * (size_t)ref + 2]);
            ok = tri_test_unbounded(rc, mk3(v0.x, v0.y, v0.z), mk3(v1.x, v1.y, v1.z), mk3(v2.x, v2.y, v2.z), tc);
            // candidates are exactly the triangles the reference loop could ever accept (its tests at the initial tMax)
            ok = ok && !tri_rejected_by_tmax(tc.det, tc.tScaled, tMax0) && tc.t < tMax0;
            if (!ANY) ok = ok && !(tc.t > os.bound);
        }
        unsigned cm = __ballot_sync(CRT_FULL, ok);
        if constexpr (ANY) {
            if (cm) { hit.ref = 1; return true; }
        } else {
            while (cm) {
                const int cl = __ffs(cm) - 1;
                cm &= cm - 1;
                const float t = __shfl_sync(CRT_FULL, tc.t, cl);
                const int r = (int)__shfl_sync(CRT_FULL, ref, cl);
                if (r == hit.ref) continue;                        // the same triangle met again in another leaf
                if (t < os.tbest) {
                    if (hit.ref >= 0) os.t2 = fminf(os.t2, os.tbest);
                    os.tbest = t; os.bound = fast_bound(t);
                    hit.ref = r; hit.t = t;
                    hit.b0 = __shfl_sync(CRT_FULL, tc.b0, cl);
                    hit.b1 = __shfl_sync(CRT_FULL, tc.b1, cl);
                    hit.b2 = __shfl_sync(CRT_FULL, tc.b2, cl);
                } else if (!(t > os.bound)) {
                    os.t2 = fminf(os.t2, t);
                }
            }
        }
    }
    return false;
}

// returns 0: hit/miss final; 1: order-sensitive, needs the exact pass
template <bool ANY, bool STATS>
CRT_D int trace_ordered_warp(const DeviceScene& S, const RayConst& rc, float tMax0, uint4* stk, WarpHit& hit, TraceStats* st) {
    const int lane = threadIdx.x & 31, g = lane >> 3, c = lane & 7;
    // octant visit order: child bits are (x: bit0 = +x, z: bit1 = +z, y: bit2 = -y), crt_host.cpp split()
    const int flip = (rc.d.x < 0 ? 1 : 0) | (rc.d.z < 0 ? 2 : 0) | (rc.d.y > 0 ? 4 : 0);
    const unsigned lt_mask = (1u << lane) - 1u;
    OrderedState os;
    os.tbest = tMax0; os.bound = ANY ? tMax0 : fast_bound(tMax0); os.t2 = INFINITY;
    hit.ref = -1; hit.t = 0; hit.b0 = hit.b1 = hit.b2 = 0;
    int sp = 0;
    {   // root
        float4 lo = __ldg(&S.nodes[0]), hi = __ldg(&S.nodes[1]);
        float m;
        bool pinf = slab_unbounded(rc, lo, hi, m);
        if (STATS && lane == 0) st->nodes++;
        if (!pinf || m > os.bound) return 0;
        if (lane == 0) stk[0] = make_uint4(__float_as_uint(lo.w), __float_as_uint(hi.w), __float_as_uint(m), 0u);
        sp = 1;
        __syncwarp();
    }
    while (sp > 0) {
        const int ng = min(4, sp);
        uint4 e = make_uint4(0, 0, 0, 0);
        if (g < ng) e = stk[sp - 1 - g];
        const bool expandable = g < ng && !(e.y & CRT_LEAF_FLAG) && !(__uint_as_float(e.z) > os.bound);
        const unsigned gm = __ballot_sync(CRT_FULL, expandable);
        const uint32_t top_a = __shfl_sync(CRT_FULL, e.x, 0), top_b = __shfl_sync(CRT_FULL, e.y, 0);
        const float top_t = __uint_as_float(__shfl_sync(CRT_FULL, e.z, 0));
        if (top_t > os.bound) { --sp; continue; }
        if (top_b & CRT_LEAF_FLAG) {
            --sp;
            const int count = (int)(top_b & CRT_LEAF_COUNT_MASK);
            if (STATS && lane == 0) st->leaves++;
            if (top_b & CRT_LEAF_PACKETS) {
                // fat leaf: test the packet boxes 32 at a time, then only the packets the ray can touch
                const uint32_t pk0 = __ldg(&S.leaf_refs[top_a - 2]);
                const int npk = (int)__ldg(&S.leaf_refs[top_a - 1]);
                for (int pb = 0; pb < npk; pb += 32) {
                    const int pi = pb + lane;
                    float4 lo = make_float4(0, 0, 0, 0), hi = lo;
                    bool pass = false;
                    if (pi < npk) {
                        lo = __ldg(&S.pk_boxes[2 * (size_t)(pk0 + pi)]);
                        hi = __ldg(&S.pk_boxes[2 * (size_t)(pk0 + pi) + 1]);
                        float m;
                        pass = slab_unbounded(rc, lo, hi, m) && !(m > os.bound);
                    }
                    unsigned pm = __ballot_sync(CRT_FULL, pass);
                    if (STATS && lane == 0) st->nodes += min(32, npk - pb);
                    while (pm) {
                        const int pl = __ffs(pm) - 1;
                        pm &= pm - 1;
                        const uint32_t first = __float_as_uint(__shfl_sync(CRT_FULL, lo.w, pl));
                        const int pcnt = (int)__float_as_uint(__shfl_sync(CRT_FULL, hi.w, pl));
                        if (STATS && lane == 0) st->tris += pcnt;
                        if (ordered_test_refs<ANY>(S, rc, tMax0, S.pk_refs + first, pcnt, os, hit)) return 0;
                    }
                }
            } else {
                if (top_b & CRT_LEAF_TIGHT) {
                    const float4* tb = reinterpret_cast<const float4*>(S.leaf_refs + top_a - 8);
                    float m;
                    if (!slab_unbounded(rc, __ldg(&tb[0]), __ldg(&tb[1]), m) || m > os.bound) continue;
                }
                if (STATS && lane == 0) st->tris += count;
                if (ordered_test_refs<ANY>(S, rc, tMax0, S.leaf_refs + top_a, count, os, hit)) return 0;
            }
            continue;
        }
        // expand the leading run of internal entries (nearest first): up to 4 nodes = 32 child boxes at once
        int k = 1;
        if (gm & 0x100u) { k = 2; if (gm & 0x10000u) { k = 3; if (gm & 0x1000000u) k = 4; } }
        sp -= k;
        bool pass = false;
        float4 lo = make_float4(0, 0, 0, 0), hi = lo;
        float m = 0;
        if (g < k) {
            const uint32_t node_idx = e.x + (uint32_t)(c ^ flip);
            lo = __ldg(&S.nodes[2 * (size_t)node_idx]);
            hi = __ldg(&S.nodes[2 * (size_t)node_idx + 1]);
            pass = slab_unbounded(rc, lo, hi, m) && !(m > os.bound);
            const uint32_t b = __float_as_uint(hi.w);
            if ((b & (CRT_LEAF_FLAG | CRT_LEAF_COUNT_MASK)) == CRT_LEAF_FLAG) pass = false;                      // empty leaf: nothing to test
        }
        if (STATS) { if (lane == 0) st->nodes += 8 * k; }
        const unsigned pm = __ballot_sync(CRT_FULL, pass);
        const int npass = __popc(pm);
        if (sp + npass > CRT_FAST_STACK) return 1;                    // stack overflow: let the exact kernel do this ray
        __syncwarp();
        if (pass) stk[sp + npass - 1 - __popc(pm & lt_mask)] = make_uint4(__float_as_uint(lo.w), __float_as_uint(hi.w), __float_as_uint(m), 0u);
        sp += npass;
        if (STATS && lane == 0) st->max_queue = max(st->max_queue, (unsigned)sp);
        __syncwarp();
    }
    if (ANY) return 0;
    return (hit.ref >= 0 && !(os.t2 > os.bound)) ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------------------
// Ordered traversal, four rays per warp (trace_mode 1; superseded as the default by the one-ray-per-lane kernel below).
//
// The octree descent is latency- and issue-bound with only 8 useful lanes per node (8 children), so a warp keeps
// FOUR rays in flight: ray slot s owns lanes 8s..8s+7, its own stack in shared memory and its own traversal state
// (replicated in its 8 lanes).  One node step pops the top entry of every slot and tests 4 x 8 child boxes at once.
// A slot that pops a non-empty leaf parks it as "pending"; pending leaves are then processed one slot at a time by
// all 32 lanes (32 triangles or 32 packet boxes per step), with the ray constants broadcast from the owning slot.
// Semantics (candidate set, slack bound, near-tie detection) are exactly trace_ordered_warp's.
#define CRT_MR_STACK 64

struct SlotRay {            // per-slot ray constants + state, identical in the slot's 8 lanes
    f3 o, inv_d;
    float Sx, Sy, Sz;
    int kz, flip;
    float tMax0, tbest, bound, t2;
    int href; float ht, hb0, hb1, hb2;
    int sp, out_idx;
    uint32_t leaf_a, leaf_b;   // pending leaf (leaf_b == 0: none)
    int status;                // 0 idle, 1 traversing, 2 finished (result ready), 3 finished, needs the exact pass
};

CRT_D bool slab_unbounded_oi(f3 o, f3 inv_d, float4 lo, float4 hi, float& min_t_out) {
    RayConst rc;
    rc.o = o; rc.inv_d = inv_d;
    return slab_unbounded(rc, lo, hi, min_t_out);
}

template <bool ANY, bool STATS, int GROUP_SHIFT = 3>
CRT_D void multi_leaf_phase(const DeviceScene& S, SlotRay& r, int src_lane, TraceStats* st, float4* hit_tb = nullptr) {
    // hit_tb != nullptr (k_trace_wide): an accepted candidate's (t, b0, b1, b2) goes straight to the output record (accepts are rare, a few
    // per ray) instead of living in four registers of every lane for the whole traversal
    const int lane = threadIdx.x & 31;
    // broadcast the owning slot's ray and state (uniform in all 32 lanes from here on)
    RayConst rc;
    rc.o.x = __shfl_sync(CRT_FULL, r.o.x, src_lane); rc.o.y = __shfl_sync(CRT_FULL, r.o.y, src_lane); rc.o.z = __shfl_sync(CRT_FULL, r.o.z, src_lane);
    rc.inv_d.x = __shfl_sync(CRT_FULL, r.inv_d.x, src_lane); rc.inv_d.y = __shfl_sync(CRT_FULL, r.inv_d.y, src_lane); rc.inv_d.z = __shfl_sync(CRT_FULL, r.inv_d.z, src_lane);
    rc.Sx = __shfl_sync(CRT_FULL, r.Sx, src_lane); rc.Sy = __shfl_sync(CRT_FULL, r.Sy, src_lane); rc.Sz = __shfl_sync(CRT_FULL, r.Sz, src_lane);
    rc.kz = __shfl_sync(CRT_FULL, r.kz, src_lane);
    rc.kx = rc.kz + 1; if (rc.kx == 3) rc.kx = 0;
    rc.ky = rc.kx + 1; if (rc.ky == 3) rc.ky = 0;
    rc.d = mk3(0, 0, 0);
    const float tMax0 = __shfl_sync(CRT_FULL, r.tMax0, src_lane);
    OrderedState os;
    os.tbest = __shfl_sync(CRT_FULL, r.tbest, src_lane);
    os.bound = __shfl_sync(CRT_FULL, r.bound, src_lane);
    os.t2 = __shfl_sync(CRT_FULL, r.t2, src_lane);
    WarpHit hit;
    hit.ref = __shfl_sync(CRT_FULL, r.href, src_lane);
    hit.t = 0; hit.b0 = hit.b1 = hit.b2 = 0;
    const int ref_in = hit.ref;
    const float tbest_in = os.tbest;
    const uint32_t leaf_a = __shfl_sync(CRT_FULL, r.leaf_a, src_lane), leaf_b = __shfl_sync(CRT_FULL, r.leaf_b, src_lane);
    const int count = (int)(leaf_b & CRT_LEAF_COUNT_MASK);
    bool any_hit = false;
    if (STATS && lane == 0) st->leaves++;
    if (leaf_b & CRT_LEAF_PACKETS) {
        const uint32_t pk0 = __ldg(&S.leaf_refs[leaf_a - 2]);
        const int npk = (int)__ldg(&S.leaf_refs[leaf_a - 1]);
        for (int pb = 0; pb < npk && !any_hit; pb += 32) {
            const int pi = pb + lane;
            float4 lo = make_float4(0, 0, 0, 0), hi = lo;
            bool pass = false;
            if (pi < npk) {
                lo = __ldg(&S.pk_boxes[2 * (size_t)(pk0 + pi)]);
                hi = __ldg(&S.pk_boxes[2 * (size_t)(pk0 + pi) + 1]);
                float m;
                pass = slab_unbounded(rc, lo, hi, m) && !(m > os.bound);
            }
            unsigned pm = __ballot_sync(CRT_FULL, pass);
            if (STATS && lane == 0) st->nodes += min(32, npk - pb);
            while (pm && !any_hit) {
                const int pl = __ffs(pm) - 1;
                pm &= pm - 1;
                const uint32_t first = __float_as_uint(__shfl_sync(CRT_FULL, lo.w, pl));
                const int pcnt = (int)__float_as_uint(__shfl_sync(CRT_FULL, hi.w, pl));
                if (STATS && lane == 0) st->tris += pcnt;
                any_hit = ordered_test_refs<ANY>(S, rc, tMax0, S.pk_refs + first, pcnt, os, hit);
            }
        }
    } else {
        bool skip = false;
        if (leaf_b & CRT_LEAF_TIGHT) {
            const float4* tb = reinterpret_cast<const float4*>(S.leaf_refs + leaf_a - 8);
            float m;
            skip = !slab_unbounded(rc, __ldg(&tb[0]), __ldg(&tb[1]), m) || m > os.bound;
        }
        if (!skip) {
            if (STATS && lane == 0) st->tris += count;
            any_hit = ordered_test_refs<ANY>(S, rc, tMax0, S.leaf_refs + leaf_a, count, os, hit);
        }
    }
    // write the state back to the owning slot
    if ((lane >> GROUP_SHIFT) == (src_lane >> GROUP_SHIFT)) {
        r.leaf_b = 0;
        if (ANY) { if (any_hit) { r.href = 1; r.status = 2; } }
        else {
            r.t2 = os.t2;
            if (hit.ref != ref_in || os.tbest != tbest_in) {
                r.tbest = os.tbest; r.bound = os.bound;
                r.href = hit.ref;
                if (hit_tb) hit_tb[r.out_idx] = make_float4(hit.t, hit.b0, hit.b1, hit.b2);
                else { r.ht = hit.t; r.hb0 = hit.b0; r.hb1 = hit.b1; r.hb2 = hit.b2; }
            }
        }
    }
}

// Leaf phase for ORDINARY pending leaves of all four slots at once: their reference lists are concatenated and dealt to
// the 32 lanes, so a batch is full even when the individual leaves are small; every lane fetches the constants of the ray
// its triangle belongs to by shuffle.  Candidates are folded into the owning slot's state by that slot's lanes.
template <bool ANY, bool STATS>
CRT_D void multi_leaf_merged(const DeviceScene& S, SlotRay& r, TraceStats* st) {
    const int lane = threadIdx.x & 31, g = lane >> 3;
    // 1. every slot culls its own pending leaf against the padded box of the leaf's triangles (all four slots in parallel)
    // (the padded box of the leaf's triangles was already tested when the leaf was pushed: S.node_tight)
    const bool mine = r.status == 1 && r.leaf_b != 0 && !(r.leaf_b & CRT_LEAF_PACKETS);
    const int my_cnt = mine ? (int)(r.leaf_b & CRT_LEAF_COUNT_MASK) : 0;
    const int c0 = __shfl_sync(CRT_FULL, my_cnt, 0), c1 = __shfl_sync(CRT_FULL, my_cnt, 8), c2 = __shfl_sync(CRT_FULL, my_cnt, 16), c3 = __shfl_sync(CRT_FULL, my_cnt, 24);
    const int p1 = c0, p2 = c0 + c1, p3 = p2 + c2, total = p3 + c3;
    if (total == 0) return;
    const uint32_t a0 = __shfl_sync(CRT_FULL, r.leaf_a, 0), a1 = __shfl_sync(CRT_FULL, r.leaf_a, 8), a2 = __shfl_sync(CRT_FULL, r.leaf_a, 16), a3 = __shfl_sync(CRT_FULL, r.leaf_a, 24);
    if (STATS && lane == 0) { st->tris += total; st->leaves += (c0 > 0) + (c1 > 0) + (c2 > 0) + (c3 > 0); }
    for (int base = 0; base < total; base += 32) {
        const int j = base + lane;
        const bool valid = j < total;
        const int slot = (j >= p1) + (j >= p2) + (j >= p3);
        const int first = slot == 0 ? 0 : (slot == 1 ? p1 : (slot == 2 ? p2 : p3));
        const uint32_t a = slot == 0 ? a0 : (slot == 1 ? a1 : (slot == 2 ? a2 : a3));
        const int src = slot << 3;
        // the owning ray's constants (valid lanes only use them; all lanes take part in the shuffles)
        const f3 o = mk3(__shfl_sync(CRT_FULL, r.o.x, src), __shfl_sync(CRT_FULL, r.o.y, src), __shfl_sync(CRT_FULL, r.o.z, src));
        const float Sx = __shfl_sync(CRT_FULL, r.Sx, src), Sy = __shfl_sync(CRT_FULL, r.Sy, src), Sz = __shfl_sync(CRT_FULL, r.Sz, src);
        const int kz = __shfl_sync(CRT_FULL, r.kz, src);
        const float tMax0 = __shfl_sync(CRT_FULL, r.tMax0, src), bound = __shfl_sync(CRT_FULL, r.bound, src);
        TriCand tc;
        tc.det = tc.tScaled = tc.t = tc.b0 = tc.b1 = tc.b2 = 0;
        bool ok = false;
        uint32_t ref = 0;
        if (valid) {
            ref = __ldg(&S.leaf_refs[a + (uint32_t)(j - first)]);
            float4 v0 = __ldg(&S.tris[3 * (size_t)ref]);
            float4 v1 = __ldg(&S.tris[3 * (size_t)ref + 1]);
            float4 v2 = __ldg(&S.tris[3 * (size_t)ref + 2]);
            ok = tri_test_unbounded_dyn(o, Sx, Sy, Sz, kz, mk3(v0.x, v0.y, v0.z), mk3(v1.x, v1.y, v1.z), mk3(v2.x, v2.y, v2.z), tc);
            ok = ok && !tri_rejected_by_tmax(tc.det, tc.tScaled, tMax0) && tc.t < tMax0;
            if (!ANY) ok = ok && !(tc.t > bound);
        }
        unsigned cm = __ballot_sync(CRT_FULL, ok);
        while (cm) {
            const int cl = __ffs(cm) - 1;
            cm &= cm - 1;
            const int sl = __shfl_sync(CRT_FULL, slot, cl);
            const float t = __shfl_sync(CRT_FULL, tc.t, cl);
            const int rr = (int)__shfl_sync(CRT_FULL, ref, cl);
            const float b0 = __shfl_sync(CRT_FULL, tc.b0, cl), b1 = __shfl_sync(CRT_FULL, tc.b1, cl), b2 = __shfl_sync(CRT_FULL, tc.b2, cl);
            if (g != sl) continue;
            if (ANY) { r.href = 1; r.status = 2; continue; }
            if (rr == r.href) continue;                                // the same triangle met again in another leaf
            if (t < r.tbest) {
                if (r.href >= 0) r.t2 = fminf(r.t2, r.tbest);
                r.tbest = t; r.bound = fast_bound(t);
                r.href = rr; r.ht = t; r.hb0 = b0; r.hb1 = b1; r.hb2 = b2;
            } else if (!(t > r.bound)) {
                r.t2 = fminf(r.t2, t);
            }
        }
    }
    if (mine) r.leaf_b = 0;
}

// ------------------------------------------------------------------------------------------------------------
// Ordered traversal, one ray per LANE for the descent (trace_mode 3, the production kernel).  Each lane walks the octree for its own
// ray with a small stack in shared memory (8 child boxes tested serially per node step, 32 rays per warp-instruction);
// parked leaves of all lanes are then tested by the whole warp in merged 32-triangle batches exactly like
// multi_leaf_merged.  Semantics are those of trace_ordered_warp.
#ifndef CRT_WIDE_STACK
#define CRT_WIDE_STACK 16          // entries per lane (8 B each): 32 KB per CTA; deeper stacks cost L1 (shared carve-out) -- overflow goes to the exact kernel
#endif

template <bool ANY, bool STATS>
CRT_D void wide_leaf_merged(const DeviceScene& S, SlotRay& r, TraceStats* st, float4* hit_tb) {
    const int lane = threadIdx.x & 31;
    const bool mine = r.status == 1 && r.leaf_b != 0 && !(r.leaf_b & CRT_LEAF_PACKETS);
    const int my_cnt = mine ? (int)(r.leaf_b & CRT_LEAF_COUNT_MASK) : 0;
    int incl = my_cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(CRT_FULL, incl, o); if (lane >= o) incl += v; }
    const int excl = incl - my_cnt;
    const int total = __shfl_sync(CRT_FULL, incl, 31);
    if (total == 0) return;
    if (STATS) { st->tris += my_cnt; st->leaves += mine ? 1 : 0; }
    for (int base = 0; base < total; base += 32) {
        const int j = base + lane;
        const bool valid = j < total;
        // owner = number of lanes whose inclusive count is <= j (lanes with count 0 are skipped automatically)
        int owner = 0;
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
            const int v = __shfl_sync(CRT_FULL, incl, owner + step - 1);
            if (v <= j) owner += step;
        }
        owner = min(owner, 31);
        const int first = __shfl_sync(CRT_FULL, excl, owner);
        const uint32_t a = __shfl_sync(CRT_FULL, r.leaf_a, owner);
        const f3 o = mk3(__shfl_sync(CRT_FULL, r.o.x, owner), __shfl_sync(CRT_FULL, r.o.y, owner), __shfl_sync(CRT_FULL, r.o.z, owner));
        const float Sx = __shfl_sync(CRT_FULL, r.Sx, owner), Sy = __shfl_sync(CRT_FULL, r.Sy, owner), Sz = __shfl_sync(CRT_FULL, r.Sz, owner);
        const int kz = __shfl_sync(CRT_FULL, r.kz, owner);
        const float tMax0 = __shfl_sync(CRT_FULL, r.tMax0, owner), bound = __shfl_sync(CRT_FULL, r.bound, owner);
        TriCand tc;
        tc.det = tc.tScaled = tc.t = tc.b0 = tc.b1 = tc.b2 = 0;
        bool ok = false;
        uint32_t ref = 0;
        if (valid) {
            ref = __ldg(&S.leaf_refs[a + (uint32_t)(j - first)]);
            float4 v0 = __ldg(&S.tris[3 * (size_t)ref]);
            float4 v1 = __ldg(&S.tris[3 * (size_t)ref + 1]);
            float4 v2 = __ldg(&S.tris[3 * (size_t)ref + 2]);
            ok = tri_test_unbounded_dyn(o, Sx, Sy, Sz, kz, mk3(v0.x, v0.y, v0.z), mk3(v1.x, v1.y, v1.z), mk3(v2.x, v2.y, v2.z), tc);
            ok = ok && !tri_rejected_by_tmax(tc.det, tc.tScaled, tMax0) && tc.t < tMax0;
            if (!ANY) ok = ok && !(tc.t > bound);
        }
        unsigned cm = __ballot_sync(CRT_FULL, ok);
        while (cm) {
            const int cl = __ffs(cm) - 1;
            cm &= cm - 1;
            const int sl = __shfl_sync(CRT_FULL, owner, cl);
            const float t = __shfl_sync(CRT_FULL, tc.t, cl);
            const int rr = (int)__shfl_sync(CRT_FULL, ref, cl);
            const float b0 = __shfl_sync(CRT_FULL, tc.b0, cl), b1 = __shfl_sync(CRT_FULL, tc.b1, cl), b2 = __shfl_sync(CRT_FULL, tc.b2, cl);
            if (lane != sl) continue;
            if (ANY) { r.href = 1; r.status = 2; continue; }
            if (rr == r.href) continue;
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
    if (mine) r.leaf_b = 0;
}

}  // namespace crt
