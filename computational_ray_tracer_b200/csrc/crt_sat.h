// crt_sat.h -- triangle / box overlap and octree cell geometry, shared by the host builder and the GPU builder
// (identical fp32 operation sequences on both sides: -ffp-contract=off / --fmad=false).
#pragma once
#include "crt_math.h"

namespace crt {

// ------------------------------------------------------------------ Akenine-Moller triangle/box overlap
// ThirdParty/AABB_triangle_Moller.h:196-474, expression for expression.  `never_rejects` reproduces the
// reference's AxisTest_Z0 whose rejecting branch also returns true (:334-345).
CRT_HD bool sat_axis(float pa, float pb, float rad, bool never_rejects = false) {
    float lo, hi;
    if (pa < pb) { lo = pa; hi = pb; } else { lo = pb; hi = pa; }
    if (lo > rad || hi < -rad) return never_rejects;
    return true;
}
CRT_HD bool plane_box_overlap(f3 normal, f3 vert, f3 maxbox) {
    float vmin[3], vmax[3];
    for (int q = 0; q < 3; ++q) {
        float v = comp(vert, q), mb = comp(maxbox, q);
        if (comp(normal, q) > 0.0f) { vmin[q] = -mb - v; vmax[q] = mb - v; }
        else { vmin[q] = mb - v; vmax[q] = -mb - v; }
    }
    if (dot3(normal, mk3(vmin[0], vmin[1], vmin[2])) > 0.0f) return false;
    if (dot3(normal, mk3(vmax[0], vmax[1], vmax[2])) >= 0.0f) return true;
    return false;
}
CRT_HD bool tri_box_overlap(f3 c, f3 h, const f3* t) {
    const f3 v0 = t[0] - c, v1 = t[1] - c, v2 = t[2] - c;
    const f3 e0 = v1 - v0, e1 = v2 - v1, e2 = v0 - v2;
    float fx = fabsf(e0.x), fy = fabsf(e0.y), fz = fabsf(e0.z);
    // edge 0: X01, Y02, Z12
    if (!sat_axis(e0.z * v0.y - e0.y * v0.z, e0.z * v2.y - e0.y * v2.z, fz * h.y + fy * h.z)) return false;
    if (!sat_axis(-e0.z * v0.x + e0.x * v0.z, -e0.z * v2.x + e0.x * v2.z, fz * h.x + fx * h.z)) return false;
    if (!sat_axis(e0.y * v2.x - e0.x * v2.y, e0.y * v1.x - e0.x * v1.y, fy * h.x + fx * h.y)) return false;
    fx = fabsf(e1.x); fy = fabsf(e1.y); fz = fabsf(e1.z);
    // edge 1: X01, Y02, Z0 (the axis that never rejects)
    if (!sat_axis(e1.z * v0.y - e1.y * v0.z, e1.z * v2.y - e1.y * v2.z, fz * h.y + fy * h.z)) return false;
    if (!sat_axis(-e1.z * v0.x + e1.x * v0.z, -e1.z * v2.x + e1.x * v2.z, fz * h.x + fx * h.z)) return false;
    if (!sat_axis(e1.y * v0.x - e1.x * v0.y, e1.y * v1.x - e1.x * v1.y, fy * h.x + fx * h.y, true)) return false;
    fx = fabsf(e2.x); fy = fabsf(e2.y); fz = fabsf(e2.z);
    // edge 2: X2, Y1, Z12
    if (!sat_axis(e2.z * v0.y - e2.y * v0.z, e2.z * v1.y - e2.y * v1.z, fz * h.y + fy * h.z)) return false;
    if (!sat_axis(-e2.z * v0.x + e2.x * v0.z, -e2.z * v1.x + e2.x * v1.z, fz * h.x + fx * h.z)) return false;
    if (!sat_axis(e2.y * v2.x - e2.x * v2.y, e2.y * v1.x - e2.x * v1.y, fy * h.x + fx * h.y)) return false;
    // the three box axes
    for (int a = 0; a < 3; ++a) {
        float x0 = comp(v0, a), x1 = comp(v1, a), x2 = comp(v2, a), lo = x0, hi = x0;
        if (x1 < lo) lo = x1;
        if (x1 > hi) hi = x1;
        if (x2 < lo) lo = x2;
        if (x2 > hi) hi = x2;
        if (lo > comp(h, a) || hi < -comp(h, a)) return false;
    }
    return plane_box_overlap(cross3(e0, e1), v0, h);
}
// Octtree_Model::tri_boundsIntersection (Octtree_Model.h:361-366)
CRT_HD bool tri_in_bounds(const f3* t, const float* bmin, const float* bmax) {
    f3 half = mk3(bmax[0] - bmin[0], bmax[1] - bmin[1], bmax[2] - bmin[2]) / 2.0f;
    f3 c = mk3(bmin[0], bmin[1], bmin[2]) + half;
    return tri_box_overlap(c, half, t);
}


// The child cells of a node (Octtree_Model.h:282-300): 8 octants, half extent padded by 0.01 on the outer faces only.
// child bits: 0 = +x ("right"), 1 = +z ("back"), 2 = -y ("bottom").
CRT_HD void child_cell(const float* bmin, const float* bmax, int k, float* lo, float* hi) {
    f3 hd = mk3(bmax[0] - bmin[0], bmax[1] - bmin[1], bmax[2] - bmin[2]) / 2.0f;
    const f3 C = mk3(bmin[0], bmin[1], bmin[2]) + hd;
    hd = hd + mk3(0.01f, 0.01f, 0.01f);
    const bool right = k & 1, back = (k >> 1) & 1, bottom = (k >> 2) & 1;
    f3 l = C + mk3(right ? 0.0f : -hd.x, bottom ? -hd.y : 0.0f, back ? 0.0f : -hd.z);
    f3 h = C + mk3(right ? hd.x : 0.0f, bottom ? 0.0f : hd.y, back ? hd.z : 0.0f);
    lo[0] = l.x; lo[1] = l.y; lo[2] = l.z; hi[0] = h.x; hi[1] = h.y; hi[2] = h.z;
}
// 8-bit mask of the child cells a triangle overlaps (bit k = child k)
CRT_HD unsigned child_overlap_mask(const float* bmin, const float* bmax, const f3* tri) {
    unsigned m = 0;
    for (int k = 0; k < 8; ++k) {
        float lo[3], hi[3];
        child_cell(bmin, bmax, k, lo, hi);
        if (tri_in_bounds(tri, lo, hi)) m |= 1u << k;
    }
    return m;
}

}  // namespace crt
