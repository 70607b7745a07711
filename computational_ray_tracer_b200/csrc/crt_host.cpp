// crt_host.cpp -- host half of libcrt_b200 (no CUDA): Octtree_Model build + flatten, TriModel helpers,
// camera matrices.  Behaviour follows the reference files cited at each function; the code is organised
// for the flat device layout, not after the reference's object graph.
#include "crt_host.h"
#include "crt_sat.h"

#include <algorithm>
#include <atomic>
#include <cstring>
#include <deque>
#include <thread>

namespace crt {

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }

// ------------------------------------------------------------------ matrices (glm evaluation order)
void m4_identity(float* m) { for (int i = 0; i < 16; ++i) m[i] = (i % 5 == 0) ? 1.0f : 0.0f; }
// result column j = ((a.c0*b0j + a.c1*b1j) + a.c2*b2j) + a.c3*b3j
void m4_mul(const float* a, const float* b, float* out) {
    float r[16];
    for (int j = 0; j < 4; ++j)
        for (int i = 0; i < 4; ++i)
            r[4 * j + i] = ((a[i] * b[4 * j] + a[4 + i] * b[4 * j + 1]) + a[8 + i] * b[4 * j + 2]) + a[12 + i] * b[4 * j + 3];
    std::memcpy(out, r, sizeof r);
}
void m3_mul(const float* a, const float* b, float* out) {
    float r[9];
    for (int j = 0; j < 3; ++j)
        for (int i = 0; i < 3; ++i) r[3 * j + i] = (a[i] * b[3 * j] + a[3 + i] * b[3 * j + 1]) + a[6 + i] * b[3 * j + 2];
    std::memcpy(out, r, sizeof r);
}
void m3_inverse(const float* m, float* out) {
#define E(c, r) m[3 * (c) + (r)]
    float ood = 1.0f / (+E(0, 0) * (E(1, 1) * E(2, 2) - E(2, 1) * E(1, 2)) - E(1, 0) * (E(0, 1) * E(2, 2) - E(2, 1) * E(0, 2)) +
                        E(2, 0) * (E(0, 1) * E(1, 2) - E(1, 1) * E(0, 2)));
    float r[9];
    r[0] = +(E(1, 1) * E(2, 2) - E(2, 1) * E(1, 2)) * ood;
    r[3] = -(E(1, 0) * E(2, 2) - E(2, 0) * E(1, 2)) * ood;
    r[6] = +(E(1, 0) * E(2, 1) - E(2, 0) * E(1, 1)) * ood;
    r[1] = -(E(0, 1) * E(2, 2) - E(2, 1) * E(0, 2)) * ood;
    r[4] = +(E(0, 0) * E(2, 2) - E(2, 0) * E(0, 2)) * ood;
    r[7] = -(E(0, 0) * E(2, 1) - E(2, 0) * E(0, 1)) * ood;
    r[2] = +(E(0, 1) * E(1, 2) - E(1, 1) * E(0, 2)) * ood;
    r[5] = -(E(0, 0) * E(1, 2) - E(1, 0) * E(0, 2)) * ood;
    r[8] = +(E(0, 0) * E(1, 1) - E(1, 0) * E(0, 1)) * ood;
#undef E
    std::memcpy(out, r, sizeof r);
}
// 4x4 inverse by cofactors, sub-determinants grouped the way glm::inverse groups them
void m4_inverse(const float* m, float* out) {
#define E(c, r) m[4 * (c) + (r)]
    const float s00 = E(2, 2) * E(3, 3) - E(3, 2) * E(2, 3), s02 = E(1, 2) * E(3, 3) - E(3, 2) * E(1, 3), s03 = E(1, 2) * E(2, 3) - E(2, 2) * E(1, 3);
    const float s04 = E(2, 1) * E(3, 3) - E(3, 1) * E(2, 3), s06 = E(1, 1) * E(3, 3) - E(3, 1) * E(1, 3), s07 = E(1, 1) * E(2, 3) - E(2, 1) * E(1, 3);
    const float s08 = E(2, 1) * E(3, 2) - E(3, 1) * E(2, 2), s10 = E(1, 1) * E(3, 2) - E(3, 1) * E(1, 2), s11 = E(1, 1) * E(2, 2) - E(2, 1) * E(1, 2);
    const float s12 = E(2, 0) * E(3, 3) - E(3, 0) * E(2, 3), s14 = E(1, 0) * E(3, 3) - E(3, 0) * E(1, 3), s15 = E(1, 0) * E(2, 3) - E(2, 0) * E(1, 3);
    const float s16 = E(2, 0) * E(3, 2) - E(3, 0) * E(2, 2), s18 = E(1, 0) * E(3, 2) - E(3, 0) * E(1, 2), s19 = E(1, 0) * E(2, 2) - E(2, 0) * E(1, 2);
    const float s20 = E(2, 0) * E(3, 1) - E(3, 0) * E(2, 1), s22 = E(1, 0) * E(3, 1) - E(3, 0) * E(1, 1), s23 = E(1, 0) * E(2, 1) - E(2, 0) * E(1, 1);
    const float F0[4] = {s00, s00, s02, s03}, F1[4] = {s04, s04, s06, s07}, F2[4] = {s08, s08, s10, s11};
    const float F3[4] = {s12, s12, s14, s15}, F4[4] = {s16, s16, s18, s19}, F5[4] = {s20, s20, s22, s23};
    const float V0[4] = {E(1, 0), E(0, 0), E(0, 0), E(0, 0)}, V1[4] = {E(1, 1), E(0, 1), E(0, 1), E(0, 1)};
    const float V2[4] = {E(1, 2), E(0, 2), E(0, 2), E(0, 2)}, V3[4] = {E(1, 3), E(0, 3), E(0, 3), E(0, 3)};
    float inv[16];
    for (int k = 0; k < 4; ++k) {
        float sa = (k & 1) ? -1.0f : 1.0f, sb = -sa;
        inv[0 + k] = ((V1[k] * F0[k] - V2[k] * F1[k]) + V3[k] * F2[k]) * sa;
        inv[4 + k] = ((V0[k] * F0[k] - V2[k] * F3[k]) + V3[k] * F4[k]) * sb;
        inv[8 + k] = ((V0[k] * F1[k] - V1[k] * F3[k]) + V3[k] * F5[k]) * sa;
        inv[12 + k] = ((V0[k] * F2[k] - V1[k] * F4[k]) + V2[k] * F5[k]) * sb;
    }
    float d0 = E(0, 0) * inv[0], d1 = E(0, 1) * inv[4], d2 = E(0, 2) * inv[8], d3 = E(0, 3) * inv[12];
    float ood = 1.0f / ((d0 + d1) + (d2 + d3));
#undef E
    for (int i = 0; i < 16; ++i) out[i] = inv[i] * ood;
}
// Shape::Shape (Shapes.h:175-182): ObjectToRender = rigid * permute(y<->z)
void shape_matrices(const float* rigid16, float* o2r, float* r2o) {
    const float perm[16] = {1, 0, 0, 0, 0, 0, 1, 0, 0, 1, 0, 0, 0, 0, 0, 1};
    m4_mul(rigid16, perm, o2r);
    m4_inverse(o2r, r2o);
}
void normal_matrix(const float* m16, float* out9) {
    float inv[16];
    m4_inverse(m16, inv);
    for (int c = 0; c < 3; ++c)
        for (int r = 0; r < 3; ++r) out9[3 * c + r] = inv[4 * r + c];   // transpose of the upper-left block
}
static void m4_translate(const float* m, f3 v, float* out) {
    float r[16];
    std::memcpy(r, m, sizeof r);
    for (int i = 0; i < 4; ++i) r[12 + i] = ((m[i] * v.x + m[4 + i] * v.y) + m[8 + i] * v.z) + m[12 + i];
    std::memcpy(out, r, sizeof r);
}
static void m4_scale(const float* m, f3 v, float* out) {
    float r[16];
    for (int i = 0; i < 4; ++i) { r[i] = m[i] * v.x; r[4 + i] = m[4 + i] * v.y; r[8 + i] = m[8 + i] * v.z; r[12 + i] = m[12 + i]; }
    std::memcpy(out, r, sizeof r);
}

}  // namespace crt

using namespace crt;

static const int kLeafCapacity = 40;
static int g_build_algorithm = 0;    // Octtree_Model::TRIANGLE_CAPACITY (Octtree_Model.h:388)

// Octtree_Model::AddTriangle (Octtree_Model.h:180-277): breadth-first over every node the triangle overlaps;
// a leaf that reaches the capacity is split on the spot, its new children are NOT visited for this triangle
// (the split already re-binned it).
void crt_octree::add_triangle(uint32_t gid) {
    const f3* t = &world_pos[3 * (size_t)gid];
    static thread_local std::vector<int> fifo;
    fifo.clear();
    fifo.push_back(0);
    for (size_t head = 0; head < fifo.size(); ++head) {
        int id = fifo[head];
        if (!tri_in_bounds(t, nodes[id].bmin, nodes[id].bmax)) continue;
        if (nodes[id].leaf) {
            nodes[id].tris.push_back(gid);
            if ((int)nodes[id].tris.size() >= kLeafCapacity) split(id);
        } else {
            for (int k = 0; k < 8; ++k) fifo.push_back(nodes[id].child[k]);
        }
    }
}

// Octtree_Model::Split (Octtree_Model.h:279-358)
void crt_octree::split(int id) {
    const float* bmin = nodes[id].bmin;
    const float* bmax = nodes[id].bmax;
    f3 hd = mk3(bmax[0] - bmin[0], bmax[1] - bmin[1], bmax[2] - bmin[2]) / 2.0f;
    const f3 C = mk3(bmin[0], bmin[1], bmin[2]) + hd;
    hd = hd + mk3(0.01f, 0.01f, 0.01f);            // padding, outer faces only
    // child order: top{FL,FR,BL,BR}, bottom{FL,FR,BL,BR}; "top" = +y, "front" = -z, "left" = -x
    HostOctreeNode kids[8];
    for (int k = 0; k < 8; ++k) {
        const bool right = k & 1, back = (k >> 1) & 1, bottom = (k >> 2) & 1;
        f3 lo = C + mk3(right ? 0.0f : -hd.x, bottom ? -hd.y : 0.0f, back ? 0.0f : -hd.z);
        f3 hi = C + mk3(right ? hd.x : 0.0f, bottom ? 0.0f : hd.y, back ? hd.z : 0.0f);
        kids[k].bmin[0] = lo.x; kids[k].bmin[1] = lo.y; kids[k].bmin[2] = lo.z;
        kids[k].bmax[0] = hi.x; kids[k].bmax[1] = hi.y; kids[k].bmax[2] = hi.z;
    }
    const std::vector<uint32_t>& parent_tris = nodes[id].tris;
    if (nodes[id].memo) {
        // an earlier attempt was aborted: only the triangles added since then can change the outcome
        uint8_t mask = nodes[id].full_mask;
        for (size_t i = nodes[id].memo_count; i < parent_tris.size() && mask; ++i) {
            const f3* t = &world_pos[3 * (size_t)parent_tris[i]];
            for (int k = 0; k < 8; ++k)
                if ((mask >> k & 1) && !tri_in_bounds(t, kids[k].bmin, kids[k].bmax)) mask &= (uint8_t)~(1u << k);
        }
        nodes[id].full_mask = mask;
        nodes[id].memo_count = (uint32_t)parent_tris.size();
        if (mask) return;                                           // some child would still receive everything
    }
    for (uint32_t gid : parent_tris) {
        const f3* t = &world_pos[3 * (size_t)gid];
        for (int k = 0; k < 8; ++k)
            if (tri_in_bounds(t, kids[k].bmin, kids[k].bmax)) kids[k].tris.push_back(gid);
    }
    uint8_t full = 0;
    for (int k = 0; k < 8; ++k)
        if (kids[k].tris.size() == parent_tris.size()) full |= (uint8_t)(1u << k);
    if (full) {                                                     // a child swallowed everything: stay a fat leaf
        nodes[id].memo = true; nodes[id].full_mask = full; nodes[id].memo_count = (uint32_t)parent_tris.size();
        return;
    }
    for (int k = 0; k < 8; ++k) {
        kids[k].parent = id;
        nodes.push_back(std::move(kids[k]));
        nodes[id].child[k] = (int)nodes.size() - 1;
    }
    nodes[id].tris.clear();
    nodes[id].tris.shrink_to_fit();
    nodes[id].leaf = false;
}

static void child_cells(const float* bmin, const float* bmax, float lo[8][3], float hi[8][3]) {
    for (int k = 0; k < 8; ++k) child_cell(bmin, bmax, k, lo[k], hi[k]);
}

// Top-down construction of the tree the incremental insertion produces.
//
// Inserting triangles one at a time looks order dependent, but the outcome is a function of each node's triangle SET:
// a triangle reaches a node iff it overlaps the cells of the node and of all its ancestors; a leaf's list is in
// insertion (global id) order because splits re-bin in list order and later arrivals are appended; and a split attempt
// happens exactly when an insertion lands in a leaf that then holds >= 40 triangles (Octtree_Model.h:216-223), succeeding
// iff no child cell overlaps ALL triangles present (:332-340).  "Some child overlaps all" can only get rarer as the set
// grows, so the node splits at the FIRST triangle g (in id order) that (1) arrives after the node was created (the re-binning
// that creates a child never splits it, whatever its size), (2) brings the count to >= 40 and (3) leaves no child
// overlapping the whole prefix.  Children are created at time g and receive every triangle of the node that overlaps them.
// Node numbering differs from the incremental builder's creation order; the flattened (breadth-first) layout -- all the
// device ever sees -- is identical, which tests/test_cpu_host_parity.py checks array by array.
void crt_octree::build_topdown() {
    struct Work { int node; int64_t tau; };
    std::vector<Work> queue;
    {
        HostOctreeNode& root = nodes[0];
        const uint32_t total = mesh_first.back();
        for (uint32_t gid = 0; gid < total; ++gid)
            if (tri_in_bounds(&world_pos[3 * (size_t)gid], root.bmin, root.bmax)) root.tris.push_back(gid);
    }
    queue.push_back({0, -1});
    std::vector<uint8_t> masks;
    for (size_t qi = 0; qi < queue.size(); ++qi) {
        const int id = queue[qi].node;
        const int64_t tau = queue[qi].tau;
        const size_t n = nodes[id].tris.size();
        if ((int)n < kLeafCapacity) continue;
        float lo[8][3], hi[8][3];
        child_cells(nodes[id].bmin, nodes[id].bmax, lo, hi);
        masks.assign(n, 0);
        uint8_t running = 0xff;
        int64_t split_at = -1;
        for (size_t i = 0; i < n; ++i) {
            const f3* t = &world_pos[3 * (size_t)nodes[id].tris[i]];
            uint8_t m = 0;
            for (int k = 0; k < 8; ++k)
                if (tri_in_bounds(t, lo[k], hi[k])) m |= (uint8_t)(1u << k);
            masks[i] = m;
            running &= m;
            if (split_at < 0 && (int)(i + 1) >= kLeafCapacity && (int64_t)nodes[id].tris[i] > tau && running == 0) split_at = (int64_t)i;
        }
        if (split_at < 0) continue;                                   // stays a (possibly fat) leaf
        const int64_t g = (int64_t)nodes[id].tris[(size_t)split_at];
        const int first = (int)nodes.size();
        for (int k = 0; k < 8; ++k) {
            HostOctreeNode kid;
            for (int a = 0; a < 3; ++a) { kid.bmin[a] = lo[k][a]; kid.bmax[a] = hi[k][a]; }
            kid.parent = id;
            nodes.push_back(std::move(kid));
        }
        for (size_t i = 0; i < n; ++i)
            for (int k = 0; k < 8; ++k)
                if (masks[i] >> k & 1) nodes[first + k].tris.push_back(nodes[id].tris[i]);
        for (int k = 0; k < 8; ++k) { nodes[id].child[k] = first + k; queue.push_back({first + k, g}); }
        nodes[id].tris.clear();
        nodes[id].tris.shrink_to_fit();
        nodes[id].leaf = false;
    }
}

// Packets of <= packet_size triangles with padded boxes for one leaf (crt_host.h, "Triangle packets"): the leaf's triangles are
// split recursively at the median of their centroids along the widest axis (the left part a multiple of packet_size, so only the
// last packet can be short) -- on the thin surface patches an octree leaf holds, this gives boxes that a ray misses about twice as
// often as Morton-ordered runs do.
static const float kPacketPad = 0x1p-15f;      // of the largest coordinate magnitude: >> the few-ulp slack of the watertight test
static void split_packets(const std::vector<f3>& cen, uint32_t* idx, size_t n, uint32_t packet_size, std::vector<uint32_t>& order) {
    if (n <= packet_size) { order.insert(order.end(), idx, idx + n); return; }
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (size_t i = 0; i < n; ++i)
        for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], comp(cen[idx[i]], a)); hi[a] = std::max(hi[a], comp(cen[idx[i]], a)); }
    int ax = 0;
    if (hi[1] - lo[1] > hi[ax] - lo[ax]) ax = 1;
    if (hi[2] - lo[2] > hi[ax] - lo[ax]) ax = 2;
    size_t k = ((n / 2 + packet_size - 1) / packet_size) * packet_size;
    if (k >= n) k = n / 2;
    std::stable_sort(idx, idx + n, [&](uint32_t x, uint32_t y) { return comp(cen[x], ax) < comp(cen[y], ax); });
    split_packets(cen, idx, k, packet_size, order);
    split_packets(cen, idx + k, n - k, packet_size, order);
}
uint32_t crt_octree::build_packets(const std::vector<uint32_t>& tris, uint32_t packet_size, FlatOctree* out) const {
    std::vector<f3> cen(tris.size());
    for (size_t i = 0; i < tris.size(); ++i) {
        const f3* t = &world_pos[3 * (size_t)tris[i]];
        cen[i] = mk3((t[0].x + t[1].x + t[2].x) / 3, (t[0].y + t[1].y + t[2].y) / 3, (t[0].z + t[1].z + t[2].z) / 3);
    }
    std::vector<uint32_t> idx(tris.size()), order;
    for (size_t i = 0; i < idx.size(); ++i) idx[i] = (uint32_t)i;
    order.reserve(tris.size());
    const uint32_t super_size = packet_size * CRT_SUPERPACKET;
    if (tris.size() > super_size) {
        // fat leaf: first the super-packets (<= 32 sub-packets' worth of triangles each), then every super-packet into sub-packets, so
        // that the sub-packets of one super-packet are consecutive
        std::vector<uint32_t> coarse;
        coarse.reserve(tris.size());
        split_packets(cen, idx.data(), idx.size(), super_size, coarse);
        for (size_t base = 0; base < coarse.size(); base += super_size)
            split_packets(cen, coarse.data() + base, std::min<size_t>(super_size, coarse.size() - base), packet_size, order);
    } else {
        split_packets(cen, idx.data(), idx.size(), packet_size, order);
    }
    auto emit_box = [&](const float* lo, const float* hi, uint32_t first, uint32_t cnt) {
        float mag = 0;
        for (int a = 0; a < 3; ++a) mag = std::max(mag, std::max(std::fabs(lo[a]), std::fabs(hi[a])));
        const float pad = std::max(mag * kPacketPad, 1e-6f);
        float rec[8] = {lo[0] - pad, lo[1] - pad, lo[2] - pad, 0, hi[0] + pad, hi[1] + pad, hi[2] + pad, 0};
        std::memcpy(&rec[3], &first, 4); std::memcpy(&rec[7], &cnt, 4);
        out->pk_boxes.insert(out->pk_boxes.end(), rec, rec + 8);
    };
    const uint32_t first_packet = (uint32_t)(out->pk_boxes.size() / 8);
    uint32_t n_packets = 0;
    std::vector<float> plo, phi;        // unpadded sub-packet bounds, for the super-packet boxes
    for (size_t base = 0; base < order.size(); base += packet_size, ++n_packets) {
        const size_t end = std::min(order.size(), base + packet_size);
        float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        const uint32_t first = (uint32_t)out->pk_refs.size();
        for (size_t k = base; k < end; ++k) {
            const uint32_t gid = tris[order[k]];
            out->pk_refs.push_back(gid);
            const f3* t = &world_pos[3 * (size_t)gid];
            for (int v = 0; v < 3; ++v)
                for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], comp(t[v], a)); hi[a] = std::max(hi[a], comp(t[v], a)); }
        }
        emit_box(lo, hi, first, (uint32_t)(end - base));
        plo.insert(plo.end(), lo, lo + 3); phi.insert(phi.end(), hi, hi + 3);
    }
    if (tris.size() <= super_size) return n_packets;
    uint32_t n_super = 0;
    for (uint32_t base = 0; base < n_packets; base += CRT_SUPERPACKET, ++n_super) {
        const uint32_t end = std::min(n_packets, base + CRT_SUPERPACKET);
        float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        for (uint32_t k = base; k < end; ++k)
            for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], plo[3 * k + a]); hi[a] = std::max(hi[a], phi[3 * k + a]); }
        emit_box(lo, hi, first_packet + base, end - base);
    }
    return n_super;
}

// Linearise: nodes renumbered in breadth-first order (the order Octtree_Model::Traverse pops them, so a
// ray's visit sequence is ascending in the new ids and 8 siblings are contiguous).
void crt_octree::flatten(const std::vector<uint8_t>& skip, FlatOctree* out) const {
    out->nodes.clear(); out->leaf_refs.clear(); out->pk_boxes.clear(); out->pk_refs.clear(); out->node_tight.clear();
    out->bfs_of_ref.assign(nodes.size(), -1);
    std::vector<int> order;
    std::vector<int> depth;
    order.reserve(nodes.size());
    order.push_back(0); depth.push_back(1);
    for (size_t head = 0; head < order.size(); ++head) {
        const HostOctreeNode& n = nodes[order[head]];
        out->bfs_of_ref[order[head]] = (int)head;
        out->depth = std::max(out->depth, depth[head]);
        if (!n.leaf) for (int k = 0; k < 8; ++k) { order.push_back(n.child[k]); depth.push_back(depth[head] + 1); }
    }
    out->nodes.resize(8 * order.size());
    // Leaves first, in parallel: the surviving references of every leaf and their packets (the per-leaf median splits are the expensive
    // part of a flatten -- 2 M leaves at 10 M triangles).  Each worker fills a chunk-local FlatOctree for a contiguous range of BFS ids;
    // the chunks are then appended in order, so the layout does not depend on the number of threads.
    struct LeafRec { size_t node; uint32_t kept_at, cnt, first_box, n_boxes; bool fat; };
    struct Chunk { FlatOctree local; std::vector<uint32_t> kept; std::vector<LeafRec> leaves; std::vector<std::pair<uint32_t, uint32_t>> super_ranges; };
    const unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    const size_t n_chunks = order.size() < 4096 ? 1 : (size_t)hw * 4;
    std::vector<Chunk> chunks(n_chunks);
    auto work = [&](size_t ci) {
        Chunk& ch = chunks[ci];
        const size_t lo = order.size() * ci / n_chunks, hi = order.size() * (ci + 1) / n_chunks;
        std::vector<uint32_t> kept;
        for (size_t i = lo; i < hi; ++i) {
            const HostOctreeNode& n = nodes[order[i]];
            if (!n.leaf) continue;
            kept.clear();
            for (uint32_t gid : n.tris)
                if (skip.empty() || !skip[gid]) kept.push_back(gid);
            LeafRec r{i, (uint32_t)ch.kept.size(), (uint32_t)kept.size(), 0, 0, false};
            if (r.cnt > 0) {
                // header: (first box, box count) -- the leaf's sub-packets, or for a fat leaf its super-packets (which follow its sub-packets)
                r.fat = r.cnt > (uint32_t)CRT_SUBPACKET * CRT_SUPERPACKET;
                const uint32_t before = (uint32_t)(ch.local.pk_boxes.size() / 8);
                r.n_boxes = build_packets(kept, (uint32_t)CRT_SUBPACKET, &ch.local);
                r.first_box = r.fat ? (uint32_t)(ch.local.pk_boxes.size() / 8) - r.n_boxes : before;
                if (r.fat) ch.super_ranges.push_back({r.first_box, r.n_boxes});
                ch.kept.insert(ch.kept.end(), kept.begin(), kept.end());
            }
            ch.leaves.push_back(r);
        }
    };
    if (n_chunks == 1) work(0);
    else {
        std::vector<std::thread> pool;
        std::atomic<size_t> next{0};
        for (unsigned t = 0; t < hw; ++t) pool.emplace_back([&] { for (size_t ci; (ci = next.fetch_add(1)) < n_chunks;) work(ci); });
        for (auto& t : pool) t.join();
    }
    for (Chunk& ch : chunks) {
        const uint32_t ref_base = (uint32_t)out->pk_refs.size(), box_base = (uint32_t)(out->pk_boxes.size() / 8);
        // relocate the chunk-local indices: a sub-packet's `first` counts pk_refs, a super-packet's counts boxes
        std::vector<uint8_t> is_super(ch.local.pk_boxes.size() / 8, 0);
        for (auto& sr : ch.super_ranges) for (uint32_t k = 0; k < sr.second; ++k) is_super[sr.first + k] = 1;
        for (size_t bx = 0; bx < is_super.size(); ++bx) {
            uint32_t first;
            std::memcpy(&first, &ch.local.pk_boxes[8 * bx + 3], 4);
            first += is_super[bx] ? box_base : ref_base;
            std::memcpy(&ch.local.pk_boxes[8 * bx + 3], &first, 4);
        }
        out->pk_refs.insert(out->pk_refs.end(), ch.local.pk_refs.begin(), ch.local.pk_refs.end());
        out->pk_boxes.insert(out->pk_boxes.end(), ch.local.pk_boxes.begin(), ch.local.pk_boxes.end());
        for (const LeafRec& r : ch.leaves) {
            uint32_t b = 0x80000000u | r.cnt;
            if (r.cnt > 0) {
                out->leaf_refs.push_back(box_base + r.first_box);
                out->leaf_refs.push_back(r.n_boxes);
                b |= r.fat ? CRT_PACKET_FLAG : CRT_SUBPK_FLAG;
            }
            const uint32_t a = (uint32_t)out->leaf_refs.size();
            out->leaf_refs.insert(out->leaf_refs.end(), ch.kept.begin() + r.kept_at, ch.kept.begin() + r.kept_at + r.cnt);
            std::memcpy(&out->nodes[8 * r.node + 3], &a, 4); std::memcpy(&out->nodes[8 * r.node + 7], &b, 4);
        }
        ch = Chunk();          // release the chunk's memory as soon as it is merged
    }
    size_t next_child = 1;
    for (size_t i = 0; i < order.size(); ++i) {
        const HostOctreeNode& n = nodes[order[i]];
        float* d = &out->nodes[8 * i];
        if (!n.leaf) {
            const uint32_t a = (uint32_t)next_child, b = 0;
            next_child += 8;
            std::memcpy(&d[3], &a, 4); std::memcpy(&d[7], &b, 4);
        }
        d[0] = n.bmin[0]; d[1] = n.bmin[1]; d[2] = n.bmin[2];
        d[4] = n.bmax[0]; d[5] = n.bmax[1]; d[6] = n.bmax[2];
    }
    // internal nodes: the low 8 bits of b flag the children that hold anything at all (an internal child with a non-zero
    // mask, or a leaf with references), so a traversal can skip empty octants without touching their records
    for (size_t i = order.size(); i-- > 0;) {
        uint32_t a, b;
        std::memcpy(&a, &out->nodes[8 * i + 3], 4); std::memcpy(&b, &out->nodes[8 * i + 7], 4);
        if (b & 0x80000000u) continue;
        uint32_t mask = 0;
        for (int k = 0; k < 8; ++k) {
            uint32_t cb;
            std::memcpy(&cb, &out->nodes[8 * (size_t)(a + k) + 7], 4);
            const bool nonempty = (cb & 0x80000000u) ? (cb & 0x1fffffffu) != 0 : (cb & 0xffu) != 0;
            if (nonempty) mask |= 1u << k;
        }
        b = mask;
        std::memcpy(&out->nodes[8 * i + 7], &b, 4);
    }
    // subtree bounds for the ordered traversal: the octree's cells are much larger than the surface patch beneath them
    // (a cell is kept whenever a triangle touches it anywhere), so every node also gets the padded box of the triangles
    // stored in its subtree.  Children follow parents in BFS order: one reverse sweep folds them upwards.
    out->node_tight.assign(8 * order.size(), 0.0f);
    for (size_t i = order.size(); i-- > 0;) {
        const HostOctreeNode& n = nodes[order[i]];
        float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        if (n.leaf) {
            for (uint32_t gid : n.tris) {
                if (!skip.empty() && skip[gid]) continue;
                const f3* t = &world_pos[3 * (size_t)gid];
                for (int v = 0; v < 3; ++v)
                    for (int a2 = 0; a2 < 3; ++a2) { lo[a2] = std::min(lo[a2], comp(t[v], a2)); hi[a2] = std::max(hi[a2], comp(t[v], a2)); }
            }
            if (lo[0] <= hi[0]) {
                float mag = 0;
                for (int a2 = 0; a2 < 3; ++a2) mag = std::max(mag, std::max(std::fabs(lo[a2]), std::fabs(hi[a2])));
                const float pad = std::max(mag * kPacketPad, 1e-6f);
                for (int a2 = 0; a2 < 3; ++a2) { lo[a2] -= pad; hi[a2] += pad; }
            }
        } else {
            uint32_t first;
            std::memcpy(&first, &out->nodes[8 * i + 3], 4);
            for (int k = 0; k < 8; ++k) {
                const float* c = &out->node_tight[8 * (size_t)(first + k)];
                for (int a2 = 0; a2 < 3; ++a2) { lo[a2] = std::min(lo[a2], c[a2]); hi[a2] = std::max(hi[a2], c[4 + a2]); }
            }
        }
        float* d = &out->node_tight[8 * i];
        d[0] = lo[0]; d[1] = lo[1]; d[2] = lo[2]; d[4] = hi[0]; d[5] = hi[1]; d[6] = hi[2];
    }
    // pad the reference list so 128-bit loads past the end of the last leaf stay in bounds
    for (int i = 0; i < 4; ++i) out->leaf_refs.push_back(0);
}

namespace crt {
crt_octree* octree_prepare(const crt_mesh_desc* meshes, uint32_t n_meshes, const float* o2r, int precomputed_world) {
    if (!meshes || !n_meshes || !o2r) { set_error("octree_build: bad arguments"); return nullptr; }
    auto* oct = new crt_octree;
    oct->mesh_first.assign(n_meshes + 1, 0);
    for (uint32_t m = 0; m < n_meshes; ++m) oct->mesh_first[m + 1] = oct->mesh_first[m] + meshes[m].n_triangles;
    oct->world_pos.resize(3 * (size_t)oct->mesh_first[n_meshes]);
    for (uint32_t m = 0; m < n_meshes; ++m)
        for (uint32_t t = 0; t < meshes[m].n_triangles; ++t)
            for (int k = 0; k < 3; ++k) {
                uint32_t vi = meshes[m].indices[3 * (size_t)t + k];
                if (vi >= meshes[m].n_vertices) { set_error("octree_build: index out of range"); delete oct; return nullptr; }
                const float* p = &meshes[m].positions[3 * (size_t)vi];
                f3 w = mk3(p[0], p[1], p[2]);
                if (!precomputed_world) w = xform_point(o2r, w);        // Octtree_Model.h:192-197
                oct->world_pos[3 * ((size_t)oct->mesh_first[m] + t) + k] = w;
            }
    HostOctreeNode root;
    float b[6];
    crt_model_bounds(meshes, n_meshes, o2r, precomputed_world, b);
    for (int a = 0; a < 3; ++a) { root.bmin[a] = b[a]; root.bmax[a] = b[3 + a]; }
    root.parent = 0;
    oct->nodes.reserve(10000);
    oct->nodes.push_back(std::move(root));
    return oct;
}
}  // namespace crt

// ------------------------------------------------------------------ C ABI: host-only entry points
extern "C" {

const char* crt_last_error(void) { return crt::g_err.c_str(); }
int crt_version(void) { return 100; }

// TriModel ctor bounds + TriModel::Bounds (Shapes.h:1282-1300,1390-1397) and Bounds3::Transform (:59-98)
int crt_model_bounds(const crt_mesh_desc* meshes, uint32_t n_meshes, const float* o2r, int precomputed_world, float* out) {
    const float big = FLT_MAX, tiny = FLT_MIN;      // the max side starts at FLT_MIN (smallest positive), SURVEY 5.1-4
    float mn[3] = {big, big, big}, mx[3] = {tiny, tiny, tiny};
    for (uint32_t m = 0; m < n_meshes; ++m)
        for (uint32_t v = 0; v < meshes[m].n_vertices; ++v)
            for (int a = 0; a < 3; ++a) {
                float p = meshes[m].positions[3 * (size_t)v + a];
                mn[a] = std::min(mn[a], p);
                mx[a] = std::max(mx[a], p);
            }
    if (!precomputed_world) {
        const float cx[2] = {mn[0], mx[0]}, cy[2] = {mn[1], mx[1]}, cz[2] = {mn[2], mx[2]};
        float tmn[3] = {big, big, big}, tmx[3] = {tiny, tiny, tiny};
        for (int i = 0; i < 8; ++i) {
            f3 p = xform_point(o2r, mk3(cx[i & 1], cy[(i >> 1) & 1], cz[(i >> 2) & 1]));
            tmn[0] = std::min(tmn[0], p.x); tmx[0] = std::max(tmx[0], p.x);
            tmn[1] = std::min(tmn[1], p.y); tmx[1] = std::max(tmx[1], p.y);
            tmn[2] = std::min(tmn[2], p.z); tmx[2] = std::max(tmx[2], p.z);
        }
        std::memcpy(mn, tmn, sizeof mn); std::memcpy(mx, tmx, sizeof mx);
    }
    for (int a = 0; a < 3; ++a) { out[a] = mn[a]; out[3 + a] = mx[a]; }
    return 0;
}

int crt_model_compute_backface(const crt_mesh_desc* mesh, const float* look3, const float* o2r, int precomputed_world, uint8_t* out) {
    if (!mesh->normals) { set_error("compute_backface: mesh has no normals"); return 1; }
    f3 look = normalize3(mk3(look3[0], look3[1], look3[2]));
    float nm[9];
    if (!precomputed_world) normal_matrix(o2r, nm);
    for (uint32_t t = 0; t < mesh->n_triangles; ++t) {
        const float* a = &mesh->normals[3 * (size_t)mesh->indices[3 * (size_t)t]];
        const float* b = &mesh->normals[3 * (size_t)mesh->indices[3 * (size_t)t + 1]];
        const float* c = &mesh->normals[3 * (size_t)mesh->indices[3 * (size_t)t + 2]];
        f3 N = normalize3(((mk3(a[0], a[1], a[2]) + mk3(b[0], b[1], b[2])) + mk3(c[0], c[1], c[2])) / 3.0f);
        if (!precomputed_world) N = normalize3(mul_m3_v3(nm, N));
        out[t] = dot3(look, N) > 0 ? 1 : 0;
    }
    return 0;
}

int crt_octree_build(const crt_mesh_desc* meshes, uint32_t n_meshes, const float* o2r, int precomputed_world, crt_octree** out) {
    if (!out) { set_error("octree_build: bad arguments"); return 1; }
    crt_octree* oct = crt::octree_prepare(meshes, n_meshes, o2r, precomputed_world);
    if (!oct) return 1;
    const uint32_t total = oct->mesh_first[n_meshes];
    if (g_build_algorithm == 1) oct->build_topdown();
    else for (uint32_t gid = 0; gid < total; ++gid) oct->add_triangle(gid);     // mesh-major, triangle-minor (Octtree_Model.h:54-62)
    *out = oct;
    return 0;
}
// The device layout of an octree (what crt_scene_set_model uploads when nothing is culled), for builder-equivalence tests.
int crt_octree_flat_sizes(const crt_octree* oct, uint64_t* sizes5) {
    if (!oct || !sizes5) { set_error("octree_flat_sizes: bad arguments"); return 1; }
    FlatOctree f;
    oct->flatten({}, &f);
    sizes5[0] = f.nodes.size(); sizes5[1] = f.leaf_refs.size(); sizes5[2] = f.node_tight.size(); sizes5[3] = f.pk_boxes.size(); sizes5[4] = f.pk_refs.size();
    return 0;
}
int crt_octree_flat_copy(const crt_octree* oct, float* nodes, uint32_t* leaf_refs, float* node_tight, float* pk_boxes, uint32_t* pk_refs) {
    if (!oct) { set_error("octree_flat_copy: bad arguments"); return 1; }
    FlatOctree f;
    oct->flatten({}, &f);
    if (nodes) std::memcpy(nodes, f.nodes.data(), f.nodes.size() * 4);
    if (leaf_refs) std::memcpy(leaf_refs, f.leaf_refs.data(), f.leaf_refs.size() * 4);
    if (node_tight) std::memcpy(node_tight, f.node_tight.data(), f.node_tight.size() * 4);
    if (pk_boxes) std::memcpy(pk_boxes, f.pk_boxes.data(), f.pk_boxes.size() * 4);
    if (pk_refs) std::memcpy(pk_refs, f.pk_refs.data(), f.pk_refs.size() * 4);
    return 0;
}
int crt_octree_set_build_algorithm(int algorithm) {
    if (algorithm < 0 || algorithm > 1) { set_error("octree_set_build_algorithm: 0 incremental insertion (reference order), 1 top-down"); return 1; }
    g_build_algorithm = algorithm;
    return 0;
}
void crt_octree_destroy(crt_octree* oct) { delete oct; }
int crt_octree_node_count(const crt_octree* oct) { return (int)oct->nodes.size(); }

int crt_octree_get_stats(const crt_octree* oct, crt_octree_stats* s) {
    std::memset(s, 0, sizeof *s);
    s->nodes = (int32_t)oct->nodes.size();
    std::vector<std::pair<int, int>> q{{0, 1}};
    for (size_t h = 0; h < q.size(); ++h) {
        const HostOctreeNode& n = oct->nodes[q[h].first];
        s->real_nodes++;
        s->depth = std::max(s->depth, q[h].second);
        if (!n.leaf) for (int k = 0; k < 8; ++k) q.push_back({n.child[k], q[h].second + 1});
        else {
            s->refs += (int64_t)n.tris.size();
            s->max_leaf = std::max<int32_t>(s->max_leaf, (int32_t)n.tris.size());
            s->leaves++;
            if (n.tris.empty()) s->empty_leaves++;
        }
    }
    s->avg_leaf = s->refs / (float)s->leaves;
    return 0;
}
int crt_octree_get_node(const crt_octree* oct, int i, float* bounds6, int32_t* leaf, int32_t* child8, int32_t* pairs, int32_t cap, int32_t* n_pairs) {
    if (i < 0 || i >= (int)oct->nodes.size()) { set_error("octree_get_node: index out of range"); return 1; }
    const HostOctreeNode& n = oct->nodes[i];
    if (bounds6) for (int a = 0; a < 3; ++a) { bounds6[a] = n.bmin[a]; bounds6[3 + a] = n.bmax[a]; }
    if (leaf) *leaf = n.leaf ? 1 : 0;
    if (child8) for (int k = 0; k < 8; ++k) child8[k] = n.leaf ? -1 : n.child[k];
    if (n_pairs) *n_pairs = (int32_t)n.tris.size();
    if (pairs)
        for (size_t j = 0; j < n.tris.size() && (int32_t)j < cap; ++j) {
            uint32_t gid = n.tris[j];
            uint32_t m = (uint32_t)(std::upper_bound(oct->mesh_first.begin(), oct->mesh_first.end(), gid) - oct->mesh_first.begin()) - 1;
            pairs[2 * j] = (int32_t)m;
            pairs[2 * j + 1] = (int32_t)(gid - oct->mesh_first[m]);
        }
    return 0;
}

// CameraBase / PerspectiveCamera / OrthographicCamera (Cameras.h:77-142, :213-245, :248-311)
int crt_camera_matrices(int kind, float near_, float far_, float sw, float sh, float fov, const float* pos, const float* look, const float* /*right*/,
                        const float* worldup, float resx, float resy, float* r2c, float* c2w) {
    float I[16], A[16], B[16], s2n[16], n2r[16], s2r[16], r2s[16];
    m4_identity(I);
    if (kind == 0) {                                   // PerspectiveCamera passes its own sensor size (Cameras.h:255)
        sw = 2 * near_ * tanf(fov * 0.01745329251994329576923690768489f / 2.0f);
        sh = 2 * near_ * tanf(fov * 0.01745329251994329576923690768489f / 2.0f) * (resx / resy);
    }
    m4_scale(I, mk3(1.0f / sw, 1.0f / sh, 1), A);
    m4_translate(I, mk3(sw / 2.0f, sh / 2.0f, 0), B);
    m4_mul(A, B, s2n);
    m4_scale(I, mk3(resx, -resy, 1), A);
    m4_translate(I, mk3(0, -1, 0), B);
    m4_mul(A, B, n2r);
    m4_mul(n2r, s2n, s2r);
    m4_inverse(s2r, r2s);
    // calculateWorldCameraMatrices (:130-142)
    f3 dir = normalize3(mk3(look[0], look[1], look[2]));
    f3 right = normalize3(cross3(mk3(worldup[0], worldup[1], worldup[2]), dir));
    f3 up = cross3(dir, right);
    const float cw[16] = {right.x, right.y, right.z, 0, up.x, up.y, up.z, 0, dir.x, dir.y, dir.z, 0, pos[0], pos[1], pos[2], 1};
    std::memcpy(c2w, cw, sizeof cw);
    float c2s[16], c2s_inv[16];
    if (kind == 0) {                                   // calculuateMatrix (:303-310)
        float invTanAng = 1.0f / tanf(fov * 0.01745329251994329576923690768489f / 2.0f);
        const float persp[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, far_ / (far_ - near_), 1, 0, 0, -(far_ * near_ / (far_ - near_)), 0};
        m4_scale(persp, mk3(invTanAng, invTanAng, 1), c2s);
    } else if (kind == 1) {                            // OrthographicCamera (:222-227)
        m4_scale(I, mk3(1, 1, (float)(1.0 / (double)(far_ - near_))), A);
        m4_translate(I, mk3(0, 0, -near_), B);
        m4_mul(A, B, c2s);
    } else if (kind == 2) {                            // PinholeCamera (:313-359) only uses M_RastertoScreen
        std::memcpy(r2c, r2s, sizeof r2s);
        return 0;
    } else { set_error("camera_matrices: unknown kind"); return 1; }
    m4_inverse(c2s, c2s_inv);
    m4_mul(c2s_inv, r2s, r2c);
    return 0;
}
int crt_shape_matrices(const float* rigid16, float* o2r, float* r2o) { shape_matrices(rigid16, o2r, r2o); return 0; }

}  // extern "C"
