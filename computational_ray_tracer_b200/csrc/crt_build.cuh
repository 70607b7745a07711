// crt_build.cuh -- Octtree_Model::CreateOcttree on the GPU.
//
// The reference inserts triangles one at a time (Octtree_Model.h:33-63,180-358); crt_host.cpp restates that loop and is the
// parity path.  The tree it produces is, however, a function of each node's triangle SET (see build_topdown() in
// crt_host.cpp for the argument and tests/test_cpu_host_parity.py for the proof by comparison), so it can be built level by
// level with every (triangle, child cell) overlap test of a level running in parallel:
//
//   k_build_masks    one thread per (node, triangle) reference: 8-bit mask of the child cells the triangle overlaps
//                    (crt_sat.h: the same Akenine-Moller test, same fp32 operation order as the host)
//   k_build_decide   one warp per node: first reference index at which the node splits (prefix-AND of the masks reaches 0,
//                    the count is >= 40 and the triangle arrived after the node was created), then the child counts
//   exclusive scans  (thrust) over split flags and child counts -> next level's node numbers and reference offsets
//   k_build_scatter  one warp per splitting node: stable distribution of its references to the 8 children (global-id order
//                    is preserved, which is the per-leaf order the reference's insertion loop produces)
//
// Each level's nodes and references are copied back and appended to the host-side crt_octree (breadth-first numbering), so
// everything downstream (flatten, packets, subtree bounds, stats, GetNode) is shared with the host builders.
#pragma once
#include <thrust/execution_policy.h>
#include <thrust/scan.h>

#include "crt_sat.h"

namespace crt {

struct BuildNode {
    float lo[3], hi[3];
    long long tau;           // global id of the triangle whose insertion created this node (-1 for the root)
    unsigned start, count;   // range in the level's reference array
};

#define CRT_BUILD_CAPACITY 40     // Octtree_Model::TRIANGLE_CAPACITY (Octtree_Model.h:388)

__global__ void k_build_root_filter(const float* world_pos, unsigned n_tris, BuildNode root, unsigned* flags) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_tris) return;
    const f3* t = reinterpret_cast<const f3*>(world_pos + 9 * (size_t)i);
    f3 tri[3] = {t[0], t[1], t[2]};
    flags[i] = tri_in_bounds(tri, root.lo, root.hi) ? 1u : 0u;
}
__global__ void k_build_root_compact(const unsigned* flags, const unsigned* offsets, unsigned n_tris, unsigned* refs, unsigned* ref_node) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_tris || !flags[i]) return;
    refs[offsets[i]] = i;
    ref_node[offsets[i]] = 0;
}

__global__ void __launch_bounds__(256) k_build_masks(const BuildNode* nodes, const unsigned* refs, const unsigned* ref_node, unsigned n_refs,
                                                       const float* world_pos, unsigned char* masks) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_refs) return;
    const BuildNode nd = nodes[ref_node[i]];
    if (nd.count < CRT_BUILD_CAPACITY) { masks[i] = 0; return; }
    const f3* t = reinterpret_cast<const f3*>(world_pos + 9 * (size_t)refs[i]);
    f3 tri[3] = {t[0], t[1], t[2]};
    masks[i] = (unsigned char)child_overlap_mask(nd.lo, nd.hi, tri);
}

// one warp per node: split decision + child counts
__global__ void __launch_bounds__(256) k_build_decide(const BuildNode* nodes, unsigned n_nodes, const unsigned* refs, const unsigned char* masks,
                                                        unsigned* split_flag, long long* split_gid, unsigned* child_cnt) {
    const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n_nodes) return;
    const BuildNode nd = nodes[warp];
    if (lane < 8) child_cnt[8 * (size_t)warp + lane] = 0;
    if (lane == 0) { split_flag[warp] = 0; split_gid[warp] = -1; }
    if (nd.count < CRT_BUILD_CAPACITY) return;
    // i0: first index where the AND of masks[0..i] is 0
    unsigned running = 0xffu;
    long long i0 = -1;
    for (unsigned base = 0; base < nd.count && i0 < 0; base += 32) {
        const unsigned i = base + lane;
        unsigned m = i < nd.count ? masks[nd.start + i] : 0xffu;
        // inclusive prefix AND across the warp
        for (int o = 1; o < 32; o <<= 1) { unsigned v = __shfl_up_sync(0xffffffffu, m, o); if ((int)lane >= o) m &= v; }
        m &= running;
        const unsigned z = __ballot_sync(0xffffffffu, m == 0 && i < nd.count);
        if (z) i0 = (long long)base + (__ffs(z) - 1);
        running = __shfl_sync(0xffffffffu, m, 31);
    }
    if (i0 < 0) return;                                                // some child overlaps everything: (fat) leaf
    // i_tau: first index whose triangle arrived after the node was created (references are in ascending global id)
    unsigned lo = 0, hi = nd.count;
    while (lo < hi) { unsigned mid = (lo + hi) >> 1; if ((long long)refs[nd.start + mid] > nd.tau) hi = mid; else lo = mid + 1; }
    long long at = i0;
    if (at < CRT_BUILD_CAPACITY - 1) at = CRT_BUILD_CAPACITY - 1;
    if (at < (long long)lo) at = (long long)lo;
    if (at >= (long long)nd.count) return;                             // no insertion after creation ever attempted the split
    // child counts
    unsigned cnt = 0;                                                   // lane k < 8 accumulates child k
    for (unsigned base = 0; base < nd.count; base += 32) {
        const unsigned i = base + lane;
        const unsigned m = i < nd.count ? masks[nd.start + i] : 0u;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const unsigned b = __ballot_sync(0xffffffffu, (m >> k) & 1u);
            if ((int)lane == k) cnt += __popc(b);
        }
    }
    if (lane < 8) child_cnt[8 * (size_t)warp + lane] = cnt;
    if (lane == 0) { split_flag[warp] = 1; split_gid[warp] = (long long)refs[nd.start + (unsigned)at]; }
}

// one warp per splitting node: children records + stable scatter of the references
__global__ void __launch_bounds__(256) k_build_scatter(const BuildNode* nodes, unsigned n_nodes, const unsigned* refs, const unsigned char* masks,
                                                         const unsigned* split_flag, const unsigned* split_rank, const long long* split_gid,
                                                         const unsigned* child_cnt, const unsigned* child_start,
                                                         BuildNode* next_nodes, unsigned* next_refs, unsigned* next_ref_node) {
    const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n_nodes || !split_flag[warp]) return;
    const BuildNode nd = nodes[warp];
    const unsigned first_child = 8 * split_rank[warp];
    if (lane < 8) {
        BuildNode c;
        child_cell(nd.lo, nd.hi, (int)lane, c.lo, c.hi);
        c.tau = split_gid[warp];
        c.start = child_start[8 * (size_t)warp + lane];
        c.count = child_cnt[8 * (size_t)warp + lane];
        next_nodes[first_child + lane] = c;
    }
    unsigned run = lane < 8 ? child_start[8 * (size_t)warp + lane] : 0;   // lane k: next free slot of child k
    for (unsigned base = 0; base < nd.count; base += 32) {
        const unsigned i = base + lane;
        const unsigned m = i < nd.count ? masks[nd.start + i] : 0u;
        const unsigned gid = i < nd.count ? refs[nd.start + i] : 0u;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const unsigned b = __ballot_sync(0xffffffffu, (m >> k) & 1u);
            const unsigned dst0 = __shfl_sync(0xffffffffu, run, k);
            if ((m >> k) & 1u) {
                const unsigned dst = dst0 + __popc(b & ((1u << lane) - 1u));
                next_refs[dst] = gid;
                next_ref_node[dst] = first_child + k;
            }
            if ((int)lane == k) run += __popc(b);
        }
    }
}

}  // namespace crt
