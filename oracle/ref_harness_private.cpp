// ORACLE SUPPORT (test infrastructure).  A second translation unit of oracle/_ref/libcrt_ref.so that reads two PRIVATE members of the
// reference's classes -- TriModel::back_facing (Shapes.h:1490) and Octtree_Model::octtree (Octtree_Model.h:415) -- which no public
// accessor exposes.  The reference sources are still compiled unmodified; this file only defines `private` as `public` AFTER every
// standard header has been read (oracle/refshim/ref_prelude.h is force-included first), which changes access checking and nothing
// else: class layouts are identical to those in ref_harness.cpp (same member order, no virtual changes), so the objects created there
// can be read here.
// AABB_triangle_Moller.h defines its functions non-inline: give this unit's copies another namespace name so the two units link
#define Moller Moller_private_unit
#define private public
#define protected public
#include "RayTracer/Sampling.h"
#include "RayTracer/Octtree_Model.h"
#undef private
#undef protected

#include <cstdint>

extern "C" {

// TriModel::back_facing[mesh] after ComputeBackFace (Shapes.h:1339-1380): one byte per triangle; returns the count (0 = table empty)
int ref_private_backfacing(const void* tri_model, int mesh, unsigned char* out) {
    const TriModel* m = (const TriModel*)tri_model;
    if (m->back_facing.empty()) return 0;
    const std::vector<bool>& v = m->back_facing[mesh];
    for (size_t i = 0; i < v.size(); ++i) out[i] = v[i] ? 1 : 0;
    return (int)v.size();
}
// number of leaf nodes / empty leaves / deepest level of the octree, walked from the private node array (what PrintInfo prints, :134-176)
void ref_private_octree_stats(const void* oct_model, int* nodes, int* leaves, int* empty_leaves, int* max_leaf) {
    const Octtree_Model* o = (const Octtree_Model*)oct_model;
    *nodes = (int)o->octtree.size(); *leaves = *empty_leaves = *max_leaf = 0;
    for (const auto& n : o->octtree) {
        if (!n.leaf) continue;
        ++*leaves;
        if (n.triangle_info.empty()) ++*empty_leaves;
        if ((int)n.triangle_info.size() > *max_leaf) *max_leaf = (int)n.triangle_info.size();
    }
}

}  // extern "C"
