// crt_host.h -- host-side components of libcrt_b200: octree builder/flattener, camera and colour
// constants, spectrum tables.  Pure C++ (no CUDA), compiled with -ffp-contract=off.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/crt_b200.h"
#include "crt_math.h"

namespace crt {

void set_error(const std::string& msg);

// ---- device-facing flat layouts (DESIGN.md "data layout in HBM") -----------------------------------
// Node: 32 bytes = 2 x float4.  lo = (pmin.xyz, a), hi = (pmax.xyz, b).
//   internal node: a = index of the first of its 8 children (children are contiguous, BFS numbering),
//                  b = 8-bit mask of the children that are not empty (bit 31 clear)
//   leaf:          a = offset of its reference list in leaf_refs, b = 0x80000000 | reference count
//                  | 0x20000000 (sub-packets) or 0x40000000 (fat leaf: super-packets over its sub-packets), see below
// Triangle: 48 bytes = 3 x float4 = (p0.xyz, material), (p1.xyz, mesh_id), (p2.xyz, tri_id), world space.
//
// Triangle packets (ours, not the reference's).  The octree's cells are much larger than the surface patch inside them (a cell is
// kept whenever a triangle touches it anywhere, leaves hold up to 40 triangles, and the reference's split rule leaves "fat" leaves
// of up to thousands: Octtree_Model.h:332-340 aborts a split when one child would receive everything).  For the ordered traversal
// every non-empty leaf also gets its triangles regrouped (recursive median split of the centroids) into sub-packets of
// <= CRT_SUBPACKET triangles with a padded bounding box each (b |= CRT_SUBPK_FLAG); a fat leaf -- more than CRT_SUPERPACKET
// sub-packets -- additionally gets super-packets, the boxes of runs of CRT_SUPERPACKET consecutive sub-packets (b |= CRT_PACKET_FLAG).
// Packets only let that traversal skip triangles whose box the ray misses; the leaf's reference-order list (what the exact BFS
// kernel walks) is unchanged.  A non-empty leaf's list in leaf_refs is preceded by two header words: index of its first box (sub-packet,
// or super-packet for a fat leaf) in pk_boxes, and the number of such boxes.
#ifndef CRT_SUBPACKET
#define CRT_SUBPACKET 4
#endif
#define CRT_SUPERPACKET 32
#define CRT_PACKET_FLAG 0x40000000u
#define CRT_SUBPK_FLAG 0x20000000u
struct FlatOctree {
    std::vector<float> nodes;         // 8 floats per node (bit patterns for a/b)
    std::vector<uint32_t> leaf_refs;  // global triangle ids
    std::vector<float> node_tight;    // 8 floats per node (BFS order): padded bounding box of every triangle stored beneath
                                      // the node (min.xyz, -, max.xyz, -); inverted (never hit) for empty subtrees
    std::vector<float> pk_boxes;      // 8 floats per box: sub-packet (pmin.xyz, first index into pk_refs), (pmax.xyz, triangle count);
                                      // super-packet (pmin.xyz, first sub-packet box), (pmax.xyz, sub-packet count)
    std::vector<uint32_t> pk_refs;    // global triangle ids, Morton order within each leaf
    std::vector<int32_t> bfs_of_ref;  // reference-order node id -> BFS id (for tests)
    int depth = 0;
};

struct HostOctreeNode {
    float bmin[3], bmax[3];
    bool leaf = true;
    int parent = -1;
    int child[8] = {-1, -1, -1, -1, -1, -1, -1, -1};
    std::vector<uint32_t> tris;       // global triangle ids, insertion order
    // memo of aborted splits (ours): children that contain ALL of tris[0..memo_count).  The reference re-bins the whole
    // leaf on every insertion into a leaf that could not be split (O(size) each, Octtree_Model.h:216-223,332-340);
    // the outcome only depends on whether some child still contains everything, which is updated per new triangle.
    bool memo = false;
    uint8_t full_mask = 0;
    uint32_t memo_count = 0;
};

}  // namespace crt

// Octtree_Model (RayTracer/Octtree_Model.h)
struct crt_octree {
    std::vector<crt::HostOctreeNode> nodes;     // reference (creation) order
    std::vector<uint32_t> mesh_first;           // global id of each mesh's triangle 0, size n_meshes+1
    std::vector<crt::f3> world_pos;             // 3 per global triangle (octree-space vertices)
    void add_triangle(uint32_t gid);
    void build_topdown();                       // same tree as inserting gid 0..n-1 with add_triangle (see crt_host.cpp)
    void split(int id);
    void flatten(const std::vector<uint8_t>& skip, crt::FlatOctree* out) const;
    uint32_t build_packets(const std::vector<uint32_t>& tris, uint32_t packet_size, crt::FlatOctree* out) const;   // returns the packet count
};

namespace crt {

// Shared first half of every builder: global triangle numbering, world-space vertices (Octtree_Model.h:188-197) and the
// root node with the model's bounds.  Returns nullptr (error set) on bad input.
crt_octree* octree_prepare(const crt_mesh_desc* meshes, uint32_t n_meshes, const float* o2r, int precomputed_world);

// spectral tables restated from the data (see crt_spectra.cpp)
struct PiecewiseLinear {
    std::vector<float> lambdas, values;
    float query(float lambda) const;
    static PiecewiseLinear from_interleaved(const float* samples, int count, bool normalize);
};
struct HostSpectra {
    float X[471], Y[471], Z[471];     // CIE 1931 matching curves on 360..830 nm (DenselySampledSpectrum)
    float D65dense[471];              // sRGB colour space illuminant (DenselySampledSpectrum of normalised D65)
    PiecewiseLinear illum[6];         // normalised A, D50, D65, F1, F2, F11
    float XYZFromSensorRGB[9], RGBFromXYZ[9], XYZFromRGB[9], white[2];
};
const HostSpectra& host_spectra();
// Measured PixelSensor (pixelsensor.h:37-68): XYZFromSensorRGB (column-major 9) from response curves and a sensor illuminant, all
// sampled at the 471 integer wavelengths, by least squares over the 24 Macbeth swatches
void measured_sensor_matrix(const float* r471, const float* g471, const float* b471, const float* illum471, float* out9);
const float* named_table(const char* name, int* n);
const float* swatch_table(int i, int* n);
int named_table_count();

// small glm-order matrix helpers (column-major float[16])
void m4_identity(float* m);
void m4_mul(const float* a, const float* b, float* out);
void m4_inverse(const float* m, float* out);
void m3_inverse(const float* m, float* out);
void m3_mul(const float* a, const float* b, float* out);
void shape_matrices(const float* rigid16, float* o2r, float* r2o);
// transpose(inverse(mat3(M))) as used by LocalSurfaceInfo::Transform (Shapes.h:147-152)
void normal_matrix(const float* m16, float* out9);

}  // namespace crt
