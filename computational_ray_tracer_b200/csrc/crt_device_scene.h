// crt_device_scene.h -- POD views of the device-resident scene, passed to kernels by value.
#pragma once
#include <cuda_runtime.h>

#include "crt_math.h"
#include "crt_sampling.h"

namespace crt {

#define CRT_LEAF_FLAG 0x80000000u
#define CRT_LEAF_PACKETS 0x40000000u     // leaf also has Morton-ordered triangle packets (crt_host.h)
#define CRT_LEAF_SUBPK 0x20000000u       // ordinary leaf: list preceded by (first sub-packet, sub-packet count), crt_host.h
#define CRT_LEAF_COUNT_MASK 0x1fffffffu
#define CRT_NLAMBDA 8            // NSpectrumSamples, ThirdParty/pbrv4/spectrum.h:19

// Spectrum record (device): kind + parameters, data in one float pool.
enum SpectrumKind { SPEC_CONSTANT = 0, SPEC_PIECEWISE = 1, SPEC_DENSE = 2, SPEC_SIGMOID = 3, SPEC_SIGMOID_ILLUM = 4, SPEC_SIGMOID_UNBOUNDED = 5 };
struct DevSpectrum {
    int kind;
    int offset;      // PIECEWISE: lambdas at pool[offset..offset+n), values at pool[offset+n..offset+2n); DENSE: 471 values
    int n;
    float c0, c1, c2, scale;   // CONSTANT: c0; SIGMOID*: polynomial + scale
    // PIECEWISE (ours): pool[offset+2n .. +lut_n) holds, per whole nanometre from lut_base on, the index of the last knot at or below it
    // (int bits), so FindInterval's binary search becomes one lookup and a step or two forward -- the same interval, the same lerp
    int lut_base, lut_n;
};
enum MaterialKind { MAT_LAMBERT = 0, MAT_DIELECTRIC = 1, MAT_CONDUCTOR = 2 };
struct DevMaterial {
    int type, refl, eta, k, emit;
    float emit_scale;
    int two_sided, eta_constant;
};
enum ShapeKind { SHAPE_SPHERE = 0, SHAPE_CYLINDER = 1, SHAPE_DISK = 2, SHAPE_TRIANGLE_SIMPLE = 3 };
struct DevShape {
    int kind, material;
    float o2r[16], r2o[16];    // ObjectToRender / RenderToObject (Shapes.h:175-182)
    float nmat[9];             // transpose(inverse(ObjectToRender)) upper 3x3 (Shapes.h:150)
    float p[12];               // sphere: r,zmin,zmax,thetamin,thetamax,phimax ; cylinder: r,zmin,zmax,phimax ;
                               // disk: h,inner,outer,phimax ; trianglesimple: p1,p2,p3
};
// Padded world-space bounds of a shape (ours): the integrator's loops over the shape list test this box first and run the
// reference's intersection routine only for shapes the ray can reach (a shape whose box is missed cannot report a hit).
struct DevShapeBox { float4 lo, hi; };
struct DevLight {              // emissive triangle
    float p0[3], p1[3], p2[3], n[3];
    float area;
    int material;
};

// Lights.h:5-8 (intent notes; defined by oracle_render.h DeltaLight): kind 0 point light at v, radiant intensity scale * spectrum (1 / r^2
// falloff); kind 1 sun, v = unit direction towards the light, irradiance scale * spectrum
struct DevDeltaLight { int kind; float v[3]; int spectrum; float scale; };

struct DeviceScene {
    // triangle model + octree
    const float4* nodes;       // 2 per node
    const uint32_t* leaf_refs;
    const float4* node_tight;  // 2 per node: padded bounds of the triangles beneath it (ordered traversal only)
    const float4* pk_boxes;    // 2 per packet
    const uint32_t* pk_refs;
    const float4* tris;        // 3 per triangle: (p0,mat) (p1,mesh) (p2,tri)
    const float4* tri_nrm;     // 3 per triangle (vertex normals) or nullptr
    const float4* tri_uv;      // 2 per triangle (u0 v0 u1 v1 | u2 v2 - -) or nullptr: MeshCache::Mesh::texcoords (AssetManager.h:20-47)
    const float4* tri_tan;     // 3 per triangle (vertex tangents) or nullptr
    const float4* tri_bitan;   // 3 per triangle (vertex bitangents) or nullptr
    int n_nodes, n_tris;
    int has_model;
    int root_leaf;             // set per launch by the path integrator: the octree is one leaf of a few triangles and the shading kernels traverse it
                               // themselves (trace_root_leaf) instead of reading the records of a traversal launch
    int retransform_surface;   // Triangle::CalculateLocalSurface applies ObjectToRender even to precomputed world positions
    float model_o2r[16];
    // analytic shapes
    const DevShape* shapes;
    const DevShapeBox* shape_boxes;
    int n_shapes;
    // threaded bounding-volume hierarchy over the shape boxes (ours): 2 x float4 per node in depth-first order,
    // (lo.xyz, bits(index of the node to continue at when this box is missed)), (hi.xyz, bits(shape id) or -1 for an inner node)
    const float4* shape_bvh;
    int n_shape_nodes;
    // shading data
    const DevMaterial* materials;
    const DevSpectrum* spectra;
    const float* pool;
    int n_materials, n_spectra;
    const DevLight* lights;
    const float* light_cdf;
    int n_lights;
    float light_total;
    const DevDeltaLight* delta_lights;
    int n_delta;
    // global tables
    const float* cieX; const float* cieY; const float* cieZ; const float* d65dense;   // 471 each; cieX/Y/Z = the film sensor's r_bar/g_bar/b_bar
                                                                                      // (the CIE observer for the default XYZ sensor)
    float imaging_ratio;                                                               // PixelSensor::imagingRatio
    const float* f1_lambdas; const float* f1_values; int f1_n;                         // normalised illuminant F1
};

struct DevCamera {
    float r2c[16], c2w[16];
    float lens_radius, focal_distance;
    int kind;
};

}  // namespace crt
