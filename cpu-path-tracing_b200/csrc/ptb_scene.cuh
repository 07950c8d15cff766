// ptb_scene.cuh -- device-side scene layout.
//
// Input records are the reference's own AoS FP64 structs (pt::sphere 88 B,
// /root/reference/src/sphere.hpp:10-17; pt::camera 176 B, src/camera.hpp:23-32).
// The FP64 parity kernels read them as they are.  The FP32 throughput kernels read
// a host-packed SoA form built once per ptb_upload_scene (ptb_api.cu: pack_scene):
//
//   * all positions are translated by -origin_shift (the camera look region) so
//     FP32 coordinates stay O(scene extent);
//   * spheres are split into two geometry classes, each tested with the form that
//     is numerically stable for it in binary32 ("hard parts"):
//       small : (c, r^2)                       classic oc-form of sphere.cpp:8-12
//       big   : (g = -k c, k = 1/2R, K = k(c.c - R^2), 2R)
//               the same quadratic divided by 2R and expanded about the shifted
//               origin, coefficients computed in FP64 on the host; near root by
//               the cancellation-free form c'/(s - hb').  Needed for the R = 1e6
//               "wall" spheres of box_scene.hpp:16-47, where oc.oc - r^2 has no
//               correct digits in binary32.
//   * inside each class the spheres that can only ever be hit at their NEAR root come
//     first: an opaque sphere (diffuse / specular) that the camera lens lies outside of
//     is never left from its inside, so the far-root branch of sphere.cpp:21-27 is dead
//     for it; dielectric spheres and spheres containing the
//     camera keep both roots;
//   * "list position" = index into the concatenation [small near-only, small both,
//     big near-only, big both].  Shading data is stored per LIST POSITION (four float4
//     planes), `order[]` maps a list position back to the caller's sphere index.
#pragma once

#include "ptb_types.h"

namespace ptb {

constexpr int kMaxConstSpheres = 64; // geometry lists that live in __constant__ memory
constexpr int kSmemShadeSpheres = 64; // shading planes staged in shared memory
constexpr int kBvhStack = 64;         // traversal stack entries (ptb_path_f32.cuh); deeper trees fall back to the scan

struct RawSphere // == pt::sphere, 88 bytes
{
    double radius;
    double px, py, pz;
    double er, eg, eb;
    double cr, cg, cb;
    int32_t reflection;
    int32_t pad_;
};
static_assert(sizeof(RawSphere) == 88, "pt::sphere layout");

struct RawCamera // == pt::camera, 176 bytes
{
    double pos[3];
    double llc[3];
    double ax[3];
    double ay[3];
    double u[3], v[3], w[3];
    double lens_radius;
};
static_assert(sizeof(RawCamera) == 176, "pt::camera layout");

struct SmallGeo
{
    float cx, cy, cz, r2;
};

struct BigGeo
{
    float gx, gy, gz, k; // g = -k*c, k = 1/(2R)
    float K, two_r, pad0, pad1; // K = k*(c.c - R^2), two_r = 2R
};

// FP32 camera, positions relative to origin_shift
struct CameraF32
{
    float px, py, pz;    // position
    float rx, ry, rz;    // lower_left_corner - position
    float ax, ay, az;    // cam_x_axis
    float bx, by, bz;    // cam_y_axis
    float lens_radius;
    float inv_w, inv_h;  // 1/width, 1/height
    float sub_len;       // 1/num_subpixels
};

// camera of the stand-alone smallpt fork (sandbox/main.cpp:235-237), shifted FP32 frame
struct SmallptCamF32
{
    float ox, oy, oz; // camera position
    float dx, dy, dz; // unit viewing direction
    float cxx;        // cx = (w * .5135 / h, 0, 0)
    float cyx, cyy, cyz;
    float push;       // 140
    float inv_w, inv_h;
    float pad_[3];
};

// Both cameras as the FP32 kernels take them: a KERNEL ARGUMENT (RenderParamsF32 / ProbeParams), so that nothing a
// render reads except the precompiled kernels' sphere coefficients lives in a per-device __constant__ symbol -- the
// run-time compiled kernels (coefficients as literals) read no shared symbol at all, and contexts on one GPU that use
// them do not have to take turns.
struct CameraPair
{
    CameraF32 cam;
    SmallptCamF32 sbcam;
};

// Everything the FP32 kernels read through the constant cache.
struct ConstSceneF32
{
    int n_small_near; // small spheres tested at the near root only
    int n_small;      // all small spheres (near-only first)
    int n_big_near;
    int n_big;
    int n_total;
    int pad_[3];
    SmallGeo small_geo[kMaxConstSpheres];
    BigGeo big_geo[kMaxConstSpheres];
    // The two coefficients an axis sphere's test reads (its one non-zero g component, K), packed back to back in
    // list order so that the unrolled scan fetches them with 16-byte uniform loads instead of one load each
    // (ptb_path_f32.cuh: key_big_axis / key_big_pair)
    alignas(16) float axis_coef[2 * kMaxConstSpheres];
    int order[2 * kMaxConstSpheres]; // list position -> original sphere index
};

// Shading planes, indexed by LIST POSITION (global memory; staged to shared memory when
// n <= kSmemShadeSpheres)
//   a = (-c/R xyz, 1/R)            outward normal = P * a.w + a.xyz
//   b = (emission rgb, reflection as int bits)
//   c = (color rgb, p = max(color))
//   d = (color/p rgb, 0)           Russian-roulette survivor weight, src/main.cpp:131-132
struct ShadePlanes
{
    float4 const* a;
    float4 const* b;
    float4 const* c;
    float4 const* d;
};

// Geometry lists for scenes too large for constant memory (same records, global memory)
struct GeoLists
{
    SmallGeo const* small_geo;
    BigGeo const* big_geo;
    int const* order; // list position -> original sphere index
    // Bounding-volume hierarchy over the small spheres (ptb_bvh.hpp), nullptr = scan the lists.
    float4 const* bvh_nodes;   // 4 float4 per node
    SmallGeo const* bvh_geo;   // spheres in leaf order
    int const* bvh_pos;        // leaf slot -> list position, bit 31 set = both roots count (sphere.cpp:21-27)
    int bvh_root;              // child code of the root
};

} // namespace ptb
