// Device-side records and small math helpers shared by every kernel of the backend.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rayrs_b200.h"

// Bounds checks of our own (compute-sanitizer is closed on this GPU pool): built with -DRRS_DEBUG_CHECKS every
// stack push, queue slot, node and primitive index is asserted on the device; the GPU test suite is run once
// against that build (scripts/gpu_cmd_debugchecks.sh).  Compiled out otherwise.
#ifdef RRS_DEBUG_CHECKS
#include <cassert>
#define RRS_CHECK(cond) assert(cond)
#else
#define RRS_CHECK(cond) ((void)0)
#endif

namespace rrs {

// ---------------------------------------------------------------------------------------
// HBM layout
// ---------------------------------------------------------------------------------------
// Primitive record, 64 B = two 32-byte halves, each fetched by ONE 256-bit load (a 48-byte record read
// as 3 x 128 bit cost three L1 wavefronts per lane; the L1 data pipe is the busiest unit of the
// traversal, profiles/r01d_c4_batched_metrics.csv).
//   meta = type | material << 2          (a.w)
//   triangle : a = (p1, meta)  b = (p2, obj_id)  c = (p3, emission)
//   sphere   : a = (centre, meta) b = (r^2, index into sphere64, -, obj_id) c = (-, -, -, emission)
//   plane    : a = (pos, umin, umax, meta) b = (vmin, vmax, axis, obj_id) c = (-, -, -, emission)
struct __align__(32) DPrim {
    float4 a, b, c, pad;
};
static_assert(sizeof(DPrim) == 64, "DPrim must be 64 bytes");

// Material record, 48 B.
//   m0 = (color rgb, tag)
//   m1 = (spec / r0 rgb, alpha^2)
//   m2 = (ior, fresnel kind, r0 of the dielectric ((1-ior)/(1+ior))^2, -)
struct DMat {
    float4 m0, m1, m2;
};

// 32-byte node: the whole binary node — both children's boxes and references — in ONE 256-bit load
// (LDG.E.ENL2.256).  Boxes are fp16, rounded outward from the fp32 boxes of RrsNode (lo down, hi up;
// out-of-range values become +-inf), so the test stays conservative: a looser box costs a few extra
// node visits near the leaves, never a wrong hit.  Each axis is one half2 word (lo, hi), which lets the
// traversal pick near/far planes with a single PRMT per axis instead of two selects.
//   w[0..2] = child 0 (lo.x,hi.x) (lo.y,hi.y) (lo.z,hi.z)   w[3..5] = child 1   w[6] = ref0   w[7] = ref1
struct __align__(32) DNode16 {
    uint32_t w[8];
};
static_assert(sizeof(DNode16) == 32, "DNode16 must be 32 bytes");

#define RRS_NO_PRIM 0xFFFFFFFFu
// origin word of a ray: primitive index (28 bits) | RRS_ORG64 when the ray also carries an f64 origin
#define RRS_ORG64 0x80000000u
#define RRS_PRIM_MASK 0x0FFFFFFFu

// ---------------------------------------------------------------------------------------
// float3 helpers (no operator overloading on CUDA's builtin float3 to keep call sites explicit
// about rounding order)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
__device__ __forceinline__ float3 add3(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 sub3(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 mul3(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ float3 scale3(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float dot3(float3 a, float3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
__device__ __forceinline__ float3 cross3(float3 a, float3 b) {
    return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float3 normalize3(float3 a) { return scale3(a, rsqrtf(dot3(a, a))); }
__device__ __forceinline__ float3 neg3(float3 a) { return f3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float3 madd3(float3 a, float s, float3 b) {  // a*s + b
    return f3(fmaf(a.x, s, b.x), fmaf(a.y, s, b.y), fmaf(a.z, s, b.z));
}
__device__ __forceinline__ float3 xyz(float4 v) { return f3(v.x, v.y, v.z); }

// ---------------------------------------------------------------------------------------
// Counter-based RNG: Philox4x32-7 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11).  counter =
// (pixel, sample, slot, 0), key = seed.  slot 0 = camera jitter, slot b+1 = bounce b (words 0..2 material draws in the
// reference's call order, word 3 = Russian roulette).  Keyed by the GLOBAL sample index, so the set of paths is
// independent of how samples are split across GPUs.
// Seven rounds is the smallest Philox4x32 the paper reports as passing BigCrush (10 is its default with a safety margin);
// the reference's own generator (ChaCha20 behind rand::random) is unpinned by any reference test, so the choice is
// ours, and the RNG is a quarter of the thread instructions of the sphere-series kernels.  Both round counts are
// checked against Random123's published known-answer vectors (tests/test_oracle_kat.py, tests/test_gpu_shading.py).
// ---------------------------------------------------------------------------------------
#ifndef RRS_PHILOX_ROUNDS
#define RRS_PHILOX_ROUNDS 7
#endif
__host__ __device__ __forceinline__ void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                    uint32_t k0, uint32_t k1, uint32_t out[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < RRS_PHILOX_ROUNDS; ++r) {
#ifdef __CUDA_ARCH__
        uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
#else
        uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// uniform in [0,1) with 24 bits: the same value the oracle uses in its sample-matched mode
__host__ __device__ __forceinline__ float u01(uint32_t w) { return (float)(w >> 8) * (1.0f / 16777216.0f); }

__device__ __forceinline__ float4 rng_uniforms(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t slot) {
    uint32_t w[4];
    philox4x32(pixel, sample, slot, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), w);
    return make_float4(u01(w[0]), u01(w[1]), u01(w[2]), u01(w[3]));
}

// ---------------------------------------------------------------------------------------
// Scene as the kernels see it
// ---------------------------------------------------------------------------------------
#define RRS_BRUTE_MAX 8

struct DScene {
    const DPrim* prims;
    const DNode16* nodes;
    const DMat* mats;
    const float4* emis;      // (strength*color, -)
    const float4* hdri;      // RGBA f32 texels
    uint32_t n_prims, n_nodes, n_mats;
    uint32_t hdri_w, hdri_h;
    float tmin, tmax;
    double tmin64, tmax64;
    // exact f64 (centre xyz, r^2) of every sphere; non-null only when the scene holds a sphere with a
    // transmissive material — see "sphere re-entry" in intersect.cuh
    const double4* sphere64;
    // f64 vertices (9 doubles per primitive index) when some triangle vertex is not exactly representable in fp32,
    // else nullptr — see triangle_t64 in intersect.cuh
    const double* tri64;
    uint32_t stack_entries;  // per-thread traversal stack size (entries, including the sentinel)
    uint32_t has_triangles;  // 0: no triangle in the scene (the per-leaf shear setup is skipped)
    uint32_t root;           // node the traversal starts at (the virtual root's only child when that is an inner node)
    uint32_t brute_count;    // > 0: that many reachable primitives in total -> no BVH, test them all (intersect.cuh)
    uint32_t brute_prim[8];  // their indices: spheres, then planes, then triangles, DFS order inside a group
    uint32_t brute_spheres, brute_planes;  // group sizes (triangles: the rest)
    uint32_t refill_lanes;   // extend refills a warp with new rays once this many lanes are idle
    uint32_t brute_box_on;   // the sphere group of the brute-force list has a bounding box worth testing first
    float brute_box[6];      // lo xyz, hi xyz of that group, padded outward
};

struct DCamera {
    float3 origin, e_x, e_y, z;
    float inv_ppc, half_w, half_h;
    uint32_t W, H;
};

}  // namespace rrs
