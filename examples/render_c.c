/* A compiled host over the C ABI, no Python and no C++ host mirror in between: what the Rust shim of INTEGRATION.md
 * does, written in C99.  It flattens `diffuse_single_sphere` (rayrs-lib/src/test_scenes.rs:14-44,60-63) by hand —
 * two primitives in DFS order, the virtual root node, two materials, the derived camera fields of Camera::new
 * (lib.rs:113-132) — hands it to librayrs_b200.so and writes the mean-radiance image as raw float32 RGB.
 *
 *   gcc -std=c99 -O2 -Iinclude examples/render_c.c -Lrayrs_b200 -lrayrs_b200 -lm -Wl,-rpath,$PWD/rayrs_b200 -o render_c
 *   ./render_c hdri.f32 HDRI_W HDRI_H  W H SPP  out.f32
 *
 * hdri.f32: HDRI_H x HDRI_W x 3 float32 (already clipped, rayrs/src/main.rs:43).  tests/test_c_example.py builds it,
 * runs it on the GPU and compares the image with the one the Python face renders from the host mirror's flattening.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rayrs_b200.h"

typedef struct { double x, y, z; } V3;
static V3 v3(double x, double y, double z) { V3 v = {x, y, z}; return v; }
static V3 sub(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
static V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
/* vecmath.rs:525-527,690-698: unit(v) = v * (1.0 / mag) — reciprocal, then multiply */
static V3 unit(V3 a) { double s = 1.0 / sqrt(a.x * a.x + a.y * a.y + a.z * a.z); return v3(a.x * s, a.y * s, a.z * s); }
static void put3(double* d, V3 v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; }

/* Camera::new, lib.rs:113-132 (the z axis is scaled by width / tan(fov/2): the reference's FOV quirk stays on the host) */
static RrsCamera camera_new(V3 origin, V3 up, V3 lookat, double fov, double width, double height, unsigned ppi) {
    RrsCamera c;
    memset(&c, 0, sizeof c);
    V3 z = unit(sub(lookat, origin));
    V3 x = unit(cross(up, z));
    V3 y = unit(cross(z, x));
    const double pi = 3.14159265358979323846264338327950288;
    double s = width / tan(fov * (pi / 180.0) / 2.0);
    put3(c.origin, origin);
    put3(c.e_x, x);
    put3(c.e_y, y);
    put3(c.z_scaled, v3(s * z.x, s * z.y, s * z.z));
    c.width = width;
    c.height = height;
    c.ppc = (uint32_t)floor((double)ppi * 2.54 + 0.5);
    c.x_pixels = (uint32_t)floor(width * (double)c.ppc + 0.5);
    c.y_pixels = (uint32_t)floor(height * (double)c.ppc + 0.5);
    return c;
}

static int fail(const char* what) {
    fprintf(stderr, "%s: %s\n", what, rrs_last_error());
    return 1;
}

int main(int argc, char** argv) {
    if (argc != 8) {
        fprintf(stderr, "usage: %s hdri.f32 HDRI_W HDRI_H W H SPP out.f32\n", argv[0]);
        return 2;
    }
    const unsigned hw = (unsigned)atoi(argv[2]), hh = (unsigned)atoi(argv[3]);
    const unsigned W = (unsigned)atoi(argv[4]), H = (unsigned)atoi(argv[5]), spp = (unsigned)atoi(argv[6]);
    float* hdri = (float*)malloc(sizeof(float) * 3u * hw * hh);
    FILE* f = fopen(argv[1], "rb");
    if (!f || fread(hdri, sizeof(float), 3u * (size_t)hw * hh, f) != 3u * (size_t)hw * hh) {
        fprintf(stderr, "cannot read %s\n", argv[1]);
        return 2;
    }
    fclose(f);

    /* ---- the scene, flattened as INTEGRATION.md section 2 describes ---- */
    RrsMaterial mats[2];
    memset(mats, 0, sizeof mats);
    mats[0].tag = RRS_MAT_COOK_TORRANCE; /* floor: test_scenes.rs:15-21 */
    mats[0].fresnel_kind = RRS_FRESNEL_METALLIC;
    mats[0].color[0] = mats[0].color[1] = mats[0].color[2] = 1.0;
    mats[0].spec_color[0] = mats[0].spec_color[1] = mats[0].spec_color[2] = 0.8; /* r0 */
    mats[0].alpha = 0.5;
    mats[1].tag = RRS_MAT_LAMBERTIAN; /* sphere: test_scenes.rs:60-63 */
    mats[1].color[0] = mats[1].color[1] = mats[1].color[2] = 0.8;

    RrsPrim prims[2]; /* DFS leaf order of the reference tree: two objects -> one leaf group [floor, sphere] */
    memset(prims, 0, sizeof prims);
    prims[0].type = RRS_PLANE;
    prims[0].obj_id = 0;
    prims[0].material = 0;
    prims[0].emission = -1;
    prims[0].v[0] = (double)RRS_AXIS_Y;
    prims[0].v[1] = -25.0; prims[0].v[2] = 25.0; prims[0].v[3] = -25.0; prims[0].v[4] = 25.0;
    prims[0].v[5] = 0.0;
    prims[1].type = RRS_SPHERE;
    prims[1].obj_id = 1;
    prims[1].material = 1;
    prims[1].emission = -1;
    prims[1].v[0] = 1.0; /* radius^2 */
    prims[1].v[1] = 0.0; prims[1].v[2] = 1.0; prims[1].v[3] = 0.0;

    /* node 0 = virtual root: child 0 is the reference root (a leaf group of 2 with the union box of
     * plane [-25,25] x {0} x [-25,25] and sphere [-1,1] x [0,2] x [-1,1]), child 1 empty (inverted box) */
    RrsNode node;
    RrsNodeF64 node64;
    memset(&node, 0, sizeof node);
    memset(&node64, 0, sizeof node64);
    const double lo[3] = {-25.0, 0.0, -25.0}, hi[3] = {25.0, 2.0, 25.0};
    for (int k = 0; k < 3; ++k) {
        node.lo0[k] = (float)lo[k]; node.hi0[k] = (float)hi[k]; /* exactly representable: no outward rounding needed */
        node64.lo0[k] = lo[k]; node64.hi0[k] = hi[k];
        node.lo1[k] = INFINITY; node.hi1[k] = -INFINITY;
        node64.lo1[k] = INFINITY; node64.hi1[k] = -INFINITY;
    }
    node.ref0 = node64.ref0 = RRS_MAKE_LEAF(0, 2);
    node.ref1 = node64.ref1 = RRS_REF_EMPTY;

    RrsSceneDesc desc;
    memset(&desc, 0, sizeof desc);
    desc.abi_version = RRS_ABI_VERSION;
    desc.n_prims = 2; desc.prims = prims;
    desc.n_nodes = 1; desc.nodes = &node; desc.nodes_f64 = &node64;
    desc.max_depth = 1;
    desc.n_materials = 2; desc.materials = mats;
    desc.n_emissions = 0; desc.emissions = NULL;
    desc.hdri_width = hw; desc.hdri_height = hh; desc.hdri_rgb = hdri;
    desc.t_min = 1e-6; desc.t_max = 1e6; /* rayrs/src/main.rs:52 */

    RrsScene* scene = NULL;
    if (rrs_scene_create(&desc, 0, &scene) != RRS_OK) return fail("rrs_scene_create");

    /* film of W x H pixels at 100 ppi (ppc = 254), camera of test_scenes.rs:26-34 */
    RrsCamera cam = camera_new(v3(0, 5, 10), v3(0, 1, 0), v3(0, 1, 0), 50.0, (double)W / 254.0, (double)H / 254.0, 100);
    if (cam.x_pixels != W || cam.y_pixels != H) {
        fprintf(stderr, "film rounding: %u x %u\n", cam.x_pixels, cam.y_pixels);
        return 1;
    }
    RrsRenderParams p;
    memset(&p, 0, sizeof p);
    p.width = W; p.height = H; p.spp = spp; p.sample_offset = 0; p.spp_total = spp;
    p.max_bounces = 50; /* main.rs:77 */
    p.seed = 0x5EEDB200ull;

    float* rgb = (float*)malloc(sizeof(float) * 3u * (size_t)W * H);
    if (rrs_render(scene, &cam, &p, rgb) != RRS_OK) return fail("rrs_render");
    RrsStats st;
    if (rrs_stats(scene, &st) != RRS_OK) return fail("rrs_stats");
    double mean = 0.0;
    for (size_t i = 0; i < 3u * (size_t)W * H; ++i) mean += rgb[i];
    mean /= 3.0 * (double)W * H;
    printf("render_c: %ux%u, %u spp: %llu rays in %.3f ms (%.1f Mrays/s), mean radiance %.6f, nan %llu negative %llu\n", W, H,
           spp, (unsigned long long)st.rays, st.device_ms, (double)st.rays / st.device_ms / 1e3, mean,
           (unsigned long long)st.nan_pixels, (unsigned long long)st.negative_pixels);

    /* closest hit through the parity entry: the ray down the optical axis meets the sphere (object 1) */
    RrsRay ray;
    put3(ray.origin, v3(0, 5, 10));
    put3(ray.direction, v3(0, -4, -10));
    int32_t id = -2;
    double t = 0.0;
    if (rrs_intersect(scene, &ray, 1, &id, &t, 32) != RRS_OK) return fail("rrs_intersect");
    printf("render_c: axis ray hits object %d at t = %.9f\n", id, t);

    f = fopen(argv[7], "wb");
    if (!f || fwrite(rgb, sizeof(float), 3u * (size_t)W * H, f) != 3u * (size_t)W * H) {
        fprintf(stderr, "cannot write %s\n", argv[7]);
        return 2;
    }
    fclose(f);
    rrs_scene_destroy(scene);
    free(rgb);
    free(hdri);
    return 0;
}
