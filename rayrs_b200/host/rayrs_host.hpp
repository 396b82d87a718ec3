// rayrs_host — C++ host mirror of the rayrs-lib scene / material / camera / BVH API.
//
// In the real integration this layer is rayrs-lib itself (Rust) plus a small `gpu` module
// (INTEGRATION.md).  No Rust toolchain exists in this image, so the host side above the C ABI
// is written in C++ with the same type names, constructor arguments and panics-as-exceptions
// as the reference, so that callers and tests read like the reference's own:
//
//   Camera::new / x_pixels / y_pixels            rayrs-lib/src/lib.rs:99-177
//   Object::{sphere,plane,triangle,from_triangles,from_spheres,box_geom}   lib.rs:321-507
//   Material::* constructors                      rayrs-lib/src/material.rs:595-901
//   BvhHeuristic::{Midpoint,Sah}, Bvh::build      rayrs-lib/src/bvh.rs:187-210,227-389
//   Scene::new                                    lib.rs:227-245
//   render_gpu(&Camera,&Scene,spp,max_bounces)    replaces rayrs/src/main.rs:57-101
//
// BVH construction stays on the host and produces the SAME tree as bvh.rs (prefix/suffix
// boxes instead of the reference's O(splits*n) rescans; min/max are exact so every SAH cost
// is bit-identical), then flattens it into the 64-byte node array of include/rayrs_b200.h.
#pragma once
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <algorithm>
#include <thread>
#include <vector>

#include "../../include/rayrs_b200.h"

namespace rayrs {

// The reference panics (assert!/unwrap) on bad construction; the mirror throws.
struct Panic : std::runtime_error {
    using std::runtime_error::runtime_error;
};

struct Vec3 {
    double x = 0, y = 0, z = 0;
    Vec3() = default;
    Vec3(double x_, double y_, double z_) : x(x_), y(y_), z(z_) {}
    static Vec3 ones() { return Vec3(1, 1, 1); }
    static Vec3 zeros() { return Vec3(0, 0, 0); }
    static Vec3 unit_x() { return Vec3(1, 0, 0); }
    static Vec3 unit_y() { return Vec3(0, 1, 0); }
    static Vec3 unit_z() { return Vec3(0, 0, 1); }
    bool xyz_in_range_inclusive(double lo, double hi) const {
        return x >= lo && x <= hi && y >= lo && y <= hi && z >= lo && z <= hi;
    }
};
Vec3 operator+(Vec3 a, Vec3 b);
Vec3 operator-(Vec3 a, Vec3 b);
Vec3 operator*(Vec3 a, double s);
Vec3 operator*(double s, Vec3 a);
Vec3 operator/(Vec3 a, double s);  // reciprocal multiply, vecmath.rs:690-698
double dot(Vec3 a, Vec3 b);
Vec3 cross(Vec3 a, Vec3 b);
Vec3 unit(Vec3 a);

enum class Axis { X = 0, XRev = 1, Y = 2, YRev = 3, Z = 4, ZRev = 5 };

struct AxisAlignedBoundingBox {
    double xmin, xmax, ymin, ymax, zmin, zmax;
    Vec3 center() const;
    double surface_area() const;
    AxisAlignedBoundingBox expand(const AxisAlignedBoundingBox& o) const;
    bool degenerate() const { return !(xmax > xmin && ymax > ymin && zmax > zmin); }
};

struct Fresnel {
    RrsFresnelKind kind;
    double ior;
    Vec3 r0;
    static Fresnel SchlickDielectric(double ior) { return Fresnel{RRS_FRESNEL_DIELECTRIC, ior, Vec3()}; }
    static Fresnel SchlickMetallic(Vec3 r0) { return Fresnel{RRS_FRESNEL_METALLIC, 0., r0}; }
};

// enum Material with the constructors of the wrapped structs
struct Material {
    RrsMaterial m{};
    static Material LambertianDiffuse(Vec3 color);
    static Material Reflect(Vec3 color);
    static Material Refract(Vec3 color, double ior);
    static Material Glass(Vec3 color, double ior);
    static Material CookTorrance(Vec3 color, double alpha, Fresnel fresnel);
    static Material CookTorranceRefract(Vec3 color, double alpha, double ior);
    static Material CookTorranceGlass(Vec3 color, double alpha, double ior);
    static Material Plastic(Vec3 color, Vec3 spec_color, double alpha, double ior);
    static Material NoReflect();
};

struct Emission {
    bool dark = true;
    double strength = 0;
    Vec3 color;
    static Emission Dark() { return Emission{}; }
    static Emission Emissive(double strength, Vec3 color);  // Emission::new material.rs:1062-1070
};

struct Geometry {
    RrsPrimType type;
    double v[9];  // layout of RrsPrim::v
    AxisAlignedBoundingBox bbox() const;  // geometry.rs:687-733
};

struct Object {
    Geometry geom;
    Material mat;
    Emission emission;
    static Object sphere(double radius, Vec3 origin, Material mat, Emission emission);
    static Object plane(Axis axis, double umin, double umax, double vmin, double vmax, double pos, Material mat,
                        Emission emission);
    static Object triangle(Vec3 p1, Vec3 p2, Vec3 p3, Material mat, Emission emission);
    static std::vector<Object> box_geom(Vec3 lower_left, Vec3 upper_right, Material mat, Emission emission);
    // lib.rs:407-415: one object per triangle of a loaded mesh (mesh_io.hpp), all with the same material
    static std::vector<Object> from_triangles(const std::vector<struct Triangle>& tris, Material mat, Emission emission);
    // lib.rs:423-431: the same for spheres (e.g. one per vertex of a mesh, wavefront_obj::load_obj_file_spheres)
    static std::vector<Object> from_spheres(const std::vector<struct Sphere>& spheres, Material mat, Emission emission);
    AxisAlignedBoundingBox bbox() const { return geom.bbox(); }
};

// A read-only view of a Vec<Object>: a std::vector converts implicitly; the C API hands over a buffer it filled on all host
// threads without the serial zero-fill a std::vector<Object>(n) costs (0.4 s for 4M objects).
struct ObjectSpan {
    const Object* ptr = nullptr;
    size_t count = 0;
    ObjectSpan() = default;
    ObjectSpan(const Object* p, size_t n) : ptr(p), count(n) {}
    ObjectSpan(const std::vector<Object>& v) : ptr(v.data()), count(v.size()) {}
    size_t size() const { return count; }
    bool empty() const { return count == 0; }
    const Object& operator[](size_t i) const { return ptr[i]; }
};

struct BvhHeuristic {
    enum Kind { kMidpoint = 0, kSah = 1 } kind;
    uint32_t splits;
    static BvhHeuristic Midpoint() { return BvhHeuristic{kMidpoint, 0}; }
    static BvhHeuristic Sah(uint32_t splits) { return BvhHeuristic{kSah, splits}; }
};

// The reference tree, flattened.
struct FlatBvh {
    std::vector<RrsNode> nodes;
    std::vector<RrsNodeF64> nodes_f64;
    std::vector<uint32_t> prim_order;  // DFS leaf order -> index into the object list
    uint32_t max_depth = 0;
    uint32_t dead_nodes = 0;           // zero-extent nodes dropped
    // pre-order dump in the oracle's format: Node -> -(nchildren), LeafNode -> object index;
    // boxes: 6 doubles (xmin,xmax,ymin,ymax,zmin,zmax) per Node in the same order
    std::vector<int64_t> topology;
    std::vector<double> boxes;
};

// where the build time goes (seconds), filled when a pointer is handed to Bvh::build
struct BvhBuildTiming {
    double boxes = 0, recursive = 0, numbering = 0, flatten = 0, depth = 0, topology = 0;
    double device = 0;  // of `recursive`: the device time of rrs_bvh_build when the tree was built on the GPU
};

struct Bvh {
    // Bvh::build bvh.rs:199-210.  threads: worker threads for independent subtrees and chunked sorts
    // (0 = one per hardware thread).
    // build_device >= 0: the tree is built on that GPU (rrs_bvh_build: the same tree, level-synchronous) instead of
    // by the recursive host build; numbering and flattening are the same code either way.
    // want_topology: also produce the pre-order dump (FlatBvh::topology / boxes) the tests compare with the oracle's.
    static FlatBvh build(BvhHeuristic heuristic, ObjectSpan objects, uint32_t bfs_nodes = 1023,
                         int threads = 0, BvhBuildTiming* timing = nullptr, int build_device = -1, bool want_topology = true);
};

struct Image {
    size_t width = 0, height = 0;
    std::vector<Vec3> pixels;  // row-major, row 0 = top
    static Image from_pixels(size_t w, size_t h, std::vector<Vec3> px) { return Image{w, h, std::move(px)}; }
};

class Camera {
public:
    // Camera::new lib.rs:99-133
    Camera(Vec3 origin, Vec3 up, Vec3 lookat, double fov, double width, double height, uint32_t ppi);
    size_t x_pixels() const;  // lib.rs:153-155
    size_t y_pixels() const;  // lib.rs:175-177
    RrsCamera derived() const;

private:
    Vec3 origin_, e_x_, e_y_, z_;
    double width_, height_;
    uint32_t ppc_;
};

// Device side of Scene::new: which GPUs hold the scene, and the library's measurement switches.
struct SceneOptions {
    std::vector<int> devices{0};  // the BVH is built and flattened ONCE and uploaded to each of them
    bool with_f64 = true;         // also upload the f64 twins (rrs_intersect precision 64)
    bool upload = true;           // false: host half only (no GPU needed)
    uint32_t flags = 0;           // RRS_SCENE_*
    uint32_t refill_lanes = 0;    // 0 = library default
    int bvh_threads = 0;          // 0 = one per hardware thread
    bool device_build = false;    // build the BVH on devices[0] (rrs_bvh_build) instead of on the host
    bool topology = true;         // keep the oracle-format dump of the tree (tests); a renderer does not need it
};

// f(begin, end) over [0, n) on up to one thread per hardware thread (chunks of at least min_chunk items)
template <typename F>
void parallel_chunks(size_t n, size_t min_chunk, F f) {
    size_t threads = std::max<size_t>(1, std::thread::hardware_concurrency());
    threads = std::min(threads, std::max<size_t>(1, n / std::max<size_t>(1, min_chunk)));
    if (threads <= 1) {
        f((size_t)0, n);
        return;
    }
    std::vector<std::thread> pool;
    for (size_t t = 0; t < threads; ++t) pool.emplace_back([=]() { f(n * t / threads, n * (t + 1) / threads); });
    for (auto& th : pool) th.join();
}

class Scene {
public:
    // Scene::new lib.rs:227-245 (+ upload to the GPU `device`)
    Scene(ObjectSpan objects, double z_near, double z_far, BvhHeuristic heuristic, const Image& hdri, int device = 0,
          bool with_f64 = true, bool upload = true);
    Scene(ObjectSpan objects, double z_near, double z_far, BvhHeuristic heuristic, const Image& hdri, const SceneOptions& opt);
    ~Scene();
    Scene(const Scene&) = delete;
    Scene& operator=(const Scene&) = delete;

    RrsScene* handle() const { return handles_.empty() ? nullptr : handles_[0]; }
    const std::vector<RrsScene*>& handles() const { return handles_; }
    const std::vector<int>& devices() const { return devices_; }
    RrsComm* comm() const;  // made on first use (rrs_comm_init_all over devices())
    const BvhBuildTiming& build_timing() const { return timing_; }
    const FlatBvh& bvh() const { return bvh_; }
    const std::vector<RrsPrim>& prims() const { return prims_; }
    const std::vector<RrsMaterial>& materials() const { return materials_; }
    double build_seconds() const { return build_seconds_; }

private:
    FlatBvh bvh_;
    std::vector<RrsPrim> prims_;
    std::vector<RrsMaterial> materials_;
    std::vector<RrsEmission> emissions_;
    std::vector<RrsScene*> handles_;
    std::vector<int> devices_;
    mutable RrsComm* comm_ = nullptr;
    BvhBuildTiming timing_;
    double build_seconds_ = 0;
    void init(ObjectSpan objects, double z_near, double z_far, BvhHeuristic heuristic, const Image& hdri, const SceneOptions& opt);
};

struct RenderOptions {
    uint64_t seed = 0x5EEDB200ull;
    uint32_t sample_offset = 0;
    uint32_t spp_total = 0;
    uint32_t queue_capacity = 0;
    uint32_t flags = 0;
};

// Drop-in for the tile loop of rayrs/src/main.rs:57-101.  A scene that lives on several GPUs (SceneOptions::devices)
// is rendered by all of them: samples split, one NCCL reduce (rrs_render_multi).
Image render_gpu(const Camera& c, const Scene& s, uint32_t spp, uint32_t max_bounces, const RenderOptions& opt = RenderOptions());
// same, into a caller-provided float buffer (height*width*3)
void render_gpu_into(const Camera& c, const Scene& s, uint32_t spp, uint32_t max_bounces, const RenderOptions& opt,
                     float* out_rgb);

}  // namespace rayrs
