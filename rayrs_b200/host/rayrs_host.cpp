// Host mirror implementation: constructors with the reference's assert!s, the BVH builder
// (same tree as rayrs-lib/src/bvh.rs:227-389) and its flattening into RrsNode[].
#include "rayrs_host.hpp"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstring>
#include <functional>
#include <future>
#include <limits>
#include <mutex>
#include <cstdio>
#include <thread>

namespace rayrs {

// ---------------------------------------------------------------------------------------
// vecmath (rayrs-lib/src/vecmath.rs)
// ---------------------------------------------------------------------------------------
Vec3 operator+(Vec3 a, Vec3 b) { return Vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
Vec3 operator-(Vec3 a, Vec3 b) { return Vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
Vec3 operator*(Vec3 a, double s) { return Vec3(a.x * s, a.y * s, a.z * s); }
Vec3 operator*(double s, Vec3 a) { return Vec3(s * a.x, s * a.y, s * a.z); }
Vec3 operator/(Vec3 a, double s) {
    double inv = 1. / s;
    return a * inv;
}
double dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
Vec3 cross(Vec3 a, Vec3 b) { return Vec3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
Vec3 unit(Vec3 a) { return a / std::sqrt(dot(a, a)); }

// ---------------------------------------------------------------------------------------
// AABB (geometry.rs:577-582,640-645,674-683)
// ---------------------------------------------------------------------------------------
Vec3 AxisAlignedBoundingBox::center() const {
    return Vec3((xmax - xmin) / 2. + xmin, (ymax - ymin) / 2. + ymin, (zmax - zmin) / 2. + zmin);
}
double AxisAlignedBoundingBox::surface_area() const {
    double x = xmax - xmin, y = ymax - ymin, z = zmax - zmin;
    return 2. * x * y + 2. * y * z + 2. * x * z;
}
AxisAlignedBoundingBox AxisAlignedBoundingBox::expand(const AxisAlignedBoundingBox& o) const {
    return AxisAlignedBoundingBox{std::fmin(xmin, o.xmin), std::fmax(xmax, o.xmax), std::fmin(ymin, o.ymin),
                                  std::fmax(ymax, o.ymax), std::fmin(zmin, o.zmin), std::fmax(zmax, o.zmax)};
}

// ---------------------------------------------------------------------------------------
// Materials (material.rs:595-901): same assertions, as exceptions
// ---------------------------------------------------------------------------------------
static void require(bool ok, const char* what) {
    if (!ok) throw Panic(what);
}
static void set3(double* d, Vec3 v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; }

Material Material::LambertianDiffuse(Vec3 color) {
    require(color.xyz_in_range_inclusive(0., 1.), "LambertianDiffuse::new: color not in [0,1]");
    Material r;
    r.m.tag = RRS_MAT_LAMBERTIAN;
    set3(r.m.color, color);
    return r;
}
Material Material::Reflect(Vec3 color) {
    require(color.xyz_in_range_inclusive(0., 1.), "Reflect::new: color not in [0,1]");
    Material r;
    r.m.tag = RRS_MAT_REFLECT;
    set3(r.m.color, color);
    return r;
}
Material Material::Refract(Vec3 color, double ior) {
    require(color.xyz_in_range_inclusive(0., 1.), "Refract::new: color not in [0,1]");
    require(ior > 0. && std::isfinite(ior), "Refract::new: ior must be positive and finite");
    Material r;
    r.m.tag = RRS_MAT_REFRACT;
    set3(r.m.color, color);
    r.m.ior = ior;
    return r;
}
Material Material::Glass(Vec3 color, double ior) {
    require(color.xyz_in_range_inclusive(0., 1.), "Glass::new: color not in [0,1]");
    require(ior > 0. && std::isfinite(ior), "Glass::new: ior must be positive and finite");
    Material r;
    r.m.tag = RRS_MAT_GLASS;
    set3(r.m.color, color);
    r.m.ior = ior;
    return r;
}
Material Material::CookTorrance(Vec3 color, double alpha, Fresnel fresnel) {
    require(color.xyz_in_range_inclusive(0., 1.), "CookTorrance::new: color not in [0,1]");
    require(alpha > 0. && std::isfinite(alpha), "CookTorrance::new: alpha must be positive and finite");
    Material r;
    r.m.tag = RRS_MAT_COOK_TORRANCE;
    set3(r.m.color, color);
    r.m.alpha = alpha;
    r.m.fresnel_kind = fresnel.kind;
    if (fresnel.kind == RRS_FRESNEL_METALLIC) set3(r.m.spec_color, fresnel.r0);
    else r.m.ior = fresnel.ior;
    return r;
}
Material Material::CookTorranceRefract(Vec3 color, double alpha, double ior) {
    require(color.xyz_in_range_inclusive(0., 1.), "CookTorranceRefract::new: color not in [0,1]");
    require(alpha > 0. && std::isfinite(alpha), "CookTorranceRefract::new: alpha must be positive and finite");
    require(ior > 0. && std::isfinite(ior), "CookTorranceRefract::new: ior must be positive and finite");
    Material r;
    r.m.tag = RRS_MAT_COOK_TORRANCE_REFRACT;
    set3(r.m.color, color);
    r.m.alpha = alpha;
    r.m.ior = ior;
    return r;
}
Material Material::CookTorranceGlass(Vec3 color, double alpha, double ior) {
    require(color.xyz_in_range_inclusive(0., 1.), "CookTorranceGlass::new: color not in [0,1]");
    require(alpha > 0. && std::isfinite(alpha), "CookTorranceGlass::new: alpha must be positive and finite");
    require(ior > 0. && std::isfinite(ior), "CookTorranceGlass::new: ior must be positive and finite");
    Material r;
    r.m.tag = RRS_MAT_COOK_TORRANCE_GLASS;
    set3(r.m.color, color);
    r.m.alpha = alpha;
    r.m.ior = ior;
    return r;
}
Material Material::Plastic(Vec3 color, Vec3 spec_color, double alpha, double ior) {
    require(color.xyz_in_range_inclusive(0., 1.), "Plastic::new: color not in [0,1]");
    require(spec_color.xyz_in_range_inclusive(0., 1.), "Plastic::new: spec_color not in [0,1]");
    require(alpha > 0. && std::isfinite(alpha), "Plastic::new: alpha must be positive and finite");
    require(ior > 0. && std::isfinite(ior), "Plastic::new: ior must be positive and finite");
    Material r;
    r.m.tag = RRS_MAT_PLASTIC;
    set3(r.m.color, color);
    set3(r.m.spec_color, spec_color);
    r.m.alpha = alpha;
    r.m.ior = ior;
    return r;
}
Material Material::NoReflect() {
    Material r;
    r.m.tag = RRS_MAT_NO_REFLECT;
    return r;
}

Emission Emission::Emissive(double strength, Vec3 color) {
    require(strength >= 0., "Emission::new: strength must be >= 0");
    require(color.xyz_in_range_inclusive(0., 1.), "RGB values need to be between 0 and 1");
    Emission e;
    e.dark = false;
    e.strength = strength;
    e.color = color;
    return e;
}

// ---------------------------------------------------------------------------------------
// Geometry / Object (geometry.rs, lib.rs:321-507)
// ---------------------------------------------------------------------------------------
AxisAlignedBoundingBox Geometry::bbox() const {
    if (type == RRS_SPHERE) {  // geometry.rs:687-696 (bbox uses sqrt(radius2))
        double r = std::sqrt(v[0]);
        return AxisAlignedBoundingBox{v[1] - r, v[1] + r, v[2] - r, v[2] + r, v[3] - r, v[3] + r};
    }
    if (type == RRS_PLANE) {  // geometry.rs:699-718
        int axis = ((int)v[0]) >> 1;
        if (axis == 0) return AxisAlignedBoundingBox{v[5], v[5], v[1], v[2], v[3], v[4]};
        if (axis == 1) return AxisAlignedBoundingBox{v[1], v[2], v[5], v[5], v[3], v[4]};
        return AxisAlignedBoundingBox{v[1], v[2], v[3], v[4], v[5], v[5]};
    }
    // geometry.rs:721-733
    return AxisAlignedBoundingBox{std::fmin(v[0], std::fmin(v[3], v[6])), std::fmax(v[0], std::fmax(v[3], v[6])),
                                  std::fmin(v[1], std::fmin(v[4], v[7])), std::fmax(v[1], std::fmax(v[4], v[7])),
                                  std::fmin(v[2], std::fmin(v[5], v[8])), std::fmax(v[2], std::fmax(v[5], v[8]))};
}

Object Object::sphere(double radius, Vec3 origin, Material mat, Emission emission) {
    require(radius > 0., "Radius has to be positive");  // geometry.rs:97
    Object o;
    o.geom.type = RRS_SPHERE;
    std::memset(o.geom.v, 0, sizeof(o.geom.v));
    o.geom.v[0] = radius * radius;  // Sphere stores radius2 (geometry.rs:98-101)
    o.geom.v[1] = origin.x; o.geom.v[2] = origin.y; o.geom.v[3] = origin.z;
    o.mat = mat;
    o.emission = emission;
    return o;
}
Object Object::plane(Axis axis, double umin, double umax, double vmin, double vmax, double pos, Material mat,
                     Emission emission) {
    require(umin < umax && vmin < vmax, "Plane cannot be constructed with these ranges");  // geometry.rs:205-212
    Object o;
    o.geom.type = RRS_PLANE;
    std::memset(o.geom.v, 0, sizeof(o.geom.v));
    o.geom.v[0] = (double)(int)axis;
    o.geom.v[1] = umin; o.geom.v[2] = umax; o.geom.v[3] = vmin; o.geom.v[4] = vmax; o.geom.v[5] = pos;
    o.mat = mat;
    o.emission = emission;
    return o;
}
Object Object::triangle(Vec3 p1, Vec3 p2, Vec3 p3, Material mat, Emission emission) {
    Object o;
    o.geom.type = RRS_TRIANGLE;
    o.geom.v[0] = p1.x; o.geom.v[1] = p1.y; o.geom.v[2] = p1.z;
    o.geom.v[3] = p2.x; o.geom.v[4] = p2.y; o.geom.v[5] = p2.z;
    o.geom.v[6] = p3.x; o.geom.v[7] = p3.y; o.geom.v[8] = p3.z;
    o.mat = mat;
    o.emission = emission;
    return o;
}
std::vector<Object> Object::box_geom(Vec3 ll, Vec3 ur, Material mat, Emission emission) {
    // lib.rs:438-507, same six planes in the same order (note: the last one really is at
    // lower_left.y like the one before it)
    return {
        Object::plane(Axis::X, ll.y, ur.y, ll.z, ur.z, ll.x, mat, emission),
        Object::plane(Axis::XRev, ll.y, ur.y, ll.z, ur.z, ur.x, mat, emission),
        Object::plane(Axis::ZRev, ll.x, ur.x, ll.y, ur.y, ll.z, mat, emission),
        Object::plane(Axis::Z, ll.x, ur.x, ll.y, ur.y, ur.z, mat, emission),
        Object::plane(Axis::YRev, ll.x, ur.x, ll.z, ur.z, ll.y, mat, emission),
        Object::plane(Axis::Y, ll.x, ur.x, ll.z, ur.z, ll.y, mat, emission),
    };
}

// ---------------------------------------------------------------------------------------
// BVH build (bvh.rs:227-389) on index ranges
// ---------------------------------------------------------------------------------------
namespace {

struct BNode {
    AxisAlignedBoundingBox box;
    int32_t child[2] = {-1, -1};  // build-node index, or -1
    uint32_t first = 0, count = 0; // leaf group / bare leaf: range in `order`
    uint8_t kind = 0;              // 0 binary Node, 1 group Node (<=4 LeafNodes), 2 bare LeafNode
};

struct Builder {
    const std::vector<AxisAlignedBoundingBox>& boxes;
    const std::vector<Vec3>& centers;
    std::vector<uint32_t>& order;
    BvhHeuristic heur;
    std::vector<BNode> nodes;
    std::mutex mu;
    std::atomic<int> spare_threads{0};

    Builder(const std::vector<AxisAlignedBoundingBox>& b, const std::vector<Vec3>& c, std::vector<uint32_t>& o,
            BvhHeuristic h)
        : boxes(b), centers(c), order(o), heur(h) {}

    int32_t alloc(const BNode& n) {
        std::lock_guard<std::mutex> g(mu);
        nodes.push_back(n);
        return (int32_t)nodes.size() - 1;
    }
    static double key(const Vec3& c, int axis) { return axis == 0 ? c.x : (axis == 1 ? c.y : c.z); }

    struct KeyIdx {
        double k;
        uint32_t i;
    };
    int borrow_threads(int want) {
        int got = 0;
        while (got < want) {
            int v = spare_threads.load();
            if (v <= 0) break;
            if (spare_threads.compare_exchange_weak(v, v - 1)) ++got;
        }
        return got;
    }
    // stable sort by key; ranges above 64K entries use up to 8 chunks on borrowed threads + stable pairwise merges
    void stable_sort_pairs(std::vector<KeyIdx>& v) {
        auto less = [](const KeyIdx& a, const KeyIdx& b) { return a.k < b.k; };
        const size_t n = v.size();
        int extra = n > 65536 ? borrow_threads(7) : 0;
        int chunks = extra >= 7 ? 8 : (extra >= 3 ? 4 : (extra >= 1 ? 2 : 1));
        if (chunks == 1) {
            spare_threads.fetch_add(extra);
            std::stable_sort(v.begin(), v.end(), less);
            return;
        }
        spare_threads.fetch_add(extra - (chunks - 1));  // keep only what the chunking uses
        std::vector<size_t> cut(chunks + 1);
        for (int c = 0; c <= chunks; ++c) cut[c] = n * (size_t)c / (size_t)chunks;
        {
            std::vector<std::future<void>> jobs;
            for (int c = 1; c < chunks; ++c)
                jobs.push_back(std::async(std::launch::async, [&, c]() { std::stable_sort(v.begin() + cut[c], v.begin() + cut[c + 1], less); }));
            std::stable_sort(v.begin() + cut[0], v.begin() + cut[1], less);
            for (auto& j : jobs) j.get();
        }
        for (int width = 1; width < chunks; width *= 2) {
            std::vector<std::future<void>> jobs;
            for (int c = 0; c + width < chunks; c += 2 * width) {
                const size_t a = cut[c], m = cut[c + width], b = cut[std::min(chunks, c + 2 * width)];
                auto merge = [&v, a, m, b, less]() { std::inplace_merge(v.begin() + a, v.begin() + m, v.begin() + b, less); };
                if (c == 0) continue;
                jobs.push_back(std::async(std::launch::async, merge));
            }
            std::inplace_merge(v.begin() + cut[0], v.begin() + cut[width], v.begin() + cut[std::min(chunks, 2 * width)], less);
            for (auto& j : jobs) j.get();
        }
        spare_threads.fetch_add(chunks - 1);
    }

    AxisAlignedBoundingBox range_box(size_t lo, size_t hi) const {
        AxisAlignedBoundingBox b = boxes[order[lo]];
        for (size_t i = lo + 1; i < hi; ++i) b = b.expand(boxes[order[i]]);
        return b;
    }

    // Scratch of one node, reused down the recursion of a thread (everything in it is dead before the children are
    // built); forked subtrees run on their own threads and get their own.
    struct Scratch {
        std::vector<KeyIdx> tmp;
        std::vector<double> keys;
        std::vector<AxisAlignedBoundingBox> nb, pre, suf;
        void release() { *this = Scratch(); }
    };
    static Scratch& scratch() {
        thread_local Scratch s;
        return s;
    }

    // known_box: the node's box when the parent already folded it (same elements, same order as range_box -> the
    // same bits), nullptr at the root
    int32_t build(size_t lo, size_t hi, int depth, const AxisAlignedBoundingBox* known_box = nullptr) {
        const size_t n = hi - lo;
        if (n == 0) throw Panic("Having a BVH for 0 objects does not make sense");  // bvh.rs:229
        BNode node;
        node.box = known_box ? *known_box : range_box(lo, hi);
        if (n <= 4) {  // bvh.rs:304-315
            node.kind = 1;
            node.first = (uint32_t)lo;
            node.count = (uint32_t)n;
            return alloc(node);
        }
        const AxisAlignedBoundingBox& bb = node.box;
        double x = bb.xmax - bb.xmin, y = bb.ymax - bb.ymin, z = bb.zmax - bb.zmin;
        int axis;
        double amin, alen;
        if (x >= y && x >= z) { axis = 0; amin = bb.xmin; alen = x; }
        else if (y >= z) { axis = 1; amin = bb.ymin; alen = y; }
        else { axis = 2; amin = bb.zmin; alen = z; }
        // BvhData::sort — Rust's sort_by is stable.  Sorted as (key, index) pairs: the comparator reads the key
        // next to the index instead of chasing it through `centers` (the same order as a stable sort of the
        // indices with that comparator), and a large range is cut into chunks sorted on the spare threads and
        // merged stably.
        Scratch& sx = scratch();
        std::vector<double>& keys = sx.keys;
        std::vector<AxisAlignedBoundingBox>& nb = sx.nb;  // the node's boxes in sorted order: ONE gather through `order`
        {
            std::vector<KeyIdx>& tmp = sx.tmp;
            tmp.resize(n);
            for (size_t k = 0; k < n; ++k) {
                const uint32_t i = order[lo + k];
                tmp[k] = KeyIdx{key(centers[i], axis), i};
                if (std::isnan(tmp[k].k)) throw Panic("partial_cmp().unwrap() on NaN centre");  // bvh.rs:104
            }
            stable_sort_pairs(tmp);
            keys.resize(n);
            nb.resize(n);
            for (size_t k = 0; k < n; ++k) {
                order[lo + k] = tmp[k].i;
                keys[k] = tmp[k].k;
                nb[k] = boxes[tmp[k].i];
            }
        }
        long ind = -1;
        if (heur.kind == BvhHeuristic::kSah) {
            // bvh.rs:258-271 with calculate_sah bvh.rs:15-38.  left(k) = [0,k), right(k) = [k,n)
            std::vector<AxisAlignedBoundingBox>& suf = sx.suf;
            suf.resize(n);
            suf[n - 1] = nb[n - 1];
            for (size_t k = n - 1; k-- > 0;) suf[k] = nb[k].expand(suf[k + 1]);
            std::vector<AxisAlignedBoundingBox>& pre = sx.pre;  // pre[k] = box of [0,k], so left(k) = pre[k-1]
            pre.resize(n);
            pre[0] = nb[0];
            for (size_t k = 1; k < n; ++k) pre[k] = pre[k - 1].expand(nb[k]);
            const double surface_area = bb.surface_area();
            const double split_dist = alen / (double)(heur.splits - 1);
            double min_sah = std::numeric_limits<double>::infinity();
            auto thr_of = [&](long i) { return amin + (double)i * split_dist; };  // bvh.rs:262, same expression
            auto consider = [&](size_t k) {  // calculate_sah bvh.rs:15-38 for left = [0,k), right = [k,n)
                double p_left = k > 0 ? pre[k - 1].surface_area() / surface_area : 0.;
                double p_right = suf[k].surface_area() / surface_area;
                double sah = 0.3 + 1. * (p_left * (double)k + p_right * (double)(n - k));
                if (sah < min_sah) {
                    min_sah = sah;
                    ind = (long)k;
                }
            };
            if (n < (size_t)heur.splits && split_dist > 0. && std::isfinite(split_dist)) {
                // Few primitives, many thresholds: most of the `splits` thresholds fall between the same two
                // centres and repeat a split already costed.  Walk the DISTINCT split indices instead, in the order
                // the threshold loop meets them (k(i) = #centres <= thr_i never decreases with i, and the strict <
                // keeps the first of equal costs, so the outcome is the loop's).  For the next centre the smallest i
                // with thr_i >= centre is guessed by a division and then settled with the loop's own expression.
                const long splits = (long)heur.splits;
                long i = 1;
                size_t k = std::upper_bound(keys.begin(), keys.end(), thr_of(1)) - keys.begin();
                while (k < n) {
                    consider(k);
                    const double c = keys[k];  // the next centre a threshold has to reach
                    long g = (long)std::ceil((c - amin) / split_dist);
                    g = std::max(g, i + 1);
                    g = std::min(g, splits + 1);
                    while (g > i + 1 && thr_of(g - 1) >= c) --g;
                    while (g <= splits && thr_of(g) < c) ++g;
                    if (g > splits) break;
                    i = g;
                    k = std::upper_bound(keys.begin() + k, keys.end(), thr_of(i)) - keys.begin();
                }
            } else {
                long last = -2;
                for (uint32_t i = 1; i < heur.splits + 1; ++i) {
                    double thr = thr_of((long)i);
                    size_t k = std::upper_bound(keys.begin(), keys.end(), thr) - keys.begin();  // first centre > thr
                    if (k >= n) continue;           // split_index -> None
                    if ((long)k == last) continue;  // same split => same cost; strict < keeps the first
                    last = (long)k;
                    consider(k);
                }
            }
        } else {
            // bvh.rs:337-350
            double split = axis == 0 ? bb.center().x : (axis == 1 ? bb.center().y : bb.center().z);
            size_t k = std::upper_bound(keys.begin(), keys.end(), split) - keys.begin();
            ind = k >= n ? -1 : (long)k;
        }
        if (ind < 0 || ind == 0 || ind == (long)n - 1) ind = (long)(n / 2);  // bvh.rs:279-287
        // the children's boxes, folded left to right over their elements in this node's sorted order — what their own
        // range_box would gather again
        AxisAlignedBoundingBox box_l = nb[0], box_r = nb[(size_t)ind];
        for (size_t k = 1; k < (size_t)ind; ++k) box_l = box_l.expand(nb[k]);
        for (size_t k = (size_t)ind + 1; k < n; ++k) box_r = box_r.expand(nb[k]);
        if (n > (1u << 16)) sx.release();  // do not hold the top levels' buffers while the subtrees are built
        node.kind = 0;
        const size_t mid = lo + (size_t)ind;
        auto side = [&](size_t a, size_t b) -> int32_t {
            if (b - a > 1) return build(a, b, depth + 1, a == lo ? &box_l : &box_r);
            BNode leaf;
            leaf.kind = 2;
            leaf.first = (uint32_t)a;
            leaf.count = 1;
            leaf.box = boxes[order[a]];
            return alloc(leaf);
        };
        // independent subtrees: run the left one on another thread while there are spares
        bool forked = false;
        std::future<int32_t> fut;
        if (n > 20000) {
            int v = spare_threads.load();
            while (v > 0 && !spare_threads.compare_exchange_weak(v, v - 1)) {
            }
            forked = v > 0;
        }
        if (forked) fut = std::async(std::launch::async, [&, lo, mid]() { return side(lo, mid); });
        int32_t right = side(mid, hi);
        int32_t left;
        if (forked) {
            left = fut.get();
            spare_threads.fetch_add(1);
        } else {
            left = side(lo, mid);
        }
        node.child[0] = left;
        node.child[1] = right;
        return alloc(node);
    }
};

float round_down(double x) {
    float f = (float)x;
    if ((double)f > x) f = std::nextafterf(f, -INFINITY);
    return f;
}
float round_up(double x) {
    float f = (float)x;
    if ((double)f < x) f = std::nextafterf(f, INFINITY);
    return f;
}

struct Flattener {
    const std::vector<BNode>& bn;
    FlatBvh& out;

    void set_child(uint32_t node, int ch, uint32_t ref, bool bare, const AxisAlignedBoundingBox* f64box,
                   const AxisAlignedBoundingBox* f32box) {
        RrsNode& n = out.nodes[node];
        RrsNodeF64& d = out.nodes_f64[node];
        float* lo = ch ? n.lo1 : n.lo0;
        float* hi = ch ? n.hi1 : n.hi0;
        double* dlo = ch ? d.lo1 : d.lo0;
        double* dhi = ch ? d.hi1 : d.hi0;
        (ch ? n.ref1 : n.ref0) = ref;
        (ch ? d.ref1 : d.ref0) = ref;
        if (bare) {
            n.flags |= 1u << ch;
            d.flags |= 1u << ch;
        }
        if (ref == RRS_REF_EMPTY || !f32box) {
            for (int k = 0; k < 3; ++k) {
                lo[k] = INFINITY; hi[k] = -INFINITY;
                dlo[k] = INFINITY; dhi[k] = -INFINITY;
            }
            return;
        }
        lo[0] = round_down(f32box->xmin); lo[1] = round_down(f32box->ymin); lo[2] = round_down(f32box->zmin);
        hi[0] = round_up(f32box->xmax); hi[1] = round_up(f32box->ymax); hi[2] = round_up(f32box->zmax);
        dlo[0] = f64box->xmin; dlo[1] = f64box->ymin; dlo[2] = f64box->zmin;
        dhi[0] = f64box->xmax; dhi[1] = f64box->ymax; dhi[2] = f64box->zmax;
    }

    // reference of a child as seen from its parent (whose own box is `parent_box`)
    // returns 1 when the child was a dead node (dropped), else 0
    uint32_t attach(uint32_t node, int ch, int32_t b, const AxisAlignedBoundingBox& parent_box,
                    const std::vector<int32_t>& flat_index) {
        const BNode& c = bn[b];
        if (c.kind == 2) {
            // bare LeafNode: no box in the reference; cull with the primitive's own box when
            // that box is a real volume, else with the parent's
            const AxisAlignedBoundingBox* fb = c.box.degenerate() ? &parent_box : &c.box;
            set_child(node, ch, RRS_MAKE_LEAF(c.first, 1), true, &c.box, fb);
            return 0;
        }
        if (c.box.degenerate()) {  // slab test can never accept a zero-extent box
            set_child(node, ch, RRS_REF_EMPTY, false, nullptr, nullptr);
            return 1;
        }
        if (c.kind == 1) {
            set_child(node, ch, RRS_MAKE_LEAF(c.first, c.count), false, &c.box, &c.box);
            return 0;
        }
        set_child(node, ch, (uint32_t)flat_index[b], false, &c.box, &c.box);
        return 0;
    }
};

void dump_topology(const std::vector<BNode>& bn, int32_t b, const std::vector<uint32_t>& order, FlatBvh& out) {
    const BNode& n = bn[b];
    if (n.kind == 2) {
        out.topology.push_back(order[n.first]);
        return;
    }
    const double box[6] = {n.box.xmin, n.box.xmax, n.box.ymin, n.box.ymax, n.box.zmin, n.box.zmax};
    out.boxes.insert(out.boxes.end(), box, box + 6);
    if (n.kind == 1) {
        out.topology.push_back(-(int64_t)n.count);
        for (uint32_t k = 0; k < n.count; ++k) out.topology.push_back(order[n.first + k]);
        return;
    }
    out.topology.push_back(-2);
    dump_topology(bn, n.child[0], order, out);
    dump_topology(bn, n.child[1], order, out);
}

}  // namespace

FlatBvh Bvh::build(BvhHeuristic heuristic, ObjectSpan objects, uint32_t bfs_nodes, int threads,
                   BvhBuildTiming* timing, int build_device, bool want_topology) {
    if (objects.empty()) throw Panic("Having a BVH for 0 objects does not make sense");
    if (heuristic.kind == BvhHeuristic::kSah && heuristic.splits < 2) throw Panic("Sah needs at least 2 splits");
    const size_t n = objects.size();
    auto tick = std::chrono::steady_clock::now();
    auto lap = [&](double BvhBuildTiming::*field) {
        auto now = std::chrono::steady_clock::now();
        if (timing) timing->*field = std::chrono::duration<double>(now - tick).count();
        tick = now;
    };
    std::vector<AxisAlignedBoundingBox> boxes(n);
    std::vector<Vec3> centers(n);
    parallel_chunks(n, 1 << 16, [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; ++i) {
            boxes[i] = objects[i].bbox();
            centers[i] = boxes[i].center();  // BvhData::new bvh.rs:87-98
        }
    });
    lap(&BvhBuildTiming::boxes);
    FlatBvh out;
    out.prim_order.resize(n);
    for (size_t i = 0; i < n; ++i) out.prim_order[i] = (uint32_t)i;
    Builder b(boxes, centers, out.prim_order, heuristic);
    int32_t root = 0;
    if (build_device >= 0) {
        // the same tree from the GPU (csrc/bvh_build.cu); its records are the builder's BNodes
        static_assert(sizeof(AxisAlignedBoundingBox) == 6 * sizeof(double), "boxes are handed over as n x 6 doubles");
        std::unique_ptr<RrsBuildNode[]> dn(new RrsBuildNode[2 * n + 2]);  // uninitialised: only n_nodes records get touched
        uint32_t n_nodes = 0;
        double dev_s = 0.;
        int rc = rrs_bvh_build(&boxes[0].xmin, (uint32_t)n, heuristic.kind == BvhHeuristic::kSah ? 1u : 0u, heuristic.splits, build_device,
                               out.prim_order.data(), dn.get(), &n_nodes, &dev_s);
        if (rc != RRS_OK) throw Panic(std::string("rrs_bvh_build failed: ") + rrs_last_error());
        if (timing) timing->device = dev_s;
        b.nodes.resize(n_nodes);
        parallel_chunks(n_nodes, 1 << 16, [&](size_t lo, size_t hi) {
            for (size_t i = lo; i < hi; ++i) {
                BNode& t = b.nodes[i];
                const RrsBuildNode& s = dn[i];
                t.box = AxisAlignedBoundingBox{s.box[0], s.box[1], s.box[2], s.box[3], s.box[4], s.box[5]};
                t.child[0] = s.child[0];
                t.child[1] = s.child[1];
                t.first = s.first;
                t.count = s.count;
                t.kind = (uint8_t)s.kind;
            }
        });
        root = 0;
    } else {
        if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
        b.spare_threads = threads - 1;
        b.nodes.reserve(n);
        root = b.build(0, n, 0);
    }
    const std::vector<BNode>& bn = b.nodes;
    lap(&BvhBuildTiming::recursive);

    // ---- flat numbering: virtual root = 0, then binary Nodes breadth-first for the first
    // `bfs_nodes` (hot top of the tree contiguous), depth-first below (subtrees contiguous)
    std::vector<int32_t> flat_index(bn.size(), -1);
    std::vector<int32_t> flat_to_build;
    flat_to_build.push_back(-1);  // virtual root
    {
        std::vector<int32_t> frontier;
        if (bn[root].kind == 0 && !bn[root].box.degenerate()) frontier.push_back(root);
        size_t head = 0;
        while (head < frontier.size() && flat_to_build.size() < (size_t)bfs_nodes + 1) {
            int32_t cur = frontier[head++];
            flat_index[cur] = (int32_t)flat_to_build.size();
            flat_to_build.push_back(cur);
            for (int c = 0; c < 2; ++c) {
                int32_t ch = bn[cur].child[c];
                if (bn[ch].kind == 0 && !bn[ch].box.degenerate()) frontier.push_back(ch);
            }
        }
        // the rest depth-first
        std::vector<int32_t> stack(frontier.begin() + head, frontier.end());
        std::reverse(stack.begin(), stack.end());
        while (!stack.empty()) {
            int32_t cur = stack.back();
            stack.pop_back();
            flat_index[cur] = (int32_t)flat_to_build.size();
            flat_to_build.push_back(cur);
            for (int c = 1; c >= 0; --c) {
                int32_t ch = bn[cur].child[c];
                if (bn[ch].kind == 0 && !bn[ch].box.degenerate()) stack.push_back(ch);
            }
        }
    }
    lap(&BvhBuildTiming::numbering);
    out.nodes.assign(flat_to_build.size(), RrsNode{});
    out.nodes_f64.assign(flat_to_build.size(), RrsNodeF64{});
    Flattener fl{bn, out};
    // virtual root
    {
        AxisAlignedBoundingBox rb = bn[root].box;
        if (rb.degenerate()) {
            out.dead_nodes++;
            fl.set_child(0, 0, RRS_REF_EMPTY, false, nullptr, nullptr);
        } else if (bn[root].kind == 1) {
            fl.set_child(0, 0, RRS_MAKE_LEAF(bn[root].first, bn[root].count), false, &rb, &rb);
        } else {
            fl.set_child(0, 0, (uint32_t)flat_index[root], false, &rb, &rb);
        }
        fl.set_child(0, 1, RRS_REF_EMPTY, false, nullptr, nullptr);
    }
    {
        // every flat node is written by exactly one iteration; dead subtrees are counted per chunk
        std::atomic<uint32_t> dead{0};
        parallel_chunks(flat_to_build.size() - 1, 1 << 15, [&](size_t lo, size_t hi) {
            Flattener part{bn, out};
            uint32_t mine = 0;
            for (size_t f = lo + 1; f < hi + 1; ++f) {
                const BNode& nd = bn[flat_to_build[f]];
                mine += part.attach((uint32_t)f, 0, nd.child[0], nd.box, flat_index);
                mine += part.attach((uint32_t)f, 1, nd.child[1], nd.box, flat_index);
            }
            dead += mine;
        });
        out.dead_nodes += dead.load();
    }
    lap(&BvhBuildTiming::flatten);
    // depth (stack bound): longest chain of flat nodes
    {
        std::vector<uint32_t> depth(out.nodes.size(), 0);
        uint32_t mx = 1;
        depth[0] = 1;
        // parents always precede... not in DFS/BFS mix; do an explicit walk
        std::vector<uint32_t> st{0};
        while (!st.empty()) {
            uint32_t f = st.back();
            st.pop_back();
            mx = std::max(mx, depth[f]);
            const uint32_t refs[2] = {out.nodes[f].ref0, out.nodes[f].ref1};
            for (uint32_t r : refs)
                if (r != RRS_REF_EMPTY && !(r & RRS_REF_LEAF)) {
                    depth[r] = depth[f] + 1;
                    st.push_back(r);
                }
        }
        out.max_depth = mx;
    }
    lap(&BvhBuildTiming::depth);
    if (want_topology) dump_topology(bn, root, out.prim_order, out);
    lap(&BvhBuildTiming::topology);
    return out;
}

// ---------------------------------------------------------------------------------------
// Camera (lib.rs:99-177)
// ---------------------------------------------------------------------------------------
Camera::Camera(Vec3 origin, Vec3 up, Vec3 lookat, double fov, double width, double height, uint32_t ppi) {
    require(fov > 0. && fov < 180., "Camera::new: fov must be within (0, 180)");
    require(width > 0., "Camera::new: width must be positive");
    require(height > 0., "Camera::new: height must be positive");
    require(!(origin.x == lookat.x && origin.y == lookat.y && origin.z == lookat.z), "Camera::new: origin == lookat");
    ppc_ = (uint32_t)std::round((double)ppi * 2.54);
    Vec3 z = unit(lookat - origin);
    Vec3 x = unit(cross(up, z));
    Vec3 y = unit(cross(z, x));
    origin_ = origin;
    e_x_ = x;
    e_y_ = y;
    width_ = width;
    height_ = height;
    const double pi = 3.14159265358979323846264338327950288;
    double rad = fov * (pi / 180.0);
    z_ = (width / std::tan(rad / 2.)) * z;  // lib.rs:131 (sic: width, not width/2)
}
size_t Camera::x_pixels() const { return (size_t)std::round(width_ * (double)ppc_); }
size_t Camera::y_pixels() const { return (size_t)std::round(height_ * (double)ppc_); }
RrsCamera Camera::derived() const {
    RrsCamera c{};
    set3(c.origin, origin_);
    set3(c.e_x, e_x_);
    set3(c.e_y, e_y_);
    set3(c.z_scaled, z_);
    c.width = width_;
    c.height = height_;
    c.ppc = ppc_;
    c.x_pixels = (uint32_t)x_pixels();
    c.y_pixels = (uint32_t)y_pixels();
    return c;
}

// ---------------------------------------------------------------------------------------
// Scene (lib.rs:227-245)
// ---------------------------------------------------------------------------------------
Scene::Scene(ObjectSpan objects, double z_near, double z_far, BvhHeuristic heuristic, const Image& hdri, int device, bool with_f64,
             bool upload) {
    SceneOptions opt;
    opt.devices = {device};
    opt.with_f64 = with_f64;
    opt.upload = upload;
    init(objects, z_near, z_far, heuristic, hdri, opt);
}

Scene::Scene(ObjectSpan objects, double z_near, double z_far, BvhHeuristic heuristic, const Image& hdri, const SceneOptions& opt) {
    init(objects, z_near, z_far, heuristic, hdri, opt);
}

void Scene::init(ObjectSpan objects, double z_near, double z_far, BvhHeuristic heuristic, const Image& hdri, const SceneOptions& opt) {
    require(z_near >= 0., "Scene::new: z_near must be >= 0");
    require(z_far > z_near, "Scene::new: z_far must be > z_near");
    require(!opt.devices.empty(), "Scene::new: no device");
    auto t0 = std::chrono::steady_clock::now();
    bvh_ = Bvh::build(heuristic, objects, 1023, opt.bvh_threads, &timing_, opt.device_build ? opt.devices[0] : -1, opt.topology);
    build_seconds_ = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

    // materials / emissions: objects carry them by value (mat.clone() in lib.rs:407-415); identical ones share one table
    // entry, numbered by first appearance in primitive (DFS leaf) order.  Three passes so that the two over all objects
    // run on every host thread: (1) per chunk of objects, the chunk's distinct materials / emissions and each object's
    // index into them; (2) the chunk tables merged into canonical ones and renumbered by first appearance along
    // prim_order; (3) the primitive table filled (a random gather over the object list).
    const size_t n = objects.size();
    constexpr size_t kChunk = 1 << 16;
    const size_t n_chunks = (n + kChunk - 1) / kChunk;
    struct ChunkTables {
        std::vector<RrsMaterial> mats;
        std::vector<RrsEmission> emis;
        std::vector<uint32_t> mat_canon, emi_canon;
    };
    std::vector<ChunkTables> chunk(n_chunks);
    std::unique_ptr<uint32_t[]> local_mat(new uint32_t[n]);
    std::unique_ptr<int32_t[]> local_emi(new int32_t[n]);
    auto find_or_add = [](auto& table, const auto& rec) -> uint32_t {
        for (size_t i = 0; i < table.size(); ++i)
            if (std::memcmp(&table[i], &rec, sizeof(rec)) == 0) return (uint32_t)i;
        table.push_back(rec);
        return (uint32_t)table.size() - 1;
    };
    parallel_chunks(n_chunks, 1, [&](size_t c_lo, size_t c_hi) {
        for (size_t c = c_lo; c < c_hi; ++c) {
            ChunkTables& ct = chunk[c];
            uint32_t last = 0;
            const RrsMaterial* last_ptr = nullptr;  // meshes: consecutive objects usually share the material
            for (size_t i = c * kChunk, e = std::min(n, (c + 1) * kChunk); i < e; ++i) {
                const Object& o = objects[i];
                if (!last_ptr || std::memcmp(last_ptr, &o.mat.m, sizeof(RrsMaterial)) != 0) {
                    last = find_or_add(ct.mats, o.mat.m);
                    last_ptr = &o.mat.m;
                }
                local_mat[i] = last;
                local_emi[i] = o.emission.dark ? -1
                                               : (int32_t)find_or_add(ct.emis, RrsEmission{o.emission.strength,
                                                                                           {o.emission.color.x, o.emission.color.y, o.emission.color.z}});
            }
        }
    });
    std::vector<RrsMaterial> canon_mats;
    std::vector<RrsEmission> canon_emis;
    for (ChunkTables& ct : chunk) {
        for (const RrsMaterial& m : ct.mats) ct.mat_canon.push_back(find_or_add(canon_mats, m));
        for (const RrsEmission& e : ct.emis) ct.emi_canon.push_back(find_or_add(canon_emis, e));
    }
    constexpr uint32_t kUnseen = 0xFFFFFFFFu;
    std::vector<uint32_t> mat_final(canon_mats.size(), kUnseen), emi_final(canon_emis.size(), kUnseen);
    size_t unseen = canon_mats.size() + canon_emis.size();
    for (size_t k = 0; k < n && unseen; ++k) {
        const uint32_t obj = bvh_.prim_order[k];
        const ChunkTables& ct = chunk[obj / kChunk];
        const uint32_t cm = ct.mat_canon[local_mat[obj]];
        if (mat_final[cm] == kUnseen) {
            mat_final[cm] = (uint32_t)materials_.size();
            materials_.push_back(canon_mats[cm]);
            --unseen;
        }
        if (local_emi[obj] >= 0) {
            const uint32_t ce = ct.emi_canon[local_emi[obj]];
            if (emi_final[ce] == kUnseen) {
                emi_final[ce] = (uint32_t)emissions_.size();
                emissions_.push_back(canon_emis[ce]);
                --unseen;
            }
        }
    }
    prims_.resize(n);
    parallel_chunks(n, kChunk, [&](size_t lo, size_t hi) {
        for (size_t k = lo; k < hi; ++k) {
            const uint32_t obj = bvh_.prim_order[k];
            const Object& o = objects[obj];
            const ChunkTables& ct = chunk[obj / kChunk];
            RrsPrim& p = prims_[k];
            p.type = o.geom.type;
            p.obj_id = obj;
            p.material = mat_final[ct.mat_canon[local_mat[obj]]];
            p.emission = local_emi[obj] < 0 ? -1 : (int32_t)emi_final[ct.emi_canon[local_emi[obj]]];
            std::memcpy(p.v, o.geom.v, sizeof(p.v));
        }
    });
    if (!opt.upload) return;
    std::vector<float> rgb(hdri.pixels.size() * 3);
    for (size_t i = 0; i < hdri.pixels.size(); ++i) {
        rgb[3 * i] = (float)hdri.pixels[i].x;
        rgb[3 * i + 1] = (float)hdri.pixels[i].y;
        rgb[3 * i + 2] = (float)hdri.pixels[i].z;
    }
    RrsSceneDesc d{};
    d.abi_version = RRS_ABI_VERSION;
    d.n_prims = (uint32_t)prims_.size();
    d.prims = prims_.data();
    d.n_nodes = (uint32_t)bvh_.nodes.size();
    d.nodes = bvh_.nodes.data();
    d.nodes_f64 = opt.with_f64 ? bvh_.nodes_f64.data() : nullptr;
    d.max_depth = bvh_.max_depth;
    d.n_materials = (uint32_t)materials_.size();
    d.materials = materials_.data();
    d.n_emissions = (uint32_t)emissions_.size();
    d.emissions = emissions_.data();
    d.hdri_width = (uint32_t)hdri.width;
    d.hdri_height = (uint32_t)hdri.height;
    d.hdri_rgb = rgb.data();
    d.t_min = z_near;
    d.t_max = z_far;
    d.flags = opt.flags;
    d.refill_lanes = opt.refill_lanes;
    devices_ = opt.devices;
    handles_.assign(devices_.size(), nullptr);
    int rc = rrs_scene_create_multi(&d, devices_.data(), (int)devices_.size(), handles_.data());
    if (rc != RRS_OK) {
        handles_.clear();
        throw Panic(std::string("rrs_scene_create failed: ") + rrs_last_error());
    }
}

Scene::~Scene() {
    if (comm_) rrs_comm_destroy(comm_);
    for (RrsScene* h : handles_)
        if (h) rrs_scene_destroy(h);
}

RrsComm* Scene::comm() const {
    if (!comm_) {
        int rc = rrs_comm_init_all(devices_.data(), (int)devices_.size(), &comm_);
        if (rc != RRS_OK) throw Panic(std::string("rrs_comm_init_all failed: ") + rrs_last_error());
    }
    return comm_;
}

void render_gpu_into(const Camera& c, const Scene& s, uint32_t spp, uint32_t max_bounces, const RenderOptions& opt,
                     float* out_rgb) {
    if (!s.handle()) throw Panic("render_gpu: scene has no device half");
    RrsCamera cam = c.derived();
    RrsRenderParams p{};
    p.width = cam.x_pixels;
    p.height = cam.y_pixels;
    p.spp = spp;
    p.sample_offset = opt.sample_offset;
    p.spp_total = opt.spp_total;
    p.max_bounces = max_bounces;
    p.seed = opt.seed;
    p.queue_capacity = opt.queue_capacity;
    p.flags = opt.flags;
    if (s.handles().size() > 1) {
        // the scene lives on several GPUs: sample split + one reduce, still one call (rrs_render_multi)
        int rc = rrs_render_multi(s.handles().data(), (int)s.handles().size(), s.comm(), &cam, &p, out_rgb, 0, nullptr);
        if (rc != RRS_OK) throw Panic(std::string("rrs_render_multi failed: ") + rrs_last_error());
        return;
    }
    int rc = rrs_render(s.handle(), &cam, &p, out_rgb);
    if (rc != RRS_OK) throw Panic(std::string("rrs_render failed: ") + rrs_last_error());
}

Image render_gpu(const Camera& c, const Scene& s, uint32_t spp, uint32_t max_bounces, const RenderOptions& opt) {
    size_t w = c.x_pixels(), h = c.y_pixels();
    std::vector<float> buf(w * h * 3);
    render_gpu_into(c, s, spp, max_bounces, opt, buf.data());
    std::vector<Vec3> px(w * h);
    for (size_t i = 0; i < w * h; ++i) px[i] = Vec3(buf[3 * i], buf[3 * i + 1], buf[3 * i + 2]);
    return Image::from_pixels(w, h, std::move(px));  // image.rs:167-173
}

}  // namespace rayrs
