// extern "C" surface of the host mirror for the Python tests / bench (ctypes).  It only
// marshals flat tables into the C++ API of rayrs_host.hpp; all logic lives there.
//
// Table formats (shared with the test-side oracle so both see the same numbers):
//   objects  n x 12 doubles  [type, material, emission(-1 = Dark), payload x 9]
//       sphere   (0): radius, ox, oy, oz
//       plane    (1): axis (0..5 = X,XRev,Y,YRev,Z,ZRev), umin, umax, vmin, vmax, pos
//       triangle (2): p1 xyz, p2 xyz, p3 xyz
//   materials n x 12 doubles [tag, color rgb, alpha, ior, fresnel kind, r0/spec rgb, -, -]
//   emissions n x 4 doubles  [strength, color rgb]
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <type_traits>
#include <new>
#include <memory>
#include <string>

#include "mesh_io.hpp"
#include "rayrs_host.hpp"

using namespace rayrs;

namespace {

thread_local std::string g_err;

Material material_from_row(const double* r) {
    Vec3 color(r[1], r[2], r[3]);
    Vec3 aux(r[7], r[8], r[9]);
    switch ((int)r[0]) {
        case RRS_MAT_LAMBERTIAN: return Material::LambertianDiffuse(color);
        case RRS_MAT_REFLECT: return Material::Reflect(color);
        case RRS_MAT_REFRACT: return Material::Refract(color, r[5]);
        case RRS_MAT_GLASS: return Material::Glass(color, r[5]);
        case RRS_MAT_COOK_TORRANCE:
            return Material::CookTorrance(color, r[4], (int)r[6] == RRS_FRESNEL_METALLIC ? Fresnel::SchlickMetallic(aux)
                                                                                          : Fresnel::SchlickDielectric(r[5]));
        case RRS_MAT_COOK_TORRANCE_REFRACT: return Material::CookTorranceRefract(color, r[4], r[5]);
        case RRS_MAT_COOK_TORRANCE_GLASS: return Material::CookTorranceGlass(color, r[4], r[5]);
        case RRS_MAT_PLASTIC: return Material::Plastic(color, aux, r[4], r[5]);
        case RRS_MAT_NO_REFLECT: return Material::NoReflect();
        default: throw Panic("unknown material tag");
    }
}

}  // namespace

extern "C" {

const char* rrh_last_error(void) { return g_err.c_str(); }

void* rrh_scene_new(const double* objs, uint64_t n_obj, const double* mats, uint64_t n_mat, const double* emis,
                    uint64_t n_emis, int heuristic, uint32_t splits, const double* hdri, uint64_t hw, uint64_t hh,
                    double tmin, double tmax, int device, int with_f64, int upload, uint32_t scene_flags,
                    uint32_t refill_lanes, int bvh_threads, const int* devices, int n_devices, int device_build, int topology) {
    try {
        std::vector<Material> mt;
        for (uint64_t i = 0; i < n_mat; ++i) mt.push_back(material_from_row(mats + 12 * i));
        std::vector<Emission> em;
        for (uint64_t i = 0; i < n_emis; ++i)
            em.push_back(Emission::Emissive(emis[4 * i], Vec3(emis[4 * i + 1], emis[4 * i + 2], emis[4 * i + 3])));
        // one Object per table row; rows are independent, so large meshes are converted on all host threads — into raw
        // storage, so that the first touch of the 190 B x n buffer is spread over the threads too (Object is trivially
        // destructible: releasing the storage is all the clean-up there is)
        static_assert(std::is_trivially_destructible<Object>::value, "the object buffer is released without destructor calls");
        struct RawFree { void operator()(Object* p) const { ::operator delete(static_cast<void*>(p)); } };
        std::unique_ptr<Object, RawFree> storage(static_cast<Object*>(::operator new(std::max<uint64_t>(n_obj, 1) * sizeof(Object))));
        Object* objects = storage.get();
        std::mutex err_mu;
        std::string err;
        parallel_chunks(n_obj, 1 << 16, [&](size_t lo, size_t hi) {
            try {
                for (size_t i = lo; i < hi; ++i) {
                    const double* r = objs + 12 * i;
                    int mi = (int)r[1], ei = (int)r[2];
                    if (mi < 0 || (uint64_t)mi >= n_mat) throw Panic("object material index out of range");
                    if (ei >= (int)n_emis) throw Panic("object emission index out of range");
                    Emission e = ei < 0 ? Emission::Dark() : em[ei];
                    switch ((int)r[0]) {
                        case 0: new (objects + i) Object(Object::sphere(r[3], Vec3(r[4], r[5], r[6]), mt[mi], e)); break;
                        case 1: new (objects + i) Object(Object::plane((Axis)(int)r[3], r[4], r[5], r[6], r[7], r[8], mt[mi], e)); break;
                        case 2:
                            new (objects + i) Object(Object::triangle(Vec3(r[3], r[4], r[5]), Vec3(r[6], r[7], r[8]), Vec3(r[9], r[10], r[11]), mt[mi], e));
                            break;
                        default: throw Panic("unknown object type");
                    }
                }
            } catch (const std::exception& ex) {
                std::lock_guard<std::mutex> g(err_mu);
                if (err.empty()) err = ex.what();
            }
        });
        if (!err.empty()) throw Panic(err);
        Image img;
        img.width = hw;
        img.height = hh;
        img.pixels.resize(hw * hh);
        for (uint64_t i = 0; i < hw * hh; ++i) img.pixels[i] = Vec3(hdri[3 * i], hdri[3 * i + 1], hdri[3 * i + 2]);
        BvhHeuristic h = heuristic == 0 ? BvhHeuristic::Midpoint() : BvhHeuristic::Sah(splits);
        SceneOptions opt;
        opt.devices = (devices && n_devices > 0) ? std::vector<int>(devices, devices + n_devices) : std::vector<int>{device};
        opt.with_f64 = with_f64 != 0;
        opt.upload = upload != 0;
        opt.flags = scene_flags;
        opt.refill_lanes = refill_lanes;
        opt.bvh_threads = bvh_threads;
        opt.device_build = device_build != 0;
        opt.topology = topology != 0;
        return new Scene(ObjectSpan(objects, n_obj), tmin, tmax, h, img, opt);
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}

void rrh_scene_free(void* s) { delete static_cast<Scene*>(s); }
RrsScene* rrh_scene_handle(void* s) { return static_cast<Scene*>(s)->handle(); }
// the handle on the i-th device the scene was uploaded to (nullptr past the end)
RrsScene* rrh_scene_handle_at(void* s, uint32_t i) {
    const auto& h = static_cast<Scene*>(s)->handles();
    return i < h.size() ? h[i] : nullptr;
}
// the scene's own communicator over its devices (rrs_comm_init_all), made on first use; nullptr on failure
RrsComm* rrh_scene_comm(void* s) {
    try {
        return static_cast<Scene*>(s)->comm();
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}
// [boxes, tree build, numbering, flatten, depth, topology dump, device part of the tree build] seconds of the BVH build
void rrh_scene_build_timing(void* s, double* out7) {
    const BvhBuildTiming& t = static_cast<Scene*>(s)->build_timing();
    out7[0] = t.boxes; out7[1] = t.recursive; out7[2] = t.numbering; out7[3] = t.flatten; out7[4] = t.depth; out7[5] = t.topology;
    out7[6] = t.device;
}

// info: [n_nodes, n_prims, max_depth, dead_nodes, n_materials, topology_len, n_boxes]
void rrh_scene_info(void* s, uint64_t* info7, double* build_seconds) {
    const Scene* sc = static_cast<Scene*>(s);
    info7[0] = sc->bvh().nodes.size();
    info7[1] = sc->prims().size();
    info7[2] = sc->bvh().max_depth;
    info7[3] = sc->bvh().dead_nodes;
    info7[4] = sc->materials().size();
    info7[5] = sc->bvh().topology.size();
    info7[6] = sc->bvh().boxes.size() / 6;
    if (build_seconds) *build_seconds = sc->build_seconds();
}
void rrh_scene_copy(void* s, RrsNode* nodes, RrsNodeF64* nodes_f64, uint32_t* prim_order, int64_t* topology, double* boxes,
                    RrsPrim* prims) {
    const Scene* sc = static_cast<Scene*>(s);
    const FlatBvh& b = sc->bvh();
    if (nodes) std::memcpy(nodes, b.nodes.data(), sizeof(RrsNode) * b.nodes.size());
    if (nodes_f64) std::memcpy(nodes_f64, b.nodes_f64.data(), sizeof(RrsNodeF64) * b.nodes_f64.size());
    if (prim_order) std::memcpy(prim_order, b.prim_order.data(), sizeof(uint32_t) * b.prim_order.size());
    if (topology) std::memcpy(topology, b.topology.data(), sizeof(int64_t) * b.topology.size());
    if (boxes) std::memcpy(boxes, b.boxes.data(), sizeof(double) * b.boxes.size());
    if (prims) std::memcpy(prims, sc->prims().data(), sizeof(RrsPrim) * sc->prims().size());
}

int rrh_camera_new(const double* origin, const double* up, const double* lookat, double fov, double width, double height,
                   uint32_t ppi, RrsCamera* out) {
    try {
        Camera c(Vec3(origin[0], origin[1], origin[2]), Vec3(up[0], up[1], up[2]), Vec3(lookat[0], lookat[1], lookat[2]), fov,
                 width, height, ppi);
        *out = c.derived();
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

// ---- mesh files (mesh_io.hpp) ----
// kind 0 = PLY (load_ply_file), 1 = OBJ (load_obj_file); returns n x 9 doubles (p1, p2, p3), free with rrh_free
double* rrh_load_mesh(const char* path, int kind, uint64_t* n_tris) {
    try {
        std::vector<Triangle> t = kind == 0 ? load_ply_file(path) : load_obj_file(path);
        double* out = static_cast<double*>(std::malloc(sizeof(double) * 9 * std::max<size_t>(t.size(), 1)));
        for (size_t i = 0; i < t.size(); ++i) {
            const Vec3* p[3] = {&t[i].p1, &t[i].p2, &t[i].p3};
            for (int k = 0; k < 3; ++k) {
                out[9 * i + 3 * k] = p[k]->x;
                out[9 * i + 3 * k + 1] = p[k]->y;
                out[9 * i + 3 * k + 2] = p[k]->z;
            }
        }
        *n_tris = t.size();
        return out;
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}
// load_obj_file_spheres: returns n x 3 doubles (centres), free with rrh_free
double* rrh_load_obj_spheres(const char* path, double radius, uint64_t* n) {
    try {
        std::vector<Sphere> sp = load_obj_file_spheres(path, radius);
        double* out = static_cast<double*>(std::malloc(sizeof(double) * 3 * std::max<size_t>(sp.size(), 1)));
        for (size_t i = 0; i < sp.size(); ++i) {
            out[3 * i] = sp[i].origin.x;
            out[3 * i + 1] = sp[i].origin.y;
            out[3 * i + 2] = sp[i].origin.z;
        }
        *n = sp.size();
        return out;
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}
void rrh_free(void* p) { std::free(p); }

// format 0 = ascii, 1 = binary_big_endian, 2 = binary_little_endian
int rrh_write_ply(const char* path, const float* xyz, uint64_t n_vertices, const int32_t* faces, uint64_t n_faces, int format) {
    try {
        ply::write_ply(path, std::vector<float>(xyz, xyz + 3 * n_vertices), std::vector<int32_t>(faces, faces + 3 * n_faces),
                       (ply::PlyFormat)format);
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

// Header of an in-memory PLY as text, one line per item (tests):
//   "format <ascii|binary_big_endian|binary_little_endian> <version>", "comment <line>",
//   "element <name> <length>", "property <type#> <name>", "list <lentype#> <elemtype#> <name>"
int rrh_ply_describe(const char* bytes, uint64_t n, char* out, uint64_t cap) {
    try {
        ply::Ply p = ply::Ply::parse(std::string(bytes, bytes + n));
        static const char* fmt[] = {"ascii", "binary_big_endian", "binary_little_endian"};
        std::string s = std::string("format ") + fmt[(int)p.header.format] + " " + p.header.version + "\n";
        for (const std::string& c : p.header.comments) s += c + "\n";
        for (const ply::PlyElement& e : p.header.elements) {
            s += "element " + e.name + " " + std::to_string(e.length) + "\n";
            for (const ply::PlyProperty& pr : e.properties)
                s += pr.is_list ? "list " + std::to_string((int)pr.lentype) + " " + std::to_string((int)pr.typ) + " " + pr.name + "\n"
                                : "property " + std::to_string((int)pr.typ) + " " + pr.name + "\n";
        }
        if (s.size() + 1 > cap) throw Panic("rrh_ply_describe: buffer too small");
        std::memcpy(out, s.c_str(), s.size() + 1);
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

}  // extern "C"
