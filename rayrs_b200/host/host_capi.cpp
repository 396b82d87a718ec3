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
#include <cstring>
#include <string>

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
                    double tmin, double tmax, int device, int with_f64, int upload) {
    try {
        std::vector<Material> mt;
        for (uint64_t i = 0; i < n_mat; ++i) mt.push_back(material_from_row(mats + 12 * i));
        std::vector<Emission> em;
        for (uint64_t i = 0; i < n_emis; ++i)
            em.push_back(Emission::Emissive(emis[4 * i], Vec3(emis[4 * i + 1], emis[4 * i + 2], emis[4 * i + 3])));
        std::vector<Object> objects;
        objects.reserve(n_obj);
        for (uint64_t i = 0; i < n_obj; ++i) {
            const double* r = objs + 12 * i;
            int mi = (int)r[1], ei = (int)r[2];
            if (mi < 0 || (uint64_t)mi >= n_mat) throw Panic("object material index out of range");
            if (ei >= (int)n_emis) throw Panic("object emission index out of range");
            Emission e = ei < 0 ? Emission::Dark() : em[ei];
            switch ((int)r[0]) {
                case 0: objects.push_back(Object::sphere(r[3], Vec3(r[4], r[5], r[6]), mt[mi], e)); break;
                case 1: objects.push_back(Object::plane((Axis)(int)r[3], r[4], r[5], r[6], r[7], r[8], mt[mi], e)); break;
                case 2:
                    objects.push_back(Object::triangle(Vec3(r[3], r[4], r[5]), Vec3(r[6], r[7], r[8]),
                                                       Vec3(r[9], r[10], r[11]), mt[mi], e));
                    break;
                default: throw Panic("unknown object type");
            }
        }
        Image img;
        img.width = hw;
        img.height = hh;
        img.pixels.resize(hw * hh);
        for (uint64_t i = 0; i < hw * hh; ++i) img.pixels[i] = Vec3(hdri[3 * i], hdri[3 * i + 1], hdri[3 * i + 2]);
        BvhHeuristic h = heuristic == 0 ? BvhHeuristic::Midpoint() : BvhHeuristic::Sah(splits);
        return new Scene(objects, tmin, tmax, h, img, device, with_f64 != 0, upload != 0);
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}

void rrh_scene_free(void* s) { delete static_cast<Scene*>(s); }
RrsScene* rrh_scene_handle(void* s) { return static_cast<Scene*>(s)->handle(); }

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

}  // extern "C"
