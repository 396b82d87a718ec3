/* A compiled host over the C ABI, no Python and no C++ host mirror in between: what the Rust shim of rust/gpu.rs
 * does, written in C99.
 *
 *   gcc -std=c99 -O2 -Iinclude examples/render_c.c -Lrayrs_b200 -lrayrs_b200 -lm -Wl,-rpath,$PWD/rayrs_b200 -o render_c
 *   ./render_c SCENE hdri.f32 HDRI_W HDRI_H  W H SPP  NGPUS out.f32  [flat.bin  [rays.f64 NRAYS hits.bin]]
 *
 * SCENE  single  `diffuse_single_sphere` (rayrs-lib/src/test_scenes.rs:14-44,60-63) flattened BY HAND: two primitives,
 *                the virtual root node, two materials.
 *        row7    `cook_torrance_spheres_metallic` (test_scenes.rs:178-224: floor + seven spheres, 8 objects > one leaf
 *                group) through a pointer-free restatement of the reference's own tree build (BvhTree::build_midpoint,
 *                bvh.rs:318-389, over index ranges) and the flattening algorithm of INTEGRATION.md section 2 written
 *                in C: DFS primitive order, binary `Node`s numbered breadth-first behind a virtual root, leaf groups
 *                as leaf runs, bare `LeafNode` children flagged and boxed by their primitive, zero-extent nodes
 *                dropped, outward f32 rounding, max_depth.  So a host that is NOT this repo's exercises every
 *                RrsNode rule.
 *        X.bin   any object list through the same build + flattener: u32 n_objects, n_materials | n_objects x RrsPrim
 *                (obj_id ignored: the position in the list) | n_materials x RrsMaterial | 7 doubles: camera origin,
 *                lookat, fov (up = +y).  The tests use it for trees with bare LeafNode children and dead nodes.
 * NGPUS  1: rrs_scene_create + rrs_render.  >1: rrs_scene_create_multi + rrs_comm_init_all + rrs_render_multi
 *        (samples split over the GPUs, one NCCL reduce inside the library) — still one render call.
 * flat.bin   (optional) the flattened arrays, for comparison with the host mirror's:
 *            u32 n_prims, n_nodes, max_depth | n_prims x RrsPrim | n_nodes x RrsNode | n_nodes x RrsNodeF64
 * rays.f64   (optional) NRAYS x 6 doubles (origin, direction) -> hits.bin: NRAYS x i32 object ids, NRAYS x f64 t
 *            (rrs_intersect, precision 32)
 * hdri.f32: HDRI_H x HDRI_W x 3 float32 (already clipped, rayrs/src/main.rs:43).  tests/test_c_example.py drives it.
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

/* ---------------------------------------------------------------------------------------------------------------
 * The host's side of a scene: objects as the reference constructs them, and what Scene::new hands to the GPU.
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct { double lo[3], hi[3]; } Box; /* AxisAlignedBoundingBox, geometry.rs:394-456 */

typedef struct {
    RrsPrim prim; /* type / material / emission / payload exactly as RrsPrim carries them; obj_id = index in the object list */
} HostObject;

/* Hittable::bbox: sphere geometry.rs:687-696 (sqrt of the stored radius^2), plane :699-718 (zero extent on its axis) */
static Box object_bbox(const RrsPrim* p) {
    Box b;
    if (p->type == RRS_SPHERE) {
        double r = sqrt(p->v[0]);
        for (int k = 0; k < 3; ++k) { b.lo[k] = p->v[1 + k] - r; b.hi[k] = p->v[1 + k] + r; }
    } else if (p->type == RRS_PLANE) {
        int axis = ((int)p->v[0]) >> 1, u = axis == 0 ? 1 : 0, v = axis == 2 ? 1 : 2;
        b.lo[axis] = b.hi[axis] = p->v[5];
        b.lo[u] = p->v[1]; b.hi[u] = p->v[2];
        b.lo[v] = p->v[3]; b.hi[v] = p->v[4];
    } else {
        for (int k = 0; k < 3; ++k) {
            b.lo[k] = fmin(p->v[k], fmin(p->v[3 + k], p->v[6 + k]));
            b.hi[k] = fmax(p->v[k], fmax(p->v[3 + k], p->v[6 + k]));
        }
    }
    return b;
}
static Box box_expand(Box a, Box b) {
    for (int k = 0; k < 3; ++k) { a.lo[k] = fmin(a.lo[k], b.lo[k]); a.hi[k] = fmax(a.hi[k], b.hi[k]); }
    return a;
}
/* a box with zero extent on an axis can never pass the reference's slab test (geometry.rs:474,491,508: tmax <= tmin) */
static int box_degenerate(const Box* b) { return !(b->hi[0] > b->lo[0] && b->hi[1] > b->lo[1] && b->hi[2] > b->lo[2]); }

/* The reference tree (bvh.rs:216-224) over ranges of one index array: kind 0 = Node with two children,
 * 1 = Node holding <= 4 LeafNodes (a leaf group), 2 = bare LeafNode (the 1-object side of a split). */
typedef struct { Box box; int child[2]; unsigned first, count; int kind; } TreeNode;
typedef struct {
    const Box* boxes; /* per object */
    unsigned* order;  /* object indices; after the build: DFS leaf order */
    TreeNode* nodes;
    int n_nodes;
} Build;

static Box range_box(const Build* b, unsigned lo, unsigned hi) { /* from_object_list: fold left to right */
    Box r = b->boxes[b->order[lo]];
    for (unsigned i = lo + 1; i < hi; ++i) r = box_expand(r, b->boxes[b->order[i]]);
    return r;
}
static double centre(const Box* b, int k) { return (b->lo[k] + b->hi[k]) / 2.0; } /* geometry.rs:577-582 */

/* BvhTree::build_midpoint, bvh.rs:318-389 */
static int build_midpoint(Build* b, unsigned lo, unsigned hi) {
    TreeNode nd;
    memset(&nd, 0, sizeof nd);
    nd.child[0] = nd.child[1] = -1;
    unsigned n = hi - lo;
    nd.box = range_box(b, lo, hi);
    if (n <= 4) { /* bvh.rs:379-387 */
        nd.kind = 1; nd.first = lo; nd.count = n;
        b->nodes[b->n_nodes] = nd;
        return b->n_nodes++;
    }
    double ext[3] = {nd.box.hi[0] - nd.box.lo[0], nd.box.hi[1] - nd.box.lo[1], nd.box.hi[2] - nd.box.lo[2]};
    int axis = (ext[0] >= ext[1] && ext[0] >= ext[2]) ? 0 : (ext[1] >= ext[2] ? 1 : 2);
    /* BvhData::sort (bvh.rs:97-136): Rust's sort_by is stable -> insertion sort on the centre along `axis` */
    for (unsigned i = lo + 1; i < hi; ++i) {
        unsigned o = b->order[i];
        double key = centre(&b->boxes[o], axis);
        unsigned j = i;
        while (j > lo && centre(&b->boxes[b->order[j - 1]], axis) > key) { b->order[j] = b->order[j - 1]; --j; }
        b->order[j] = o;
    }
    /* split_index (bvh.rs:138-143, split_ind :7-13): first object whose centre is > the box centre, or None */
    double split = centre(&nd.box, axis);
    long ind = -1;
    for (unsigned i = lo; i < hi; ++i)
        if (centre(&b->boxes[b->order[i]], axis) > split) { ind = (long)(i - lo); break; }
    if (ind < 0 || ind == 0 || ind == (long)n - 1) ind = (long)(n / 2); /* bvh.rs:350-358 */
    unsigned mid = lo + (unsigned)ind;
    nd.kind = 0;
    int self = b->n_nodes++; /* reserve: children are appended behind */
    for (int side = 0; side < 2; ++side) {
        unsigned a = side ? mid : lo, e = side ? hi : mid;
        if (e - a > 1) {
            nd.child[side] = build_midpoint(b, a, e);
        } else { /* BvhTree::LeafNode(obj): no box of its own */
            TreeNode leaf;
            memset(&leaf, 0, sizeof leaf);
            leaf.kind = 2; leaf.first = a; leaf.count = 1; leaf.box = b->boxes[b->order[a]];
            leaf.child[0] = leaf.child[1] = -1;
            b->nodes[b->n_nodes] = leaf;
            nd.child[side] = b->n_nodes++;
        }
    }
    b->nodes[self] = nd;
    return self;
}

static float round_down(double x) { float f = (float)x; if ((double)f > x) f = nextafterf(f, -INFINITY); return f; }
static float round_up(double x) { float f = (float)x; if ((double)f < x) f = nextafterf(f, INFINITY); return f; }

/* one child slot of an RrsNode (+ its f64 twin): `cull` is the box the fp32 traversal tests, `exact` the reference's */
static void set_child(RrsNode* n, RrsNodeF64* d, int ch, uint32_t ref, int bare, const Box* exact, const Box* cull) {
    float *lo = ch ? n->lo1 : n->lo0, *hi = ch ? n->hi1 : n->hi0;
    double *dlo = ch ? d->lo1 : d->lo0, *dhi = ch ? d->hi1 : d->hi0;
    if (ch) { n->ref1 = ref; d->ref1 = ref; } else { n->ref0 = ref; d->ref0 = ref; }
    if (bare) { n->flags |= 1u << ch; d->flags |= 1u << ch; }
    for (int k = 0; k < 3; ++k) {
        if (ref == RRS_REF_EMPTY) { /* an empty child carries the inverted infinite box */
            lo[k] = INFINITY; hi[k] = -INFINITY; dlo[k] = INFINITY; dhi[k] = -INFINITY;
        } else { /* fp32 boxes are rounded OUTWARD from the reference's f64 boxes */
            lo[k] = round_down(cull->lo[k]); hi[k] = round_up(cull->hi[k]);
            dlo[k] = exact->lo[k]; dhi[k] = exact->hi[k];
        }
    }
}

/* reference of tree node `t` as seen from its parent, whose box is `parent` */
static void attach(RrsNode* n, RrsNodeF64* d, int ch, const TreeNode* tree, int t, const Box* parent, const int* flat_index) {
    const TreeNode* c = &tree[t];
    if (c->kind == 2) { /* bare LeafNode: the reference tests it whenever the parent is entered; cull with its own box */
        set_child(n, d, ch, RRS_MAKE_LEAF(c->first, 1), 1, &c->box, box_degenerate(&c->box) ? parent : &c->box);
    } else if (box_degenerate(&c->box)) { /* a Node the reference can never enter: dead subtree */
        set_child(n, d, ch, RRS_REF_EMPTY, 0, NULL, NULL);
    } else if (c->kind == 1) {
        set_child(n, d, ch, RRS_MAKE_LEAF(c->first, c->count), 0, &c->box, &c->box);
    } else {
        set_child(n, d, ch, (uint32_t)flat_index[t], 0, &c->box, &c->box);
    }
}

typedef struct {
    RrsPrim* prims; unsigned n_prims;
    RrsNode* nodes; RrsNodeF64* nodes64; unsigned n_nodes, max_depth;
} Flat;

/* Scene::new's BVH half: build the reference tree, then flatten it for the GPU. */
static Flat build_and_flatten(const HostObject* objs, unsigned n) {
    Box* boxes = (Box*)malloc(sizeof(Box) * n);
    unsigned* order = (unsigned*)malloc(sizeof(unsigned) * n);
    for (unsigned i = 0; i < n; ++i) { boxes[i] = object_bbox(&objs[i].prim); order[i] = i; }
    Build b;
    b.boxes = boxes; b.order = order; b.n_nodes = 0;
    b.nodes = (TreeNode*)malloc(sizeof(TreeNode) * (2 * n + 2));
    int root = build_midpoint(&b, 0, n);
    Flat f;
    memset(&f, 0, sizeof f);
    /* primitives in DFS leaf order (= the order array: ranges are split in place): index = tie-break priority */
    f.n_prims = n;
    f.prims = (RrsPrim*)malloc(sizeof(RrsPrim) * n);
    for (unsigned k = 0; k < n; ++k) { f.prims[k] = objs[order[k]].prim; f.prims[k].obj_id = order[k]; }
    /* flat numbering: 0 = virtual root, then the binary Nodes with a real box, breadth-first */
    int* flat_index = (int*)malloc(sizeof(int) * (size_t)b.n_nodes);
    int* flat_to_tree = (int*)malloc(sizeof(int) * (size_t)(b.n_nodes + 1));
    for (int i = 0; i < b.n_nodes; ++i) flat_index[i] = -1;
    unsigned count = 1, head = 1;
    flat_to_tree[0] = -1;
    if (b.nodes[root].kind == 0 && !box_degenerate(&b.nodes[root].box)) { flat_index[root] = 1; flat_to_tree[count++] = root; }
    while (head < count) {
        const TreeNode* t = &b.nodes[flat_to_tree[head++]];
        for (int c = 0; c < 2; ++c) {
            const TreeNode* ch = &b.nodes[t->child[c]];
            if (ch->kind == 0 && !box_degenerate(&ch->box)) { flat_index[t->child[c]] = (int)count; flat_to_tree[count++] = t->child[c]; }
        }
    }
    f.n_nodes = count;
    f.nodes = (RrsNode*)calloc(count, sizeof(RrsNode));
    f.nodes64 = (RrsNodeF64*)calloc(count, sizeof(RrsNodeF64));
    /* virtual root: child 0 = the reference root with its own box (its slab test is the root's), child 1 empty */
    {
        const TreeNode* r = &b.nodes[root];
        if (box_degenerate(&r->box)) set_child(&f.nodes[0], &f.nodes64[0], 0, RRS_REF_EMPTY, 0, NULL, NULL);
        else if (r->kind == 1) set_child(&f.nodes[0], &f.nodes64[0], 0, RRS_MAKE_LEAF(r->first, r->count), 0, &r->box, &r->box);
        else set_child(&f.nodes[0], &f.nodes64[0], 0, (uint32_t)flat_index[root], 0, &r->box, &r->box);
        set_child(&f.nodes[0], &f.nodes64[0], 1, RRS_REF_EMPTY, 0, NULL, NULL);
    }
    for (unsigned k = 1; k < count; ++k) {
        const TreeNode* t = &b.nodes[flat_to_tree[k]];
        attach(&f.nodes[k], &f.nodes64[k], 0, b.nodes, t->child[0], &t->box, flat_index);
        attach(&f.nodes[k], &f.nodes64[k], 1, b.nodes, t->child[1], &t->box, flat_index);
    }
    /* max_depth: the longest chain of RrsNodes from node 0 (parents precede children in breadth-first numbering) */
    unsigned* depth = (unsigned*)calloc(count, sizeof(unsigned));
    depth[0] = 1;
    f.max_depth = 1;
    for (unsigned k = 0; k < count; ++k) {
        const uint32_t refs[2] = {f.nodes[k].ref0, f.nodes[k].ref1};
        for (int c = 0; c < 2; ++c)
            if (refs[c] != RRS_REF_EMPTY && !(refs[c] & RRS_REF_LEAF)) {
                depth[refs[c]] = depth[k] + 1;
                if (depth[refs[c]] > f.max_depth) f.max_depth = depth[refs[c]];
            }
    }
    free(depth); free(flat_index); free(flat_to_tree); free(b.nodes); free(boxes); free(order);
    return f;
}

static RrsMaterial metal(double alpha, double r0) { /* Material::CookTorrance(.., Fresnel::SchlickMetallic(r0)) */
    RrsMaterial m;
    memset(&m, 0, sizeof m);
    m.tag = RRS_MAT_COOK_TORRANCE;
    m.fresnel_kind = RRS_FRESNEL_METALLIC;
    m.color[0] = m.color[1] = m.color[2] = 1.0;
    m.spec_color[0] = m.spec_color[1] = m.spec_color[2] = r0;
    m.alpha = alpha;
    return m;
}
static HostObject floor_plane(unsigned material) { /* test_scenes.rs:15-21 */
    HostObject o;
    memset(&o, 0, sizeof o);
    o.prim.type = RRS_PLANE; o.prim.material = material; o.prim.emission = -1;
    o.prim.v[0] = (double)RRS_AXIS_Y;
    o.prim.v[1] = -25.0; o.prim.v[2] = 25.0; o.prim.v[3] = -25.0; o.prim.v[4] = 25.0; o.prim.v[5] = 0.0;
    return o;
}
static HostObject sphere(double radius, V3 c, unsigned material) {
    HostObject o;
    memset(&o, 0, sizeof o);
    o.prim.type = RRS_SPHERE; o.prim.material = material; o.prim.emission = -1;
    o.prim.v[0] = radius * radius; /* Sphere stores radius^2, geometry.rs:98-101 */
    o.prim.v[1] = c.x; o.prim.v[2] = c.y; o.prim.v[3] = c.z;
    return o;
}

int main(int argc, char** argv) {
    if (argc != 10 && argc != 11 && argc != 14) {
        fprintf(stderr, "usage: %s single|row7|scene.bin hdri.f32 HDRI_W HDRI_H W H SPP NGPUS out.f32 [flat.bin [rays.f64 NRAYS hits.bin]]\n", argv[0]);
        return 2;
    }
    const int row7 = strcmp(argv[1], "row7") == 0;
    const size_t name_len = strlen(argv[1]);
    const int from_file = name_len > 4 && strcmp(argv[1] + name_len - 4, ".bin") == 0;
    const unsigned hw = (unsigned)atoi(argv[3]), hh = (unsigned)atoi(argv[4]);
    const unsigned W = (unsigned)atoi(argv[5]), H = (unsigned)atoi(argv[6]), spp = (unsigned)atoi(argv[7]);
    const int ngpus = atoi(argv[8]);
    if (ngpus < 1 || ngpus > 16) { fprintf(stderr, "NGPUS must be within 1..16\n"); return 2; }
    float* hdri = (float*)malloc(sizeof(float) * 3u * hw * hh);
    FILE* f = fopen(argv[2], "rb");
    if (!f || fread(hdri, sizeof(float), 3u * (size_t)hw * hh, f) != 3u * (size_t)hw * hh) {
        fprintf(stderr, "cannot read %s\n", argv[2]);
        return 2;
    }
    fclose(f);

    RrsMaterial mats[8];
    RrsMaterial* file_mats = NULL;
    RrsSceneDesc desc;
    memset(&desc, 0, sizeof desc);
    Flat flat;
    memset(&flat, 0, sizeof flat);
    RrsPrim prims2[2];
    RrsNode node;
    RrsNodeF64 node64;
    RrsCamera cam;
    if (from_file) {
        /* ---- a caller's object list: tree build + flattener ---- */
        uint32_t head[2];
        double camf[7];
        f = fopen(argv[1], "rb");
        if (!f || fread(head, sizeof head, 1, f) != 1 || head[0] == 0 || head[1] == 0) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
        HostObject* objs = (HostObject*)malloc(sizeof(HostObject) * head[0]);
        file_mats = (RrsMaterial*)malloc(sizeof(RrsMaterial) * head[1]);
        for (uint32_t i = 0; i < head[0]; ++i)
            if (fread(&objs[i].prim, sizeof(RrsPrim), 1, f) != 1) { fprintf(stderr, "short object table\n"); return 2; }
        if (fread(file_mats, sizeof(RrsMaterial), head[1], f) != head[1] || fread(camf, sizeof camf, 1, f) != 1) { fprintf(stderr, "short scene file\n"); return 2; }
        fclose(f);
        flat = build_and_flatten(objs, head[0]);
        free(objs);
        desc.n_prims = flat.n_prims; desc.prims = flat.prims;
        desc.n_nodes = flat.n_nodes; desc.nodes = flat.nodes; desc.nodes_f64 = flat.nodes64;
        desc.max_depth = flat.max_depth;
        desc.n_materials = head[1];
        cam = camera_new(v3(camf[0], camf[1], camf[2]), v3(0, 1, 0), v3(camf[3], camf[4], camf[5]), camf[6], (double)W / 254.0, (double)H / 254.0, 100);
    } else if (!row7) {
        /* ---- diffuse_single_sphere, flattened by hand ---- */
        mats[0] = metal(0.5, 0.8);  /* floor: test_scenes.rs:15-21 */
        memset(&mats[1], 0, sizeof mats[1]);
        mats[1].tag = RRS_MAT_LAMBERTIAN; /* sphere: test_scenes.rs:60-63 */
        mats[1].color[0] = mats[1].color[1] = mats[1].color[2] = 0.8;
        /* DFS leaf order of the reference tree: two objects -> one leaf group [floor, sphere] */
        prims2[0] = floor_plane(0).prim; prims2[0].obj_id = 0;
        prims2[1] = sphere(1.0, v3(0, 1, 0), 1).prim; prims2[1].obj_id = 1;
        /* node 0 = virtual root: child 0 is the reference root (a leaf group of 2 with the union box of
         * plane [-25,25] x {0} x [-25,25] and sphere [-1,1] x [0,2] x [-1,1]), child 1 empty (inverted box) */
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
        desc.n_prims = 2; desc.prims = prims2;
        desc.n_nodes = 1; desc.nodes = &node; desc.nodes_f64 = &node64;
        desc.max_depth = 1;
        desc.n_materials = 2;
        /* film of W x H pixels at 100 ppi (ppc = 254), camera of test_scenes.rs:26-34 */
        cam = camera_new(v3(0, 5, 10), v3(0, 1, 0), v3(0, 1, 0), 50.0, (double)W / 254.0, (double)H / 254.0, 100);
    } else {
        /* ---- cook_torrance_spheres_metallic (test_scenes.rs:178-224): floor + 7 spheres at x = 2.2 (i - 3) with
         * alpha = 0.01 (4 i + 1), through the tree build and the flattener above ---- */
        HostObject objs[8];
        mats[0] = metal(0.5, 0.8);
        objs[0] = floor_plane(0);
        for (int i = 0; i < 7; ++i) {
            mats[1 + i] = metal(0.01 * (double)(4 * i + 1), 0.8);
            objs[1 + i] = sphere(1.0, v3(2.2 * (double)(i - 3), 1.0, 0.0), (unsigned)(1 + i));
        }
        flat = build_and_flatten(objs, 8);
        desc.n_prims = flat.n_prims; desc.prims = flat.prims;
        desc.n_nodes = flat.n_nodes; desc.nodes = flat.nodes; desc.nodes_f64 = flat.nodes64;
        desc.max_depth = flat.max_depth;
        desc.n_materials = 8;
        cam = camera_new(v3(0, 10, 20), v3(0, 1, 0), v3(0, 1, 0), 72.0, (double)W / 254.0, (double)H / 254.0, 100); /* :194-202 */
    }
    desc.abi_version = RRS_ABI_VERSION;
    desc.materials = from_file ? file_mats : mats;
    desc.n_emissions = 0; desc.emissions = NULL;
    desc.hdri_width = hw; desc.hdri_height = hh; desc.hdri_rgb = hdri;
    desc.t_min = 1e-6; desc.t_max = 1e6; /* rayrs/src/main.rs:52 */

    if (argc >= 11) { /* the flattened arrays, for comparison with another host's */
        f = fopen(argv[10], "wb");
        uint32_t head[3] = {desc.n_prims, desc.n_nodes, desc.max_depth};
        if (!f || fwrite(head, sizeof head, 1, f) != 1 || fwrite(desc.prims, sizeof(RrsPrim), desc.n_prims, f) != desc.n_prims ||
            fwrite(desc.nodes, sizeof(RrsNode), desc.n_nodes, f) != desc.n_nodes ||
            fwrite(desc.nodes_f64, sizeof(RrsNodeF64), desc.n_nodes, f) != desc.n_nodes) {
            fprintf(stderr, "cannot write %s\n", argv[10]);
            return 2;
        }
        fclose(f);
    }

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

    RrsScene* scenes[16];
    RrsComm* comm = NULL;
    int devices[16];
    for (int i = 0; i < ngpus; ++i) { devices[i] = i; scenes[i] = NULL; }
    if (ngpus == 1) {
        if (rrs_scene_create(&desc, 0, &scenes[0]) != RRS_OK) return fail("rrs_scene_create");
        if (rrs_render(scenes[0], &cam, &p, rgb) != RRS_OK) return fail("rrs_render");
    } else {
        /* the scene is converted once and uploaded to every GPU; the samples are split inside the one render call */
        if (rrs_scene_create_multi(&desc, devices, ngpus, scenes) != RRS_OK) return fail("rrs_scene_create_multi");
        if (rrs_comm_init_all(devices, ngpus, &comm) != RRS_OK) return fail("rrs_comm_init_all");
        if (rrs_render_multi(scenes, ngpus, comm, &cam, &p, rgb, 0, NULL) != RRS_OK) return fail("rrs_render_multi");
    }
    unsigned long long rays = 0;
    RrsStats st;
    for (int i = ngpus - 1; i >= 0; --i) { /* ends on device 0, whose statistics carry the resolve's census */
        if (rrs_stats(scenes[i], &st) != RRS_OK) return fail("rrs_stats");
        rays += st.rays;
    }
    double mean = 0.0;
    for (size_t i = 0; i < 3u * (size_t)W * H; ++i) mean += rgb[i];
    mean /= 3.0 * (double)W * H;
    printf("render_c: %s on %d GPU(s), %ux%u, %u spp: %llu rays, %.3f ms on device 0, mean radiance %.6f, nan %llu negative %llu census %llu\n",
           argv[1], ngpus, W, H, spp, rays, st.device_ms, mean, (unsigned long long)st.nan_pixels,
           (unsigned long long)st.negative_pixels, (unsigned long long)st.census_mismatch_pixels);

    /* closest hit through the parity entry: the ray down the optical axis meets the middle sphere */
    RrsRay ray;
    put3(ray.origin, v3(cam.origin[0], cam.origin[1], cam.origin[2]));
    put3(ray.direction, v3(-cam.origin[0], 1.0 - cam.origin[1], -cam.origin[2]));
    int32_t id = -2;
    double t = 0.0;
    if (rrs_intersect(scenes[0], &ray, 1, &id, &t, 32) != RRS_OK) return fail("rrs_intersect");
    printf("render_c: axis ray hits object %d at t = %.9f\n", id, t);

    if (argc == 14) { /* a caller's ray set through rrs_intersect */
        const size_t n = (size_t)atol(argv[12]);
        RrsRay* rays_in = (RrsRay*)malloc(sizeof(RrsRay) * (n ? n : 1));
        int32_t* ids = (int32_t*)malloc(sizeof(int32_t) * (n ? n : 1));
        double* ts = (double*)malloc(sizeof(double) * (n ? n : 1));
        f = fopen(argv[11], "rb");
        if (!f || fread(rays_in, sizeof(RrsRay), n, f) != n) { fprintf(stderr, "cannot read %s\n", argv[11]); return 2; }
        fclose(f);
        if (rrs_intersect(scenes[0], rays_in, n, ids, ts, 32) != RRS_OK) return fail("rrs_intersect");
        f = fopen(argv[13], "wb");
        if (!f || fwrite(ids, sizeof(int32_t), n, f) != n || fwrite(ts, sizeof(double), n, f) != n) { fprintf(stderr, "cannot write %s\n", argv[13]); return 2; }
        fclose(f);
        free(rays_in); free(ids); free(ts);
    }

    f = fopen(argv[9], "wb");
    if (!f || fwrite(rgb, sizeof(float), 3u * (size_t)W * H, f) != 3u * (size_t)W * H) {
        fprintf(stderr, "cannot write %s\n", argv[9]);
        return 2;
    }
    fclose(f);
    if (comm) rrs_comm_destroy(comm);
    for (int i = 0; i < ngpus; ++i) rrs_scene_destroy(scenes[i]);
    free(rgb); free(hdri);
    free(flat.prims); free(flat.nodes); free(flat.nodes64); free(file_mats);
    return 0;
}
