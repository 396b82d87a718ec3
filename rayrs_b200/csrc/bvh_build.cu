// BvhTree::build_sah / build_midpoint (rayrs-lib/src/bvh.rs:227-389) on the device, producing THE SAME TREE as the
// reference's recursive build (and as the host mirror's, rayrs_b200/host/rayrs_host.cpp, which the tests compare it
// with node for node): same boxes bit for bit, same split indices, same primitive order.
//
// The recursion becomes a level-synchronous sweep.  Every node of the current level that holds more than 4 objects
// is split at once:
//   sort      BvhData::sort (bvh.rs:97-136): stable sort of the node's objects by the centre of their box along the
//             node's longest axis.  One pass over ALL n positions per level: a stable radix sort by the f64 key
//             (finished ranges carry key 0 and keep their order), then a stable radix sort by the rank of the range a
//             position belongs to, which puts every range back in its place (ranges are contiguous and ordered).
//   sweep     calculate_sah (bvh.rs:15-38) for every split the reference's threshold loop can reach: prefix / suffix
//             box unions over each range (segmented scans), surface areas, the cost 0.3 + (p_left k + p_right (n-k)),
//             all in f64 with the reference's operation order (this file is compiled with -fmad=false).  A split index
//             k is REACHED when some threshold min + i * split_dist (i = 1..splits, bvh.rs:259-271) separates centre
//             k-1 from centre k; the loop keeps the first strict minimum, i.e. the smallest k of the smallest cost.
//   split     bvh.rs:279-287: None, 0 or len-1 -> len / 2; sides of one object become bare LeafNodes, sides of <= 4
//             objects leaf groups, the others the next level's nodes.  The children's boxes are the prefix / suffix
//             unions at the split (min / max are exact, so the fold order of from_object_list cannot show).
// CUB (shipped with the CUDA toolkit) supplies the device-wide radix sort, scan-by-key and reduce-by-key; everything
// specific to the reference's build is in the kernels below.  Scene setup, not the render hot path: it exists
// because the host build dominated the end-to-end time of the mesh configurations (3.5 s of a 4.6 s setup against a
// 1.2 s render for configuration 5 on eight GPUs).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_reduce.cuh>
#include <cub/device/device_scan.cuh>
#include <thrust/iterator/reverse_iterator.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "wavefront.cuh"

namespace rrs {
int api_fail(int code, const std::string& msg);  // api.cu
}
using namespace rrs;

namespace {

struct Box6 {  // xmin, xmax, ymin, ymax, zmin, zmax — the field order of the host mirror's AxisAlignedBoundingBox
    double v[6];
};

__host__ __device__ inline Box6 box_union(const Box6& a, const Box6& b) {  // AxisAlignedBoundingBox::expand geometry.rs:660-668
    Box6 r;
    r.v[0] = fmin(a.v[0], b.v[0]);
    r.v[1] = fmax(a.v[1], b.v[1]);
    r.v[2] = fmin(a.v[2], b.v[2]);
    r.v[3] = fmax(a.v[3], b.v[3]);
    r.v[4] = fmin(a.v[4], b.v[4]);
    r.v[5] = fmax(a.v[5], b.v[5]);
    return r;
}
struct BoxUnion {
    __host__ __device__ Box6 operator()(const Box6& a, const Box6& b) const { return box_union(a, b); }
};
__device__ inline double surface_area(const Box6& b) {  // geometry.rs:640-645, same association
    const double x = b.v[1] - b.v[0], y = b.v[3] - b.v[2], z = b.v[5] - b.v[4];
    return 2. * x * y + 2. * y * z + 2. * x * z;
}
__device__ inline double centre_of(const Box6& b, int axis) {  // geometry.rs:577-582: (max - min) / 2 + min
    return (b.v[2 * axis + 1] - b.v[2 * axis]) / 2. + b.v[2 * axis];
}

// Candidate split of a range: the SAH cost and the split index.  Ordering = the reference's loop: smaller cost wins,
// equal costs keep the first one met, and the loop meets split indices in increasing order.
struct Cand {
    double sah;
    uint32_t k;
    uint32_t pad;
};
struct CandMin {
    __host__ __device__ Cand operator()(const Cand& a, const Cand& b) const {
        return (b.sah < a.sah || (b.sah == a.sah && b.k < a.k)) ? b : a;
    }
};

// One node of the level being split.
struct Active {
    Box6 box;
    uint32_t lo, hi;    // range of positions
    uint32_t out;       // index of its record in the output array
    int32_t axis;       // longest axis (ties: x, then y — bvh.rs:248-257)
    double amin, alen;  // box minimum and extent along it
};

__global__ void k_iota(uint32_t* p, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}

// per level: the node a position belongs to (-1: a finished range)
__global__ void k_mark_ranges(const Active* __restrict__ act, uint32_t n_act, int32_t* __restrict__ node_of, uint32_t* __restrict__ start_flag) {
    const uint32_t a = blockIdx.x;
    if (a >= n_act) return;
    const uint32_t lo = act[a].lo, hi = act[a].hi;
    for (uint32_t p = lo + threadIdx.x; p < hi; p += blockDim.x) node_of[p] = (int32_t)a;
    if (threadIdx.x == 0) start_flag[lo] = 1;
}

// sortable image of an f64 key: unsigned order == numeric order; -0.0 is folded into +0.0 first (the reference compares
// with partial_cmp, for which they are equal, and the stable sort then keeps their order)
__device__ inline unsigned long long sortable(double x) {
    x = x + 0.0;
    unsigned long long u = (unsigned long long)__double_as_longlong(x);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}

__global__ void k_keys(const Box6* __restrict__ boxes, const uint32_t* __restrict__ order, const int32_t* __restrict__ node_of,
                       const Active* __restrict__ act, uint32_t n, unsigned long long* __restrict__ keys) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int32_t a = node_of[p];
    keys[p] = a < 0 ? 0ull : sortable(centre_of(boxes[order[p]], act[a].axis));
}

__global__ void k_gather_u32(const uint32_t* __restrict__ src, const uint32_t* __restrict__ idx, uint32_t n, uint32_t* __restrict__ dst) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}

// after the sort: boxes and keys of the active ranges in sorted order
__global__ void k_gather_sorted(const Box6* __restrict__ boxes, const uint32_t* __restrict__ order, const int32_t* __restrict__ node_of,
                                const Active* __restrict__ act, uint32_t n, Box6* __restrict__ nb, double* __restrict__ key) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const Box6 b = boxes[order[p]];
    nb[p] = b;
    const int32_t a = node_of[p];
    key[p] = a < 0 ? 0. : centre_of(b, act[a].axis);
}

// calculate_sah for the split "left = [lo, p), right = [p, hi)" of position p's node, if the threshold loop reaches it
__global__ void k_candidates(const Active* __restrict__ act, const int32_t* __restrict__ node_of, const double* __restrict__ key,
                             const Box6* __restrict__ pre, const Box6* __restrict__ suf, uint32_t n, uint32_t splits,
                             Cand* __restrict__ cand) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    Cand c;
    c.sah = INFINITY;
    c.k = 0xFFFFFFFFu;
    c.pad = 0;
    const int32_t a = node_of[p];
    if (a >= 0) {
        const Active& nd = act[a];
        const uint32_t k = p - nd.lo, cnt = nd.hi - nd.lo;
        const double split_dist = nd.alen / (double)(splits - 1);  // bvh.rs:259
        const double amin = nd.amin;
        // the loop reaches split k iff some i in 1..=splits has centre[k-1] <= min + i * split_dist < centre[k]
        // (split_index = first centre > threshold, bvh.rs:7-13); thresholds grow with i, so test the smallest i that
        // satisfies the left inequality
        bool reached = false;
        const double right = key[p];
        if (k == 0) {
            reached = (amin + 1. * split_dist) < right;
        } else {
            const double left = key[p - 1];
            if (split_dist > 0. && isfinite(split_dist)) {
                double g = ceil((left - amin) / split_dist);
                long i = g < 1. ? 1 : (g > (double)splits + 1. ? (long)splits + 1 : (long)g);
                // the division only guesses; the loop's own expression settles it (off by one at most)
                for (int it = 0; it < 8 && i > 1 && amin + (double)(i - 1) * split_dist >= left; ++it) --i;
                for (int it = 0; it < 8 && i <= (long)splits && amin + (double)i * split_dist < left; ++it) ++i;
                reached = i <= (long)splits && (amin + (double)i * split_dist) < right;
            } else {
                const double thr = amin + 1. * split_dist;  // every threshold is the same value
                reached = left <= thr && thr < right;
            }
        }
        if (reached) {
            const double sa = surface_area(nd.box);
            const double p_left = k > 0 ? surface_area(pre[p - 1]) / sa : 0.;
            const double p_right = surface_area(suf[p]) / sa;
            const double sah = 0.3 + 1. * (p_left * (double)k + p_right * (double)(cnt - k));
            if (sah == sah) {  // a NaN cost never passes `sah < min_sah`
                c.sah = sah;
                c.k = k;
            }
        }
    }
    cand[p] = c;
}

// One record of the output tree (the host mirror's BNode): kind 0 binary Node, 1 leaf-group Node (<= 4 objects),
// 2 bare LeafNode.
struct OutNode {
    double box[6];
    int32_t child[2];
    uint32_t first, count;
    uint32_t kind, pad;
};
static_assert(sizeof(OutNode) == sizeof(RrsBuildNode), "RrsBuildNode layout");

struct Counters {
    uint32_t n_out;       // records written
    uint32_t n_next;      // nodes of the next level
};

__device__ inline void set_axis(Active& a) {  // bvh.rs:248-257
    const double x = a.box.v[1] - a.box.v[0], y = a.box.v[3] - a.box.v[2], z = a.box.v[5] - a.box.v[4];
    if (x >= y && x >= z) { a.axis = 0; a.amin = a.box.v[0]; a.alen = x; }
    else if (y >= z) { a.axis = 1; a.amin = a.box.v[2]; a.alen = y; }
    else { a.axis = 2; a.amin = a.box.v[4]; a.alen = z; }
}

// split every node of the level: pick the index, emit the children
__global__ void k_split(const Active* __restrict__ act, uint32_t n_act, const uint32_t* __restrict__ rank, const Cand* __restrict__ best,
                        const double* __restrict__ key, const Box6* __restrict__ pre, const Box6* __restrict__ suf,
                        const Box6* __restrict__ nb, int heuristic, OutNode* __restrict__ out, Active* __restrict__ next,
                        Counters* __restrict__ cnt, uint32_t* __restrict__ start_flag) {
    const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n_act) return;
    const Active nd = act[a];
    const uint32_t n = nd.hi - nd.lo;
    long ind = -1;
    if (heuristic == 1) {
        const Cand c = best[rank[nd.lo]];
        if (c.k != 0xFFFFFFFFu && c.sah < INFINITY) ind = (long)c.k;
    } else {
        // build_midpoint bvh.rs:337-350: first centre > centre of the node's box
        const double split = centre_of(nd.box, nd.axis);
        uint32_t lo = 0, hi = n;  // upper_bound over the sorted keys
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (key[nd.lo + mid] <= split) lo = mid + 1;
            else hi = mid;
        }
        ind = lo >= n ? -1 : (long)lo;
    }
    if (ind < 0 || ind == 0 || ind == (long)n - 1) ind = (long)(n / 2);  // bvh.rs:279-287 / 350-358
    const uint32_t mid = nd.lo + (uint32_t)ind;
    start_flag[mid] = 1;
    OutNode rec;
    for (int k = 0; k < 6; ++k) rec.box[k] = nd.box.v[k];
    rec.first = nd.lo;
    rec.count = n;
    rec.kind = 0;
    rec.pad = 0;
    for (int side = 0; side < 2; ++side) {
        const uint32_t lo = side ? mid : nd.lo, hi = side ? nd.hi : mid;
        const uint32_t id = atomicAdd(&cnt->n_out, 1u);
        rec.child[side] = (int32_t)id;
        OutNode ch;
        ch.child[0] = ch.child[1] = -1;
        ch.first = lo;
        ch.count = hi - lo;
        ch.pad = 0;
        const Box6 cb = (hi - lo == 1) ? nb[lo] : (side ? suf[mid] : pre[mid - 1]);
        for (int k = 0; k < 6; ++k) ch.box[k] = cb.v[k];
        if (hi - lo == 1) {
            ch.kind = 2;  // BvhTree::LeafNode
        } else if (hi - lo <= 4) {
            ch.kind = 1;  // a Node of LeafNodes, bvh.rs:304-315
        } else {
            ch.kind = 0;
            Active nx;
            nx.box = cb;
            nx.lo = lo;
            nx.hi = hi;
            nx.out = id;
            set_axis(nx);
            next[atomicAdd(&cnt->n_next, 1u)] = nx;
        }
        out[id] = ch;
    }
    // the node's own record: box / range were written when it was created; fill in the children
    out[nd.out].child[0] = rec.child[0];
    out[nd.out].child[1] = rec.child[1];
}

__global__ void k_root(const Box6* __restrict__ root_box, uint32_t n, OutNode* __restrict__ out, Active* __restrict__ act, Counters* __restrict__ cnt) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    OutNode r;
    for (int k = 0; k < 6; ++k) r.box[k] = root_box->v[k];
    r.child[0] = r.child[1] = -1;
    r.first = 0;
    r.count = n;
    r.kind = n <= 4 ? 1u : 0u;
    r.pad = 0;
    out[0] = r;
    cnt->n_out = 1;
    cnt->n_next = 0;
    if (n > 4) {
        Active a;
        a.box = *root_box;
        a.lo = 0;
        a.hi = n;
        a.out = 0;
        set_axis(a);
        act[0] = a;
        cnt->n_next = 1;
    }
}

// One device allocation carved into the build's arrays (two dozen cudaMallocs of a few hundred MB each cost more than
// the build itself); freed on every return path.
struct Arena {
    char* base = nullptr;
    size_t size = 0, used = 0;
    ~Arena() { if (base) cudaFree(base); }
    template <typename T>
    size_t reserve(size_t count) {  // first pass: sizes only
        size = (size + 255) & ~(size_t)255;
        size_t off = size;
        size += sizeof(T) * std::max<size_t>(count, 1);
        return off;
    }
    template <typename T>
    T* at(size_t off) const { return reinterpret_cast<T*>(base + off); }
};

#define BUILD_CHECK(expr)                                                                              \
    do {                                                                                               \
        cudaError_t _e = (expr);                                                                       \
        if (_e != cudaSuccess) return api_fail(RRS_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

}  // namespace

extern "C" int rrs_bvh_build(const double* boxes, uint32_t n, uint32_t heuristic, uint32_t splits, int device, uint32_t* prim_order,
                             RrsBuildNode* nodes, uint32_t* n_nodes, double* seconds) {
    if (!boxes || !prim_order || !nodes || !n_nodes) return api_fail(RRS_ERR_INVALID, "null argument");
    if (n == 0) return api_fail(RRS_ERR_INVALID, "Having a BVH for 0 objects does not make sense (bvh.rs:229)");
    if (n >= (1u << 28)) return api_fail(RRS_ERR_INVALID, "too many objects (max 2^28-1)");
    if (heuristic > 1) return api_fail(RRS_ERR_INVALID, "heuristic must be 0 (Midpoint) or 1 (Sah)");
    if (heuristic == 1 && splits < 2) return api_fail(RRS_ERR_INVALID, "Sah needs at least 2 splits");
    for (size_t i = 0; i < (size_t)n * 6; ++i)
        if (std::isnan(boxes[i])) return api_fail(RRS_ERR_INVALID, "NaN in a bounding box (bvh.rs:104 partial_cmp().unwrap() panics)");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return api_fail(RRS_ERR_NO_DEVICE, "no CUDA device: rayrs_b200 has no CPU fallback");
    }
    if (device < 0 || device >= ndev) return api_fail(RRS_ERR_INVALID, "device index out of range");
    cudaDeviceProp prop;
    BUILD_CHECK(cudaGetDeviceProperties(&prop, device));
    if (!(prop.major == 10 && prop.minor == 0)) return api_fail(RRS_ERR_NO_DEVICE, "device is not sm_100: kernels are built for sm_100a only");
    BUILD_CHECK(cudaSetDevice(device));

    Arena mem;
    const size_t max_nodes = 2 * (size_t)n + 2;
    const size_t o_boxes = mem.reserve<Box6>(n), o_nb = mem.reserve<Box6>(n), o_pre = mem.reserve<Box6>(n), o_suf = mem.reserve<Box6>(n);
    const size_t o_root = mem.reserve<Box6>(1);
    const size_t o_order = mem.reserve<uint32_t>(n), o_order2 = mem.reserve<uint32_t>(n), o_pos = mem.reserve<uint32_t>(n), o_pos2 = mem.reserve<uint32_t>(n);
    const size_t o_rank = mem.reserve<uint32_t>(n), o_rank_sorted = mem.reserve<uint32_t>(n), o_rank_tmp = mem.reserve<uint32_t>(n);
    const size_t o_unique = mem.reserve<uint32_t>(n), o_nruns = mem.reserve<uint32_t>(1);
    const size_t o_keys = mem.reserve<unsigned long long>(n), o_keys2 = mem.reserve<unsigned long long>(n), o_key = mem.reserve<double>(n);
    const size_t o_node_of = mem.reserve<int32_t>(n), o_flag = mem.reserve<uint32_t>(n);
    const size_t o_cand = mem.reserve<Cand>(n), o_best = mem.reserve<Cand>((size_t)n + 1);
    const size_t o_act = mem.reserve<Active>(n / 5 + 2), o_next = mem.reserve<Active>(n / 5 + 2);
    const size_t o_out = mem.reserve<OutNode>(max_nodes), o_cnt = mem.reserve<Counters>(1);

    // CUB scratch, sized once for the largest call (size queries only read the pointer TYPES)
    size_t tmp_bytes = 0, b = 0;
    typedef thrust::reverse_iterator<const uint32_t*> RevKey;
    typedef thrust::reverse_iterator<const Box6*> RevIn;
    typedef thrust::reverse_iterator<Box6*> RevOut;
    {
        unsigned long long* k64 = nullptr;
        uint32_t* u32 = nullptr;
        Box6* bx = nullptr;
        Cand* cd = nullptr;
        cub::DeviceRadixSort::SortPairs(nullptr, b, k64, k64, u32, u32, n, 0, 64);
        tmp_bytes = std::max(tmp_bytes, b);
        cub::DeviceRadixSort::SortPairs(nullptr, b, u32, u32, u32, u32, n, 0, 32);
        tmp_bytes = std::max(tmp_bytes, b);
        cub::DeviceScan::InclusiveSum(nullptr, b, u32, u32, n);
        tmp_bytes = std::max(tmp_bytes, b);
        cub::DeviceScan::InclusiveScanByKey(nullptr, b, (const uint32_t*)u32, (const Box6*)bx, bx, BoxUnion(), n);
        tmp_bytes = std::max(tmp_bytes, b);
        cub::DeviceScan::InclusiveScanByKey(nullptr, b, RevKey(u32), RevIn(bx), RevOut(bx), BoxUnion(), n);
        tmp_bytes = std::max(tmp_bytes, b);
        cub::DeviceReduce::ReduceByKey(nullptr, b, (const uint32_t*)u32, u32, (const Cand*)cd, cd, u32, CandMin(), n);
        tmp_bytes = std::max(tmp_bytes, b);
    }
    Box6 init_box;
    for (int k = 0; k < 3; ++k) { init_box.v[2 * k] = INFINITY; init_box.v[2 * k + 1] = -INFINITY; }
    cub::DeviceReduce::Reduce(nullptr, b, (const Box6*)nullptr, (Box6*)nullptr, n, BoxUnion(), init_box);
    tmp_bytes = std::max(tmp_bytes, b);
    const size_t o_tmp = mem.reserve<char>(tmp_bytes);
    BUILD_CHECK(cudaMalloc(&mem.base, mem.size));
    Box6 *d_boxes = mem.at<Box6>(o_boxes), *d_nb = mem.at<Box6>(o_nb), *d_pre = mem.at<Box6>(o_pre), *d_suf = mem.at<Box6>(o_suf), *d_root = mem.at<Box6>(o_root);
    uint32_t *d_order = mem.at<uint32_t>(o_order), *d_order2 = mem.at<uint32_t>(o_order2), *d_pos = mem.at<uint32_t>(o_pos), *d_pos2 = mem.at<uint32_t>(o_pos2);
    uint32_t *d_rank = mem.at<uint32_t>(o_rank), *d_rank_sorted = mem.at<uint32_t>(o_rank_sorted), *d_rank_tmp = mem.at<uint32_t>(o_rank_tmp);
    uint32_t *d_unique = mem.at<uint32_t>(o_unique), *d_nruns = mem.at<uint32_t>(o_nruns);
    unsigned long long *d_keys = mem.at<unsigned long long>(o_keys), *d_keys2 = mem.at<unsigned long long>(o_keys2);
    double* d_key = mem.at<double>(o_key);
    int32_t* d_node_of = mem.at<int32_t>(o_node_of);
    uint32_t* d_flag = mem.at<uint32_t>(o_flag);
    Cand *d_cand = mem.at<Cand>(o_cand), *d_best = mem.at<Cand>(o_best);
    Active *d_act = mem.at<Active>(o_act), *d_next = mem.at<Active>(o_next);
    OutNode* d_out = mem.at<OutNode>(o_out);
    Counters* d_cnt = mem.at<Counters>(o_cnt);
    void* d_tmp = mem.at<char>(o_tmp);

    cudaEvent_t ev0, ev1;
    cudaEventCreate(&ev0);
    cudaEventCreate(&ev1);
    BUILD_CHECK(cudaMemcpy(d_boxes, boxes, sizeof(Box6) * n, cudaMemcpyHostToDevice));
    cudaEventRecord(ev0);
    const int T = 256;
    const unsigned G = (n + T - 1) / T;
    k_iota<<<G, T>>>(d_order, n);
    // the root's box: from_object_list over every object (bvh.rs:231)
    b = tmp_bytes;
    BUILD_CHECK(cub::DeviceReduce::Reduce(d_tmp, b, (const Box6*)d_boxes, d_root, n, BoxUnion(), init_box));
    k_root<<<1, 1>>>(d_root, n, d_out, d_act, d_cnt);
    BUILD_CHECK(cudaMemset(d_flag, 0, sizeof(uint32_t) * n));
    Counters h_cnt{};
    BUILD_CHECK(cudaMemcpy(&h_cnt, d_cnt, sizeof(h_cnt), cudaMemcpyDeviceToHost));
    uint32_t n_act = h_cnt.n_next;
    uint32_t levels = 0;
    while (n_act > 0) {
        if (++levels > 4096) return api_fail(RRS_ERR_CUDA, "BVH build did not terminate");
        // which node each position belongs to; where ranges start (finished ranges keep their flags from earlier levels)
        BUILD_CHECK(cudaMemset(d_node_of, 0xFF, sizeof(int32_t) * n));
        k_mark_ranges<<<n_act, 128>>>(d_act, n_act, d_node_of, d_flag);
        b = tmp_bytes;
        BUILD_CHECK(cub::DeviceScan::InclusiveSum(d_tmp, b, d_flag, d_rank, n));  // rank of the range of each position (1-based)
        // ---- BvhData::sort for every node of the level at once ----
        k_keys<<<G, T>>>(d_boxes, d_order, d_node_of, d_act, n, d_keys);
        k_iota<<<G, T>>>(d_pos, n);
        b = tmp_bytes;
        BUILD_CHECK(cub::DeviceRadixSort::SortPairs(d_tmp, b, d_keys, d_keys2, d_pos, d_pos2, n, 0, 64));
        k_gather_u32<<<G, T>>>(d_rank, d_pos2, n, d_rank_tmp);
        b = tmp_bytes;
        BUILD_CHECK(cub::DeviceRadixSort::SortPairs(d_tmp, b, d_rank_tmp, d_rank_sorted, d_pos2, d_pos, n, 0, 32));
        k_gather_u32<<<G, T>>>(d_order, d_pos, n, d_order2);
        std::swap(d_order, d_order2);
        k_gather_sorted<<<G, T>>>(d_boxes, d_order, d_node_of, d_act, n, d_nb, d_key);
        // ---- prefix / suffix boxes of every range ----
        b = tmp_bytes;
        BUILD_CHECK(cub::DeviceScan::InclusiveScanByKey(d_tmp, b, (const uint32_t*)d_rank, (const Box6*)d_nb, d_pre, BoxUnion(), n));
        b = tmp_bytes;
        BUILD_CHECK(cub::DeviceScan::InclusiveScanByKey(d_tmp, b, RevKey(d_rank + n), RevIn(d_nb + n), RevOut(d_suf + n), BoxUnion(), n));
        if (heuristic == 1) {
            k_candidates<<<G, T>>>(d_act, d_node_of, d_key, d_pre, d_suf, n, splits, d_cand);
            b = tmp_bytes;
            // one aggregate per run of equal ranks, in rank order: aggregate r-1 belongs to rank r
            BUILD_CHECK(cub::DeviceReduce::ReduceByKey(d_tmp, b, (const uint32_t*)d_rank, d_unique, (const Cand*)d_cand, d_best + 1, d_nruns,
                                                       CandMin(), n));
        }
        BUILD_CHECK(cudaMemset(&d_cnt->n_next, 0, sizeof(uint32_t)));
        k_split<<<(n_act + 127) / 128, 128>>>(d_act, n_act, d_rank, d_best, d_key, d_pre, d_suf, d_nb, (int)heuristic, d_out, d_next, d_cnt, d_flag);
        BUILD_CHECK(cudaGetLastError());
        BUILD_CHECK(cudaMemcpy(&h_cnt, d_cnt, sizeof(h_cnt), cudaMemcpyDeviceToHost));
        if (h_cnt.n_out > max_nodes) return api_fail(RRS_ERR_CUDA, "BVH build overflowed its node array");
        std::swap(d_act, d_next);
        n_act = h_cnt.n_next;
    }
    cudaEventRecord(ev1);
    BUILD_CHECK(cudaMemcpy(prim_order, d_order, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost));
    BUILD_CHECK(cudaMemcpy(nodes, d_out, sizeof(OutNode) * h_cnt.n_out, cudaMemcpyDeviceToHost));
    *n_nodes = h_cnt.n_out;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev0, ev1);
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    if (seconds) *seconds = (double)ms * 1e-3;
    return RRS_OK;
}
