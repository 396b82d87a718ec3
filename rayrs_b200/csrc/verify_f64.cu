// fp64 verification traversal (rrs_intersect precision=64).  NOT the production path: it
// walks the flattened tree in the reference's own visiting order with the reference's own
// arithmetic — full DFS, children left to right, every box tested against the ORIGINAL
// (tmin, tmax), no pruning (bvh.rs:391-415), f64, no FMA contraction (this TU is compiled
// with --fmad=false), IEEE division and square root — so that primitive IDs can be compared
// bit-for-bit with the CPU restatement on EVERY ray.  It proves that flattening preserved
// the reference semantics; the fp32 kernel in intersect.cuh is what renders.
#include <cmath>
#include <vector>

#include "literal_f64.cuh"
#include "wavefront.cuh"

namespace rrs {

#define VF_STACK 192

__global__ void k_intersect64(const RrsPrim* __restrict__ prims, const RrsNodeF64* __restrict__ nodes,
                              const RrsRay* __restrict__ rays, uint32_t n, double tmin, double tmax,
                              int32_t* __restrict__ obj_id, double* __restrict__ t_out, int* __restrict__ overflow) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    D3 o = d3(rays[i].origin[0], rays[i].origin[1], rays[i].origin[2]);
    D3 d = d3(rays[i].direction[0], rays[i].direction[1], rays[i].direction[2]);
    // stack entry: node << 1 | child
    uint32_t stack[VF_STACK];
    int sp = 0;
    stack[sp++] = 0u;  // virtual root, child 0
    bool have = false;
    double tb = 0.;
    int32_t best = -1;
    while (sp > 0) {
        uint32_t e = stack[--sp];
        const RrsNodeF64& nd = nodes[e >> 1];
        int ch = e & 1;
        uint32_t ref = ch ? nd.ref1 : nd.ref0;
        if (ref == RRS_REF_EMPTY) continue;
        bool bare = (nd.flags >> ch) & 1u;
        if (!bare) {
            const double* lo = ch ? nd.lo1 : nd.lo0;
            const double* hi = ch ? nd.hi1 : nd.hi0;
            if (!aabb64(lo, hi, o, d, tmin, tmax)) continue;
        }
        if (ref & RRS_REF_LEAF) {
            uint32_t first = ref & 0x0FFFFFFFu, count = ((ref >> 28) & 7u) + 1u;
            for (uint32_t k = 0; k < count; ++k) {
                double t;
                if (prim64(prims[first + k], o, d, t) && t > tmin && t < tmax) {  // bvh.rs:404-413
                    // RayIntersection::update bvh.rs:50-72
                    if (!have || (t > tmin && t < tb)) {
                        have = true;
                        tb = t;
                        best = (int32_t)prims[first + k].obj_id;
                    }
                }
            }
        } else {
            if (sp + 2 > VF_STACK) { *overflow = 1; break; }
            stack[sp++] = (ref << 1) | 1u;  // right child later
            stack[sp++] = (ref << 1);       // left child first
        }
    }
    obj_id[i] = have ? best : -1;
    t_out[i] = have ? tb : INFINITY;
}

int vf_intersect64(SceneImpl* s, const RrsRay* rays, size_t n, int32_t* obj_id, double* t, std::string& err) {
    RRS_CUDA_CHECK(cudaSetDevice(s->device), err);
    if (!s->nodes_f64 || !s->prims_f64) { err = "scene was created without f64 nodes"; return RRS_ERR_INVALID; }
    if (n == 0) return RRS_OK;
    DevBuf<RrsRay> d_r;
    DevBuf<int32_t> d_id;
    DevBuf<double> d_t;
    DevBuf<int> d_ovf;
    RRS_CUDA_CHECK(d_r.alloc(n), err);
    RRS_CUDA_CHECK(d_id.alloc(n), err);
    RRS_CUDA_CHECK(d_t.alloc(n), err);
    RRS_CUDA_CHECK(d_ovf.alloc(1), err);
    RRS_CUDA_CHECK(cudaMemset(d_ovf, 0, sizeof(int)), err);
    RRS_CUDA_CHECK(cudaMemcpy(d_r, rays, sizeof(RrsRay) * n, cudaMemcpyHostToDevice), err);
    k_intersect64<<<(unsigned)((n + 63) / 64), 64>>>(s->prims_f64, s->nodes_f64, d_r, (uint32_t)n, s->tmin, s->tmax, d_id,
                                                     d_t, d_ovf);
    RRS_CUDA_CHECK(cudaGetLastError(), err);
    int ovf = 0;
    RRS_CUDA_CHECK(cudaMemcpy(&ovf, d_ovf, sizeof(int), cudaMemcpyDeviceToHost), err);
    RRS_CUDA_CHECK(cudaMemcpy(obj_id, d_id, sizeof(int32_t) * n, cudaMemcpyDeviceToHost), err);
    RRS_CUDA_CHECK(cudaMemcpy(t, d_t, sizeof(double) * n, cudaMemcpyDeviceToHost), err);
    if (ovf) { err = "fp64 verification stack overflow"; return RRS_ERR_TOO_DEEP; }
    return RRS_OK;
}

}  // namespace rrs
