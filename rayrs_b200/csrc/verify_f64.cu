// fp64 verification traversal (rrs_intersect precision=64).  NOT the production path: it
// walks the flattened tree in the reference's own visiting order with the reference's own
// arithmetic — full DFS, children left to right, every box tested against the ORIGINAL
// (tmin, tmax), no pruning (bvh.rs:391-415), f64, no FMA contraction (this TU is compiled
// with --fmad=false), IEEE division and square root — so that primitive IDs can be compared
// bit-for-bit with the CPU restatement on EVERY ray.  It proves that flattening preserved
// the reference semantics; the fp32 kernel in intersect.cuh is what renders.
#include <cmath>
#include <vector>

#include "wavefront.cuh"

namespace rrs {

struct D3 {
    double x, y, z;
};
__device__ __forceinline__ D3 d3(double x, double y, double z) { return D3{x, y, z}; }
__device__ __forceinline__ D3 sub(D3 a, D3 b) { return d3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ D3 add(D3 a, D3 b) { return d3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ D3 mul(D3 a, double s) { return d3(a.x * s, a.y * s, a.z * s); }
// vecmath.rs:533-535 — left to right, no contraction
__device__ __forceinline__ double dot(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ D3 cross(D3 a, D3 b) {
    return d3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

// geometry.rs:458-513
__device__ __forceinline__ bool aabb64(const double* lo, const double* hi, D3 o, D3 d, double tmin, double tmax) {
    const double oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        double mx = hi[k] - oo[k], mn = lo[k] - oo[k], inv = 1. / dd[k];
        double t0, t1;
        if (inv < 0.) { t0 = mx * inv; t1 = mn * inv; } else { t0 = mn * inv; t1 = mx * inv; }
        tmin = fmax(tmin, t0);
        tmax = fmin(tmax, t1);
        if (tmax <= tmin) return false;
    }
    return true;
}

__device__ __forceinline__ bool prim64(const RrsPrim& p, D3 o, D3 d, double& t) {
    if (p.type == RRS_SPHERE) {  // geometry.rs:106-132
        D3 od = sub(o, d3(p.v[1], p.v[2], p.v[3]));
        double a = dot(d, d);
        double b = 2. * dot(d, od);
        double c = dot(od, od) - p.v[0];
        double desc = b * b - 4. * a * c;
        if (desc > 0.) {
            double t1 = (-b - sqrt(desc)) / (2. * a);
            double t2 = (-b + sqrt(desc)) / (2. * a);
            if (t1 < 0.) {
                if (t2 < 0.) return false;
                t = t2;
                return true;
            }
            t = t1;
            return true;
        }
        return false;
    }
    if (p.type == RRS_PLANE) {  // geometry.rs:229-271
        int axis = ((int)p.v[0]) >> 1;
        double ok = axis == 0 ? o.x : (axis == 1 ? o.y : o.z);
        double dk = axis == 0 ? d.x : (axis == 1 ? d.y : d.z);
        if (dk != 0.) {
            double tt = (p.v[5] - ok) / dk;
            D3 q = add(o, mul(d, tt));
            double u = axis == 0 ? q.y : q.x;
            double v = axis == 2 ? q.y : q.z;
            if (p.v[1] <= u && u < p.v[2] && p.v[3] <= v && v < p.v[4]) {
                t = tt;
                return true;
            }
        }
        return false;
    }
    // geometry.rs:341-375 (e1, e2 as Triangle::new derives them)
    D3 p1 = d3(p.v[0], p.v[1], p.v[2]);
    D3 e1 = sub(d3(p.v[3], p.v[4], p.v[5]), p1);
    D3 e2 = sub(d3(p.v[6], p.v[7], p.v[8]), p1);
    D3 T = sub(o, p1);
    D3 P = cross(d, e2);
    D3 Q = cross(T, e1);
    double den = dot(P, e1);
    double dist = dot(Q, e2) / den;
    double u = dot(P, T) / den;
    double v = dot(Q, d) / den;
    if (dist < 0. || u < 0. || v < 0. || u + v > 1.) return false;
    t = dist;
    return true;
}

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
