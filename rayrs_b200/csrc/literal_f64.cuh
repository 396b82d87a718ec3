// f64 device helpers that restate the reference's arithmetic operation for operation (vecmath.rs, geometry.rs) for the
// verification paths: the fp64 traversal (verify_f64.cu) and the Pdf::Hittable hook (nee.cu).  Every TU that includes
// this header is compiled with --fmad=false (Rust never contracts a*b+c) — see EXTRA in rayrs_b200/build.py.
// NOT the production path: the fp32 kernels in intersect.cuh / shading.cuh are what renders.
// The functions are __host__ __device__ so that tests/native/nee_host_check.cu can run the very same code on the CPU.
#pragma once
#include "device_types.cuh"

namespace rrs {

struct D3 {
    double x, y, z;
};
__host__ __device__ __forceinline__ D3 d3(double x, double y, double z) { return D3{x, y, z}; }
__host__ __device__ __forceinline__ D3 sub(D3 a, D3 b) { return d3(a.x - b.x, a.y - b.y, a.z - b.z); }
__host__ __device__ __forceinline__ D3 add(D3 a, D3 b) { return d3(a.x + b.x, a.y + b.y, a.z + b.z); }
__host__ __device__ __forceinline__ D3 mul(D3 a, double s) { return d3(a.x * s, a.y * s, a.z * s); }
// vecmath.rs:533-535 — left to right, no contraction
__host__ __device__ __forceinline__ double dot(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__host__ __device__ __forceinline__ D3 cross(D3 a, D3 b) {
    return d3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

// geometry.rs:458-513
__host__ __device__ __forceinline__ bool aabb64(const double* lo, const double* hi, D3 o, D3 d, double tmin, double tmax) {
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

__host__ __device__ __forceinline__ bool prim64(const RrsPrim& p, D3 o, D3 d, double& t) {
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

}  // namespace rrs
