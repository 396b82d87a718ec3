// The f64 half of the Pdf::Hittable hook (nee.cu): Hittable::{area, sample}, Pdf::Hittable::{value, generate}, Pdf::Cosine
// generate and LambertianDiffuse::scatter under Pdf::Mix — the reference's arithmetic operation for operation.
// __host__ __device__: the kernel in nee.cu and the CPU harness tests/native/nee_host_check.cu run the same code.
#pragma once
#include <cmath>

#include "literal_f64.cuh"

namespace rrs {

#define NEE_PI 3.14159265358979323846
#define NEE_FRAC_1_PI 0.318309886183790671537767526745028724

__host__ __device__ __forceinline__ double mag64(D3 a) { return sqrt(dot(a, a)); }
// vecmath.rs:525-527 with Div<f64> = multiplication by the reciprocal (vecmath.rs:690-698)
__host__ __device__ __forceinline__ D3 unit64(D3 a) { return mul(a, 1. / mag64(a)); }

// Vec3::orthonormal_basis vecmath.rs:341-352
__host__ __device__ __forceinline__ void orthonormal_basis64(D3 n, D3& e1, D3& e2) {
    if (fabs(n.x) > fabs(n.y))
        e1 = unit64(d3(n.z, 0., -n.x));
    else
        e1 = unit64(d3(0., n.z, -n.y));
    e2 = unit64(cross(n, e1));
}

// Pdf::Cosine generate material.rs:982-993
__host__ __device__ __forceinline__ D3 cosine_generate64(D3 n, double u, double uphi) {
    D3 e1, e2;
    orthonormal_basis64(n, e1, e2);
    double phi = 2. * NEE_PI * uphi;
    double x = cos(phi) * sqrt(u);
    double y = sin(phi) * sqrt(u);
    double z = sqrt(1. - u);
    return add(add(mul(e1, x), mul(e2, y)), mul(n, z));
}

// Hittable::area geometry.rs:138-140 (sphere), 284-286 (plane), 381-383 with Triangle::new :341-355 (triangle)
__host__ __device__ __forceinline__ double hittable_area64(const RrsPrim& p) {
    if (p.type == RRS_SPHERE) return 4. * NEE_PI * p.v[0];
    if (p.type == RRS_PLANE) return (p.v[2] - p.v[1]) * (p.v[4] - p.v[3]);
    D3 p1 = d3(p.v[0], p.v[1], p.v[2]);
    D3 e1 = sub(d3(p.v[3], p.v[4], p.v[5]), p1);
    D3 e2 = sub(d3(p.v[6], p.v[7], p.v[8]), p1);
    return mag64(cross(e1, e2)) / 2.;
}

// Hittable::sample geometry.rs:142-152 (sphere), 288-299 (plane), 385-387 (triangle: the reference's stub)
__host__ __device__ __forceinline__ D3 hittable_sample64(const RrsPrim& p, double ua, double ub) {
    if (p.type == RRS_SPHERE) {
        double u = ua;
        double phi = 2. * NEE_PI * ub;
        double x = cos(phi) * 2. * sqrt(u * (1. - u));
        double y = sin(phi) * 2. * sqrt(u * (1. - u));
        double z = 1. - 2. * u;
        return add(mul(d3(x, y, z), sqrt(p.v[0])), d3(p.v[1], p.v[2], p.v[3]));
    }
    if (p.type == RRS_PLANE) {
        double u = ua * (p.v[2] - p.v[1]) + p.v[1];
        double v = ub * (p.v[4] - p.v[3]) + p.v[3];
        int axis = ((int)p.v[0]) >> 1;
        return axis == 0 ? d3(p.v[5], u, v) : (axis == 1 ? d3(u, p.v[5], v) : d3(u, v, p.v[5]));
    }
    return d3(0., 0., 0.);
}

// Pdf::Hittable value material.rs:943-950
__host__ __device__ __forceinline__ double pdf_hittable_value64(const RrsPrim& g, D3 position, D3 n, D3 l) {
    double t;
    if (prim64(g, position, l, t)) {
        D3 d = sub(position, add(position, mul(l, t)));
        return dot(d, d) / (dot(n, l) * hittable_area64(g));
    }
    return 0.;
}

// LambertianDiffuse::scatter with pdf = Some(Pdf::Hittable(light)) (material.rs:259-281): the lobe is sampled from and
// weighted with Pdf::Mix(MixKind::Constant(0.5), Pdf::Hittable(light), Pdf::Cosine).  ua picks the side of the mix
// (generate :1028-1034), ub / uc are the two draws of the chosen generator.  out7 = [1, color rgb, direction xyz].
__host__ __device__ __forceinline__ void lambert_scatter_pdf64(double cr, double cg, double cb, const RrsPrim& light, D3 position,
                                                               D3 nrm, double ua, double ub, double uc, double* out7) {
    const double factor = 0.5;  // MixKind::Constant(0.5).value material.rs:1529-1534
    D3 l;
    if (ua < factor)  // -> Pdf::Hittable generate :1027
        l = unit64(sub(hittable_sample64(light, ub, uc), position));
    else
        l = cosine_generate64(nrm, ub, uc);
    const double nl = dot(nrm, l);
    double pdfv;
    if (nl < 0.)  // Pdf::Mix value :952-954
        pdfv = INFINITY;
    else
        pdfv = factor * pdf_hittable_value64(light, position, nrm, l) + (1. - factor) * (nl * NEE_FRAC_1_PI);
    const double inv = 1. / pdfv;  // Vec3 / f64 multiplies by the reciprocal (vecmath.rs:690-698)
    out7[0] = 1.;
    out7[1] = cr * NEE_FRAC_1_PI * nl * inv;  // brdf = color / pi (material.rs:1233-1243); brdf * cos / pdf
    out7[2] = cg * NEE_FRAC_1_PI * nl * inv;
    out7[3] = cb * NEE_FRAC_1_PI * nl * inv;
    out7[4] = l.x; out7[5] = l.y; out7[6] = l.z;
}

}  // namespace rrs
