// Material::evaluate with a caller's pdf (rrs_material_evaluate_pdf) — the reference's dormant next-event-estimation
// hook.  `Material::evaluate(position, normal, view, pdf: Option<Pdf>)` (material.rs:91-109) hands the pdf to every
// `Bsdf::scatter`; the one arm that looks at it is `LambertianDiffuse::scatter` (material.rs:259-281), reached
// directly or as Plastic's diffuse lobe (material.rs:588): it samples and weights the lobe with
// `Pdf::Mix(MixKind::Constant(0.5), pdf, Pdf::Cosine)` (generate :1028-1034, value :951-959).  The pdf a light-sampling
// caller passes is `Pdf::Hittable(&geometry)` (value :943-950, generate :1027) over `Hittable::{intersect, area, sample}`
// (geometry.rs:106-152,229-299,359-387).
//
// `radiance()` passes `None` (lib.rs:532), so NO render reaches this code — neither the reference's nor ours — and
// k_wavefront carries none of it.  Until a render does, the hook is evaluated with the reference's own arithmetic: f64,
// operation for operation, no FMA contraction (this TU is compiled with --fmad=false like verify_f64.cu), from the
// caller's f64 primitive record, so that it can be held to the oracle at rounding level; the arms that ignore the pdf
// run the production fp32 `material_evaluate` and return what rrs_material_evaluate returns.  Quirks of the reference
// are kept as written: `Sphere::sample` is not uniform on the sphere (its own FIXME), `Triangle::sample` returns the
// origin, and `Pdf::Hittable::value` divides by the cosine at the SHADED point, not at the light.
#include <cmath>
#include <vector>

#include "nee_f64.cuh"
#include "shading.cuh"
#include "wavefront.cuh"

namespace rrs {

// pnv: n x 9 (position, unit normal, unit view); u: n x 4 draws in call order; out: n x 7 [scatter flag, color rgb, direction xyz]
__global__ void k_material_evaluate_pdf(DMat m, RrsPrim light, const double* __restrict__ pnv, const double* __restrict__ u,
                                        uint32_t n, double* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* q = pnv + 9 * (size_t)i;
    const double* uu = u + 4 * (size_t)i;
    double* o = out + 7 * (size_t)i;
    const uint32_t tag = __float_as_uint(m.m0.w);
    const float3 nf = f3((float)q[3], (float)q[4], (float)q[5]);
    const float3 vf = f3((float)q[6], (float)q[7], (float)q[8]);
    // which arm runs: the diffuse lobe looks at the pdf, everything else ignores it.  Plastic picks its lobe with the
    // first draw against the Schlick term (material.rs:575-579) exactly as the production shading does (shading.cuh).
    bool diffuse = tag == RRS_MAT_LAMBERTIAN;
    int first = 0;
    if (tag == RRS_MAT_PLASTIC) {
        diffuse = !((float)uu[0] < schlick_scalar(m.m2.z, dot3(nf, vf)));
        first = 1;
    }
    if (!diffuse) {
        ScatterOut so = material_evaluate<true>(m, nf, vf, (float)uu[0], (float)uu[1], (float)uu[2]);
        o[0] = so.scatter ? 1. : 0.;
        o[1] = so.color.x; o[2] = so.color.y; o[3] = so.color.z;
        o[4] = so.dir.x; o[5] = so.dir.y; o[6] = so.dir.z;
        return;
    }
    // LambertianDiffuse::scatter with Some(pdf): f64, nee_f64.cuh (the albedo is the DMat's fp32 value)
    lambert_scatter_pdf64((double)m.m0.x, (double)m.m0.y, (double)m.m0.z, light, d3(q[0], q[1], q[2]), d3(q[3], q[4], q[5]),
                          uu[first], uu[first + 1], uu[first + 2], o);
}

int nee_material_evaluate_pdf(SceneImpl* s, uint32_t material, const RrsPrim* light, const double* pnv, const double* u,
                              size_t n, double* out, std::string& err) {
    RRS_CUDA_CHECK(cudaSetDevice(s->device), err);
    if (material >= s->d.n_mats) { err = "material index out of range"; return RRS_ERR_INVALID; }
    if (light->type != RRS_SPHERE && light->type != RRS_PLANE && light->type != RRS_TRIANGLE) {
        err = "light: unknown primitive type";
        return RRS_ERR_INVALID;
    }
    if (n == 0) return RRS_OK;
    if (n > 0xFFFFFFFFull) { err = "batch too large"; return RRS_ERR_INVALID; }
    DMat m;
    RRS_CUDA_CHECK(cudaMemcpy(&m, s->mats + material, sizeof(DMat), cudaMemcpyDeviceToHost), err);
    DevBuf<double> d_q, d_u, d_out;
    RRS_CUDA_CHECK(d_q.alloc(9 * n), err);
    RRS_CUDA_CHECK(d_u.alloc(4 * n), err);
    RRS_CUDA_CHECK(d_out.alloc(7 * n), err);
    RRS_CUDA_CHECK(cudaMemcpy(d_q, pnv, sizeof(double) * 9 * n, cudaMemcpyHostToDevice), err);
    RRS_CUDA_CHECK(cudaMemcpy(d_u, u, sizeof(double) * 4 * n, cudaMemcpyHostToDevice), err);
    k_material_evaluate_pdf<<<(unsigned)((n + 127) / 128), 128>>>(m, *light, d_q, d_u, (uint32_t)n, d_out);
    RRS_CUDA_CHECK(cudaGetLastError(), err);
    RRS_CUDA_CHECK(cudaMemcpy(out, d_out, sizeof(double) * 7 * n, cudaMemcpyDeviceToHost), err);
    return RRS_OK;
}

}  // namespace rrs
