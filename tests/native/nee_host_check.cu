// TEST INFRASTRUCTURE: the f64 half of the Pdf::Hittable hook (rayrs_b200/csrc/nee_f64.cuh, __host__ __device__) compiled
// for the HOST, so that tests/test_pdf_hook.py can hold the very code k_material_evaluate_pdf runs to the oracle on a
// machine without a GPU.  Built by the test with nvcc (host code: -Xcompiler -ffp-contract=off); not part of the library.
#include <stdint.h>

#include "../../rayrs_b200/csrc/nee_f64.cuh"

extern "C" {

// color3: the albedo; pnv: n x 9 (position, unit normal, unit view — view unused); u3: n x 3 (side of the mix, two draws)
void nee_host_lambert(const double* color3, const RrsPrim* light, const double* pnv, const double* u3, uint64_t n, double* out) {
    for (uint64_t i = 0; i < n; ++i) {
        const double* q = pnv + 9 * i;
        const double* u = u3 + 3 * i;
        rrs::lambert_scatter_pdf64(color3[0], color3[1], color3[2], *light, rrs::d3(q[0], q[1], q[2]), rrs::d3(q[3], q[4], q[5]),
                                   u[0], u[1], u[2], out + 7 * i);
    }
}

// [Hittable::area, Hittable::sample(u2)]
void nee_host_area_sample(const RrsPrim* p, const double* u2, double* out4) {
    rrs::D3 s = rrs::hittable_sample64(*p, u2[0], u2[1]);
    out4[0] = rrs::hittable_area64(*p);
    out4[1] = s.x; out4[2] = s.y; out4[3] = s.z;
}

}
