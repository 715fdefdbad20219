// Host build of deplex_b200/csrc/cr_math.cuh for tests/test_cr_math_cpu.py (the header is __host__ __device__).
#include "../../deplex_b200/csrc/cr_math.cuh"

extern "C" {
void crm_sincos(const double* a, double* s, double* c, long n) {
  for (long i = 0; i < n; ++i) dpx::crm::sincos_cr(a[i], s[i], c[i]);
}
// a0 = a starting value a few ulp off the true atan2, as a GPU libm may return
void crm_atan2(const double* y, const double* x, const double* a0, double* out, long n) {
  for (long i = 0; i < n; ++i) out[i] = dpx::crm::atan2_cr(y[i], x[i], a0[i]);
}
}
