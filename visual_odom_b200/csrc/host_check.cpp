// host_check.cpp -- compiles the host/device math of pnp_math.cuh for the HOST so that the CPU test
// suite (-m "not gpu") can check the exact code the CUDA kernels run against cv2, without a GPU.
// It is a test hook of the product's own math, not a fallback: nothing in the library calls it.
#include "pnp_math.cuh"
extern "C" {
__attribute__((visibility("default"))) void vo_hostcheck_epnp5(const float* X, const float* uv, const float* K9, double* rvec, double* tvec, double* R)
{
    vomath::epnp5(X, uv, (double)K9[0], (double)K9[4], (double)K9[2], (double)K9[5], rvec, tvec, R);
}
__attribute__((visibility("default"))) void vo_hostcheck_triangulate(const float* Pl12, const float* Pr12, const float* a, const float* b, int n, float* X)
{
    double Pl[12], Pr[12];
    for (int k = 0; k < 12; k++) { Pl[k] = Pl12[k]; Pr[k] = Pr12[k]; }
    for (int i = 0; i < n; i++) vomath::triangulate_dlt(Pl, Pr, a[2 * i], a[2 * i + 1], b[2 * i], b[2 * i + 1], X + 3 * i);
}
__attribute__((visibility("default"))) void vo_hostcheck_rodrigues(const double* r, double* R, double* r_back)
{
    vomath::rodrigues_fwd(r, R);
    vomath::rodrigues_inv(R, r_back);
}
}
