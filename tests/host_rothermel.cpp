// Test-only: compiles simfire_b200/csrc/sfb_rothermel.cuh for the HOST so the operation
// order of the device function can be checked against the golden vectors without a GPU.
// Not part of the product library (the product has no CPU path).
#include "../simfire_b200/csrc/sfb_rothermel.cuh"
extern "C" void host_rate_of_spread(const signed char* dir, const float* rec, const float* particle,
                                    long long n, double* out) {
    SfbParticle fp{particle[0], particle[1], particle[2], particle[3], particle[4]};
    for (long long i = 0; i < n; ++i) out[i] = sfb_rate_of_spread_pair(dir[i], rec + 8 * i, fp);
}
