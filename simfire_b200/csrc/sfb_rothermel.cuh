// Rothermel surface rate of spread for one (source -> destination) pair.
//
// Reproduces the arithmetic of simfire/world/rothermel.py:54-134 as it is executed when
// RothermelFireManager.update calls it with the float32 vectors of fire.py:519-548:
//   * every fuel / wind / slope term is a float32 operation, rounded after each step
//     (compile with --fmad=false: NumPy never fuses a multiply with an add);
//   * `sign` (rothermel.py:118) is an int64 array, so phi_s (:119) and the final quotient
//     (:128) are float64;
//   * x**2 is a multiply, x**0.5 a square root, every other ** is a float32 power.
// The float32 transcendental functions are evaluated in float64 and rounded once, i.e.
// they return the correctly rounded float32 value (NumPy's SIMD float32 pow/exp/cos are
// within 1-2 ulp of that; the parity budget is 1e-5 relative).
//
// __host__ __device__ so that tests can compile the identical source with g++ and check
// the operation order against the golden vectors without a GPU.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define SFB_HD __host__ __device__ __forceinline__
#else
#define SFB_HD inline
#endif

struct SfbParticle {
    float h, S_T, S_e, p_p, M_f;  // FuelParticle (parameters.py:8-27), Environment.M_f (:53)
};

// theta = arctan2(y_src - y_dst, x_dst - x_src) in float32 for the neighbour order of
// fire.py:211-221 (rothermel.py:102; image y grows downwards).  West is +pi, not -pi.
SFB_HD float sfb_travel_angle(int dir) {
    switch (dir & 7) {
        case 0: return 0.0f;
        case 1: return -0.7853981256484985f;
        case 2: return -1.5707963705062866f;
        case 3: return -2.356194496154785f;
        case 4: return 3.1415927410125732f;
        case 5: return 2.356194496154785f;
        case 6: return 1.5707963705062866f;
        default: return 0.7853981256484985f;
    }
}

SFB_HD float sfb_powf(float x, float y) { return (float)pow((double)x, (double)y); }
SFB_HD float sfb_expf(float x) { return (float)exp((double)x); }
SFB_HD float sfb_cosf(float x) { return (float)cos((double)x); }
// np.minimum / np.maximum propagate NaN (fminf / fmaxf do not)
SFB_HD float sfb_minf(float a, float b) { return (a != a) ? a : ((b != b) ? b : (a < b ? a : b)); }
SFB_HD float sfb_maxf(float a, float b) { return (a != a) ? a : ((b != b) ? b : (a > b ? a : b)); }

// The part of the evaluation that depends only on the cell's fuel and on the handle-wide
// particle constants (rothermel.py:74-98, :121-123): computed once per cell when the static
// planes are uploaded (k_derive_static) and again, through the same code, by the one-shot
// entry point below.  Every member is a value the reference rounds to float32 at that point,
// so keeping it in a float32 record changes nothing.
struct SfbFuelTerms {
    float IRxi;  // I_R * xi                      (numerator factor of rothermel.py:128)
    float den;   // (p_b * eps) * Q_ig            (denominator of rothermel.py:128)
    float c, b;  // wind coefficients             (rothermel.py:96-97)
    float rpow;  // (B / B_op) ** -e              (rothermel.py:111)
    float sfac;  // 5.275 * B ** -0.3             (rothermel.py:119)
    float w_0;   // <= 0: non-burnable, R = 0     (rothermel.py:54)
    float pad;
};

SFB_HD SfbFuelTerms sfb_fuel_terms(float w_0, float delta, float M_x, float sigma, const SfbParticle& fp) {
    SfbFuelTerms t;
    // fuel-only terms, float32 (rothermel.py:74-98)
    const float eta_S = sfb_minf(0.174f * sfb_powf(fp.S_e, -0.19f), 1.0f);
    const float r_M = sfb_minf(fp.M_f / M_x, 1.0f);
    const float eta_M = ((1.0f - 2.59f * r_M) + 5.11f * (r_M * r_M)) - 3.52f * sfb_powf(r_M, 3.0f);
    const float w_n = w_0 * (1.0f - fp.S_T);
    const float p_b = w_0 / delta;
    const float B = p_b / fp.p_p;
    const float B_op = 3.348f * sfb_powf(sigma, -0.8189f);
    const float s15 = sfb_powf(sigma, 1.5f);
    const float gamma_max = s15 / (495.0f + 0.0594f * s15);
    const float A = 133.0f * sfb_powf(sigma, -0.7913f);
    const float ratio = B / B_op;
    const float gamma = (gamma_max * sfb_powf(ratio, A)) * sfb_expf(A * (1.0f - ratio));
    const float I_R = (((gamma * w_n) * fp.h) * eta_M) * eta_S;
    const float xi = sfb_expf((0.792f + 0.681f * sqrtf(sigma)) * (B + 0.1f)) / (192.0f + 0.2595f * sigma);
    t.c = 7.47f * sfb_expf(-0.133f * sfb_powf(sigma, 0.55f));
    t.b = 0.02526f * sfb_powf(sigma, 0.54f);
    const float e = 0.715f * sfb_expf(-3.59e-4f * sigma);
    t.rpow = sfb_powf(ratio, -e);
    t.sfac = 5.275f * sfb_powf(B, -0.3f);
    // heat sink, float32 (rothermel.py:121-123)
    const float eps = sfb_expf(-138.0f / sigma);
    const float Q_ig = 250.0f + 1116.0f * fp.M_f;
    t.IRxi = I_R * xi;
    t.den = (p_b * eps) * Q_ig;
    t.w_0 = w_0;
    t.pad = 0.0f;
    return t;
}

// The direction-dependent part (rothermel.py:102-119, :128-134) for a pair travelling in
// direction `dir` into a cell with fuel terms `t`, wind (U, U_dir) and slope.
SFB_HD double sfb_spread_from_terms(int dir, const SfbFuelTerms& t, float U, float U_dir, float slope_mag,
                                    float slope_dir) {
    if (!(t.w_0 > 0.0f)) return 0.0;  // rothermel.py:54-71: non-burnable pairs keep R = 0
    const float theta = sfb_travel_angle(dir);
    // wind factor, float32 (rothermel.py:102-111); np.radians(x) = x * (pi_f32 / 180_f32)
    const float deg2rad = 3.14159265358979323846f / 180.0f;
    const float psi = (90.0f - U_dir) * deg2rad;
    const float U_along = sfb_maxf(U * sfb_cosf(psi - theta), 0.0f);
    const float phi_w = (t.c * sfb_powf(U_along, t.b)) * t.rpow;
    // slope factor: float32 until the int64 sign promotes to float64 (rothermel.py:117-119)
    const float s_along = (-slope_mag) * sfb_cosf(slope_dir + theta);
    const double sign = (s_along > 0.0f) ? 1.0 : -1.0;
    const double phi_s = ((double)t.sfac * sign) * (double)(s_along * s_along);
    // rothermel.py:128: ((I_R*xi)[f32] * ((1+phi_w)[f32] + phi_s)[f64]) / ((p_b*eps)*Q_ig)[f32]
    const double num = (double)t.IRxi * ((double)(1.0f + phi_w) + phi_s);
    const double R = num / (double)t.den;
    return (R != R) ? R : (R > 0.0 ? R : 0.0);  // np.maximum(R, 0) keeps NaN
}

// rec = {w_0, delta, M_x, sigma, U, U_dir, slope_mag, slope_dir} of the DESTINATION cell
// (fire.py:481-497).  Returns ft/min, float64, >= 0 (rothermel.py:134).
SFB_HD double sfb_rate_of_spread_pair(int dir, const float* rec, const SfbParticle& fp) {
    if (!(rec[0] > 0.0f)) return 0.0;
    const SfbFuelTerms t = sfb_fuel_terms(rec[0], rec[1], rec[2], rec[3], fp);
    return sfb_spread_from_terms(dir, t, rec[4], rec[5], rec[6], rec[7]);
}
