// Per-entry covariance functions shared by the tiled builder (cov.cu) and the fused small-problem kernel
// (small.cu): value and every d/dtheta of one covariance entry, written with the reference's operation order.
#pragma once
#include "common.cuh"

namespace gpx_cov {

constexpr double PI_D = 3.141592653589793238462643383279502884;  // == np.pi

struct CovParams {
    int kind, ntheta, D;
    double th[11];
};

template <int KIND>
struct NTheta { static constexpr int value = KIND == GPX_COV_SE ? 2 : KIND == GPX_COV_LIN ? 1 : KIND == GPX_COV_PER ? 2 : 11; };

// value and derivatives of one covariance entry.  `acc` is the squared distance (or, for LIN, the
// centred dot product); `aux` is sum_d (a_d + b_d) for LIN's derivative; diag = (same_x && i == j).
template <int KIND, bool WITH_DK>
__device__ __forceinline__ double cov_eval(const CovParams& p, double acc, double aux, bool diag, double* dk) {
    if (KIND == GPX_COV_SE) {
        const double sigma = p.th[0], l = p.th[1];
        const double e = exp(-.5 * (1.0 / (l * l)) * acc);  // GP_regression.py:19
        if (WITH_DK) {
            dk[0] = 2.0 * sigma * e;                          // tune...:48
            dk[1] = (sigma * sigma) * e * (acc / (l * l * l));  // tune...:54
        }
        return (sigma * sigma) * e;
    } else if (KIND == GPX_COV_LIN) {
        if (WITH_DK) dk[0] = -(aux - 2.0 * p.D * p.th[0]);
        return acc;                                           // GP_regression.py:32
    } else if (KIND == GPX_COV_PER) {
        const double per = p.th[0], l = p.th[1];
        const double r = sqrt(acc);
        const double sn = sin(PI_D * r / per);
        const double k = exp(-2.0 * (sn * sn) / (l * l));     // GP_regression.py:49
        if (WITH_DK) {
            dk[0] = k * (2.0 * PI_D * r / (per * per * l * l)) * sin(2.0 * PI_D * r / per);
            dk[1] = k * 4.0 * (sn * sn) / (l * l * l);
        }
        return k;
    } else {  // CO2 composite, CO2_example.py:9-94
        const double* t = p.th;
        const double d = acc;
        const double r = sqrt(d);
        const double e1 = exp(-.5 * d / (t[1] * t[1]));                       // :17
        const double sn = sin(PI_D * r);
        const double q = sn / t[4];
        const double e2 = exp(-.5 * d / (t[3] * t[3]) + -2.0 * (q * q));      // :30-32
        const double u = 1.0 + .5 * d / (t[7] * (t[6] * t[6]));               // :44
        const double pw = 1.0 / pow(u, t[7]);                                 // :45
        const double e4 = exp(-.5 * d / (t[9] * t[9]));                       // :65
        const double k1 = (t[0] * t[0]) * e1;
        const double k2 = (t[2] * t[2]) * e2;
        const double k3 = (t[5] * t[5]) * pw;
        double k4 = (t[8] * t[8]) * e4;
        if (diag) k4 += t[10] * t[10];                                        // :66 (delta iff square block)
        if (WITH_DK) {
            dk[0] = 2.0 * t[0] * e1;
            dk[1] = k1 * d / (t[1] * t[1] * t[1]);
            dk[2] = 2.0 * k2 / t[2];
            dk[3] = k2 * d / (t[3] * t[3] * t[3]);
            dk[4] = k2 * 4.0 * (sn * sn) / (t[4] * t[4] * t[4]);
            dk[5] = 2.0 * k3 / t[5];
            dk[6] = k3 * d / (t[6] * t[6] * t[6] * u);
            dk[7] = k3 * (-log(u) + (u - 1.0) / u);
            dk[8] = 2.0 * t[8] * e4;
            dk[9] = (t[8] * t[8]) * e4 * d / (t[9] * t[9] * t[9]);
            dk[10] = diag ? 2.0 * t[10] : 0.0;
        }
        return ((k1 + k2) + k3) + k4;                                         // :90-93 (left-to-right sum)
    }
}

}  // namespace gpx_cov
