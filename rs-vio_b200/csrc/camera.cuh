// camera.cuh -- unprojection of tracked pixels to normalised image coordinates.
//
// Replaces, per feature, `camera.unproject_one(&Vector2::new(u, v))` of the reference
// (src/estimator/frame.rs:118,131; models built in src/datasets/mod.rs:93-163).  The model code of
// the reference is the un-vendored crate camera-intrinsic-model 0.7.2 (Cargo.toml:20): PARITY
// UNPINNED.  Restated from the published models, f64 like the reference:
//   pinhole-radtan (OpenCV, 5 coefficients):  x' = x (1 + k1 r^2 + k2 r^4 + k3 r^6) + 2 p1 x y + p2 (r^2 + 2 x^2)
//                                             y' = y (1 + ...)                      + p1 (r^2 + 2 y^2) + 2 p2 x y
//       inverted by Newton's method on the 2x2 system to f64 convergence (any convergent scheme of the
//       crate gives the same root);
//   EUCM (Khomutenko et al. 2016): closed form, m_z = (1 - beta alpha^2 r^2) / (alpha sqrt(1 - (2 alpha - 1) beta r^2) + 1 - alpha).
// Output: (x / z, y / z), what the estimator's factors consume (src/optimization/factors.rs:13-25).
#pragma once
#include <math.h>

#include "../../include/rsvio_b200.h"

namespace rsvio {

struct CameraModelDev {
    int model;
    double p[9];
};

#if defined(__CUDACC__)
#define RSVIO_HD __host__ __device__ __forceinline__
#else
#define RSVIO_HD inline
#endif

// returns false (and NaNs) when (u, v) is outside the model's domain or the iteration does not converge
RSVIO_HD bool unproject_pixel(const CameraModelDev &c, double u, double v, double &x, double &y) {
    const double xd = (u - c.p[2]) / c.p[0], yd = (v - c.p[3]) / c.p[1];
    x = y = nan("");
    if (c.model == RSVIO_CAMERA_PINHOLE_RADTAN) {
        const double k1 = c.p[4], k2 = c.p[5], p1 = c.p[6], p2 = c.p[7], k3 = c.p[8];
        double a = xd, b = yd;
        for (int it = 0; it < 30; ++it) {
            const double r2 = a * a + b * b;
            const double rad = 1.0 + r2 * (k1 + r2 * (k2 + r2 * k3));
            const double drad = k1 + r2 * (2.0 * k2 + r2 * 3.0 * k3);
            const double fx = a * rad + 2.0 * p1 * a * b + p2 * (r2 + 2.0 * a * a) - xd;
            const double fy = b * rad + p1 * (r2 + 2.0 * b * b) + 2.0 * p2 * a * b - yd;
            const double jxx = rad + 2.0 * a * a * drad + 2.0 * p1 * b + 6.0 * p2 * a;
            const double jxy = 2.0 * a * b * drad + 2.0 * p1 * a + 2.0 * p2 * b;
            const double jyy = rad + 2.0 * b * b * drad + 6.0 * p1 * b + 2.0 * p2 * a;
            const double det = jxx * jyy - jxy * jxy;
            if (!(fabs(det) > 1e-300)) return false;
            const double da = (jyy * fx - jxy * fy) / det, db = (jxx * fy - jxy * fx) / det;
            a -= da;
            b -= db;
            if (fabs(da) + fabs(db) <= 1e-14 * (1.0 + fabs(a) + fabs(b))) {  // a few ulps
                x = a;
                y = b;
                return true;
            }
        }
        return false;
    }
    if (c.model == RSVIO_CAMERA_EUCM) {
        const double alpha = c.p[4], beta = c.p[5];
        const double r2 = xd * xd + yd * yd;
        const double disc = 1.0 - (2.0 * alpha - 1.0) * beta * r2;
        if (disc < 0.0) return false;
        const double mz = (1.0 - beta * alpha * alpha * r2) / (alpha * sqrt(disc) + (1.0 - alpha));
        if (!(mz > 0.0)) return false;  // at or behind the image plane: no normalised coordinate
        x = xd / mz;
        y = yd / mz;
        return true;
    }
    return false;
}

}  // namespace rsvio
