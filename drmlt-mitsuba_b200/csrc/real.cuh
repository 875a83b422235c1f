// real.cuh -- the arithmetic type of the shading / chain logic.
//
// The reference's default build computes in double (MTS_DOUBLE_PRECISION,
// data/cmake/MitsubaBuildOptions.cmake:48-70).  Ray traversal -- the bandwidth-bound part -- runs
// in float32 on 16-byte node/triangle records; everything that turns a hit into a path vertex, a
// BSDF sample, a pdf, a MIS weight or an acceptance probability runs in `Real` = double: the hit
// is re-intersected in double against the one triangle the traversal selected, so vertex
// positions, cosines and geometric terms carry no float cancellation error and f(u) agrees with
// the reference's arithmetic to ~1e-12 instead of ~1e-4.  B200 issues FP64 at half the FP32 rate,
// and this logic is a small share of the mutation cost next to the BVH gathers.
#pragma once
#include "common.cuh"

typedef double Real;
#define R_PI 3.14159265358979323846
#define R_INV_PI 0.31830988618379067154
#define R_DELTA_EPS 1e-3                    /* DeltaEpsilon, constants.h:31 */
#define R_RCPOVERFLOW 5.56268464626800345e-309   /* 0x1p-1024, constants.h:58 (double build) */

struct R3 { Real x, y, z; };
struct R2 { Real x, y; };
DR_HD R3 r3(Real x, Real y, Real z) { R3 r; r.x = x; r.y = y; r.z = z; return r; }
DR_HD R3 r3(Real v) { return r3(v, v, v); }
DR_HD R3 r3(float3 v) { return r3((Real) v.x, (Real) v.y, (Real) v.z); }
DR_HD R2 r2(Real x, Real y) { R2 r; r.x = x; r.y = y; return r; }
DR_HD float3 to_f3(R3 v) { return make_float3((float) v.x, (float) v.y, (float) v.z); }
DR_HD R3 operator+(R3 a, R3 b) { return r3(a.x + b.x, a.y + b.y, a.z + b.z); }
DR_HD R3 operator-(R3 a, R3 b) { return r3(a.x - b.x, a.y - b.y, a.z - b.z); }
DR_HD R3 operator-(R3 a) { return r3(-a.x, -a.y, -a.z); }
DR_HD R3 operator*(R3 a, Real s) { return r3(a.x * s, a.y * s, a.z * s); }
DR_HD R3 operator*(Real s, R3 a) { return r3(a.x * s, a.y * s, a.z * s); }
DR_HD R3 operator*(R3 a, R3 b) { return r3(a.x * b.x, a.y * b.y, a.z * b.z); }
DR_HD R3 operator/(R3 a, Real s) { const Real r = 1.0 / s; return r3(a.x * r, a.y * r, a.z * r); }   // TVector3::operator/ (vector.h)
DR_HD R3 operator/(R3 a, R3 b) { return r3(a.x / b.x, a.y / b.y, a.z / b.z); }
DR_HD R3 &operator+=(R3 &a, R3 b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
DR_HD R3 &operator*=(R3 &a, R3 b) { a.x *= b.x; a.y *= b.y; a.z *= b.z; return a; }
DR_HD R3 &operator*=(R3 &a, Real s) { a.x *= s; a.y *= s; a.z *= s; return a; }
DR_HD Real dot(R3 a, R3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
DR_HD Real absdot(R3 a, R3 b) { return fabs(dot(a, b)); }
DR_HD R3 cross(R3 a, R3 b) { return r3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
DR_HD Real length(R3 a) { return sqrt(dot(a, a)); }
DR_HD R3 normalize(R3 a) { return a / sqrt(dot(a, a)); }
DR_HD bool is_zero(R3 a) { return a.x == 0.0 && a.y == 0.0 && a.z == 0.0; }
DR_HD Real max3(R3 a) { return fmax(a.x, fmax(a.y, a.z)); }
DR_HD Real luminance(R3 c) { return c.x * 0.212671f + c.y * 0.715160f + c.z * 0.072169f; }   // spectrum.h:734-736
DR_HD bool rgb_valid(R3 c) { return isfinite(c.x) && isfinite(c.y) && isfinite(c.z) && c.x >= 0.0 && c.y >= 0.0 && c.z >= 0.0; }
DR_HD Real safe_sqrt(Real v) { return sqrt(fmax(v, 0.0)); }
DR_HD Real safe_acos(Real v) { return acos(fmin(1.0, fmax(-1.0, v))); }

// src/libcore/util.cpp:600-609 coordinateSystem / frame.h:57-59 Frame(n)
DR_HD void coordinate_system(R3 a, R3 &b, R3 &c) {
    if (fabs(a.x) > fabs(a.y)) {
        Real invLen = 1.0 / sqrt(a.x * a.x + a.z * a.z);
        c = r3(a.z * invLen, 0.0, -a.x * invLen);
    } else {
        Real invLen = 1.0 / sqrt(a.y * a.y + a.z * a.z);
        c = r3(0.0, a.z * invLen, -a.y * invLen);
    }
    b = cross(c, a);
}
