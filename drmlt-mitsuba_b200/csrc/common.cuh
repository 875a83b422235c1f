// common.cuh -- float3/RGB helpers, constants and the counter-based RNG of the GPU path.
//
// Numerics: float32 everywhere (Mitsuba's single-precision constants: Epsilon = 1e-4f,
// ShadowEpsilon = 1e-3f, include/mitsuba/core/constants.h:29-31); the MIS ratio sweep is done in
// double like the reference's (src/libbidir/path.cpp:979-1025).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define DR_HD __host__ __device__ __forceinline__
#define DR_D __device__ __forceinline__

#define DR_PI 3.14159265358979323846f
#define DR_INV_PI 0.31830988618379067154f
#define DR_DELTA_EPS 1e-3f
#define DR_RCPOVERFLOW 2.93873587705571876e-39f   /* 0x1p-128f, constants.h:57 */

// ------------------------------------------------------------------ float3
DR_HD float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
DR_HD float3 f3(float v) { return make_float3(v, v, v); }
DR_HD float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
DR_HD float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
DR_HD float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
DR_HD float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
DR_HD float3 operator*(float s, float3 a) { return f3(a.x * s, a.y * s, a.z * s); }
DR_HD float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
DR_HD float3 operator/(float3 a, float s) { float r = 1.0f / s; return f3(a.x * r, a.y * r, a.z * r); }
DR_HD float3 operator/(float3 a, float3 b) { return f3(a.x / b.x, a.y / b.y, a.z / b.z); }
DR_HD float3 &operator+=(float3 &a, float3 b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
DR_HD float3 &operator*=(float3 &a, float3 b) { a.x *= b.x; a.y *= b.y; a.z *= b.z; return a; }
DR_HD float3 &operator*=(float3 &a, float s) { a.x *= s; a.y *= s; a.z *= s; return a; }
DR_HD float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
DR_HD float absdot(float3 a, float3 b) { return fabsf(dot(a, b)); }
DR_HD float3 cross(float3 a, float3 b) { return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
DR_HD float length(float3 a) { return sqrtf(dot(a, a)); }
DR_HD float3 normalize(float3 a) { return a * (1.0f / sqrtf(dot(a, a))); }
DR_HD bool is_zero(float3 a) { return a.x == 0.f && a.y == 0.f && a.z == 0.f; }
DR_HD float max3(float3 a) { return fmaxf(a.x, fmaxf(a.y, a.z)); }
// spectrum.h:734-736
DR_HD float luminance(float3 c) { return c.x * 0.212671f + c.y * 0.715160f + c.z * 0.072169f; }
DR_HD bool rgb_valid(float3 c) {
    return isfinite(c.x) && isfinite(c.y) && isfinite(c.z) && c.x >= 0.f && c.y >= 0.f && c.z >= 0.f;
}
DR_HD float safe_sqrtf(float v) { return sqrtf(fmaxf(v, 0.f)); }
DR_HD float safe_acosf(float v) { return acosf(fminf(1.f, fmaxf(-1.f, v))); }

// src/libcore/util.cpp:600-609 coordinateSystem / frame.h:57-59 Frame(n)
DR_HD void coordinate_system(float3 a, float3 &b, float3 &c) {
    if (fabsf(a.x) > fabsf(a.y)) {
        float invLen = 1.0f / sqrtf(a.x * a.x + a.z * a.z);
        c = f3(a.z * invLen, 0.0f, -a.x * invLen);
    } else {
        float invLen = 1.0f / sqrtf(a.y * a.y + a.z * a.z);
        c = f3(0.0f, a.z * invLen, -a.y * invLen);
    }
    b = cross(c, a);
}

// ------------------------------------------------------------------ Philox4x32-10, keyed uniforms
// Same address space as oracle/orc_rng.hpp (DESIGN.md "uniform address space"):
//   word(stream,a,b,j) = philox(ctr={lo(a),hi(a),b,(stream<<24)|(j>>2)}, key=seed)[j&3]
//   uniform = (word >> 8) * 2^-24
enum { S_BOOT = 1, S_RESAMPLE = 2, S_COIN = 3, S_STAGE1 = 4, S_STAGE2 = 7, S_DIRECT = 10 };

DR_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t) (((uint64_t) a * b) >> 32);
#endif
}

DR_HD uint4 philox4x32_10(uint4 c, uint64_t seed) {
    uint32_t k0 = (uint32_t) seed, k1 = (uint32_t) (seed >> 32);
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        uint32_t hi0 = mulhi32(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        uint32_t hi1 = mulhi32(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c;
}
DR_HD float u32_to_unit(uint32_t w) { return (float) (w >> 8) * (1.0f / 16777216.0f); }

// the four uniforms j = 4q .. 4q+3 of (stream, a, b)
DR_HD float4 keyed_uniform4(uint64_t seed, uint32_t stream, uint64_t a, uint32_t b, uint32_t q) {
    uint4 r = philox4x32_10(make_uint4((uint32_t) a, (uint32_t) (a >> 32), b, (stream << 24) | q), seed);
    return make_float4(u32_to_unit(r.x), u32_to_unit(r.y), u32_to_unit(r.z), u32_to_unit(r.w));
}
DR_HD float keyed_uniform(uint64_t seed, uint32_t stream, uint64_t a, uint32_t b, uint32_t j) {
    float4 r = keyed_uniform4(seed, stream, a, b, j >> 2);
    switch (j & 3) { case 0: return r.x; case 1: return r.y; case 2: return r.z; default: return r.w; }
}
