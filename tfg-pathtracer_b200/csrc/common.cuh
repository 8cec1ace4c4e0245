/*
 * common.cuh — device-side vector math, the parity-critical "exact" arithmetic, RNGs and scene views.
 *
 * Two arithmetic regimes live side by side:
 *   ex::  IEEE binary32 ops issued through __f*_rn intrinsics in the reference's source order, which nvcc never
 *         contracts into FMAs.  Used where the contract is BIT-EXACT against the reference built with
 *         --fmad=false: ray normalisation (S/Ray.hpp:14-18), Moeller-Trumbore (S/Tri.hpp:38-68), the
 *         shadow-terminator position and the ordering key (S/Tri.hpp:70-92, S/BVH.hpp:170).
 *   plain float ops (FMA contraction allowed) everywhere else: box tests, BRDF, sampling, film — compared
 *         against the reference within the image tolerance.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/eleven_b200.h"

namespace eleven {

#define EL_PI 3.14159265358979323846f   /* S/Math.hpp:7 (float) */

struct F3 { float x, y, z; };
__host__ __device__ __forceinline__ F3 f3(float x, float y, float z) { F3 r; r.x = x; r.y = y; r.z = z; return r; }
__host__ __device__ __forceinline__ F3 f3(float s) { return f3(s, s, s); }
__device__ __forceinline__ F3 operator+(F3 a, F3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ F3 operator-(F3 a, F3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ F3 operator*(F3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ F3 operator*(float s, F3 a) { return f3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ F3 operator*(F3 a, F3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ F3 operator/(F3 a, float s) { return f3(a.x / s, a.y / s, a.z / s); }
__device__ __forceinline__ float dot(F3 a, F3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ F3 cross(F3 a, F3 b) { return f3(a.y * b.z - a.z * b.y, -(a.x * b.z - a.z * b.x), a.x * b.y - a.y * b.x); }
__device__ __forceinline__ float length(F3 a) { return sqrtf(dot(a, a)); }
__device__ __forceinline__ F3 normalized(F3 a) { float l = length(a); return l == 0.f ? a : f3(a.x / l, a.y / l, a.z / l); }  // S/Vector.hpp:159-172

// ---------------------------------------------------------------------------------------------------
// exact (uncontracted, round-to-nearest) arithmetic in the reference's association order
// ---------------------------------------------------------------------------------------------------
namespace ex {
__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float sqrt(float a) { return __fsqrt_rn(a); }
__device__ __forceinline__ F3 add(F3 a, F3 b) { return f3(add(a.x, b.x), add(a.y, b.y), add(a.z, b.z)); }
__device__ __forceinline__ F3 sub(F3 a, F3 b) { return f3(sub(a.x, b.x), sub(a.y, b.y), sub(a.z, b.z)); }
__device__ __forceinline__ F3 mul(F3 a, float s) { return f3(mul(a.x, s), mul(a.y, s), mul(a.z, s)); }
__device__ __forceinline__ float dot(F3 a, F3 b) { return add(add(mul(a.x, b.x), mul(a.y, b.y)), mul(a.z, b.z)); }        // S/Vector.hpp:151
__device__ __forceinline__ F3 cross(F3 a, F3 b) {                                                                          // S/Vector.hpp:155
    return f3(sub(mul(a.y, b.z), mul(a.z, b.y)), -sub(mul(a.x, b.z), mul(a.z, b.x)), sub(mul(a.x, b.y), mul(a.y, b.x)));
}
__device__ __forceinline__ float length(F3 a) { return sqrt(add(add(mul(a.x, a.x), mul(a.y, a.y)), mul(a.z, a.z))); }    // S/Vector.hpp:93
__device__ __forceinline__ F3 normalize(F3 a) {                                                                            // S/Vector.hpp:159 / S/Ray.hpp:17
    float l = ex::length(a);
    if (l == 0.f) return a;
    return f3(div(a.x, l), div(a.y, l), div(a.z, l));
}
// a + b*s, as `a + b * s` on Vector3 (two roundings per component)
__device__ __forceinline__ F3 madd(F3 a, F3 b, float s) { return add(a, mul(b, s)); }
} // namespace ex

// ---------------------------------------------------------------------------------------------------
// RNG
// ---------------------------------------------------------------------------------------------------
struct Xorwow { uint32_t v0, v1, v2, v3, v4, d; };     // curandStateXORWOW_t without the Box-Muller fields

__device__ __forceinline__ uint32_t xorwowNext(Xorwow& s) {       // curand_kernel.h:863-876
    uint32_t t = s.v0 ^ (s.v0 >> 2);
    s.v0 = s.v1; s.v1 = s.v2; s.v2 = s.v3; s.v3 = s.v4;
    s.v4 = (s.v4 ^ (s.v4 << 4)) ^ (t ^ (t << 1));
    s.d += 362437u;
    return s.v4 + s.d;
}
// curand_uniform.h:69-72, uncontracted like the --fmad=false reference build: (0, 1]
__device__ __forceinline__ float u32ToUniform(uint32_t x) { return __fadd_rn(__fmul_rn((float)x, 2.3283064e-10f), 2.3283064e-10f / 2.0f); }
__device__ __forceinline__ float xorwowUniform(Xorwow& s) { return u32ToUniform(xorwowNext(s)); }

// Philox-4x32-10 (Salmon et al., SC'11): counter-based generator for the fast mode; key = seed, counter =
// (pixel, sample, dimension block, 0).  Stateless: any (pixel, sample) can be drawn on any GPU.
__device__ __forceinline__ uint4 philox4x32(uint4 c, uint2 k) {
#pragma unroll
    for (int i = 0; i < 10; i++) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
    }
    return c;
}

// ---------------------------------------------------------------------------------------------------
// device views of the scene
// ---------------------------------------------------------------------------------------------------
struct DevTex {
    const void* data;        // uchar4 (formats 1,2) or float4 (format 0) per texel
    int32_t width, height;
    float xTile, yTile, xOffset, yOffset;
    uint32_t format, filter;
};

struct DevMaterial {         // ElevenMaterial
    int32_t albedoTex, emissionTex, roughnessTex, metallicTex, normalTex, opacityTex;
    float albedo[3], emission[3], opacity[3];
    float roughness, metallic, clearcoatGloss, clearcoat, anisotropic, eta, transmission, specular, specularTint, sheenTint, subsurface, sheen;
};

// One texel record for the four maps of a material (albedo rgb, roughness r | normal rgb, metallic r), built at upload when the
// maps are 8-bit, unfiltered and share size / tiling / offset: ONE 8-byte gather per hit instead of four 4-byte gathers into
// four 64 MB arrays (k_shade is bound by exactly those DRAM sectors on bounce rays).
struct DevPackedMaps {
    const uint2* data;       // nullptr: the material's maps are fetched one by one
    int32_t width, height;
    float xTile, yTile, xOffset, yOffset;
    uint32_t albedoFormat, pad_;   // ELEVEN_TEX_U8_SRGB / _LINEAR of the albedo map (the other three are decoded by their own format too)
    uint32_t roughFormat, metalFormat, normalFormat, pad2_;
};

struct DevCamera { uint32_t xRes, yRes; float focalLength, sensorWidth, sensorHeight, aperture, focusDistance; float rot[3]; float pos[3]; uint32_t bokeh; };

struct AliasEntry { float prob; uint32_t alias; };

struct DevScene {
    const float4* nodes;          // Node8 as 5 x float4
    const float4* slots;          // TriSlot as 3 x float4
    const float* nodeSlack;       // per node: max shadow-terminator shift bound in its subtree
    const float4* shadeTris;      // 9 x float4 per triangle, original index order
    const int32_t* objectMaterial;
    const int32_t* triMaterial;   // per triangle (original index): material id, for the material-sorted shading queue
    const DevMaterial* materials;
    const DevTex* textures;
    const DevPackedMaps* packed;  // per material
    const float* lut;             // [2][256]: sRGB (gamma 2.2f) then linear (gamma 1.0f) fastPow tables
    DevTex hdri;                  // float4 texels: rgb + (r+g)+b
    const float* cdf;             // W*H+1
    const AliasEntry* alias;      // W*H
    float radianceSum;
    const float* lights;          // n x 6
    uint32_t lightCount, triCount, nodeCount;
    float keySlack;
    uint32_t byteMagic;           // 0x47000000, passed as DATA so that the compiler keeps it in a register (trace_engine.cuh: byteMagic15)
    DevCamera cam;
};

} // namespace eleven
