/*
 * shading.cuh — camera, texture fetch, environment, sampling routines and the Disney principled BRDF (device).
 *
 * Same estimator as the reference (so that images converge to the same result), new data layout:
 *   textures   8-bit RGBA texels (4 B, one 32-bit load) decoded through the 256-entry fastPow tables that reproduce
 *              stb_image's patched LDR decode bit for bit (S/stb_image.h:127-136,1863), instead of float RGB (12 B);
 *              float textures are padded to float4 (one 128-bit load)
 *   HDRI       float4 texels whose .w caches (r+g)+b, the quantity both the CDF and pdf() need (S/HDRI.hpp:115,148)
 * Function-by-function provenance is cited inline.
 *
 * Every function is a template on FM ("fast math").  FM = false is the parity flavour: IEEE division/sqrt, libm-accurate
 * sin/cos/pow/log and the reference's float/double promotions, compiled with --fmad=false (compared with the reference
 * under the image tolerance).  FM = true is the production flavour (the reference's own shipping build is
 * -use_fast_math, S/tfg-pathtracer.vcxproj:136-143): MUFU-based reciprocal/rsqrt/sin/cos/exp2/log2, float instead of
 * double intermediates — ncu showed IEEE division sequences and double sqrt/div as the largest instruction groups of
 * k_shade (profiles/r1_pipeline_v3_ncu_full.txt).
 */
#pragma once
#include "common.cuh"
#include "traverse.cuh"

namespace eleven {

__device__ __forceinline__ float minf_(float a, float b) { return a < b ? a : b; }          // S/Math.hpp:60 (NaN-unsafe on purpose)
__device__ __forceinline__ float maxf_(float a, float b) { return a > b ? a : b; }          // S/Math.hpp:64
__device__ __forceinline__ float clampf_(float a, float b, float c) { return a < b ? b : a > c ? c : a; }   // S/Math.hpp:29
__device__ __forceinline__ float lerpf_(float a, float b, float c) { return a + c * (b - a); }              // S/Math.hpp:43 FAST_LERP
__device__ __forceinline__ F3 lerp3(F3 a, F3 b, float c) { return f3(lerpf_(a.x, b.x, c), lerpf_(a.y, b.y, c), lerpf_(a.z, b.z, c)); }

// ---- arithmetic flavours ------------------------------------------------------------------------------
template <bool FM> struct M {
    static __device__ __forceinline__ float div(float a, float b) { return FM ? __fdividef(a, b) : a / b; }
    static __device__ __forceinline__ float rcp(float a) { return FM ? __fdividef(1.0f, a) : 1.0f / a; }
    static __device__ __forceinline__ float sqrt(float a) {
        if (FM) { float r; asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(a)); return r; }
        return sqrtf(a);
    }
    static __device__ __forceinline__ float sin(float a) { return FM ? __sinf(a) : sinf(a); }
    static __device__ __forceinline__ float cos(float a) { return FM ? __cosf(a) : cosf(a); }
    static __device__ __forceinline__ float log(float a) { return FM ? __logf(a) : logf(a); }
    static __device__ __forceinline__ float pow(float a, float b) { return FM ? __powf(a, b) : powf(a, b); }
    static __device__ __forceinline__ F3 normalized(F3 a) {                                  // S/Vector.hpp:159-172
        if (FM) { const float d = dot(a, a); return d == 0.f ? a : a * rsqrtf(d); }
        const float l = length(a); return l == 0.f ? a : f3(a.x / l, a.y / l, a.z / l);
    }
    static __device__ __forceinline__ F3 div3(F3 a, float s) { if (FM) return a * __fdividef(1.0f, s); return f3(a.x / s, a.y / s, a.z / s); }
};

// C remainder with the sign of the dividend; power-of-two sizes (the usual case) avoid the integer division sequence
__device__ __forceinline__ int modSize(int x, int n) {
    if ((n & (n - 1)) == 0) return x >= 0 ? (x & (n - 1)) : -((-x) & (n - 1));
    return x % n;
}

// ---- texture fetch: S/Texture.hpp:95-142 -----------------------------------------------------------
__device__ __forceinline__ long long texelIndex(const DevTex& t, int x, int y) {
    x = modSize((int)(t.xTile * (x + t.xOffset * t.width)), t.width);
    y = modSize((int)(t.yTile * (y + t.yOffset * t.height)), t.height);
    long long idx = (long long)y * t.width + x;                       // the reference indexes 3*(y*W+x): negative x walks into row y-1
    const long long n = (long long)t.width * t.height;
    return idx < 0 ? 0 : (idx >= n ? n - 1 : idx);                    // defined behaviour for the reference's out-of-bounds reads
}
__device__ __forceinline__ F3 texelAt(const DevTex& t, const float* __restrict__ lut, int x, int y) {
    const long long idx = texelIndex(t, x, y);
    if (t.format == ELEVEN_TEX_F32_RGB) {
        const float4 v = __ldg((const float4*)t.data + idx);
        return f3(v.x, v.y, v.z);
    }
    const uchar4 c = __ldg((const uchar4*)t.data + idx);
    const float* l = lut + (t.format == ELEVEN_TEX_U8_SRGB ? 0 : 256);
    return f3(__ldg(l + c.x), __ldg(l + c.y), __ldg(l + c.z));
}
__device__ __forceinline__ F3 texFromUV(const DevTex& t, const float* lut, float u, float v) {            // :110-112
    return texelAt(t, lut, (int)(u * t.width), (int)(v * t.height));
}
// cold path (Texture::filter defaults to NO_FILTER, S/Texture.hpp:21): kept out of line, inlining it at the five fetch
// sites was a third of k_shade's 100 KB of code and the kernel stalled on instruction fetch (ncu: no_instruction)
__device__ __noinline__ F3 texBilinear(const DevTex& t, const float* lut, float u, float v) {          // :114-135
    const float x = u * t.width, y = v * t.height;
    const float t1x = floorf(x), t1y = floorf(y), t2x = t1x + 1, t2y = t1y + 1;
    const float a = (x - t1x) / (t2x - t1x), b = (y - t1y) / (t2y - t1y);
    const F3 v1 = texelAt(t, lut, (int)t1x, (int)t1y), v2 = texelAt(t, lut, (int)t2x, (int)t1y);
    const F3 v3 = texelAt(t, lut, (int)t1x, (int)t2y), v4 = texelAt(t, lut, (int)t2x, (int)t2y);
    return lerp3(lerp3(v1, v2, a), lerp3(v3, v4, a), b);
}
__device__ __noinline__ F3 texFiltered(const DevTex& t, const float* lut, float u, float v) {          // :137-142 (cold: float or filtered maps)
    return t.filter == 0 ? texFromUV(t, lut, u, v) : texBilinear(t, lut, u, v);
}

// ---- environment: S/Texture.hpp:144-207, S/HDRI.hpp:130-162 ------------------------------------------
__device__ __forceinline__ void limitUV(float& u, float& v) {                                              // S/Math.hpp:38-41
    u += (float)(-(int)(u > 1) + -(int)(u < 0));
    v += (float)(-(int)(v > 1) + -(int)(v < 0));
}
__device__ __forceinline__ void sphericalMapping(F3 p, float& u, float& v) {                               // S/Texture.hpp:144-156 with origin 0, radius 1
    const float theta = acosf(-p.y);
    const float phi = atan2f(-p.z, p.x) + EL_PI;
    u = phi / (2 * EL_PI);
    v = theta / EL_PI;
    limitUV(u, v);
}
template <bool FM>
__device__ __forceinline__ F3 reverseSphericalMapping(float u, float v) {                                  // S/Texture.hpp:195-207
    const float phi = u * 2 * EL_PI, theta = v * EL_PI;
    const float px = M<FM>::cos(phi - EL_PI), py = -M<FM>::cos(theta), pz = -M<FM>::sin(phi - EL_PI);
    const float a = M<FM>::sqrt(1 - py * py);
    return f3(a * px, py, a * pz);
}
__device__ __forceinline__ void inverseTransformUV(const DevTex& t, float u, float v, float& nu, float& nv) {   // S/Texture.hpp:177-193
    int x = (int)(u * t.width), y = (int)(v * t.height);
    x = modSize((int)(t.xTile * (x - t.xOffset * t.width)), t.width);
    y = modSize((int)(t.yTile * (y - t.yOffset * t.height)), t.height);
    nu = (float)x / (float)t.width; nv = (float)y / (float)t.height;
    limitUV(nu, nv);
}
__device__ __forceinline__ float4 envTexelRaw(const DevTex& t, int x, int y) {
    return __ldg((const float4*)t.data + texelIndex(t, x, y));
}
// radiance seen by an escaped ray: S/kernel.cu:415-417 (nearest texel; bilinear if the texture says so)
__device__ __forceinline__ F3 envLookup(const DevScene& S, F3 dir) {
    float u, v;
    sphericalMapping(f3(-dir.x, -dir.y, -dir.z), u, v);
    if (S.hdri.filter == 0) { const float4 c = envTexelRaw(S.hdri, (int)(u * S.hdri.width), (int)(v * S.hdri.height)); return f3(c.x, c.y, c.z); }
    return texBilinear(S.hdri, nullptr, u, v);
}
// S/HDRI.hpp:130-142: the reference's own (approximate) lower bound
__device__ __forceinline__ int cdfSearch(const float* __restrict__ arr, float value, int length) {
    int from = 0, to = length - 1;
    while (to - from > 0) {
        const int m = from + (to - from) / 2;
        const float a = __ldg(arr + m);
        if (value == a) return m;
        if (value < a) to = m - 1;
        if (value > a) from = m + 1;
    }
    return to;
}
// S/HDRI.hpp:145-152; dv = the texel at (x, y), fetched by the caller (k_shade issues that load early)
template <bool FM>
__device__ __forceinline__ float hdriPdf(const DevScene& S, float4 dv, int y) {
    const float theta = (((float)y / (float)S.hdri.height)) * EL_PI;
    const float num = ((dv.w) / S.radianceSum) * S.hdri.width * S.hdri.height;
    if (FM) return __fdividef(num, 2.0f * EL_PI * __sinf(theta));
    return (float)((double)num / (2.0 * (double)EL_PI * (double)sinf(theta)));
}

// ---- sampling: S/Sampling.hpp:21-54 ------------------------------------------------------------------
template <bool FM>
__device__ __forceinline__ void uniformCircleSampling(float u1, float u2, float u3, float& x, float& y) {
    const float t = 2 * EL_PI * u1, u = u2 + u3, r = u > 1 ? 2 - u : u;
    const float s = M<FM>::sin(t), c = M<FM>::cos(t);        /* separate calls like the reference (S/Sampling.hpp:27-28) */
    x = r * c; y = r * s;
}
template <bool FM>
__device__ __forceinline__ F3 cosineSampleHemisphere(float u1, float u2) {
    const float r = M<FM>::sqrt(u1), phi = FM ? 2.0f * EL_PI * u2 : (float)(2.0 * (double)EL_PI * (double)u2);
    const float s = M<FM>::sin(phi), c = M<FM>::cos(phi);
    F3 d; d.x = r * c; d.y = r * s;
    const float zz = FM ? 1.0f - d.x * d.x - d.y * d.y : (float)(1.0 - (double)(d.x * d.x) - (double)(d.y * d.y));
    d.z = M<FM>::sqrt(maxf_(0.0f, zz));
    return d;
}
template <bool FM>
__device__ __forceinline__ F3 importanceSampleGGX(float rgh, float r1, float r2) {
    const float a = maxf_(0.001f, rgh);
    const float phi = r1 * EL_PI * 2;
    float cosTheta, sinTheta;
    if (FM) {
        cosTheta = M<true>::sqrt(__fdividef(1.0f - r2, 1.0f + (a * a - 1.0f) * r2));
        sinTheta = clampf_(M<true>::sqrt(1.0f - cosTheta * cosTheta), 0.0f, 1.0f);
    } else {
        cosTheta = (float)::sqrt((1.0 - (double)r2) / (1.0 + ((double)(a * a) - 1.0) * (double)r2));
        sinTheta = clampf_((float)::sqrt(1.0 - (double)(cosTheta * cosTheta)), 0.0f, 1.0f);
    }
    const float sp = M<FM>::sin(phi), cp = M<FM>::cos(phi);
    return f3(sinTheta * cp, sinTheta * sp, cosTheta);
}

// ---- Disney principled BRDF: S/Disney.hpp:41-253 -----------------------------------------------------
struct HitData {                                                                                           // S/kernel.h:46-69
    float metallic, roughness, clearcoatGloss, clearcoat, anisotropic, eta, transmission, specular, specularTint, sheenTint, subsurface, sheen;
    F3 emission, albedo, normal;
};
template <bool FM>
__device__ __forceinline__ void createBasis(F3 n, F3& T, F3& B) { T = M<FM>::normalized(cross(f3(0, 1, 0), n)); B = M<FM>::normalized(cross(n, T)); }   // :41-45
__device__ __forceinline__ float schlick(float u) { const float m = clampf_(1.0f - u, 0.0f, 1.0f); const float m2 = m * m; return m2 * m2 * m; }
template <bool FM>
__device__ __forceinline__ float GTR1(float NDotH, float a) {                                              // :67-73
    if (a >= 1.0f) return 1.0f / EL_PI;
    const float a2 = a * a;
    const float t = 1.0f + (a2 - 1.0f) * NDotH * NDotH;
    return M<FM>::div(a2 - 1.0f, EL_PI * M<FM>::log(a2) * t);
}
template <bool FM>
__device__ __forceinline__ float GTR2aniso(float NDotH, float HDotX, float HDotY, float ax, float ay) {    // :81-86
    const float a = M<FM>::div(HDotX, ax), b = M<FM>::div(HDotY, ay), c = a * a + b * b + NDotH * NDotH;
    return M<FM>::rcp(EL_PI * ax * ay * c * c);
}
template <bool FM>
__device__ __forceinline__ float smithG(float NDotV, float alphaG) { const float a = alphaG * alphaG, b = NDotV * NDotV; return M<FM>::rcp(NDotV + M<FM>::sqrt(a + b - a * b)); }
template <bool FM>
__device__ __forceinline__ float smithGaniso(float NDotV, float VDotX, float VDotY, float ax, float ay) {
    const float a = VDotX * ax, b = VDotY * ay, c = NDotV;
    return M<FM>::rcp(NDotV + M<FM>::sqrt(a * a + b * b + c * c));
}
// everything about the BRDF that depends on the hit point only (not on L): hoisted out of the up-to-3 Eval calls per hit
struct BrdfFrame {
    F3 N, V, T, B;
    float NDotV, ax, ay, diffuseRatio;
    F3 Cspec0, Csheen;
    float GsV, GrV, clearcoatAlpha;
    bool frontV;
};
template <bool FM>
__device__ __forceinline__ BrdfFrame makeBrdfFrame(const HitData& hd, F3 rayDir) {
    BrdfFrame f;
    f.N = hd.normal; f.V = f3(-rayDir.x, -rayDir.y, -rayDir.z);
    createBasis<FM>(hd.normal, f.T, f.B);
    f.NDotV = fabsf(dot(f.N, f.V)); f.frontV = dot(f.N, f.V) > 0.0f;
    const float aspect = M<FM>::sqrt(1.0f - hd.anisotropic * 0.9f);
    f.ax = maxf_(0.001f, M<FM>::div(hd.roughness, aspect)); f.ay = maxf_(0.001f, hd.roughness * aspect);
    f.diffuseRatio = 0.5f * (1.0f - hd.metallic);
    const F3 Cd = hd.albedo;
    const float Cdlum = 0.3f * Cd.x + 0.6f * Cd.y + 0.1f * Cd.z;
    const F3 Ctint = Cdlum > 0.0f ? M<FM>::div3(Cd, Cdlum) : f3(1.0f);
    f.Cspec0 = lerp3((hd.specular * 0.08f) * lerp3(f3(1.0f), Ctint, hd.specularTint), Cd, hd.metallic);
    f.Csheen = lerp3(f3(1.0f), Ctint, hd.sheenTint);
    // view-dependent masking terms: identical for every light direction evaluated at this hit
    f.GsV = smithGaniso<FM>(f.NDotV, dot(f.V, f.T), dot(f.V, f.B), f.ax, f.ay);
    f.GrV = smithG<FM>(f.NDotV, 0.25f);
    f.clearcoatAlpha = lerpf_(0.1f, 0.001f, hd.clearcoatGloss);
    return f;
}
template <bool FM>
__device__ __forceinline__ F3 disneyEval(const HitData& hd, const BrdfFrame& f, F3 L) {                    // :179-253
    if (!(hd.transmission < 1.0f && dot(f.N, L) > 0.0f && f.frontV)) return f3(0.f);
    const F3 H = M<FM>::normalized(L + f.V);
    const float NDotL = fabsf(dot(f.N, L)), NDotV = f.NDotV, NDotH = fabsf(dot(f.N, H)), LDotH = fabsf(dot(L, H));
    const float FL = schlick(NDotL), FV = schlick(NDotV);
    const float Fd90 = 0.5f + 2.0f * LDotH * LDotH * hd.roughness;
    const float Fd = lerpf_(1.0f, Fd90, FL) * lerpf_(1.0f, Fd90, FV);
    const float Fss90 = LDotH * LDotH * hd.roughness;
    const float Fss = lerpf_(1.0f, Fss90, FL) * lerpf_(1.0f, Fss90, FV);
    const float ss = 1.25f * (Fss * (M<FM>::rcp(NDotL + NDotV) - 0.5f) + 0.5f);
    const float Ds = GTR2aniso<FM>(NDotH, dot(H, f.T), dot(H, f.B), f.ax, f.ay);
    const float FH = schlick(LDotH);
    const F3 Fs = lerp3(f.Cspec0, f3(1.0f), FH);
    float Gs = smithGaniso<FM>(NDotL, dot(L, f.T), dot(L, f.B), f.ax, f.ay);
    Gs *= f.GsV;
    const F3 Fsheen = (FH * hd.sheen) * f.Csheen;
    // clearcoat lobe: skipped when the material has none (its weight is 0.25*clearcoat; a NaN/inf lobe times 0 cannot
    // occur because GTR1/SmithG are finite for N.L, N.V > 0, which the guard above enforces)
    float coat = 0.f;
    if (hd.clearcoat != 0.f) {
        const float Dr = GTR1<FM>(NDotH, f.clearcoatAlpha);
        const float Fr = lerpf_(0.04f, 1.0f, FH);
        const float Gr = smithG<FM>(NDotL, 0.25f) * f.GrV;
        coat = 0.25f * hd.clearcoat * Gr * Fr * Dr;
    }
    const F3 diffuse = (((1.0f / EL_PI) * lerpf_(Fd, ss, hd.subsurface)) * hd.albedo + Fsheen) * (1.0f - hd.metallic);
    const F3 spec = (Gs * Fs) * Ds;
    return f3(diffuse.x + spec.x + coat, diffuse.y + spec.y + coat, diffuse.z + spec.z + coat);
}
template <bool FM>
__device__ __forceinline__ float disneyPdf(const HitData& hd, const BrdfFrame& f, F3 L) {                  // :108-147
    if (dot(f.N, L) <= 0.0f) return 1.0f;
    const F3 H = M<FM>::normalized(L + f.V);
    const float NDotH = fabsf(dot(f.N, H));
    const float specularRatio = 1.0f - f.diffuseRatio;
    const float pdfGTR2 = GTR2aniso<FM>(NDotH, dot(H, f.T), dot(H, f.B), f.ax, f.ay) * NDotH;
    const float pdfGTR1 = GTR1<FM>(NDotH, f.clearcoatAlpha) * NDotH;
    const float ratio = M<FM>::rcp(1.0f + hd.clearcoat);
    const float pdfSpec = M<FM>::div(lerpf_(pdfGTR1, pdfGTR2, ratio), 4.0f * fabsf(dot(L, H)));
    const float pdfDiff = fabsf(dot(L, f.N)) * (1.0f / EL_PI);
    return f.diffuseRatio * pdfDiff + specularRatio * pdfSpec;
}
template <bool FM>
__device__ __forceinline__ F3 disneySample(const HitData& hd, const BrdfFrame& f, float r1, float r2, float r3) {   // :150-177
    if (r3 < f.diffuseRatio) {
        const F3 H = cosineSampleHemisphere<FM>(r1, r2);
        return f.T * H.x + f.B * H.y + f.N * H.z;
    }
    F3 H = importanceSampleGGX<FM>(hd.roughness, r1, r2);
    H = f.T * H.x + f.B * H.y + f.N * H.z;
    const F3 I = f3(-f.V.x, -f.V.y, -f.V.z);
    return I - (2 * dot(I, H)) * H;                                                                        // reflect, S/Vector.hpp:214
}

// ---- S/kernel.cu:54-119 generateHitData ---------------------------------------------------------------
template <bool FM>
__device__ __forceinline__ void generateHitData(const DevScene& S, const DevMaterial& m, const DevPackedMaps& pk, HitData& hd, F3 normal, F3 tangent, F3 bitangent, float tu, float tv) {
    if (pk.data) {                                                   // all four maps in one 8-byte record (see DevPackedMaps)
        DevTex g; g.data = nullptr; g.width = pk.width; g.height = pk.height; g.xTile = pk.xTile; g.yTile = pk.yTile; g.xOffset = pk.xOffset; g.yOffset = pk.yOffset;
        g.format = 0; g.filter = 0;
        const uint2 c = __ldg(pk.data + texelIndex(g, (int)(tu * pk.width), (int)(tv * pk.height)));
        const float* la = S.lut + (pk.albedoFormat == ELEVEN_TEX_U8_SRGB ? 0 : 256);
        const float* lr = S.lut + (pk.roughFormat == ELEVEN_TEX_U8_SRGB ? 0 : 256);
        const float* lm = S.lut + (pk.metalFormat == ELEVEN_TEX_U8_SRGB ? 0 : 256);
        const float* ln_ = S.lut + (pk.normalFormat == ELEVEN_TEX_U8_SRGB ? 0 : 256);
        hd.albedo = f3(__ldg(la + (c.x & 0xffu)), __ldg(la + ((c.x >> 8) & 0xffu)), __ldg(la + ((c.x >> 16) & 0xffu)));
        hd.roughness = __ldg(lr + (c.x >> 24));
        const F3 nc = f3(__ldg(ln_ + (c.y & 0xffu)), __ldg(ln_ + ((c.y >> 8) & 0xffu)), __ldg(ln_ + ((c.y >> 16) & 0xffu)));
        hd.metallic = __ldg(lm + (c.y >> 24));
        hd.emission = m.emissionTex < 0 ? f3(m.emission[0], m.emission[1], m.emission[2]) : texFiltered(S.textures[m.emissionTex], S.lut, tu, tv);
        const F3 ln = f3(nc.x * 2 - 1, nc.y * 2 - 1, nc.z * 2 - 1);
        hd.normal = M<FM>::normalized(ln.x * tangent - ln.y * bitangent + ln.z * normal);
        hd.roughness = M<FM>::pow(hd.roughness, 2.2f);
        hd.metallic = M<FM>::pow(hd.metallic, 2.2f);
        hd.clearcoatGloss = m.clearcoatGloss; hd.clearcoat = m.clearcoat; hd.anisotropic = m.anisotropic; hd.eta = m.eta;
        hd.transmission = m.transmission; hd.specular = m.specular; hd.specularTint = m.specularTint; hd.sheenTint = m.sheenTint;
        hd.subsurface = m.subsurface; hd.sheen = m.sheen;
        return;
    }
    // Memory-level parallelism: the texel fetches are gathers into ~1 GB of maps (DRAM latency each).  Fetched one after
    // the other, each behind the decode of the previous one, they were four of the five hottest stall sites of k_shade
    // (ncu source view, round 1).  When every bound map is 8-bit and unfiltered (what the loader produces) all texel loads
    // are issued back to back, then decoded.
    const int ids[5] = {m.albedoTex, m.emissionTex, m.roughnessTex, m.metallicTex, m.normalTex};
    bool batch = true;
    uint32_t fmt[5];
#pragma unroll
    for (int k = 0; k < 5; k++) {
        fmt[k] = 0;
        if (ids[k] >= 0) {
            const DevTex& T = S.textures[ids[k]];
            fmt[k] = T.format;
            batch = batch && T.format != ELEVEN_TEX_F32_RGB && (k == 4 || T.filter == 0);
        }
    }
    if (batch) {
        uint32_t c[5];                             // texels kept as packed words until all five loads are in flight
#pragma unroll
        for (int k = 0; k < 5; k++) {
            c[k] = 0u;
            if (ids[k] >= 0) { const DevTex& T = S.textures[ids[k]]; c[k] = __ldg((const uint32_t*)T.data + texelIndex(T, (int)(tu * T.width), (int)(tv * T.height))); }
        }
        F3 val[5];
#pragma unroll
        for (int k = 0; k < 5; k++) {
            const float* l = S.lut + (fmt[k] == ELEVEN_TEX_U8_SRGB ? 0 : 256);
            val[k] = ids[k] >= 0 ? f3(__ldg(l + (c[k] & 0xffu)), __ldg(l + ((c[k] >> 8) & 0xffu)), __ldg(l + ((c[k] >> 16) & 0xffu))) : f3(0.f);
        }
        hd.albedo = ids[0] < 0 ? f3(m.albedo[0], m.albedo[1], m.albedo[2]) : val[0];
        hd.emission = ids[1] < 0 ? f3(m.emission[0], m.emission[1], m.emission[2]) : val[1];
        hd.roughness = ids[2] < 0 ? m.roughness : val[2].x;
        hd.metallic = ids[3] < 0 ? m.metallic : val[3].x;
        if (ids[4] < 0) hd.normal = normal;
        else {
            const F3 ln = f3(val[4].x * 2 - 1, val[4].y * 2 - 1, val[4].z * 2 - 1);
            hd.normal = M<FM>::normalized(ln.x * tangent - ln.y * bitangent + ln.z * normal);
        }
    } else {
        hd.albedo = m.albedoTex < 0 ? f3(m.albedo[0], m.albedo[1], m.albedo[2]) : texFiltered(S.textures[m.albedoTex], S.lut, tu, tv);
        hd.emission = m.emissionTex < 0 ? f3(m.emission[0], m.emission[1], m.emission[2]) : texFiltered(S.textures[m.emissionTex], S.lut, tu, tv);
        hd.roughness = m.roughnessTex < 0 ? m.roughness : texFiltered(S.textures[m.roughnessTex], S.lut, tu, tv).x;
        hd.metallic = m.metallicTex < 0 ? m.metallic : texFiltered(S.textures[m.metallicTex], S.lut, tu, tv).x;
        if (m.normalTex < 0) hd.normal = normal;
        else {
            const F3 nc = texFromUV(S.textures[m.normalTex], S.lut, tu, tv);
            const F3 ln = f3(nc.x * 2 - 1, nc.y * 2 - 1, nc.z * 2 - 1);
            hd.normal = M<FM>::normalized(ln.x * tangent - ln.y * bitangent + ln.z * normal);
        }
    }
    hd.roughness = M<FM>::pow(hd.roughness, 2.2f);      // "linear to sRGB", applied to constants too (S/kernel.cu:103-104)
    hd.metallic = M<FM>::pow(hd.metallic, 2.2f);
    hd.clearcoatGloss = m.clearcoatGloss; hd.clearcoat = m.clearcoat; hd.anisotropic = m.anisotropic; hd.eta = m.eta;
    hd.transmission = m.transmission; hd.specular = m.specular; hd.specularTint = m.specularTint; hd.sheenTint = m.sheenTint;
    hd.subsurface = m.subsurface; hd.sheen = m.sheen;
}

// ---- S/kernel.cu:260-337 calculateCameraRay --------------------------------------------------------------
// The three Euler rotations are per-frame constants: sin/cos are hoisted to the host (CamRot), the reference
// recomputes six sin/cos per ray (S/kernel.cu:304-306).
struct CamRot { float sx, cx, sy, cy, sz, cz; };
template <bool FM>
__device__ __forceinline__ Ray cameraRay(const DevCamera& c, const CamRot& R, int x, int y, float r1, float r2, float r3, float r4, float r5) {
    const F3 pos = f3(c.pos[0], c.pos[1], c.pos[2]);
    const float dx = pos.x + ((float)x) / ((float)c.xRes) * c.sensorWidth;
    const float dy = pos.y + ((float)y) / ((float)c.yRes) * c.sensorHeight;
    const float odx = (float)((double)(-c.sensorWidth) / 2.0 + (double)dx);
    const float ody = (float)((double)(-c.sensorHeight) / 2.0 + (double)dy);
    const float rx = (float)((1.0 / (double)(float)c.xRes) * ((double)r1 - 0.5) * (double)c.sensorWidth);
    const float ry = (float)((1.0 / (double)(float)c.yRes) * ((double)r2 - 0.5) * (double)c.sensorHeight);
    const float SPx = odx + rx, SPy = ody + ry, SPz = pos.z + c.focalLength;
    const F3 dir = f3(SPx, SPy, SPz) - pos;
    const F3 dX = f3(dir.x, dir.y * R.cx - dir.z * R.sx, dir.y * R.sx + dir.z * R.cx);
    const F3 dY = f3(dX.x * R.cy + dX.z * R.sy, dX.y, dX.z * R.cy - dX.x * R.sy);
    const F3 dZ = f3(dY.x * R.cz - dY.y * R.sz, dY.x * R.sz + dY.y * R.cz, dY.z);
    Ray ray = makeRay(pos, dZ);
    if (c.bokeh) {
        const float diameter = c.focalLength / c.aperture;
        const float l = c.focusDistance + c.focalLength;
        const F3 focusPoint = ray.o + ray.d * l;
        float ix, iy;
        uniformCircleSampling<FM>(r3, r4, r5, ix, iy);
        ix *= diameter * 0.5f; iy *= diameter * 0.5f;
        const F3 orig = pos + f3(ix, iy, 0.f);
        ray = makeRay(orig, focusPoint - orig);
    }
    return ray;
}

} // namespace eleven
