/*
 * test_hooks.cuh — known-answer kernels for the device shading functions (eleven_test_* in include/eleven_b200.h).
 *
 * The reference debugs its BRDF and its environment sampling by printing tables from one-thread kernels
 * (printBRDFMaterial, printHDRISampling: S/kernel.cu:726-794).  These hooks are the same idea with the tables returned
 * to the caller: every kernel below calls the very device functions k_shade calls (shading.cuh), one record per thread,
 * so that tests can hold them against the golden vectors made from the reference's own headers (tests/golden/disney.npz,
 * scene_*.npz) and against the CPU oracle.
 */
#pragma once
#include "shading.cuh"

namespace eleven {

__device__ __forceinline__ HitData hitDataFromRecord(const float* p) {      // S/kernel.h:46-69 order, see eleven_test_disney
    HitData hd;
    hd.metallic = p[0]; hd.roughness = p[1]; hd.clearcoatGloss = p[2]; hd.clearcoat = p[3]; hd.anisotropic = p[4]; hd.eta = p[5];
    hd.transmission = p[6]; hd.specular = p[7]; hd.specularTint = p[8]; hd.sheenTint = p[9]; hd.subsurface = p[10]; hd.sheen = p[11];
    hd.emission = f3(p[12], p[13], p[14]); hd.albedo = f3(p[15], p[16], p[17]); hd.normal = f3(p[18], p[19], p[20]);
    return hd;
}

// DisneyEval / DisneyPdf / DisneySample (S/Disney.hpp:108-253) exactly as k_shade evaluates them: one BrdfFrame per hit
template <bool FM>
__global__ void k_testDisney(const float* __restrict__ rec, uint32_t n, float* __restrict__ evalPdf, float* __restrict__ sample) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* p = rec + 30 * (size_t)i;
    const HitData hd = hitDataFromRecord(p);
    const Ray ray = makeRay(f3(0.f), f3(p[21], p[22], p[23]));               // Ray's constructor normalises (S/Ray.hpp:14-18)
    const BrdfFrame bf = makeBrdfFrame<FM>(hd, ray.d);
    const F3 L = f3(p[24], p[25], p[26]);
    const F3 e = disneyEval<FM>(hd, bf, L);
    evalPdf[4 * i] = e.x; evalPdf[4 * i + 1] = e.y; evalPdf[4 * i + 2] = e.z; evalPdf[4 * i + 3] = disneyPdf<FM>(hd, bf, L);
    const F3 s = disneySample<FM>(hd, bf, p[27], p[28], p[29]);
    sample[3 * i] = s.x; sample[3 * i + 1] = s.y; sample[3 * i + 2] = s.z;
}

// the environment-sample half of hdriLight (S/kernel.cu:236-243): texel by CDF search or alias table, direction, pdf
template <bool FM>
__global__ void k_testHdri(const __grid_constant__ DevScene S, const float* __restrict__ r, const float* __restrict__ r2, uint32_t n, int envMode,
                           int32_t* __restrict__ xy, float* __restrict__ dir, float* __restrict__ pdf) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int EW = S.hdri.width, EH = S.hdri.height;
    int texel;
    if (envMode == ELEVEN_ENV_CDF) texel = cdfSearch(S.cdf, r[i], EW * EH);
    else {
        const uint32_t nT = (uint32_t)(EW * EH);
        const uint32_t k = min(nT - 1u, (uint32_t)(r[i] * (float)nT));
        const AliasEntry ae = S.alias[k];
        texel = r2[i] <= ae.prob ? (int)k : (int)ae.alias;
    }
    const int x = texel % EW, y = texel / EW;
    xy[2 * i] = x; xy[2 * i + 1] = y;
    const float nu = (float)x / (float)EW, nv = (float)y / (float)EH;
    float iu, iv; inverseTransformUV(S.hdri, nu, nv, iu, iv);
    const float4 ev = envTexelRaw(S.hdri, (int)(iu * EW), (int)(iv * EH));
    const F3 rsm = M<FM>::normalized(reverseSphericalMapping<FM>(iu, iv));
    dir[3 * i] = -rsm.x; dir[3 * i + 1] = -rsm.y; dir[3 * i + 2] = -rsm.z;
    pdf[i] = hdriPdf<FM>(S, ev, (int)(iv * EH));
}

__global__ void k_testEnvLookup(const __grid_constant__ DevScene S, const float* __restrict__ dirs, uint32_t n, float* __restrict__ rgb) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const F3 c = envLookup(S, f3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]));
    rgb[3 * i] = c.x; rgb[3 * i + 1] = c.y; rgb[3 * i + 2] = c.z;
}

// generateHitData (S/kernel.cu:54-119): material constants, texture fetch (nearest / bilinear, 8-bit / float, packed records), normal map
template <bool FM>
__global__ void k_testHitData(const __grid_constant__ DevScene S, const float* __restrict__ attrs, const int32_t* __restrict__ objectIds, uint32_t n,
                              float* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* a = attrs + 14 * (size_t)i;
    const int mid = S.objectMaterial[objectIds[i]];
    HitData hd;
    generateHitData<FM>(S, S.materials[mid], S.packed[mid], hd, f3(a[3], a[4], a[5]), f3(a[6], a[7], a[8]), f3(a[9], a[10], a[11]), a[12], a[13]);
    float* o = out + 21 * (size_t)i;
    o[0] = hd.metallic; o[1] = hd.roughness; o[2] = hd.clearcoatGloss; o[3] = hd.clearcoat; o[4] = hd.anisotropic; o[5] = hd.eta;
    o[6] = hd.transmission; o[7] = hd.specular; o[8] = hd.specularTint; o[9] = hd.sheenTint; o[10] = hd.subsurface; o[11] = hd.sheen;
    o[12] = hd.emission.x; o[13] = hd.emission.y; o[14] = hd.emission.z; o[15] = hd.albedo.x; o[16] = hd.albedo.y; o[17] = hd.albedo.z;
    o[18] = hd.normal.x; o[19] = hd.normal.y; o[20] = hd.normal.z;
}

} // namespace eleven
