/*
 * scene_upload.cuh — device side of eleven_scene_upload: texel re-layouts that renderSetup's host loops used to do.
 *
 * The reference uploads every texture as float RGB (S/kernel.cu:515-560).  Here the caller's bytes go to the device as they are
 * (3 B/texel for 8-bit maps, 12 B/texel for float maps) and are expanded there to the 4- and 16-byte texels the shading kernels
 * gather: a 4096^2 map is a 50 MB copy + a 67 MB device write instead of a 16.7 M-iteration host loop, a 67 MB temporary and a 67 MB
 * pageable copy (round 1: 1.3-1.7 s of eleven_scene_upload for the 12 maps of the benchmark scene, against 0.27 s of rendering for
 * 1000 spp on 8 GPUs).  The packed material-map records (shading.cuh: DevPackedMaps) are interleaved on the device as well.
 */
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace eleven {

__global__ void k_expandRgb8(const uint8_t* __restrict__ rgb, uchar4* __restrict__ out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = make_uchar4(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2], 255);
}
// float RGB -> float4 with .w = (r + g) + b, the weight HDRI::pdf and the CDF use (S/HDRI.hpp:112-125,149)
__global__ void k_expandRgbF32(const float* __restrict__ rgb, float4* __restrict__ out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const float r = rgb[3 * i], g = rgb[3 * i + 1], b = rgb[3 * i + 2]; out[i] = make_float4(r, g, b, __fadd_rn(__fadd_rn(r, g), b)); }
}
// one 8-byte record per texel: albedo rgb + roughness | normal rgb + metallic (first channel of the single-channel maps)
__global__ void k_packMaps(const uchar4* __restrict__ albedo, const uchar4* __restrict__ rough, const uchar4* __restrict__ metal,
                           const uchar4* __restrict__ normal, uint2* __restrict__ out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uchar4 a = albedo[i], r = rough[i], m = metal[i], nn = normal[i];
    out[i] = make_uint2((uint32_t)a.x | ((uint32_t)a.y << 8) | ((uint32_t)a.z << 16) | ((uint32_t)r.x << 24),
                        (uint32_t)nn.x | ((uint32_t)nn.y << 8) | ((uint32_t)nn.z << 16) | ((uint32_t)m.x << 24));
}

} // namespace eleven
