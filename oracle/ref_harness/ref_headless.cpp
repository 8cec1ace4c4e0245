/*
 * ref_headless.cpp — headless driver for THE UNMODIFIED REFERENCE RENDERER (its loader, its BVH
 * builder, its kernel.cu), replacing S/main.cpp, which needs <Windows.h>, GLFW, glad and OIDN
 * (S/main.cpp:8,14,26-33).  Glue written for this repo: loadScene -> renderSetup -> timed
 * renderCuda -> getBuffers -> raw float dumps.  All rendering code is #included / compiled from
 * /root/reference/src/tfg-pathtracer by oracle/Makefile (kernel.cu needs four one-token syntax
 * fixes to compile with nvcc 12.9 and SceneLoader.hpp a '\\' -> '/' path separator fix; both are
 * applied by sed to a temporary copy at build time, see the Makefile — nothing is copied into
 * the repo).  Output: oracle/_ref/eleven_ref_headless_{precise,fast} (git-ignored binaries).
 *
 * usage: eleven_ref_headless <scene_dir> <spp> <out_prefix> [--dump-scene file] [--external-textures] [--warmup spp] [--snapshots a,b,c]
 *   --snapshots: also write <out_prefix>.beauty@<n>.f32 when the film holds n samples (n ascending, < spp): renderCuda is called
 *   in segments, which the reference's running-mean film supports (its own --warmup path does the same)
 *   <out_prefix>.beauty.f32 / .normal.f32 / .tangent.f32 / .bitangent.f32  (W*H*4 floats each)
 *   <out_prefix>.pathcount.i32, <out_prefix>.json (timings, counters)
 */
#include <cstring>
#include <cmath>
#include <map>
#include <chrono>
#include <string>
#include <cuda_runtime.h>
#include "Texture.hpp"
#include "Definitions.h"
#include "kernel.h"
#include "SceneLoader.hpp"
#include "mikktspaceCallback.hpp"
#include "flat_scene.h"

static_assert(sizeof(Tri) == sizeof(ElevenTri), "ElevenTri must mirror Tri (S/Tri.hpp:13-19)");

static double nowMs() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
static void writeRaw(const std::string& path, const void* p, size_t bytes) {
    FILE* f = fopen(path.c_str(), "wb"); if (!f) { fprintf(stderr, "cannot write %s\n", path.c_str()); exit(3); }
    fwrite(p, 1, bytes, f); fclose(f);
}

static void dumpScene(Scene& s, const char* path, bool externalTextures) {
    FlatScene fs;
    Camera& c = s.camera;
    fs.camera.xRes = c.xRes; fs.camera.yRes = c.yRes; fs.camera.focalLength = c.focalLength;
    fs.camera.sensorWidth = c.sensorWidth; fs.camera.sensorHeight = c.sensorHeight; fs.camera.aperture = c.aperture;
    fs.camera.focusDistance = c.focusDistance;
    fs.camera.rotation[0] = c.rotation.x; fs.camera.rotation[1] = c.rotation.y; fs.camera.rotation[2] = c.rotation.z;
    fs.camera.position[0] = c.position.x; fs.camera.position[1] = c.position.y; fs.camera.position[2] = c.position.z;
    fs.camera.bokeh = c.bokeh ? 1 : 0;
    fs.tris.resize(s.tris.size());
    if (!s.tris.empty()) memcpy(fs.tris.data(), s.tris.data(), s.tris.size() * sizeof(Tri));
    for (auto& mo : s.meshObjects) fs.objectMaterial.push_back(mo.materialID);
    std::vector<int> srgb(s.textures.size(), 0);
    for (auto& m : s.materials) {
        ElevenMaterial e;
        e.albedoTextureID = m.albedoTextureID; e.emissionTextureID = m.emissionTextureID; e.roughnessTextureID = m.roughnessTextureID;
        e.metallicTextureID = m.metallicTextureID; e.normalTextureID = m.normalTextureID; e.opacityTextureID = m.opacityTextureID;
        e.albedo[0] = m.albedo.x; e.albedo[1] = m.albedo.y; e.albedo[2] = m.albedo.z;
        e.emission[0] = m.emission.x; e.emission[1] = m.emission.y; e.emission[2] = m.emission.z;
        e.opacity[0] = m.opacity.x; e.opacity[1] = m.opacity.y; e.opacity[2] = m.opacity.z;
        e.roughness = m.roughness; e.metallic = m.metallic; e.clearcoatGloss = m.clearcoatGloss; e.clearcoat = m.clearcoat;
        e.anisotropic = m.anisotropic; e.eta = m.eta; e.transmission = m.transmission; e.specular = m.specular;
        e.specularTint = m.specularTint; e.sheenTint = m.sheenTint; e.subsurface = m.subsurface; e.sheen = m.sheen;
        fs.materials.push_back(e);
        if (m.albedoTextureID >= 0) srgb[m.albedoTextureID] = 1;
    }
    std::string listing;
    for (size_t i = 0; i < s.textures.size(); i++) {
        Texture& t = s.textures[i];
        FlatTexHeader h; h.format = externalTextures ? 3u : (uint32_t)ELEVEN_TEX_F32_RGB; h.width = t.width; h.height = t.height;
        h.xTile = t.xTile; h.yTile = t.yTile; h.xOffset = t.xOffset; h.yOffset = t.yOffset; h.filter = (uint32_t)t.filter;
        fs.texHeaders.push_back(h);
        std::vector<uint8_t> d;
        if (!externalTextures) { d.resize((size_t)t.width * t.height * 12); memcpy(d.data(), t.data, d.size()); }
        fs.texData.push_back(d);
        listing += std::to_string(i) + " " + (srgb[i] ? "srgb " : "linear ") + t.path + "\n";
    }
    Texture& ht = s.hdri.texture;
    fs.hdriHeader.format = ELEVEN_TEX_F32_RGB; fs.hdriHeader.width = ht.width; fs.hdriHeader.height = ht.height;
    fs.hdriHeader.xTile = ht.xTile; fs.hdriHeader.yTile = ht.yTile; fs.hdriHeader.xOffset = ht.xOffset; fs.hdriHeader.yOffset = ht.yOffset;
    fs.hdriHeader.filter = (uint32_t)ht.filter;
    fs.hdriData.resize((size_t)ht.width * ht.height * 12); memcpy(fs.hdriData.data(), ht.data, fs.hdriData.size());
    for (auto& pl : s.pointLights) {
        ElevenPointLight e; e.position[0] = pl.position.x; e.position[1] = pl.position.y; e.position[2] = pl.position.z;
        e.radiance[0] = pl.radiance.x; e.radiance[1] = pl.radiance.y; e.radiance[2] = pl.radiance.z; fs.lights.push_back(e);
    }
    if (!fs.save(path)) { fprintf(stderr, "cannot write %s\n", path); exit(3); }
    writeRaw(std::string(path) + ".textures.txt", listing.data(), listing.size());
}

int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: %s <scene_dir> <spp> <out_prefix> [--dump-scene file] [--external-textures]\n", argv[0]); return 1; }
    std::string dir = argv[1]; int spp = atoi(argv[2]); std::string out = argv[3];
    const char* dump = 0; bool ext = false; int warm = 0; std::vector<int> snaps;
    for (int i = 4; i < argc; i++) {
        if (!strcmp(argv[i], "--dump-scene") && i + 1 < argc) dump = argv[++i];
        else if (!strcmp(argv[i], "--warmup") && i + 1 < argc) warm = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--external-textures")) ext = true;
        else if (!strcmp(argv[i], "--snapshots") && i + 1 < argc) { for (char* t = strtok(argv[++i], ","); t; t = strtok(0, ",")) snaps.push_back(atoi(t)); }
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        if (!dump) { fprintf(stderr, "no CUDA device\n"); return 4; }
    }
    double t0 = nowMs();
    Scene scene = loadScene(dir);
    double t1 = nowMs();
    if (dump) dumpScene(scene, dump, ext);
    if (ndev == 0) { printf("{\"load_ms\": %.3f, \"dump_only\": true}\n", t1 - t0); return 0; }
    int W = scene.camera.xRes, H = scene.camera.yRes;
    double t2 = nowMs();
    cudaError_t e = renderSetup(&scene);
    double t3 = nowMs();
    if (e != cudaSuccess) { fprintf(stderr, "renderSetup: %s\n", cudaGetErrorString(e)); return 5; }
    const double setupMs = t3 - t2;
    if (warm > 0) { renderCuda(&scene, warm); cudaDeviceSynchronize(); t3 = nowMs(); }     // untimed warm-up samples
    int done = 0;
    for (int sn : snaps) {
        if (sn <= done || sn >= spp) continue;
        renderCuda(&scene, sn - done); cudaDeviceSynchronize(); done = sn;
        RenderData sd; sd.pars = RenderParameters(W, H, sn);
        for (int i = 0; i < PASSES_COUNT; i++) sd.passes[i] = new float[(size_t)W * H * 4];
        std::vector<int> spc((size_t)W * H);
        getBuffers(sd, spc.data(), W * H); cudaDeviceSynchronize();
        writeRaw(out + ".beauty@" + std::to_string(sn) + ".f32", sd.passes[BEAUTY], (size_t)W * H * 16);
        for (int i = 0; i < PASSES_COUNT; i++) delete[] sd.passes[i];
    }
    renderCuda(&scene, spp - done);
    e = cudaDeviceSynchronize();
    double t4 = nowMs();
    if (e != cudaSuccess) { fprintf(stderr, "renderCuda: %s\n", cudaGetErrorString(e)); return 5; }
    RenderData data; data.pars = RenderParameters(W, H, spp);
    for (int i = 0; i < PASSES_COUNT; i++) { data.passes[i] = new float[(size_t)W * H * 4]; memset(data.passes[i], 0, (size_t)W * H * 16); }
    std::vector<int> pc((size_t)W * H);
    getBuffers(data, pc.data(), W * H);
    cudaDeviceSynchronize();
    int samples = getSamples();
    cudaDeviceSynchronize();
    writeRaw(out + ".beauty.f32", data.passes[BEAUTY], (size_t)W * H * 16);
    writeRaw(out + ".normal.f32", data.passes[NORMAL], (size_t)W * H * 16);
    writeRaw(out + ".tangent.f32", data.passes[TANGENT], (size_t)W * H * 16);
    writeRaw(out + ".bitangent.f32", data.passes[BITANGENT], (size_t)W * H * 16);
    writeRaw(out + ".pathcount.i32", pc.data(), pc.size() * 4);
    long long paths = 0; for (int v : pc) paths += v;
    double renderMs = t4 - t3;
    char js[1024];
    snprintf(js, sizeof js,
        "{\"impl\": \"reference-cuda\", \"width\": %d, \"height\": %d, \"spp\": %d, \"tris\": %d, \"load_ms\": %.3f, "
        "\"setup_ms\": %.3f, \"render_ms\": %.3f, \"samples_per_s\": %.1f, \"hit_bounces\": %lld, \"kpaths_per_s\": %.3f, \"samples_pixel0\": %d}\n",
        W, H, spp, (int)scene.tris.size(), t1 - t0, setupMs, renderMs, (double)W * H * spp / (renderMs * 1e-3), paths, paths / renderMs, samples);
    fputs(js, stdout);
    writeRaw(out + ".json", js, strlen(js));
    return 0;
}
