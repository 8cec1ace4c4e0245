/*
 * ref_host_vectors.cpp — golden-vector generator that runs THE REFERENCE'S OWN CODE on the host.
 *
 * This file is glue written for this repo; every algorithm it exercises is #included from
 * /root/reference/src/tfg-pathtracer (never copied): Tri::hit, BVH::build/transverse/intersect,
 * DisneyEval/Pdf/Sample, HDRI::generateCDF/sample/pdf, Texture mappings, stb_image's patched
 * LDR decode, and the CUDA toolkit's curand XORWOW (host path of curand_kernel.h).
 * Built by oracle/Makefile into oracle/_ref/ref_host_vectors (git-ignored).  Used only by
 * tests/golden/make_golden.py, in the build container, to pin oracle/eleven_oracle.cpp.
 *
 * Output container: sequence of { u32 nameLen; name; u32 dtype (0 f32, 1 i32, 2 u32); u64 count; data }.
 */
#include <cstring>
#include <cmath>
#include <map>
#include <chrono>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#include <curand_kernel.h>
#include "Math.hpp"      /* the patched temporary copy (see oracle/Makefile): must come first */
#include "Texture.hpp"
#include "Definitions.h"
#include "kernel.h"
#include "Disney.hpp"
#include "mikktspaceCallback.hpp"
#include "flat_scene.h"

static_assert(sizeof(Tri) == sizeof(ElevenTri), "ElevenTri must mirror Tri (S/Tri.hpp:13-19)");
static_assert(sizeof(PointLight) == sizeof(ElevenPointLight), "PointLight layout");

static FILE* g_out;
static void put(const char* name, uint32_t dtype, const void* data, uint64_t count) {
    uint32_t nl = (uint32_t)strlen(name);
    fwrite(&nl, 4, 1, g_out); fwrite(name, 1, nl, g_out); fwrite(&dtype, 4, 1, g_out); fwrite(&count, 8, 1, g_out);
    fwrite(data, 4, count, g_out);
}
static std::vector<float> readFloats(const char* path) {
    std::vector<float> v; FILE* f = fopen(path, "rb"); if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET); v.resize(n / 4);
    if (fread(v.data(), 4, v.size(), f) != v.size()) exit(2);
    fclose(f); return v;
}

static int modeXorwow() {
    const unsigned long long subs[] = {0, 1, 2, 5, 1000, 65535, 65536, 2073599, 8294399};
    std::vector<float> u; std::vector<unsigned> st;
    for (unsigned long long s : subs) {
        curandState cs; curand_init(0, s, 0, &cs);
        for (int k = 0; k < 5; k++) st.push_back(cs.v[k]);
        st.push_back(cs.d);
        for (int i = 0; i < 32; i++) { float x = curand_uniform(&cs); u.push_back(x); }
    }
    std::vector<unsigned> subs32; for (unsigned long long s : subs) subs32.push_back((unsigned)s);
    put("subsequences", 2, subs32.data(), subs32.size());
    put("uniforms", 0, u.data(), u.size());
    put("states", 2, st.data(), st.size());
    return 0;
}

// Decode a 256x1 24-bit BMP whose pixel i is (i,i,i) through stbi_loadf exactly as Texture's ctor does.
static int modeTexel() {
    std::vector<unsigned char> bmp(54 + 256 * 3, 0);
    unsigned fileSize = (unsigned)bmp.size(), off = 54, hdr = 40, w = 256, h = 1; unsigned short planes = 1, bpp = 24;
    bmp[0] = 'B'; bmp[1] = 'M'; memcpy(&bmp[2], &fileSize, 4); memcpy(&bmp[10], &off, 4); memcpy(&bmp[14], &hdr, 4);
    memcpy(&bmp[18], &w, 4); memcpy(&bmp[22], &h, 4); memcpy(&bmp[26], &planes, 2); memcpy(&bmp[28], &bpp, 2);
    for (int i = 0; i < 256; i++) { bmp[54 + 3 * i] = bmp[54 + 3 * i + 1] = bmp[54 + 3 * i + 2] = (unsigned char)i; }
    for (int pass = 0; pass < 2; pass++) {
        stbi_ldr_to_hdr_gamma(pass == 0 ? 2.2f : 1.0f);
        int ww, hh, ch; float* d = stbi_loadf_from_memory(bmp.data(), (int)bmp.size(), &ww, &hh, &ch, 0);
        if (!d || ww != 256) { fprintf(stderr, "bmp decode failed\n"); return 2; }
        std::vector<float> t(256); for (int i = 0; i < 256; i++) t[i] = d[i * ch];
        put(pass == 0 ? "srgb" : "linear", 0, t.data(), 256);
        stbi_image_free(d);
    }
    return 0;
}

static HitData hdFrom(const float* p) {
    HitData hd;
    hd.metallic = p[0]; hd.roughness = p[1]; hd.clearcoatGloss = p[2]; hd.clearcoat = p[3]; hd.anisotropic = p[4]; hd.eta = p[5];
    hd.transmission = p[6]; hd.specular = p[7]; hd.specularTint = p[8]; hd.sheenTint = p[9]; hd.subsurface = p[10]; hd.sheen = p[11];
    hd.emission = Vector3(p[12], p[13], p[14]); hd.albedo = Vector3(p[15], p[16], p[17]); hd.normal = Vector3(p[18], p[19], p[20]);
    return hd;
}
// in: N x 30 floats (hd21, rayDir3, L3, r3)
static int modeDisney(const char* inPath) {
    std::vector<float> in = readFloats(inPath); size_t n = in.size() / 30;
    std::vector<float> ev(n * 4), sm(n * 3);
    for (size_t i = 0; i < n; i++) {
        const float* p = &in[30 * i];
        HitData hd = hdFrom(p);
        Ray r(Vector3(0), Vector3(p[21], p[22], p[23]));
        Vector3 L(p[24], p[25], p[26]);
        Vector3 e = DisneyEval(r, hd, L);
        ev[4 * i] = e.x; ev[4 * i + 1] = e.y; ev[4 * i + 2] = e.z; ev[4 * i + 3] = DisneyPdf(r, hd, L);
        Vector3 s = DisneySample(r, hd, p[27], p[28], p[29]);
        sm[3 * i] = s.x; sm[3 * i + 1] = s.y; sm[3 * i + 2] = s.z;
    }
    put("eval_pdf", 0, ev.data(), ev.size());
    put("sample", 0, sm.data(), sm.size());
    return 0;
}

// in: flat scene, rays (N x 6 floats), r values for HDRI sampling (M floats), directions (K x 3) for env lookup
static int modeScene(const char* scenePath, const char* raysPath, const char* rPath, const char* dirPath) {
    FlatScene fs; if (!fs.load(scenePath)) { fprintf(stderr, "cannot load %s\n", scenePath); return 2; }
    Scene* scene = new Scene();
    scene->camera = Camera(fs.camera.xRes, fs.camera.yRes);
    for (size_t i = 0; i < fs.tris.size(); i++) { Tri t; memcpy(&t, &fs.tris[i], sizeof t); scene->tris.push_back(t); }
    // HDRI (float RGB) -> HDRI::generateCDF
    HDRI& h = scene->hdri;
    h.texture.width = fs.hdriHeader.width; h.texture.height = fs.hdriHeader.height;
    h.texture.xTile = fs.hdriHeader.xTile; h.texture.yTile = fs.hdriHeader.yTile;
    h.texture.xOffset = fs.hdriHeader.xOffset; h.texture.yOffset = fs.hdriHeader.yOffset;
    h.texture.data = (float*)fs.hdriData.data();
    h.cdf = new float[(size_t)h.texture.width * h.texture.height + 1];
    fprintf(stderr, "[ref] generateCDF %dx%d\n", h.texture.width, h.texture.height);
    h.generateCDF();
    put("cdf", 0, h.cdf, (uint64_t)h.texture.width * h.texture.height + 1);
    put("radiance_sum", 0, &h.radianceSum, 1);

    fprintf(stderr, "[ref] buildBVH %zu tris\n", scene->tris.size());
    BVH* bvh = scene->buildBVH();
    fprintf(stderr, "[ref] built\n");
    bvh->tris = scene->getTris();
    size_t nn = (size_t)(2 << BVH_DEPTH) - 1;
    std::vector<float> boxes(nn * 6); std::vector<int> meta(nn * 3);
    for (size_t i = 0; i < nn; i++) {
        const Node& n = bvh->nodes[i];
        boxes[6 * i] = n.b1.x; boxes[6 * i + 1] = n.b1.y; boxes[6 * i + 2] = n.b1.z;
        boxes[6 * i + 3] = n.b2.x; boxes[6 * i + 4] = n.b2.y; boxes[6 * i + 5] = n.b2.z;
        meta[3 * i] = n.from; meta[3 * i + 1] = n.to; meta[3 * i + 2] = n.depth;
    }
    put("bvh_boxes", 0, boxes.data(), boxes.size());
    put("bvh_meta", 1, meta.data(), meta.size());
    put("tri_indices", 1, bvh->triIndices, fs.tris.size());

    std::vector<float> rays = readFloats(raysPath); size_t nr = rays.size() / 6;
    fprintf(stderr, "[ref] tracing %zu rays\n", nr);
    std::vector<float> full(nr * 14, 0.f), rdir(nr * 3); std::vector<int> obj(nr), valid(nr);
    for (size_t i = 0; i < nr; i++) {
        Ray r(Vector3(rays[6 * i], rays[6 * i + 1], rays[6 * i + 2]), Vector3(rays[6 * i + 3], rays[6 * i + 4], rays[6 * i + 5]));
        Hit hit = Hit();
        bvh->transverse(r, hit);                      // S/BVH.hpp:120 (the renderer's traversal), on the host
        valid[i] = hit.valid ? 1 : 0; obj[i] = hit.valid ? (int)hit.objectID : -1;
        rdir[3 * i] = r.direction.x; rdir[3 * i + 1] = r.direction.y; rdir[3 * i + 2] = r.direction.z;
        if (hit.valid) {
            float* f = &full[14 * i];
            f[0] = hit.position.x; f[1] = hit.position.y; f[2] = hit.position.z;
            f[3] = hit.normal.x; f[4] = hit.normal.y; f[5] = hit.normal.z;
            f[6] = hit.tangent.x; f[7] = hit.tangent.y; f[8] = hit.tangent.z;
            f[9] = hit.bitangent.x; f[10] = hit.bitangent.y; f[11] = hit.bitangent.z;
            f[12] = hit.tu; f[13] = hit.tv;
        }
    }
    put("hit_full", 0, full.data(), full.size());
    put("hit_valid", 1, valid.data(), valid.size());
    put("hit_obj", 1, obj.data(), obj.size());
    put("ray_dir_normalised", 0, rdir.data(), rdir.size());

    // Environment NEE: the composition hdriLight performs (S/kernel.cu:236-254) from the reference's members.
    std::vector<float> rs = readFloats(rPath);
    fprintf(stderr, "[ref] hdri sampling %zu\n", rs.size());
    std::vector<int> xy(rs.size() * 2); std::vector<float> dir(rs.size() * 3), pdf(rs.size());
    for (size_t i = 0; i < rs.size(); i++) {
        Vector3 tc = h.sample(rs[i]);
        xy[2 * i] = (int)tc.x; xy[2 * i + 1] = (int)tc.y;
        float nu = tc.x / (float)h.texture.width, nv = tc.y / (float)h.texture.height;
        float iu = h.texture.inverseTransformUV(nu, nv).x, iv = h.texture.inverseTransformUV(nu, nv).y;
        Vector3 nd = -h.texture.reverseSphericalMapping(iu, iv).normalized();
        dir[3 * i] = nd.x; dir[3 * i + 1] = nd.y; dir[3 * i + 2] = nd.z;
        pdf[i] = h.pdf(iu * h.texture.width, iv * h.texture.height);
    }
    put("hdri_xy", 1, xy.data(), xy.size());
    put("hdri_dir", 0, dir.data(), dir.size());
    put("hdri_pdf", 0, pdf.data(), pdf.size());

    // Escaped-ray environment lookup (S/kernel.cu:415-417)
    std::vector<float> ds = readFloats(dirPath); size_t nd = ds.size() / 3;
    std::vector<float> uv(nd * 2), rgb(nd * 3);
    for (size_t i = 0; i < nd; i++) {
        float u, v; Vector3 d(ds[3 * i], ds[3 * i + 1], ds[3 * i + 2]);
        Texture::sphericalMapping(Vector3(), -1 * d, 1, u, v);
        Vector3 c = h.texture.getValueFromUVFiltered(u, v);
        uv[2 * i] = u; uv[2 * i + 1] = v; rgb[3 * i] = c.x; rgb[3 * i + 1] = c.y; rgb[3 * i + 2] = c.z;
    }
    put("env_uv", 0, uv.data(), uv.size());
    put("env_rgb", 0, rgb.data(), rgb.size());
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s <xorwow|texel|disney|scene> <out> [inputs...]\n", argv[0]); return 1; }
    std::string mode = argv[1];
    g_out = fopen(argv[2], "wb"); if (!g_out) return 2;
    int rc = 1;
    if (mode == "xorwow") rc = modeXorwow();
    else if (mode == "texel") rc = modeTexel();
    else if (mode == "disney" && argc >= 4) rc = modeDisney(argv[3]);
    else if (mode == "scene" && argc >= 7) rc = modeScene(argv[3], argv[4], argv[5], argv[6]);
    fclose(g_out);
    return rc;
}
