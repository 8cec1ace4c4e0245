/*
 * flat_scene.h — the "ELVNSCN1" flat binary scene container shared by the Python scene
 * generators (tfg-pathtracer_b200/scenes.py), the reference harnesses in this directory and the
 * C++ host (tfg-pathtracer_b200/host).  It is the ElevenSceneDesc of include/eleven_b200.h laid
 * out in a file: no reference code, just POD records, little-endian.
 *
 *   char  magic[8] = "ELVNSCN1"
 *   ElevenCamera                       (56 B)
 *   u32 triCount      ; ElevenTri[triCount]            (152 B each)
 *   u32 objectCount   ; i32 objectMaterial[objectCount]
 *   u32 materialCount ; ElevenMaterial[materialCount]  (108 B each)
 *   u32 textureCount  ; textureCount x { FlatTexHeader ; texels }
 *   FlatTexHeader hdri ; texels (must be F32_RGB)
 *   u32 lightCount    ; ElevenPointLight[lightCount]   (24 B each)
 * texels = width*height*3 floats (format 0) or bytes (formats 1,2); format 3 = external, no texels follow.
 */
#ifndef ELEVEN_FLAT_SCENE_H
#define ELEVEN_FLAT_SCENE_H
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <vector>
#include "../../include/eleven_b200.h"

struct FlatTexHeader { uint32_t format; int32_t width, height; float xTile, yTile, xOffset, yOffset; uint32_t filter; };

struct FlatScene {
    ElevenCamera camera;
    std::vector<ElevenTri> tris;
    std::vector<int32_t> objectMaterial;
    std::vector<ElevenMaterial> materials;
    std::vector<FlatTexHeader> texHeaders;
    std::vector<std::vector<uint8_t> > texData;
    FlatTexHeader hdriHeader;
    std::vector<uint8_t> hdriData;
    std::vector<ElevenPointLight> lights;
    std::vector<ElevenTexture> texViews;   // filled by desc()

    static size_t texBytes(const FlatTexHeader& h) {
        if (h.format == 3u) return 0;   /* external: texels live in the image file listed in <scene>.textures.txt */
        return (size_t)h.width * h.height * 3 * (h.format == ELEVEN_TEX_F32_RGB ? 4 : 1);
    }
    bool load(const char* path) {
        FILE* f = fopen(path, "rb"); if (!f) return false;
        char magic[8]; uint32_t n;
        bool ok = fread(magic, 1, 8, f) == 8 && memcmp(magic, "ELVNSCN1", 8) == 0;
        ok = ok && fread(&camera, sizeof camera, 1, f) == 1;
        ok = ok && fread(&n, 4, 1, f) == 1; if (ok) { tris.resize(n); ok = n == 0 || fread(tris.data(), sizeof(ElevenTri), n, f) == n; }
        ok = ok && fread(&n, 4, 1, f) == 1; if (ok) { objectMaterial.resize(n); ok = n == 0 || fread(objectMaterial.data(), 4, n, f) == n; }
        ok = ok && fread(&n, 4, 1, f) == 1; if (ok) { materials.resize(n); ok = n == 0 || fread(materials.data(), sizeof(ElevenMaterial), n, f) == n; }
        ok = ok && fread(&n, 4, 1, f) == 1;
        if (ok) {
            texHeaders.resize(n); texData.resize(n);
            for (uint32_t i = 0; ok && i < n; i++) {
                ok = fread(&texHeaders[i], sizeof(FlatTexHeader), 1, f) == 1;
                if (ok) { texData[i].resize(texBytes(texHeaders[i])); ok = fread(texData[i].data(), 1, texData[i].size(), f) == texData[i].size(); }
            }
        }
        ok = ok && fread(&hdriHeader, sizeof hdriHeader, 1, f) == 1;
        if (ok) { hdriData.resize(texBytes(hdriHeader)); ok = fread(hdriData.data(), 1, hdriData.size(), f) == hdriData.size(); }
        ok = ok && fread(&n, 4, 1, f) == 1; if (ok) { lights.resize(n); ok = n == 0 || fread(lights.data(), sizeof(ElevenPointLight), n, f) == n; }
        fclose(f);
        return ok;
    }
    bool save(const char* path) const {
        FILE* f = fopen(path, "wb"); if (!f) return false;
        uint32_t n;
        fwrite("ELVNSCN1", 1, 8, f); fwrite(&camera, sizeof camera, 1, f);
        n = (uint32_t)tris.size(); fwrite(&n, 4, 1, f); fwrite(tris.data(), sizeof(ElevenTri), n, f);
        n = (uint32_t)objectMaterial.size(); fwrite(&n, 4, 1, f); fwrite(objectMaterial.data(), 4, n, f);
        n = (uint32_t)materials.size(); fwrite(&n, 4, 1, f); fwrite(materials.data(), sizeof(ElevenMaterial), n, f);
        n = (uint32_t)texHeaders.size(); fwrite(&n, 4, 1, f);
        for (uint32_t i = 0; i < n; i++) { fwrite(&texHeaders[i], sizeof(FlatTexHeader), 1, f); fwrite(texData[i].data(), 1, texData[i].size(), f); }
        fwrite(&hdriHeader, sizeof hdriHeader, 1, f); fwrite(hdriData.data(), 1, hdriData.size(), f);
        n = (uint32_t)lights.size(); fwrite(&n, 4, 1, f); fwrite(lights.data(), sizeof(ElevenPointLight), n, f);
        bool ok = ferror(f) == 0; fclose(f); return ok;
    }
    static ElevenTexture view(const FlatTexHeader& h, const void* data) {
        ElevenTexture t; t.data = data; t.format = h.format; t.width = h.width; t.height = h.height;
        t.xTile = h.xTile; t.yTile = h.yTile; t.xOffset = h.xOffset; t.yOffset = h.yOffset; t.filter = h.filter; return t;
    }
    ElevenSceneDesc desc() {
        ElevenSceneDesc d; memset(&d, 0, sizeof d);
        d.camera = camera;
        d.triCount = (uint32_t)tris.size(); d.tris = tris.data();
        d.objectCount = (uint32_t)objectMaterial.size(); d.objectMaterial = objectMaterial.data();
        d.materialCount = (uint32_t)materials.size(); d.materials = materials.data();
        texViews.clear();
        for (size_t i = 0; i < texHeaders.size(); i++) texViews.push_back(view(texHeaders[i], texData[i].data()));
        d.textureCount = (uint32_t)texViews.size(); d.textures = texViews.data();
        d.hdri = view(hdriHeader, hdriData.data());
        d.pointLightCount = (uint32_t)lights.size(); d.pointLights = lights.data();
        return d;
    }
};
#endif
