/*
 * scene_loader.h — C++ host: loads a scene directory in the reference's on-disk format into the flat
 * ElevenSceneDesc of include/eleven_b200.h.
 *
 * Same inputs and semantics as the reference's loader (S/SceneLoader.hpp:7-144, S/ObjLoader.hpp:20-208),
 * written from the format description, not from its code:
 *   scene.json  camera{xRes,yRes,position,rotation,focalLength,focusDistance,aperture,bokeh}, hdri{name|color,xOffset,
 *               yOffset}, pointLights[{position,radiance}]            (tolerant parser: quotes optional on keys)
 *   scene.mtl   newmtl / Kd / Ks(.x -> specular) / Ke / Ni / d / map_Kd (sRGB) / map_Ns / refl / map_Bump (linear)
 *   scene.obj   `o` blocks; v / vn with z negated, vn normalised per use; vt; usemtl (last wins); triangular `f a/b/c`
 *   textures    24-bit .bmp kept as 8-bit RGB (the library decodes them with the reference's fastPow tables), row 0 =
 *               bottom image row (the reference flips on load, S/Texture.hpp:49); texture ids in the reference's order
 *               (per material: map_Bump, map_Kd, map_Ns, refl — its std::map iteration order), de-duplicated by path
 *   HDRI/<name>.hdr  Radiance RGBE (flat or new-RLE), not flipped; or a 1024x1024 constant colour (S/HDRI.hpp:22-38)
 *   tangents    Mikkelsen tangent space per `o` object, as the reference computes it through its vendored mikktspace
 *               (tangent_space.cpp; pinned against the reference loader's own output, tests/test_host_cli.py)
 */
#ifndef ELEVEN_SCENE_LOADER_H
#define ELEVEN_SCENE_LOADER_H

#include <string>
#include <vector>
#include "../../include/eleven_b200.h"

namespace eleven_host {

struct LoadedTexture { std::vector<unsigned char> bytes; std::vector<float> floats; ElevenTexture view; std::string path; };

struct LoadedScene {
    ElevenCamera camera;
    std::vector<ElevenTri> tris;
    std::vector<int32_t> objectMaterial;
    std::vector<std::string> objectNames, materialNames;
    std::vector<ElevenMaterial> materials;
    std::vector<LoadedTexture> textures;
    LoadedTexture hdri;
    std::vector<ElevenPointLight> lights;
    std::vector<ElevenTexture> textureViews;
    ElevenSceneDesc desc();          /* pointers into this object: keep it alive until eleven_scene_upload returns */
};

/* Loads <dir>/scene.json + scene.mtl + scene.obj (reference layout) or, if `path` is a file, an ELVNSCN1 flat container. */
bool loadScene(const std::string& path, LoadedScene& out, std::string& err);
bool saveFlat(const LoadedScene& s, const std::string& path, std::string& err);

/* Per-corner tangents + Tri::tangentsSign of one mesh object (S/ObjLoader.hpp:167-168 -> S/mikktspaceCallback.hpp). */
void computeTangentSpace(ElevenTri* tris, size_t n);

bool readBmp24(const std::string& path, int& w, int& h, std::vector<unsigned char>& rgbBottomUp, std::string& err);
bool readHdr(const std::string& path, int& w, int& h, std::vector<float>& rgbTopDown, std::string& err);
bool writeBmp24(const std::string& path, int w, int h, const unsigned char* rgba8TopDown, std::string& err);

} // namespace eleven_host
#endif
