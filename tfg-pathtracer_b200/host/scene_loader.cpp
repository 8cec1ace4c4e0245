/* scene_loader.cpp — see scene_loader.h.  Plain C++17, no third-party code. */
#include "scene_loader.h"

#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <sys/stat.h>
#include <atomic>
#include <thread>

namespace eleven_host {

// ---------------------------------------------------------------------------------------------------
// tiny JSON (objects, arrays, numbers, strings with ' or ", true/false/null, unquoted keys)
// ---------------------------------------------------------------------------------------------------
struct JVal {
    enum Kind { Null, Num, Bool, Str, Obj, Arr } kind = Null;
    double num = 0; bool b = false; std::string str;
    std::map<std::string, std::shared_ptr<JVal>> obj; std::vector<std::shared_ptr<JVal>> arr;
    const JVal* get(const std::string& k) const { auto it = obj.find(k); return it == obj.end() ? nullptr : it->second.get(); }
    double number(const std::string& k, double dflt = 0) const {
        const JVal* v = get(k); if (!v) return dflt;
        if (v->kind == Num) return v->num;
        if (v->kind == Bool) return v->b ? 1 : 0;
        if (v->kind == Str) return atof(v->str.c_str());
        return dflt;
    }
    bool boolean(const std::string& k) const { const JVal* v = get(k); return v && ((v->kind == Bool && v->b) || (v->kind == Num && v->num != 0) || (v->kind == Str && v->str == "true")); }
};
struct JParser {
    const std::string& s; size_t p = 0; std::string err;
    explicit JParser(const std::string& str) : s(str) {}
    void ws() { while (p < s.size() && (isspace((unsigned char)s[p]) || s[p] == ',')) p++; }
    std::string token() {
        ws(); std::string t;
        if (p < s.size() && (s[p] == '"' || s[p] == '\'')) { char q = s[p++]; while (p < s.size() && s[p] != q) t.push_back(s[p++]); p++; return t; }
        while (p < s.size() && (isalnum((unsigned char)s[p]) || strchr("_.+-", s[p]))) t.push_back(s[p++]);
        return t;
    }
    std::shared_ptr<JVal> value() {
        ws(); auto v = std::make_shared<JVal>();
        if (p >= s.size()) { err = "unexpected end"; return v; }
        char c = s[p];
        if (c == '{') {
            p++; v->kind = JVal::Obj;
            for (;;) { ws(); if (p >= s.size()) { err = "unterminated object"; break; } if (s[p] == '}') { p++; break; }
                std::string k = token(); ws(); if (p < s.size() && s[p] == ':') p++; else { err = "expected ':' after key " + k; break; }
                v->obj[k] = value(); if (!err.empty()) break; }
        } else if (c == '[') {
            p++; v->kind = JVal::Arr;
            for (;;) { ws(); if (p >= s.size()) { err = "unterminated array"; break; } if (s[p] == ']') { p++; break; } v->arr.push_back(value()); if (!err.empty()) break; }
        } else if (c == '"' || c == '\'') { v->kind = JVal::Str; v->str = token(); }
        else {
            std::string t = token();
            if (t == "true") { v->kind = JVal::Bool; v->b = true; } else if (t == "false") { v->kind = JVal::Bool; v->b = false; }
            else if (t == "null" || t.empty()) { v->kind = JVal::Null; if (t.empty()) { err = std::string("unexpected character '") + c + "'"; } }
            else { v->kind = JVal::Num; v->num = atof(t.c_str()); }
        }
        return v;
    }
};

static bool readFile(const std::string& path, std::string& out) {
    std::ifstream f(path, std::ios::binary); if (!f) return false;
    std::stringstream ss; ss << f.rdbuf(); out = ss.str(); return true;
}
static bool isDir(const std::string& p) { struct stat st; return stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode); }

// ---------------------------------------------------------------------------------------------------
// images
// ---------------------------------------------------------------------------------------------------
bool readBmp24(const std::string& path, int& w, int& h, std::vector<unsigned char>& rgb, std::string& err) {
    std::string raw; if (!readFile(path, raw)) { err = "cannot open " + path; return false; }
    if (raw.size() < 54 || raw[0] != 'B' || raw[1] != 'M') { err = path + ": not a BMP"; return false; }
    auto u32 = [&](size_t o) { uint32_t v; memcpy(&v, &raw[o], 4); return v; };
    auto i32 = [&](size_t o) { int32_t v; memcpy(&v, &raw[o], 4); return v; };
    auto u16 = [&](size_t o) { uint16_t v; memcpy(&v, &raw[o], 2); return v; };
    const uint32_t off = u32(10); w = i32(18); int hh = i32(22); const int bpp = u16(28); const uint32_t comp = u32(30);
    if ((bpp != 24 && bpp != 32) || (comp != 0 && comp != 3) || w <= 0 || hh == 0) { err = path + ": only uncompressed 24/32-bit BMP"; return false; }
    const bool topDown = hh < 0; h = abs(hh);
    const size_t bytes = bpp / 8, stride = ((size_t)w * bytes + 3) & ~(size_t)3;
    if (raw.size() < off + stride * h) { err = path + ": truncated"; return false; }
    rgb.resize((size_t)w * h * 3);
    for (int y = 0; y < h; y++) {
        const unsigned char* src = (const unsigned char*)&raw[off + stride * (size_t)(topDown ? h - 1 - y : y)];   // row 0 = bottom image row
        unsigned char* dst = &rgb[(size_t)y * w * 3];
        for (int x = 0; x < w; x++) { dst[3 * x] = src[bytes * x + 2]; dst[3 * x + 1] = src[bytes * x + 1]; dst[3 * x + 2] = src[bytes * x]; }
    }
    return true;
}

bool readHdr(const std::string& path, int& w, int& h, std::vector<float>& out, std::string& err) {
    std::string raw; if (!readFile(path, raw)) { err = "cannot open " + path; return false; }
    size_t p = 0;
    auto line = [&]() { size_t e = raw.find('\n', p); if (e == std::string::npos) e = raw.size(); std::string s = raw.substr(p, e - p); p = e + 1; return s; };
    std::string first = line();
    if (first != "#?RADIANCE" && first != "#?RGBE") { err = path + ": not a Radiance file"; return false; }
    bool fmt = false;
    for (;;) { if (p >= raw.size()) { err = path + ": no resolution line"; return false; } std::string s = line(); if (s.empty()) break; if (s == "FORMAT=32-bit_rle_rgbe") fmt = true; }
    if (!fmt) { err = path + ": unsupported FORMAT"; return false; }
    std::string res = line();
    if (sscanf(res.c_str(), "-Y %d +X %d", &h, &w) != 2 || w <= 0 || h <= 0) { err = path + ": unsupported orientation '" + res + "'"; return false; }
    std::vector<unsigned char> rgbe((size_t)w * h * 4);
    const unsigned char* d = (const unsigned char*)raw.data(); const size_t n = raw.size();
    bool flat = w < 8 || w >= 32768 || p + 4 > n || !(d[p] == 2 && d[p + 1] == 2 && !(d[p + 2] & 0x80));
    if (flat) { if (p + rgbe.size() > n) { err = path + ": truncated"; return false; } memcpy(rgbe.data(), d + p, rgbe.size()); }
    else for (int y = 0; y < h; y++) {
        if (p + 4 > n || d[p] != 2 || d[p + 1] != 2 || ((d[p + 2] << 8) | d[p + 3]) != w) { err = path + ": bad scanline header"; return false; }
        p += 4;
        for (int c = 0; c < 4; c++) for (int x = 0; x < w;) {
            if (p >= n) { err = path + ": truncated"; return false; }
            int cnt = d[p++];
            if (cnt > 128) { cnt -= 128; if (p >= n || x + cnt > w) { err = path + ": bad run"; return false; } unsigned char v = d[p++]; for (int k = 0; k < cnt; k++) rgbe[((size_t)y * w + x++) * 4 + c] = v; }
            else { if (p + cnt > n || x + cnt > w || cnt == 0) { err = path + ": bad literal run"; return false; } for (int k = 0; k < cnt; k++) rgbe[((size_t)y * w + x++) * 4 + c] = d[p++]; }
        }
    }
    out.resize((size_t)w * h * 3);
    for (size_t i = 0; i < (size_t)w * h; i++) {           // value = mantissa byte * 2^(e - 136), zero when e == 0
        const int e = rgbe[4 * i + 3];
        const float f = e ? ldexpf(1.0f, e - 136) : 0.f;
        out[3 * i] = rgbe[4 * i] * f; out[3 * i + 1] = rgbe[4 * i + 1] * f; out[3 * i + 2] = rgbe[4 * i + 2] * f;
    }
    return true;
}

bool writeBmp24(const std::string& path, int w, int h, const unsigned char* rgba, std::string& err) {
    const size_t stride = ((size_t)w * 3 + 3) & ~(size_t)3, size = 54 + stride * h;
    std::vector<unsigned char> f(size, 0);
    auto put32 = [&](size_t o, uint32_t v) { memcpy(&f[o], &v, 4); };
    f[0] = 'B'; f[1] = 'M'; put32(2, (uint32_t)size); put32(10, 54); put32(14, 40); put32(18, (uint32_t)w); put32(22, (uint32_t)h);
    f[26] = 1; f[28] = 24; put32(34, (uint32_t)(stride * h)); put32(38, 2835); put32(42, 2835);
    for (int y = 0; y < h; y++) {                           // film row 0 is the TOP row; BMP stores bottom-up
        const unsigned char* src = rgba + (size_t)(h - 1 - y) * w * 4; unsigned char* dst = &f[54 + stride * (size_t)y];
        for (int x = 0; x < w; x++) { dst[3 * x] = src[4 * x + 2]; dst[3 * x + 1] = src[4 * x + 1]; dst[3 * x + 2] = src[4 * x]; }
    }
    FILE* fp = fopen(path.c_str(), "wb"); if (!fp) { err = "cannot write " + path; return false; }
    const bool ok = fwrite(f.data(), 1, f.size(), fp) == f.size(); fclose(fp);
    if (!ok) err = "short write to " + path;
    return ok;
}

// ---------------------------------------------------------------------------------------------------
// flat container (oracle/ref_harness/flat_scene.h documents the layout)
// ---------------------------------------------------------------------------------------------------
struct FlatTexHeader { uint32_t format; int32_t width, height; float xTile, yTile, xOffset, yOffset; uint32_t filter; };

static bool loadFlat(const std::string& path, LoadedScene& s, std::string& err) {
    FILE* f = fopen(path.c_str(), "rb"); if (!f) { err = "cannot open " + path; return false; }
    char magic[8]; uint32_t n = 0;
    auto rd = [&](void* p, size_t sz, size_t cnt) { return cnt == 0 || fread(p, sz, cnt, f) == cnt; };
    bool ok = rd(magic, 1, 8) && memcmp(magic, "ELVNSCN1", 8) == 0 && rd(&s.camera, sizeof s.camera, 1);
    ok = ok && rd(&n, 4, 1); if (ok) { s.tris.resize(n); ok = rd(s.tris.data(), sizeof(ElevenTri), n); }
    ok = ok && rd(&n, 4, 1); if (ok) { s.objectMaterial.resize(n); ok = rd(s.objectMaterial.data(), 4, n); }
    ok = ok && rd(&n, 4, 1); if (ok) { s.materials.resize(n); ok = rd(s.materials.data(), sizeof(ElevenMaterial), n); }
    auto rdTex = [&](LoadedTexture& t) {
        FlatTexHeader h; if (!rd(&h, sizeof h, 1)) return false;
        if (h.format > ELEVEN_TEX_U8_LINEAR || h.width <= 0 || h.height <= 0) { err = "flat scene holds an external/invalid texture"; return false; }
        const size_t cnt = (size_t)h.width * h.height * 3;
        t.view = ElevenTexture{nullptr, h.format, h.width, h.height, h.xTile, h.yTile, h.xOffset, h.yOffset, h.filter};
        if (h.format == ELEVEN_TEX_F32_RGB) { t.floats.resize(cnt); return rd(t.floats.data(), 4, cnt); }
        t.bytes.resize(cnt); return rd(t.bytes.data(), 1, cnt);
    };
    ok = ok && rd(&n, 4, 1);
    if (ok) { s.textures.resize(n); for (uint32_t i = 0; ok && i < n; i++) ok = rdTex(s.textures[i]); }
    ok = ok && rdTex(s.hdri);
    ok = ok && rd(&n, 4, 1); if (ok) { s.lights.resize(n); ok = rd(s.lights.data(), sizeof(ElevenPointLight), n); }
    fclose(f);
    if (!ok && err.empty()) err = path + ": not a valid ELVNSCN1 file";
    return ok;
}

bool saveFlat(const LoadedScene& s, const std::string& path, std::string& err) {
    FILE* f = fopen(path.c_str(), "wb"); if (!f) { err = "cannot write " + path; return false; }
    auto wr = [&](const void* p, size_t sz, size_t cnt) { if (cnt) fwrite(p, sz, cnt, f); };
    auto wrN = [&](size_t n) { uint32_t v = (uint32_t)n; fwrite(&v, 4, 1, f); };
    fwrite("ELVNSCN1", 1, 8, f); wr(&s.camera, sizeof s.camera, 1);
    wrN(s.tris.size()); wr(s.tris.data(), sizeof(ElevenTri), s.tris.size());
    wrN(s.objectMaterial.size()); wr(s.objectMaterial.data(), 4, s.objectMaterial.size());
    wrN(s.materials.size()); wr(s.materials.data(), sizeof(ElevenMaterial), s.materials.size());
    auto wrTex = [&](const LoadedTexture& t) {
        FlatTexHeader h{t.view.format, t.view.width, t.view.height, t.view.xTile, t.view.yTile, t.view.xOffset, t.view.yOffset, t.view.filter};
        wr(&h, sizeof h, 1);
        if (t.view.format == ELEVEN_TEX_F32_RGB) wr(t.floats.data(), 4, t.floats.size()); else wr(t.bytes.data(), 1, t.bytes.size());
    };
    wrN(s.textures.size()); for (const auto& t : s.textures) wrTex(t);
    wrTex(s.hdri);
    wrN(s.lights.size()); wr(s.lights.data(), sizeof(ElevenPointLight), s.lights.size());
    const bool ok = ferror(f) == 0; fclose(f);
    if (!ok) err = "write error on " + path;
    return ok;
}

// ---------------------------------------------------------------------------------------------------
// scene directory
// ---------------------------------------------------------------------------------------------------
static void parseFloats(const std::string& s, float* out, int n) {
    const char* p = s.c_str();
    for (int i = 0; i < n; i++) { char* e; out[i] = strtof(p, &e); if (e == p) { out[i] = 0; break; } p = e; }
}
static std::string afterFirstSpace(const std::string& s) { size_t i = s.find(' '); return i == std::string::npos ? "" : s.substr(i + 1); }
static void rstrip(std::string& s) { while (!s.empty() && (s.back() == '\r' || s.back() == '\n' || s.back() == ' ')) s.pop_back(); }

static ElevenMaterial defaultMaterial() {          // Material.hpp:12-36
    ElevenMaterial m; memset(&m, 0, sizeof m);
    m.albedoTextureID = m.emissionTextureID = m.roughnessTextureID = m.metallicTextureID = m.normalTextureID = m.opacityTextureID = -1;
    m.albedo[0] = m.albedo[1] = m.albedo[2] = 0.5f; m.opacity[0] = m.opacity[1] = m.opacity[2] = 1.f;
    m.roughness = 1.f; m.specular = 0.5f;
    return m;
}

static bool loadDir(std::string dir, LoadedScene& s, std::string& err) {
    if (!dir.empty() && dir.back() != '/') dir += '/';
    std::string text;
    if (!readFile(dir + "scene.json", text)) { err = "cannot open " + dir + "scene.json"; return false; }
    JParser jp(text); auto root = jp.value();
    if (!jp.err.empty() || root->kind != JVal::Obj) { err = "scene.json: " + (jp.err.empty() ? std::string("not an object") : jp.err); return false; }
    // camera (S/SceneLoader.hpp:25-46; S/Camera.hpp:14,35-37)
    const JVal* cam = root->get("camera"); if (!cam) { err = "scene.json: no camera"; return false; }
    memset(&s.camera, 0, sizeof s.camera);
    s.camera.xRes = (uint32_t)cam->number("xRes"); s.camera.yRes = (uint32_t)cam->number("yRes");
    if (!s.camera.xRes || !s.camera.yRes) { err = "scene.json: camera resolution missing"; return false; }
    s.camera.focalLength = (float)cam->number("focalLength"); s.camera.focusDistance = (float)cam->number("focusDistance");
    s.camera.aperture = (float)cam->number("aperture"); s.camera.bokeh = cam->boolean("bokeh") ? 1 : 0;
    s.camera.sensorWidth = 35 * 0.001f; s.camera.sensorHeight = s.camera.sensorWidth * ((float)s.camera.yRes / (float)s.camera.xRes);
    const char* xyz[3] = {"x", "y", "z"};
    if (const JVal* p = cam->get("position")) for (int a = 0; a < 3; a++) s.camera.position[a] = (float)p->number(xyz[a]);
    if (const JVal* p = cam->get("rotation")) for (int a = 0; a < 3; a++) s.camera.rotation[a] = (float)p->number(xyz[a]);
    // environment (S/SceneLoader.hpp:51-65)
    const JVal* hd = root->get("hdri"); if (!hd) { err = "scene.json: no hdri"; return false; }
    int ew = 0, eh = 0;
    if (const JVal* nm = hd->get("name")) { if (!readHdr(dir + "HDRI/" + nm->str + ".hdr", ew, eh, s.hdri.floats, err)) return false; }
    else if (const JVal* col = hd->get("color")) {
        ew = eh = 1024; s.hdri.floats.resize((size_t)ew * eh * 3);
        const float c[3] = {(float)col->number("r"), (float)col->number("g"), (float)col->number("b")};
        for (size_t i = 0; i < (size_t)ew * eh; i++) { s.hdri.floats[3 * i] = c[0]; s.hdri.floats[3 * i + 1] = c[1]; s.hdri.floats[3 * i + 2] = c[2]; }
    } else { err = "scene.json: hdri needs name or color"; return false; }
    s.hdri.view = ElevenTexture{nullptr, ELEVEN_TEX_F32_RGB, ew, eh, 1.f, 1.f, 0.f, 0.f, 0u};
    if (hd->get("xOffset")) s.hdri.view.xOffset = (float)hd->number("xOffset");
    if (hd->get("yOffset")) s.hdri.view.xOffset = (float)hd->number("yOffset");       // the reference writes yOffset into xOffset (S/SceneLoader.hpp:64-65)
    // point lights (S/SceneLoader.hpp:127-141)
    if (const JVal* pl = root->get("pointLights")) for (auto& l : pl->arr) {
        ElevenPointLight e; memset(&e, 0, sizeof e);
        const char* rgb[3] = {"r", "g", "b"};
        if (const JVal* p = l->get("position")) for (int a = 0; a < 3; a++) e.position[a] = (float)p->number(xyz[a]);
        if (const JVal* p = l->get("radiance")) for (int a = 0; a < 3; a++) e.radiance[a] = (float)(p->get(xyz[a]) ? p->number(xyz[a]) : p->number(rgb[a]));
        s.lights.push_back(e);
    }
    // materials (S/ObjLoader.hpp:33-68, S/SceneLoader.hpp:69-107)
    std::ifstream mtl(dir + "scene.mtl"); std::string line;
    std::vector<std::map<std::string, std::string>> maps;
    while (std::getline(mtl, line)) {
        rstrip(line);
        if (line.find("newmtl") != std::string::npos) { s.materials.push_back(defaultMaterial()); s.materialNames.push_back(afterFirstSpace(line)); maps.emplace_back(); continue; }
        if (s.materials.empty() || line.size() < 2) continue;
        ElevenMaterial& m = s.materials.back(); float v[3] = {0, 0, 0};
        if (line[0] == 'K' && line[1] == 'd') { parseFloats(line.substr(2), m.albedo, 3); }
        else if (line[0] == 'K' && line[1] == 's') { parseFloats(line.substr(2), v, 3); m.specular = v[0]; }
        else if (line[0] == 'K' && line[1] == 'e') { parseFloats(line.substr(2), m.emission, 3); }
        else if (line[0] == 'N' && line[1] == 'i') { m.eta = strtof(afterFirstSpace(line).c_str(), nullptr); }
        else if (line[0] == 'd') { const float o = strtof(afterFirstSpace(line).c_str(), nullptr); m.opacity[0] = m.opacity[1] = m.opacity[2] = o; }
        for (const char* key : {"map_Kd", "map_Ns", "map_Bump", "refl"}) if (line.find(key) != std::string::npos) maps.back()[key] = afterFirstSpace(line);
    }
    if (s.materials.empty()) { s.materials.push_back(defaultMaterial()); s.materialNames.push_back("default"); maps.emplace_back(); }
    // texture ids in the reference's order (per material its std::map iteration order: map_Bump, map_Kd, map_Ns, refl; de-duplicated by path);
    // the files themselves (12 x 50 MB for ClockCC0) are read and decoded by a pool of threads below, while another thread parses the OBJ
    std::vector<std::string> texFile;
    for (size_t i = 0; i < s.materials.size(); i++) for (auto& kv : maps[i]) {
        int id = -1;
        for (size_t j = 0; j < s.textures.size(); j++) if (s.textures[j].path == kv.second) id = (int)j;
        if (id < 0) {
            LoadedTexture t;
            std::string p = kv.second; if (!p.empty() && p[0] != '/') { struct stat st; if (stat(p.c_str(), &st) != 0) p = dir + p; }   // the reference resolves against the CWD
            t.path = kv.second;
            t.view = ElevenTexture{nullptr, (uint32_t)(kv.first == "map_Kd" ? ELEVEN_TEX_U8_SRGB : ELEVEN_TEX_U8_LINEAR), 0, 0, 1.f, 1.f, 0.f, 0.f, 0u};
            id = (int)s.textures.size(); s.textures.push_back(std::move(t)); texFile.push_back(p);
        }
        ElevenMaterial& m = s.materials[i];
        if (kv.first == "map_Kd") m.albedoTextureID = id; else if (kv.first == "map_Ns") m.roughnessTextureID = id;
        else if (kv.first == "refl") m.metallicTextureID = id; else if (kv.first == "map_Bump") m.normalTextureID = id;
    }
    std::vector<std::string> texErr(s.textures.size());
    std::atomic<size_t> nextTex(0);
    auto texWorker = [&]() {
        for (;;) {
            const size_t k = nextTex.fetch_add(1); if (k >= s.textures.size()) break;
            int w = 0, h = 0;
            if (readBmp24(texFile[k], w, h, s.textures[k].bytes, texErr[k])) { s.textures[k].view.width = w; s.textures[k].view.height = h; }
        }
    };
    std::vector<std::thread> texThreads;
    for (unsigned t = 0; t < std::min<size_t>(s.textures.size(), std::max(1u, std::thread::hardware_concurrency())); t++) texThreads.emplace_back(texWorker);
    struct JoinAll { std::vector<std::thread>& v; ~JoinAll() { for (auto& t : v) if (t.joinable()) t.join(); } } joinTex{texThreads};
    // geometry (S/ObjLoader.hpp:71-171)
    std::ifstream obj(dir + "scene.obj");
    if (!obj) { err = "cannot open " + dir + "scene.obj"; return false; }
    std::vector<float> V, VT, VN; int object = -1; std::string objMtl; size_t objectFirstTri = 0;
    auto closeObject = [&]() {
        if (object < 0) return;
        // tangent frames per object, like CalcTangents::calc(&mo) at the end of parseObj (S/ObjLoader.hpp:167-168)
        computeTangentSpace(s.tris.data() + objectFirstTri, s.tris.size() - objectFirstTri);
        objectFirstTri = s.tris.size();
        int mid = 0; for (size_t j = 0; j < s.materialNames.size(); j++) if (s.materialNames[j] == objMtl) mid = (int)j;
        s.objectMaterial.push_back(mid);
    };
    while (std::getline(obj, line)) {
        rstrip(line); if (line.empty()) continue;
        if (line[0] == 'o') { closeObject(); object++; objMtl.clear(); s.objectNames.push_back(afterFirstSpace(line)); continue; }
        if (object < 0) continue;                                   // everything before the first `o` is ignored
        if (line.find("usemtl") != std::string::npos) { objMtl = afterFirstSpace(line); continue; }
        float v[3] = {0, 0, 0};
        if (line[0] == 'v' && line.size() > 1 && line[1] == ' ') { parseFloats(line.substr(2), v, 3); V.insert(V.end(), {v[0], v[1], -v[2]}); }
        else if (line[0] == 'v' && line.size() > 1 && line[1] == 't') { parseFloats(line.substr(2), v, 3); VT.insert(VT.end(), {v[0], v[1], v[2]}); }
        else if (line[0] == 'v' && line.size() > 1 && line[1] == 'n') { parseFloats(line.substr(2), v, 3); VN.insert(VN.end(), {v[0], v[1], -v[2]}); }
        else if (line[0] == 'f') {
            ElevenTri T; memset(&T, 0, sizeof T); T.objectID = object;
            const char* p = line.c_str() + 1; int corner = 0;
            while (*p && corner < 3) {
                while (*p == ' ') p++;
                if (!*p) break;
                long idx[3] = {0, 0, 0}; int k = 0;
                while (*p && *p != ' ') { if (*p == '/') { k++; p++; continue; } char* e; long val = strtol(p, &e, 10); if (e == p) { p++; continue; } if (k < 3) idx[k] = val; p = e; }
                if (idx[0] > 0 && (size_t)idx[0] * 3 <= V.size()) memcpy(T.vertices[corner], &V[(idx[0] - 1) * 3], 12);
                if (idx[1] > 0 && (size_t)idx[1] * 3 <= VT.size()) memcpy(T.uv[corner], &VT[(idx[1] - 1) * 3], 12);
                if (idx[2] > 0 && (size_t)idx[2] * 3 <= VN.size()) {
                    const float* n = &VN[(idx[2] - 1) * 3]; const float l = sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
                    for (int a = 0; a < 3; a++) T.normals[corner][a] = l == 0 ? n[a] : n[a] / l;
                }
                corner++;
            }
            if (corner == 3) s.tris.push_back(T);
        }
    }
    closeObject();
    if (s.objectMaterial.empty()) { s.objectMaterial.push_back(0); }
    for (auto& t : texThreads) t.join();
    for (const std::string& e : texErr) if (!e.empty()) { err = e; return false; }
    return true;
}

bool loadScene(const std::string& path, LoadedScene& out, std::string& err) {
    out = LoadedScene();
    return isDir(path) ? loadDir(path, out, err) : loadFlat(path, out, err);
}

ElevenSceneDesc LoadedScene::desc() {
    ElevenSceneDesc d; memset(&d, 0, sizeof d);
    d.camera = camera;
    d.triCount = (uint32_t)tris.size(); d.tris = tris.data();
    d.objectCount = (uint32_t)objectMaterial.size(); d.objectMaterial = objectMaterial.data();
    d.materialCount = (uint32_t)materials.size(); d.materials = materials.data();
    textureViews.clear();
    for (auto& t : textures) { ElevenTexture v = t.view; v.data = v.format == ELEVEN_TEX_F32_RGB ? (const void*)t.floats.data() : (const void*)t.bytes.data(); textureViews.push_back(v); }
    d.textureCount = (uint32_t)textureViews.size(); d.textures = textureViews.data();
    d.hdri = hdri.view; d.hdri.data = hdri.floats.data();
    d.pointLightCount = (uint32_t)lights.size(); d.pointLights = lights.data();
    return d;
}

} // namespace eleven_host
