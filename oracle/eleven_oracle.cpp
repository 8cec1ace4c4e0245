/*
 * eleven_oracle.cpp — CPU restatement of the reference's per-sample path-tracing algorithm.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (tfg-pathtracer_b200/, the C-ABI
 * library, the CLI) links, loads or calls this file.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may use it, and only as the checker.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function here bit-for-bit
 * (float ops) against vectors produced by the reference's own headers compiled on the host
 * (oracle/ref_harness/ref_host_vectors.cpp -> oracle/_ref/ref_host_vectors, fixtures committed
 * under tests/golden/), and the image path against raw renders of the reference CUDA build on a
 * B200 (oracle/_ref/eleven_ref_headless, fixtures under tests/golden/).
 *
 * S/ = /root/reference/src/tfg-pathtracer/.  Every function cites the lines it follows.
 * Arithmetic flavour: plain IEEE binary32 ops in the reference's source order with its
 * float/double promotions; build with -ffp-contract=off (this is the reference built with
 * nvcc --fmad=false, SURVEY §8c "precise flavour").  RNG draws are taken into named temporaries
 * in left-to-right order (what nvcc does for the device code).
 *
 * Defined behaviour where the reference has UB, shared with the CUDA path:
 *   - hdriPdf = 0 when the environment shadow ray is occluded (S/kernel.cu:344,248)
 *   - point-light index clamped to count-1 when r1 == 1 (S/kernel.cu:185)
 *   - texel linear index clamped to [0, W*H) (S/Texture.hpp:100-105 reads out of bounds)
 */
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cfloat>
#include <atomic>
#include <thread>
#include <vector>

#include "../include/eleven_b200.h"

namespace {

// ----------------------------------------------------------------------------------------------
// Vector arithmetic: S/Vector.hpp:7-239.  Component order and association as in the reference.
// ----------------------------------------------------------------------------------------------
struct V3 { float x, y, z; };
static inline V3 v3(float x, float y, float z) { V3 r = {x, y, z}; return r; }
static inline V3 v3(float s) { V3 r = {s, s, s}; return r; }
static inline V3 v3(const float* p) { V3 r = {p[0], p[1], p[2]}; return r; }
static inline V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }     // :195
static inline V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }     // :191
static inline V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }       // :203
static inline V3 operator*(float s, V3 a) { return v3(a.x * s, a.y * s, a.z * s); }       // :207
static inline V3 operator*(V3 a, V3 b) { return v3(b.x * a.x, b.y * a.y, b.z * a.z); }     // :176
static inline V3 operator/(V3 a, float s) { return v3(a.x / s, a.y / s, a.z / s); }       // :210
static inline V3 addScalar(V3 a, float s) { return v3(a.x + s, a.y + s, a.z + s); }       // :199
static inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }         // :151
static inline V3 cross(V3 a, V3 b) {                                                      // :155
    return v3((a.y * b.z - a.z * b.y), -(a.x * b.z - a.z * b.x), (a.x * b.y - a.y * b.x));
}
static inline float length(V3 a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }     // :93
static inline V3 normalized(V3 a) {                                                       // :159-172
    float l = length(a);
    if (l == 0) return a;
    return v3(a.x / l, a.y / l, a.z / l);
}
static inline V3 reflectV(V3 a, V3 b) { return a - (2 * dot(a, b)) * b; }                 // :214

static const float PIf = 3.14159265358979323846f;                                         // S/Math.hpp:7
static inline float minf(float a, float b) { return a < b ? a : b; }                      // S/Math.hpp:60
static inline float maxf(float a, float b) { return a > b ? a : b; }                      // S/Math.hpp:64
static inline float clampf(float a, float b, float c) { return a < b ? b : a > c ? c : a; } // S/Math.hpp:29
static inline float lerpf(float a, float b, float c) { return a + c * (b - a); }          // S/Math.hpp:43 (FAST_LERP)
static inline V3 lerpV(V3 a, V3 b, float c) { return v3(lerpf(a.x, b.x, c), lerpf(a.y, b.y, c), lerpf(a.z, b.z, c)); }
static inline float mapf(float a, float b, float c, float d, float e) { return d + ((a - b) / (c - b)) * (e - d); } // S/Math.hpp:21

struct Ray { V3 o, d; };
static inline Ray makeRay(V3 o, V3 d) { Ray r; r.o = o; r.d = normalized(d); return r; }  // S/Ray.hpp:14-18

struct Hit {                                                                             // S/Hit.hpp:6-13 (+tri,t,u,v,key)
    V3 position, normal, tangent, bitangent;
    bool valid; int objectID; float tu, tv;
    int tri; float t, u, v;
};
static inline Hit noHit() { Hit h; memset(&h, 0, sizeof h); h.valid = false; h.tri = -1; return h; }

struct HitData {                                                                         // S/kernel.h:46-69
    float metallic, roughness, clearcoatGloss, clearcoat, anisotropic, eta, transmission,
          specular, specularTint, sheenTint, subsurface, sheen;
    V3 emission, albedo, normal, tangent, bitangent;
};

// ----------------------------------------------------------------------------------------------
// XORWOW, the cuRAND device generator the reference seeds per pixel with curand_init(0, idx, 0)
// (S/kernel.cu:140).  cuRAND is a CUDA-toolkit header (curand_kernel.h, 11.1 at the reference,
// 12.9 here), not part of /root/reference; this restates its published algorithm: Marsaglia
// xorshift on 5 words + Weyl counter (curand_kernel.h:863-876), seed scrambling (:780-791),
// subsequence = skip-ahead by 2^67 draws.  The skip-ahead matrix is derived here from the step
// function by repeated squaring over GF(2) instead of reading the toolkit's precalc tables.
// ----------------------------------------------------------------------------------------------
struct Xorwow { uint32_t v[5]; uint32_t d; };

static inline uint32_t xorwowNext(Xorwow& s) {
    uint32_t t = s.v[0] ^ (s.v[0] >> 2);
    s.v[0] = s.v[1]; s.v[1] = s.v[2]; s.v[2] = s.v[3]; s.v[3] = s.v[4];
    s.v[4] = (s.v[4] ^ (s.v[4] << 4)) ^ (t ^ (t << 1));
    s.d += 362437u;
    return s.v[4] + s.d;
}
// curand_uniform.h:69-72: (0,1]
static inline float xorwowUniform(Xorwow& s) {
    uint32_t x = xorwowNext(s);
    return (float)x * 2.3283064e-10f + (2.3283064e-10f / 2.0f);
}

// 160x160 bit matrix; column j (as 5 words) = image of basis vector e_j.
struct BitMat { uint32_t col[160][5]; };
static void matVec(const BitMat& m, const uint32_t in[5], uint32_t out[5]) {
    uint32_t r[5] = {0, 0, 0, 0, 0};
    for (int w = 0; w < 5; w++)
        for (int b = 0; b < 32; b++)
            if (in[w] & (1u << b)) { const uint32_t* c = m.col[w * 32 + b]; for (int k = 0; k < 5; k++) r[k] ^= c[k]; }
    for (int k = 0; k < 5; k++) out[k] = r[k];
}
static void matMul(const BitMat& a, const BitMat& b, BitMat& out) {   // out = a * b (apply b, then a)
    BitMat r;
    for (int j = 0; j < 160; j++) matVec(a, b.col[j], r.col[j]);
    out = r;
}
static BitMat g_seqMat[32];      // g_seqMat[k] = step^(2^67 * 2^k)
static std::atomic<int> g_seqMatReady(0);
static void buildSeqMats() {
    static std::atomic<int> lock(0);
    if (g_seqMatReady.load()) return;
    int expected = 0;
    if (!lock.compare_exchange_strong(expected, 1)) { while (!g_seqMatReady.load()) std::this_thread::yield(); return; }
    BitMat m;
    for (int j = 0; j < 160; j++) {
        Xorwow s; memset(&s, 0, sizeof s); s.v[j / 32] = 1u << (j % 32);
        xorwowNext(s);
        for (int k = 0; k < 5; k++) m.col[j][k] = s.v[k];
    }
    for (int i = 0; i < 67; i++) matMul(m, m, m);
    g_seqMat[0] = m;
    for (int k = 1; k < 32; k++) matMul(g_seqMat[k - 1], g_seqMat[k - 1], g_seqMat[k]);
    g_seqMatReady.store(1);
}
static Xorwow xorwowInit(uint64_t seed, uint64_t subsequence) {
    buildSeqMats();
    Xorwow s;
    uint32_t s0 = ((uint32_t)seed) ^ 0xaad26b49u;
    uint32_t s1 = (uint32_t)(seed >> 32) ^ 0xf7dcefddu;
    uint32_t t0 = 1099087573u * s0;
    uint32_t t1 = 2591861531u * s1;
    s.d = 6615241u + t1 + t0;
    s.v[0] = 123456789u + t0;
    s.v[1] = 362436069u ^ t0;
    s.v[2] = 521288629u + t1;
    s.v[3] = 88675123u ^ t1;
    s.v[4] = 5783321u + t0;
    for (int k = 0; k < 32 && (subsequence >> k); k++)
        if ((subsequence >> k) & 1) matVec(g_seqMat[k], s.v, s.v);
    return s;   // the Weyl counter d is unchanged: 2^67 * 362437 = 0 mod 2^32 (curand_kernel.h:698)
}

// ----------------------------------------------------------------------------------------------
// stb_image's patched LDR->float decode: (float)(fastPow(byte/255.0f, gamma) * 1.0f)
// S/stb_image.h:127-136 (Ankerl's approximate pow) and :1863.  gamma = 2.2f (sRGB) or 1.0f.
// ----------------------------------------------------------------------------------------------
static double fastPowRef(double a, double b) {
    union { double d; int32_t x[2]; } u; u.d = a;
    u.x[1] = (int32_t)(b * (u.x[1] - 1072632447) + 1072632447);
    u.x[0] = 0;
    return u.d;
}
static float texelDecode(int byte, float gamma) {
    return (float)(fastPowRef((float)byte / 255.0f, gamma) * 1.0f);
}

// ----------------------------------------------------------------------------------------------
// Scene held by the oracle (float RGB textures exactly like the reference, S/Texture.hpp:18).
// ----------------------------------------------------------------------------------------------
struct Tex {
    std::vector<float> data; int width, height; float xTile, yTile, xOffset, yOffset; uint32_t filter;

    // S/Texture.hpp:95-108.  Linear index clamped (the reference reads out of bounds for negative UVs).
    V3 texel(int x, int y) const {
        x = (int)(xTile * (x + xOffset * width)) % width;
        y = (int)(yTile * (y + yOffset * height)) % height;
        long long idx = (long long)y * width + x;
        long long n = (long long)width * height;
        if (idx < 0) idx = 0;
        if (idx >= n) idx = n - 1;
        return v3(data[3 * idx + 0], data[3 * idx + 1], data[3 * idx + 2]);
    }
    V3 fromUV(float u, float v) const { return texel((int)(u * width), (int)(v * height)); }   // :110-112
    V3 bilinear(float u, float v) const {                                                    // :114-135
        float x = u * width, y = v * height;
        float t1x = floorf(x), t1y = floorf(y);
        float t2x = t1x + 1, t2y = t1y + 1;
        float a = (x - t1x) / (t2x - t1x), b = (y - t1y) / (t2y - t1y);
        V3 v1 = texel((int)t1x, (int)t1y), v2 = texel((int)t2x, (int)t1y);
        V3 v3_ = texel((int)t1x, (int)t2y), v4 = texel((int)t2x, (int)t2y);
        return lerpV(lerpV(v1, v2, a), lerpV(v3_, v4, a), b);
    }
    V3 filtered(float u, float v) const { return filter == 0 ? fromUV(u, v) : bilinear(u, v); } // :137-142

    // :177-193
    void inverseTransformUV(float u, float v, float& nu, float& nv) const {
        int x = (int)(u * width), y = (int)(v * height);
        x = (int)(xTile * (x - xOffset * width)) % width;
        y = (int)(yTile * (y - yOffset * height)) % height;
        nu = (float)x / (float)width; nv = (float)y / (float)height;
        nu += (float)(-(nu > 1) + -(nu < 0)); nv += (float)(-(nv > 1) + -(nv < 0));            // limitUV, S/Math.hpp:38-41
    }
};
// S/Texture.hpp:144-156 + limitUV S/Math.hpp:38-41
static void sphericalMapping(V3 p, float& u, float& v) {
    float theta = acosf(-p.y);
    float phi = atan2f(-p.z, p.x) + PIf;
    u = phi / (2 * PIf);
    v = theta / PIf;
    u += (float)(-(u > 1) + -(u < 0));
    v += (float)(-(v > 1) + -(v < 0));
}
// S/Texture.hpp:195-207
static V3 reverseSphericalMapping(float u, float v) {
    float phi = u * 2 * PIf, theta = v * PIf;
    float px = cosf(phi - PIf), py = -cosf(theta), pz = -sinf(phi - PIf);
    float a = sqrtf(1 - py * py);
    return v3(a * px, py, a * pz);
}

struct RefNode { V3 b1, b2; int from, to, depth; };                                         // S/BVH.hpp:29-48

static const int REF_DEPTH = 18;                                                            // S/Definitions.h:12
static const int REF_BINS = 14;                                                             // S/Definitions.h:13

struct Scene {
    ElevenCamera cam;
    std::vector<ElevenTri> tris;
    std::vector<int> objMat;
    std::vector<ElevenMaterial> mats;
    std::vector<Tex> tex;
    Tex hdri; std::vector<float> cdf; float radianceSum;
    std::vector<ElevenPointLight> lights;
    // reference BVH: implicit complete binary tree, pre-order numbering (S/BVH.hpp:62,177-185)
    std::vector<RefNode> nodes; std::vector<int> triIndices;
    // film (S/kernel.cu:41-45) + rng
    std::vector<float> passes[ELEVEN_PASS_COUNT];
    std::vector<uint32_t> samples, pathcount;
    std::vector<Xorwow> rng;
    std::atomic<uint64_t> raysExt, raysEnv, raysLight;
    // 0: the reference's binarySearch (S/HDRI.hpp:130-142), which returns the texel AFTER the one whose CDF interval holds r for about
    //    half of all r (it stops on `to`, one past `from`, without a final comparison) while HDRI::pdf is evaluated for the texel it
    //    returned: the texel distribution is ~(w_i + w_(i-1))/2 against a pdf of w_i.
    // 1: exact inversion (the texel i with cdf[i] < r <= cdf[i+1]).  NOT the reference: the yardstick for the product's alias-table
    //    mode (north_star), which samples the nominal weights w_i exactly and therefore shares THIS estimator's expectation.
    int envSearchExact;
    Scene() : raysExt(0), raysEnv(0), raysLight(0), envSearchExact(0) {}
};

// ----------------------------------------------------------------------------------------------
// Ray/triangle: S/Tri.hpp:38-160 (Moeller-Trumbore + smooth attributes + shadow-terminator shift)
// ----------------------------------------------------------------------------------------------
static bool triHit(const ElevenTri& T, int triIdx, const Ray& ray, Hit& hit) {
    float EPSILON = 0.0000001;
    V3 v0 = v3(T.vertices[0]), v1 = v3(T.vertices[1]), v2 = v3(T.vertices[2]);
    V3 edge1 = v1 - v0, edge2 = v2 - v0;
    V3 pvec = cross(ray.d, edge2);
    float det = dot(edge1, pvec);
    float inv_det = (float)(1.0 / det);
    if (det > -EPSILON && det < EPSILON) return false;
    V3 tvec = ray.o - v0;
    float u = dot(tvec, pvec) * inv_det;
    if (u < 0.0 || u > 1.0) return false;
    V3 qvec = cross(tvec, edge1);
    float v = dot(ray.d, qvec) * inv_det;
    if (v < 0.0 || (u + v) > 1.0) return false;
    float t = dot(edge2, qvec) * inv_det;
    if (t < 0) return false;

    V3 uv0 = v3(T.uv[0]), uv1 = v3(T.uv[1]), uv2 = v3(T.uv[2]);
    V3 tUV = uv0 + (uv1 - uv0) * u + (uv2 - uv0) * v;
    V3 geomPosition = ray.o + ray.d * t;
    V3 n0 = v3(T.normals[0]), n1 = v3(T.normals[1]), n2 = v3(T.normals[2]);
    V3 t0 = v3(T.tangents[0]), t1 = v3(T.tangents[1]), t2 = v3(T.tangents[2]);
    V3 shadingNormal = n0 + (n1 - n0) * u + (n2 - n0) * v;
    V3 shadingTangent = t0 + (t1 - t0) * u + (t2 - t0) * v;
    // projectOnPlane, S/Tri.hpp:34-36
    V3 p0 = geomPosition - dot(geomPosition - v0, n0) * n0;
    V3 p1 = geomPosition - dot(geomPosition - v1, n1) * n1;
    V3 p2 = geomPosition - dot(geomPosition - v2, n2) * n2;
    V3 shadingPosition = p0 + (p1 - p0) * u + (p2 - p0) * v;
    bool convex = dot(shadingPosition - geomPosition, shadingNormal) > 0.0f;
    hit.tangent = shadingTangent;
    hit.position = convex ? shadingPosition : geomPosition;
    hit.normal = shadingNormal;
    hit.bitangent = T.tangentsSign * cross(hit.normal, hit.tangent);
    hit.valid = true; hit.tu = tUV.x; hit.tv = tUV.y; hit.objectID = T.objectID;
    hit.tri = triIdx; hit.t = t; hit.u = u; hit.v = v;
    return true;
}

// S/BVH.hpp:69-101: slab test, 1/dir recomputed per box, NaN-unsafe min/max, no tmax clip.
static bool slab(const Ray& ray, V3 b1, V3 b2) {
    float fx = 1.0f / ray.d.x, fy = 1.0f / ray.d.y, fz = 1.0f / ray.d.z;
    float t1 = (b1.x - ray.o.x) * fx, t2 = (b2.x - ray.o.x) * fx;
    float t3 = (b1.y - ray.o.y) * fy, t4 = (b2.y - ray.o.y) * fy;
    float t5 = (b1.z - ray.o.z) * fz, t6 = (b2.z - ray.o.z) * fz;
    float tmin = maxf(maxf(minf(t1, t2), minf(t3, t4)), minf(t5, t6));
    float tmax = minf(minf(maxf(t1, t2), maxf(t3, t4)), maxf(t5, t6));
    if (tmax < 0) return false;
    if (tmin > tmax) return false;
    return true;
}

// S/BVH.hpp:159-175: strict '<' on |hit.position - origin|, first found wins ties.
static void leafTest(const Scene& S, const Ray& ray, const RefNode& n, Hit& nearest) {
    for (int i = n.from; i < n.to; i++) {
        Hit h = noHit();
        int ti = S.triIndices[i];
        if (triHit(S.tris[ti], ti, ray, h)) {
            if (!nearest.valid) nearest = h;
            else if (length(h.position - ray.o) < length(nearest.position - ray.o)) nearest = h;
        }
    }
}
// S/BVH.hpp:120-157 (iterative, explicit stack) expressed as the same depth-first order:
// left subtree before right subtree; at depth 17 the RIGHT leaf is tested before the LEFT leaf
// (:138-142); the root box is never tested (:128).
static void traverseRef(const Scene& S, const Ray& ray, int idx, int depth, Hit& nearest) {
    int l = idx + 1, r = idx + (2 << (REF_DEPTH - depth - 1));                              // :177-185
    bool lO = slab(ray, S.nodes[l].b1, S.nodes[l].b2);
    bool rO = slab(ray, S.nodes[r].b1, S.nodes[r].b2);
    if (depth == REF_DEPTH - 1) {
        if (rO) leafTest(S, ray, S.nodes[r], nearest);
        if (lO) leafTest(S, ray, S.nodes[l], nearest);
        return;
    }
    if (lO) traverseRef(S, ray, l, depth + 1, nearest);
    if (rO) traverseRef(S, ray, r, depth + 1, nearest);
}
// The reference's non-BVH path (S/kernel.cu:159-170, S/MeshObject.hpp:40-59), flattened over all
// triangles in index order; used to classify reference slab-test misses.
static void traverseBrute(const Scene& S, const Ray& ray, Hit& nearest) {
    for (size_t i = 0; i < S.tris.size(); i++) {
        Hit h = noHit();
        if (triHit(S.tris[i], (int)i, ray, h)) {
            if (!nearest.valid) nearest = h;
            else if (length(h.position - ray.o) < length(nearest.position - ray.o)) nearest = h;
        }
    }
}
static Hit throwRay(Scene& S, const Ray& ray) {                                            // S/kernel.cu:152-173
    Hit h = noHit();
    traverseRef(S, ray, 0, 0, h);
    return h;
}

// ----------------------------------------------------------------------------------------------
// Reference BVH builder: S/BVH.hpp:187-211,291-330 (buildAux), :373-460 (divideSAH),
// :492-578 (boundsUnion / boundsArea / bounds).  Always to depth 18, empty subtrees included.
// ----------------------------------------------------------------------------------------------
static float boxArea(V3 b1, V3 b2) {                                                       // :519-527
    float x = b2.x - b1.x, y = b2.y - b1.y, z = b2.z - b1.z;
    return 2 * (x * y + x * z + y * z);
}
static void boxUnion(V3 b1, V3 b2, V3 b3, V3 b4, V3& b5, V3& b6) {                          // :492-517
    if (boxArea(b1, b2) <= 0 || boxArea(b3, b4) <= 0) {
        if (boxArea(b1, b2) <= 0) { b5 = b3; b6 = b4; }
        if (boxArea(b3, b4) <= 0) { b5 = b1; b6 = b2; }
    } else {
        b5.x = minf(b1.x, minf(b2.x, minf(b3.x, b4.x)));
        b5.y = minf(b1.y, minf(b2.y, minf(b3.y, b4.y)));
        b5.z = minf(b1.z, minf(b2.z, minf(b3.z, b4.z)));
        b6.x = maxf(b1.x, maxf(b2.x, maxf(b3.x, b4.x)));
        b6.y = maxf(b1.y, maxf(b2.y, maxf(b3.y, b4.y)));
        b6.z = maxf(b1.z, maxf(b2.z, maxf(b3.z, b4.z)));
    }
}
static void triBox(const ElevenTri& T, V3& b1, V3& b2) {                                    // :529-538
    for (int a = 0; a < 3; a++) {
        float p0 = T.vertices[0][a], p1 = T.vertices[1][a], p2 = T.vertices[2][a];
        (&b1.x)[a] = minf(p0, minf(p1, p2));
        (&b2.x)[a] = maxf(p0, maxf(p1, p2));
    }
}
static void setBox(const Scene& S, const std::vector<int>& ids, V3& b1, V3& b2) {           // :540-578
    if (ids.empty()) return;
    b1 = v3(S.tris[ids[0]].vertices[0]); b2 = b1;
    for (size_t i = 0; i < ids.size(); i++) {
        const ElevenTri& T = S.tris[ids[i]];
        for (int k = 0; k < 3; k++) {
            b1.x = minf(T.vertices[k][0], b1.x); b1.y = minf(T.vertices[k][1], b1.y); b1.z = minf(T.vertices[k][2], b1.z);
        }
        for (int k = 0; k < 3; k++) {
            b2.x = maxf(T.vertices[k][0], b2.x); b2.y = maxf(T.vertices[k][1], b2.y); b2.z = maxf(T.vertices[k][2], b2.z);
        }
    }
}
static inline float centroidAxis(const ElevenTri& T, int a) {                               // S/Tri.hpp:27-32
    return (T.vertices[0][a] + T.vertices[1][a] + T.vertices[2][a]) / 3.0f;
}
static inline float comp(V3 v, int a) { return a == 0 ? v.x : a == 1 ? v.y : a == 2 ? v.z : v.x; }
// float -> int as x86 cvttss2si does it (the reference relies on it for NaN/inf, S/BVH.hpp:451)
static inline int f2i(float f) {
    if (!(f > -2147483648.0f && f < 2147483648.0f)) return INT32_MIN;
    return (int)f;
}
static void divideSAH(const Scene& S, const std::vector<int>& ids, std::vector<int>& L, std::vector<int>& R) {
    if (ids.empty()) return;
    V3 tb1 = v3(0.f), tb2 = v3(0.f);
    int bestBin = 0, bestAxis = 0; float best = FLT_MAX;
    setBox(S, ids, tb1, tb2);
    for (int axis = 0; axis < 3; axis++) {
        V3 b1s[REF_BINS], b2s[REF_BINS]; int count[REF_BINS];
        for (int i = 0; i < REF_BINS; i++) { count[i] = 0; b1s[i] = v3(0.f); b2s[i] = v3(0.f); }
        float lo = comp(tb1, axis), hi = comp(tb2, axis);
        for (size_t i = 0; i < ids.size(); i++) {
            const ElevenTri& T = S.tris[ids[i]];
            int bin = 0; V3 b1 = v3(0.f), b2 = v3(0.f);
            if (lo != hi) bin = f2i(mapf(centroidAxis(T, axis), lo, hi, 0, REF_BINS - 1));
            count[bin]++;
            triBox(T, b1, b2);
            boxUnion(b1s[bin], b2s[bin], b1, b2, b1s[bin], b2s[bin]);
        }
        for (int i = 0; i < REF_BINS; i++) {
            int c1 = 0, c2 = 0; V3 b1 = v3(0.f), b2 = v3(0.f), b3 = v3(0.f), b4 = v3(0.f);
            for (int j = 0; j < i; j++) { c1 += count[j]; boxUnion(b1, b2, b1s[j], b2s[j], b1, b2); }
            for (int k = i; k < REF_BINS; k++) { c2 += count[k]; boxUnion(b3, b4, b1s[k], b2s[k], b3, b4); }
            float h = boxArea(b1, b2) * (float)c1 + boxArea(b3, b4) * (float)c2;
            if (h < best) { best = h; bestBin = i; bestAxis = axis; }
        }
    }
    float lo = comp(tb1, bestAxis), hi = comp(tb2, bestAxis);
    for (size_t i = 0; i < ids.size(); i++) {
        int bin = f2i(mapf(centroidAxis(S.tris[ids[i]], bestAxis), lo, hi, 0, REF_BINS - 1));
        if (bin < bestBin) L.push_back(ids[i]); else R.push_back(ids[i]);
    }
}
static void buildAux(Scene& S, int depth, const std::vector<int>& ids, int& nodeIdx) {      // :291-330
    V3 b1 = v3(0.f), b2 = v3(0.f);
    setBox(S, ids, b1, b2);
    RefNode n; n.b1 = b1; n.b2 = b2; n.depth = depth; n.from = 0; n.to = 0;
    if (depth == REF_DEPTH) {
        n.from = (int)S.triIndices.size(); n.to = n.from + (int)ids.size();
        S.nodes[nodeIdx++] = n;
        for (size_t i = 0; i < ids.size(); i++) S.triIndices.push_back(ids[i]);
    } else {
        S.nodes[nodeIdx++] = n;
        std::vector<int> L, R;
        divideSAH(S, ids, L, R);
        buildAux(S, depth + 1, L, nodeIdx);
        buildAux(S, depth + 1, R, nodeIdx);
    }
}
static void buildRefBVH(Scene& S) {
    S.nodes.assign((size_t)(2 << REF_DEPTH) - 1, RefNode());
    for (auto& n : S.nodes) { n.b1 = v3(0.f); n.b2 = v3(0.f); n.from = n.to = n.depth = 0; }
    S.triIndices.clear(); S.triIndices.reserve(S.tris.size());
    std::vector<int> ids(S.tris.size());
    for (size_t i = 0; i < ids.size(); i++) ids[i] = (int)i;
    int nodeIdx = 0;
    buildAux(S, 0, ids, nodeIdx);
}

// ----------------------------------------------------------------------------------------------
// HDRI: S/HDRI.hpp:107-128 (generateCDF), :130-142 (binarySearch), :145-152 (pdf), :154-162 (sample)
// ----------------------------------------------------------------------------------------------
static void buildCDF(Scene& S) {
    const Tex& t = S.hdri;
    S.cdf.assign((size_t)t.width * t.height + 1, 0.f);
    float sum = 0; S.cdf[0] = 0;
    for (int j = 0; j < t.height; j++) for (int i = 0; i < t.width; i++) { V3 d = t.texel(i, j); sum += d.x + d.y + d.z; }
    size_t c = 0;
    for (int j = 0; j < t.height; j++) for (int i = 0; i < t.width; i++) {
        V3 d = t.texel(i, j);
        S.cdf[c + 1] = S.cdf[c] + (d.x + d.y + d.z) / sum; c++;
    }
    S.radianceSum = sum;
}
static int cdfSearch(const float* arr, float value, int length) {
    int from = 0, to = length - 1;
    while (to - from > 0) {
        int m = from + (to - from) / 2;
        if (value == arr[m]) return m;
        if (value < arr[m]) to = m - 1;
        if (value > arr[m]) from = m + 1;
    }
    return to;
}
static int cdfSearchExact(const float* arr, float value, int length) {     // smallest i with arr[i + 1] >= value
    int lo = 0, hi = length - 1;
    while (lo < hi) { int m = lo + (hi - lo) / 2; if (arr[m + 1] >= value) hi = m; else lo = m + 1; }
    return lo;
}
static float hdriPdf(const Scene& S, int x, int y) {
    V3 dv = S.hdri.texel(x, y);
    float theta = (((float)y / (float)S.hdri.height)) * PIf;
    return (float)(((dv.x + dv.y + dv.z) / S.radianceSum) * S.hdri.width * S.hdri.height / (2.0 * PIf * sinf(theta)));
}

// ----------------------------------------------------------------------------------------------
// Sampling: S/Sampling.hpp:21-54
// ----------------------------------------------------------------------------------------------
static void uniformCircleSampling(float u1, float u2, float u3, float& x, float& y) {
    float t = 2 * PIf * u1, u = u2 + u3, r = u > 1 ? 2 - u : u;
    x = r * cosf(t); y = r * sinf(t);
}
static V3 cosineSampleHemisphere(float u1, float u2) {
    V3 d; float r = sqrtf(u1); float phi = (float)(2.0 * PIf * u2);
    d.x = r * cosf(phi); d.y = r * sinf(phi);
    d.z = sqrtf(maxf(0.0, (float)(1.0 - d.x * d.x - d.y * d.y)));
    return d;
}
static V3 importanceSampleGGX(float rgh, float r1, float r2) {
    float a = maxf(0.001, rgh);
    float phi = r1 * PIf * 2;
    float cosTheta = (float)sqrt((1.0 - r2) / (1.0 + (a * a - 1.0) * r2));
    float sinTheta = clampf((float)sqrt(1.0 - (cosTheta * cosTheta)), 0.0, 1.0);
    float sinPhi = sinf(phi), cosPhi = cosf(phi);
    return v3(sinTheta * cosPhi, sinTheta * sinPhi, cosTheta);
}

// ----------------------------------------------------------------------------------------------
// Disney principled BRDF: S/Disney.hpp:41-253
// ----------------------------------------------------------------------------------------------
static void createBasis(V3 n, V3& T, V3& B) {                                              // :41-45
    T = normalized(cross(v3(0, 1, 0), n));
    B = normalized(cross(n, T));
}
static float schlick(float u) { float m = clampf((float)(1.0 - u), 0.0, 1.0); float m2 = m * m; return m2 * m2 * m; } // :47-51
static float GTR1(float NDotH, float a) {                                                  // :67-73
    if (a >= 1.0) return (float)(1.0 / PIf);
    float a2 = a * a;
    float t = (float)(1.0 + (a2 - 1.0) * NDotH * NDotH);
    return (float)((a2 - 1.0) / (PIf * logf(a2) * t));
}
static float GTR2aniso(float NDotH, float HDotX, float HDotY, float ax, float ay) {        // :81-86
    float a = HDotX / ax, b = HDotY / ay;
    float c = a * a + b * b + NDotH * NDotH;
    return (float)(1.0 / (PIf * ax * ay * c * c));
}
static float smithG(float NDotV, float alphaG) {                                           // :88-92
    float a = alphaG * alphaG, b = NDotV * NDotV;
    return (float)(1.0 / (NDotV + sqrtf(a + b - a * b)));
}
static float smithGaniso(float NDotV, float VDotX, float VDotY, float ax, float ay) {      // :94-99
    float a = VDotX * ax, b = VDotY * ay, c = NDotV;
    return (float)(1.0 / (NDotV + sqrtf(a * a + b * b + c * c)));
}
static float disneyPdf(const Ray& ray, const HitData& hd, V3 L) {                          // :108-147
    V3 N = hd.normal, V = -1 * ray.d, H = normalized(L + V), T, B;
    createBasis(hd.normal, T, B);
    float NDotH = fabsf(dot(N, H));
    if (dot(N, L) <= 0.0) return 1.0;
    float clearcoatAlpha = lerpf(0.1, 0.001, hd.clearcoatGloss);
    float diffuseRatio = (float)(0.5 * (1.0 - hd.metallic));
    float specularRatio = (float)(1.0 - diffuseRatio);
    float aspect = (float)sqrt(1.0 - hd.anisotropic * 0.9);
    float ax = maxf(0.001, hd.roughness / aspect), ay = maxf(0.001, hd.roughness * aspect);
    float pdfGTR2 = GTR2aniso(NDotH, dot(H, T), dot(H, B), ax, ay) * NDotH;
    float pdfGTR1 = GTR1(NDotH, clearcoatAlpha) * NDotH;
    float ratio = (float)(1.0 / (1.0 + hd.clearcoat));
    float pdfSpec = (float)(lerpf(pdfGTR1, pdfGTR2, ratio) / (4.0 * fabsf(dot(L, H))));
    float pdfDiff = (float)(fabsf(dot(L, N)) * (1.0 / PIf));
    return diffuseRatio * pdfDiff + specularRatio * pdfSpec;
}
static V3 disneySample(const Ray& ray, const HitData& hd, float r1, float r2, float r3) {  // :150-177
    V3 N = hd.normal, V = -1 * ray.d, T, B;
    createBasis(hd.normal, T, B);
    float diffuseRatio = (float)(0.5 * (1.0 - hd.metallic));
    if (r3 < diffuseRatio) {
        V3 H = cosineSampleHemisphere(r1, r2);
        return T * H.x + B * H.y + N * H.z;
    }
    V3 H = importanceSampleGGX(hd.roughness, r1, r2);
    H = T * H.x + B * H.y + N * H.z;
    return reflectV(-1 * V, H);
}
static V3 disneyEval(const Ray& ray, const HitData& hd, V3 L) {                            // :179-253
    V3 V = -1 * ray.d, T = v3(0.f), B = v3(0.f);
    createBasis(hd.normal, T, B);
    V3 H = normalized(L + V);
    float NDotL = fabsf(dot(hd.normal, L)), NDotV = fabsf(dot(hd.normal, V)), NDotH = fabsf(dot(hd.normal, H));
    float LDotH = fabsf(dot(L, H));
    V3 brdf = v3(0.f);
    if (hd.transmission < 1.0 && dot(hd.normal, L) > 0.0 && dot(hd.normal, V) > 0.0) {
        V3 Cd = hd.albedo;
        float Cdlum = (float)(0.3 * Cd.x + 0.6 * Cd.y + 0.1 * Cd.z);
        V3 Ctint = Cdlum > 0.0 ? Cd / Cdlum : v3(1.0f);
        V3 Cspec0 = lerpV((float)(hd.specular * 0.08) * lerpV(v3(1.0f), Ctint, hd.specularTint), Cd, hd.metallic);
        V3 Csheen = lerpV(v3(1.0f), Ctint, hd.sheenTint);
        float FL = schlick(NDotL), FV = schlick(NDotV);
        float Fd90 = (float)(0.5 + 2.0 * LDotH * LDotH * hd.roughness);
        float Fd = lerpf(1.0, Fd90, FL) * lerpf(1.0, Fd90, FV);
        float Fss90 = LDotH * LDotH * hd.roughness;
        float Fss = lerpf(1.0, Fss90, FL) * lerpf(1.0, Fss90, FV);
        float ss = (float)(1.25 * (Fss * (1.0 / (NDotL + NDotV) - 0.5) + 0.5));
        float aspect = (float)sqrt(1.0 - hd.anisotropic * 0.9);
        float ax = maxf(0.001, hd.roughness / aspect), ay = maxf(0.001, hd.roughness * aspect);
        float Ds = GTR2aniso(NDotH, dot(H, T), dot(H, B), ax, ay);
        float FH = schlick(LDotH);
        V3 Fs = lerpV(Cspec0, v3(1.0f), FH);
        float Gs = smithGaniso(NDotL, dot(L, T), dot(L, B), ax, ay);
        Gs *= smithGaniso(NDotV, dot(V, T), dot(V, B), ax, ay);
        V3 Fsheen = (FH * hd.sheen) * Csheen;
        float Dr = GTR1(NDotH, lerpf(0.1, 0.001, hd.clearcoatGloss));
        float Fr = lerpf(0.04, 1.0, FH);
        float Gr = smithG(NDotL, 0.25) * smithG(NDotV, 0.25);
        V3 diffuse = ((float)((1.0 / PIf) * lerpf(Fd, ss, hd.subsurface)) * Cd + Fsheen) * (float)(1.0 - hd.metallic);
        V3 spec = (Gs * Fs) * Ds;
        float coat = (float)(0.25 * hd.clearcoat * Gr * Fr * Dr);
        brdf = addScalar(diffuse + spec, coat);
    }
    return brdf;
}

// ----------------------------------------------------------------------------------------------
// S/kernel.cu:54-119 generateHitData
// ----------------------------------------------------------------------------------------------
static void generateHitData(const Scene& S, const ElevenMaterial& m, HitData& hd, const Hit& hit) {
    V3 normal = hit.normal, tangent = hit.tangent, bitangent = hit.bitangent;
    hd.albedo = m.albedoTextureID < 0 ? v3(m.albedo) : S.tex[m.albedoTextureID].filtered(hit.tu, hit.tv);
    hd.emission = m.emissionTextureID < 0 ? v3(m.emission) : S.tex[m.emissionTextureID].filtered(hit.tu, hit.tv);
    hd.roughness = m.roughnessTextureID < 0 ? m.roughness : S.tex[m.roughnessTextureID].filtered(hit.tu, hit.tv).x;
    hd.metallic = m.metallicTextureID < 0 ? m.metallic : S.tex[m.metallicTextureID].filtered(hit.tu, hit.tv).x;
    if (m.normalTextureID < 0) hd.normal = normal;
    else {
        V3 nc = S.tex[m.normalTextureID].fromUV(hit.tu, hit.tv);
        V3 ln = (nc * 2.0f) - v3(1.0f);
        hd.normal = normalized(ln.x * tangent - ln.y * bitangent + ln.z * normal);
    }
    hd.roughness = powf(hd.roughness, 2.2f);
    hd.metallic = powf(hd.metallic, 2.2f);
    hd.clearcoatGloss = m.clearcoatGloss; hd.clearcoat = m.clearcoat; hd.anisotropic = m.anisotropic;
    hd.eta = m.eta; hd.transmission = m.transmission; hd.specular = m.specular; hd.specularTint = m.specularTint;
    hd.sheenTint = m.sheenTint; hd.subsurface = m.subsurface; hd.sheen = m.sheen;
    hd.tangent = tangent; hd.bitangent = bitangent;
}

// S/kernel.cu:260-337
static Ray cameraRay(const ElevenCamera& c, int x, int y, float r1, float r2, float r3, float r4, float r5) {
    V3 pos = v3(c.position);
    float dx = pos.x + ((float)x) / ((float)c.xRes) * c.sensorWidth;
    float dy = pos.y + ((float)y) / ((float)c.yRes) * c.sensorHeight;
    float odx = (float)((-c.sensorWidth / 2.0) + dx);
    float ody = (float)((-c.sensorHeight / 2.0) + dy);
    float rx = (float)((1.0 / (float)c.xRes) * (r1 - 0.5) * c.sensorWidth);
    float ry = (float)((1.0 / (float)c.yRes) * (r2 - 0.5) * c.sensorHeight);
    float SPx = odx + rx, SPy = ody + ry, SPz = pos.z + c.focalLength;
    float k = (float)(PIf / 180.0);
    V3 rot = v3(c.rotation[0] * k, c.rotation[1] * k, c.rotation[2] * k);
    V3 dir = v3(SPx, SPy, SPz) - pos;
    V3 dX = v3(dir.x, dir.y * cosf(rot.x) - dir.z * sinf(rot.x), dir.y * sinf(rot.x) + dir.z * cosf(rot.x));
    V3 dY = v3(dX.x * cosf(rot.y) + dX.z * sinf(rot.y), dX.y, dX.z * cosf(rot.y) - dX.x * sinf(rot.y));
    V3 dZ = v3(dY.x * cosf(rot.z) - dY.y * sinf(rot.z), dY.x * sinf(rot.z) + dY.y * cosf(rot.z), dY.z);
    Ray ray = makeRay(pos, dZ);
    if (c.bokeh) {
        float diameter = c.focalLength / c.aperture;
        float l = c.focusDistance + c.focalLength;
        V3 focusPoint = ray.o + ray.d * l;
        float ix, iy;
        uniformCircleSampling(r3, r4, r5, ix, iy);
        ix = (float)(ix * (diameter * 0.5)); iy = (float)(iy * (diameter * 0.5));
        V3 orig = pos + v3(ix, iy, 0);
        ray = makeRay(orig, focusPoint - orig);
    }
    return ray;
}

// S/kernel.cu:210-258 (HDRIIS branch).  Returns the contribution; pdf = 0 when occluded (defined UB).
static V3 hdriLight(Scene& S, const Ray& ray, V3 point, const HitData& hd, float r1, float& pdf) {
    const Tex& t = S.hdri;
    int count = S.envSearchExact ? cdfSearchExact(S.cdf.data(), r1, t.width * t.height) : cdfSearch(S.cdf.data(), r1, t.width * t.height);
    float sx = (float)(count % t.width), sy = (float)(count / t.width);                      // HDRI::sample :154-162
    float nu = sx / (float)t.width, nv = sy / (float)t.height;
    float iu, iv; t.inverseTransformUV(nu, nv, iu, iv);
    V3 newDir = normalized(reverseSphericalMapping(iu, iv)) * -1.0f;
    Ray shadow = makeRay(point + newDir * 0.001f, newDir);
    S.raysEnv++;
    Hit sh = throwRay(S, shadow);
    if (sh.valid) {
        // UB in the reference: hdriPdf is left uninitialised here (S/kernel.cu:248,344).  Defined as 0 (mode 0).
        // Modes 1/2 exist only to probe what the compiled reference actually does (tools/, not used by tests).
        static int ubMode = getenv("ORC_HDRIPDF_UB") ? atoi(getenv("ORC_HDRIPDF_UB")) : 0;
        if (ubMode == 1) pdf = hdriPdf(S, (int)(iu * t.width), (int)(iv * t.height));
        else if (ubMode == 0) pdf = 0;                                   // modes 2,3: keep the stale value
        return v3(0.f);
    }
    V3 val = t.fromUV(iu, iv);
    V3 brdf = disneyEval(ray, hd, newDir);
    pdf = hdriPdf(S, (int)(iu * t.width), (int)(iv * t.height));
    return brdf * fabsf(dot(newDir, hd.normal)) * val / pdf;
}
// S/kernel.cu:175-205
static V3 pointLight(Scene& S, const Ray& ray, const HitData& hd, V3 point, float& pdf, float r1) {
    uint32_t n = (uint32_t)S.lights.size();
    if (n <= 0) { pdf = 0; return v3(0.f); }
    pdf = (float)(((float)n) / (2.0 * PIf));
    int li = (int)(n * r1); if (li >= (int)n) li = (int)n - 1;
    const ElevenPointLight& light = S.lights[li];
    V3 lp = v3(light.position);
    V3 newDir = normalized(lp - point);
    float dist = length(lp - point);
    Ray shadow = makeRay(point + newDir * 0.001f, newDir);
    S.raysLight++;
    Hit sh = throwRay(S, shadow);
    float shadowDist = length(sh.position - point);
    if (sh.valid && shadowDist < dist) return v3(0.f);
    V3 value = v3(light.radiance) / (dist * dist);
    V3 brdf = disneyEval(ray, hd, newDir);
    return value * brdf * fabsf(dot(newDir, hd.normal)) / pdf;
}
static thread_local bool g_debugPixel = false;
static thread_local float g_stalePdf = 0;   // probe modes 2/3 only (see hdriLight)
// S/kernel.cu:339-358
static void shade(Scene& S, const Ray& ray, const HitData& hd, const Hit& hit, V3 newDir, float r1, V3& hitLight, V3& reduction) {
    V3 brdf = disneyEval(ray, hd, newDir);
    float brdfPdf = disneyPdf(ray, hd, newDir);
    float hPdf = g_stalePdf, pPdf = 0;
    V3 hdriCalc = hdriLight(S, ray, hit.position, hd, r1, hPdf);
    g_stalePdf = hPdf;
    V3 pointCalc = pointLight(S, ray, hd, hit.position, pPdf, r1);
    V3 brdfCalc = hd.emission * (brdf * fabsf(dot(newDir, hd.normal))) / brdfPdf;
    float w1 = hPdf / (hPdf + pPdf + brdfPdf);
    float w2 = pPdf / (hPdf + pPdf + brdfPdf);
    float w3 = brdfPdf / (hPdf + pPdf + brdfPdf);
    hitLight = reduction * (w1 * hdriCalc + w2 * pointCalc + w3 * brdfCalc);
    if (g_debugPixel) fprintf(stderr, "  [orc] N=(%g %g %g) L=(%g %g %g) brdf=(%g %g %g) brdfPdf=%g hPdf=%g pPdf=%g hdriCalc=(%g %g %g) pointCalc.x=%g w=(%g %g %g) red=(%g %g %g)\n",
        hd.normal.x, hd.normal.y, hd.normal.z, newDir.x, newDir.y, newDir.z, brdf.x, brdf.y, brdf.z, brdfPdf, hPdf, pPdf, hdriCalc.x, hdriCalc.y, hdriCalc.z, pointCalc.x, w1, w2, w3, reduction.x, reduction.y, reduction.z);
    reduction = reduction * ((brdf * fabsf(dot(newDir, hd.normal))) / brdfPdf);
}

// S/kernel.cu:369-481: one pixel-sample.  (x, y) are the thread coordinates; film/rng index is
// W*(H-1-y)+x.
static void renderPixelSample(Scene& S, int x, int y, int maxBounces) {
    const ElevenCamera& c = S.cam;
    int idx = (int)(c.xRes * (c.yRes - y - 1) + x);
    Xorwow rs = S.rng[idx];
    uint32_t sa = S.samples[idx];
    { static int ubMode = getenv("ORC_HDRIPDF_UB") ? atoi(getenv("ORC_HDRIPDF_UB")) : 0; if (ubMode == 3) g_stalePdf = 0; }
    float c1 = xorwowUniform(rs), c2 = xorwowUniform(rs), c3 = xorwowUniform(rs), c4 = xorwowUniform(rs), c5 = xorwowUniform(rs);
    Ray ray = cameraRay(c, x, y, c1, c2, c3, c4, c5);
    { static const char* dp = getenv("ORC_DEBUG_PIXEL"); int dx = -1, dy = -1; if (dp) sscanf(dp, "%d,%d", &dx, &dy); g_debugPixel = dp && (int)(c.yRes - y - 1) == dy && x == dx; }
    V3 light = v3(0.f), normal = v3(0.f), tangent = v3(0.f), bitangent = v3(0.f), reduction = v3(1.f);
    int i = 0;
    for (i = 0; i < maxBounces; i++) {
        S.raysExt++;
        Hit hit = throwRay(S, ray);
        if (!hit.valid) {
            float u, v; sphericalMapping(-1 * ray.d, u, v);
            light = light + S.hdri.filtered(u, v) * reduction;
            break;
        }
        if (g_debugPixel) fprintf(stderr, "[orc] bounce %d tri %d t=%g pos=(%g %g %g) dir=(%g %g %g)\n", i, hit.tri, hit.t, hit.position.x, hit.position.y, hit.position.z, ray.d.x, ray.d.y, ray.d.z);
        const ElevenMaterial& m = S.mats[S.objMat[hit.objectID]];
        HitData hd; generateHitData(S, m, hd, hit);
        float b1 = xorwowUniform(rs), b2 = xorwowUniform(rs), b3 = xorwowUniform(rs);
        V3 bounced = disneySample(ray, hd, b1, b2, b3);
        float s1 = xorwowUniform(rs), s2 = xorwowUniform(rs), s3 = xorwowUniform(rs); (void)s2; (void)s3;
        V3 hitLight;
        shade(S, ray, hd, hit, bounced, s1, hitLight, reduction);
        light = light + hitLight;
        if (i == 0) { normal = hit.normal; tangent = hit.tangent; bitangent = hit.bitangent; }
        ray = makeRay(hit.position + bounced * 0.001f, bounced);
    }
    S.pathcount[idx] += i;
    light = v3(clampf(light.x, 0, 10), clampf(light.y, 0, 10), clampf(light.z, 0, 10));
    if (!std::isnan(light.x) && !std::isnan(light.y) && !std::isnan(light.z)) {
        float* P[4] = {&S.passes[ELEVEN_PASS_BEAUTY][4 * idx], &S.passes[ELEVEN_PASS_NORMAL][4 * idx],
                       &S.passes[ELEVEN_PASS_TANGENT][4 * idx], &S.passes[ELEVEN_PASS_BITANGENT][4 * idx]};
        V3 val[4] = {light, normal, tangent, bitangent};
        if (sa > 0) for (int p = 0; p < 4; p++) for (int k = 0; k < 3; k++) P[p][k] *= ((float)sa) / ((float)(sa + 1));
        for (int p = 0; p < 4; p++) {
            P[p][0] += val[p].x / ((float)sa + 1); P[p][1] += val[p].y / ((float)sa + 1); P[p][2] += val[p].z / ((float)sa + 1);
        }
        S.samples[idx]++;
    }
    S.rng[idx] = rs;
}

static void resetFilm(Scene& S) {                                                          // setupKernel, S/kernel.cu:121-150
    size_t n = (size_t)S.cam.xRes * S.cam.yRes;
    for (int p = 0; p < ELEVEN_PASS_COUNT; p++) {
        S.passes[p].assign(n * 4, 0.f);
        for (size_t i = 0; i < n; i++) S.passes[p][4 * i + 3] = 1.f;
    }
    S.samples.assign(n, 0); S.pathcount.assign(n, 0);
    S.rng.resize(n);
    // sequence idx+1 = M^(2^67) applied to sequence idx: one matvec per pixel
    buildSeqMats();
    Xorwow s = xorwowInit(0, 0);
    for (size_t i = 0; i < n; i++) { S.rng[i] = s; matVec(g_seqMat[0], s.v, s.v); }
    S.raysExt = 0; S.raysEnv = 0; S.raysLight = 0;
}

static Tex copyTex(const ElevenTexture& t) {
    Tex r; r.width = t.width; r.height = t.height; r.xTile = t.xTile; r.yTile = t.yTile;
    r.xOffset = t.xOffset; r.yOffset = t.yOffset; r.filter = t.filter;
    size_t n = (size_t)t.width * t.height * 3;
    r.data.resize(n);
    if (t.format == ELEVEN_TEX_F32_RGB) memcpy(r.data.data(), t.data, n * sizeof(float));
    else {
        float gamma = t.format == ELEVEN_TEX_U8_SRGB ? 2.2f : 1.0f;
        float lut[256]; for (int i = 0; i < 256; i++) lut[i] = texelDecode(i, gamma);
        const uint8_t* b = (const uint8_t*)t.data;
        for (size_t i = 0; i < n; i++) r.data[i] = lut[b[i]];
    }
    return r;
}

} // namespace

// ==============================================================================================
// C API (ctypes)
// ==============================================================================================
extern "C" {

void* orc_create(const ElevenSceneDesc* d, int buildBvh) {
    Scene* S = new Scene();
    S->cam = d->camera;
    S->tris.assign(d->tris, d->tris + d->triCount);
    S->objMat.assign(d->objectMaterial, d->objectMaterial + d->objectCount);
    S->mats.assign(d->materials, d->materials + d->materialCount);
    for (uint32_t i = 0; i < d->textureCount; i++) S->tex.push_back(copyTex(d->textures[i]));
    S->hdri = copyTex(d->hdri);
    buildCDF(*S);
    if (d->pointLightCount) S->lights.assign(d->pointLights, d->pointLights + d->pointLightCount);
    if (buildBvh) buildRefBVH(*S);
    resetFilm(*S);
    return S;
}
void orc_destroy(void* h) { delete (Scene*)h; }

int orc_bvh_node_count(void* h) { return (int)((Scene*)h)->nodes.size(); }
// out: per node 9 values as float/int bits: b1(3) b2(3) from to depth -> written into two arrays
void orc_bvh_dump(void* h, float* boxes /*n*6*/, int32_t* meta /*n*3*/, int32_t* triIndices) {
    Scene& S = *(Scene*)h;
    for (size_t i = 0; i < S.nodes.size(); i++) {
        const RefNode& n = S.nodes[i];
        boxes[6 * i + 0] = n.b1.x; boxes[6 * i + 1] = n.b1.y; boxes[6 * i + 2] = n.b1.z;
        boxes[6 * i + 3] = n.b2.x; boxes[6 * i + 4] = n.b2.y; boxes[6 * i + 5] = n.b2.z;
        meta[3 * i + 0] = n.from; meta[3 * i + 1] = n.to; meta[3 * i + 2] = n.depth;
    }
    memcpy(triIndices, S.triIndices.data(), S.triIndices.size() * sizeof(int));
}

// mode 0: reference BVH order; mode 1: brute force over all triangles in index order.
// full (optional): 14 floats per ray: position, normal, tangent, bitangent, tu, tv; objectID in objIds.
void orc_trace(void* h, const float* rays, size_t n, ElevenHit* hits, float* full, int32_t* objIds, int mode, int nthreads) {
    Scene& S = *(Scene*)h;
    if (nthreads < 1) nthreads = 1;
    std::atomic<size_t> next(0);
    auto work = [&]() {
        for (;;) {
            size_t b = next.fetch_add(256); if (b >= n) break;
            size_t e = b + 256 < n ? b + 256 : n;
            for (size_t i = b; i < e; i++) {
                Ray r = makeRay(v3(rays + 6 * i), v3(rays + 6 * i + 3));
                Hit hit = noHit();
                if (mode == 0) traverseRef(S, r, 0, 0, hit); else traverseBrute(S, r, hit);
                ElevenHit& o = hits[i];
                if (hit.valid) { o.tri = hit.tri; o.t = hit.t; o.u = hit.u; o.v = hit.v; o.key = length(hit.position - r.o); }
                else { o.tri = -1; o.t = o.u = o.v = o.key = 0; }
                if (full) {
                    float* f = full + 14 * i;
                    f[0] = hit.position.x; f[1] = hit.position.y; f[2] = hit.position.z;
                    f[3] = hit.normal.x; f[4] = hit.normal.y; f[5] = hit.normal.z;
                    f[6] = hit.tangent.x; f[7] = hit.tangent.y; f[8] = hit.tangent.z;
                    f[9] = hit.bitangent.x; f[10] = hit.bitangent.y; f[11] = hit.bitangent.z;
                    f[12] = hit.tu; f[13] = hit.tv;
                }
                if (objIds) objIds[i] = hit.valid ? hit.objectID : -1;
            }
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; t++) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
}

// Renders `spp` more samples of the pixel rows [y0, y1) (thread coordinates, like blockIdx*8+threadIdx).
void orc_render_rows(void* h, int spp, int y0, int y1, int maxBounces, int nthreads) {
    Scene& S = *(Scene*)h;
    if (nthreads < 1) nthreads = 1;
    int W = (int)S.cam.xRes;
    std::atomic<int> next(y0);
    auto work = [&]() {
        for (;;) {
            int y = next.fetch_add(1); if (y >= y1) break;
            for (int x = 0; x < W; x++) for (int s = 0; s < spp; s++) renderPixelSample(S, x, y, maxBounces);
        }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; t++) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
}
void orc_render(void* h, int spp, int maxBounces, int nthreads) {
    Scene& S = *(Scene*)h; orc_render_rows(h, spp, 0, (int)S.cam.yRes, maxBounces, nthreads);
}
void orc_reset(void* h) { resetFilm(*(Scene*)h); }
void orc_get_film(void* h, int pass, float* out) { Scene& S = *(Scene*)h; memcpy(out, S.passes[pass].data(), S.passes[pass].size() * sizeof(float)); }
void orc_get_counts(void* h, uint32_t* samples, uint32_t* pathcount) {
    Scene& S = *(Scene*)h;
    if (samples) memcpy(samples, S.samples.data(), S.samples.size() * 4);
    if (pathcount) memcpy(pathcount, S.pathcount.data(), S.pathcount.size() * 4);
}
void orc_get_ray_counts(void* h, uint64_t* out3) { Scene& S = *(Scene*)h; out3[0] = S.raysExt; out3[1] = S.raysEnv; out3[2] = S.raysLight; }

// ---- leaf functions for the golden-vector tests ------------------------------------------------
void orc_xorwow_uniforms(uint64_t seed, uint64_t subsequence, int n, float* out, uint32_t* state6) {
    Xorwow s = xorwowInit(seed, subsequence);
    if (state6) { memcpy(state6, s.v, 20); state6[5] = s.d; }
    for (int i = 0; i < n; i++) out[i] = xorwowUniform(s);
}
void orc_texel_table(float gamma, float* out256) { for (int i = 0; i < 256; i++) out256[i] = texelDecode(i, gamma); }
double orc_fastpow(double a, double b) { return fastPowRef(a, b); }

static HitData hdFrom(const float* p) {   // 12 scalars in HitData order + emission, albedo, normal (tangent/bitangent unused by the BRDF)
    HitData hd; memset(&hd, 0, sizeof hd);
    hd.metallic = p[0]; hd.roughness = p[1]; hd.clearcoatGloss = p[2]; hd.clearcoat = p[3]; hd.anisotropic = p[4]; hd.eta = p[5];
    hd.transmission = p[6]; hd.specular = p[7]; hd.specularTint = p[8]; hd.sheenTint = p[9]; hd.subsurface = p[10]; hd.sheen = p[11];
    hd.emission = v3(p + 12); hd.albedo = v3(p + 15); hd.normal = v3(p + 18);
    return hd;
}
// in: hd[21], rayDir[3] (normalised by Ray ctor), L[3]; out: eval[3], pdf
void orc_disney_eval_pdf(const float* hd21, const float* rayDir, const float* L, float* out4) {
    HitData hd = hdFrom(hd21); Ray r = makeRay(v3(0.f), v3(rayDir));
    V3 e = disneyEval(r, hd, v3(L)); out4[0] = e.x; out4[1] = e.y; out4[2] = e.z; out4[3] = disneyPdf(r, hd, v3(L));
}
void orc_disney_sample(const float* hd21, const float* rayDir, const float* r3, float* out3) {
    HitData hd = hdFrom(hd21); Ray r = makeRay(v3(0.f), v3(rayDir));
    V3 d = disneySample(r, hd, r3[0], r3[1], r3[2]); out3[0] = d.x; out3[1] = d.y; out3[2] = d.z;
}
void orc_camera_ray(void* h, int x, int y, const float* r5, float* out6) {
    Scene& S = *(Scene*)h; Ray r = cameraRay(S.cam, x, y, r5[0], r5[1], r5[2], r5[3], r5[4]);
    out6[0] = r.o.x; out6[1] = r.o.y; out6[2] = r.o.z; out6[3] = r.d.x; out6[4] = r.d.y; out6[5] = r.d.z;
}
// HDRI::sample(r) -> (x, y) and the NEE direction + pdf derived from it
void orc_hdri_sample(void* h, const float* r, int n, int32_t* xy, float* dir3, float* pdf) {
    Scene& S = *(Scene*)h; const Tex& t = S.hdri;
    for (int i = 0; i < n; i++) {
        int count = S.envSearchExact ? cdfSearchExact(S.cdf.data(), r[i], t.width * t.height) : cdfSearch(S.cdf.data(), r[i], t.width * t.height);
        int x = count % t.width, y = count / t.width;
        xy[2 * i] = x; xy[2 * i + 1] = y;
        float nu = (float)x / (float)t.width, nv = (float)y / (float)t.height, iu, iv;
        t.inverseTransformUV(nu, nv, iu, iv);
        V3 d = normalized(reverseSphericalMapping(iu, iv)) * -1.0f;
        dir3[3 * i] = d.x; dir3[3 * i + 1] = d.y; dir3[3 * i + 2] = d.z;
        pdf[i] = hdriPdf(S, (int)(iu * t.width), (int)(iv * t.height));
    }
}
// 0 = the reference's binarySearch (default), 1 = exact inversion (yardstick for the alias-table mode; see Scene::envSearchExact)
void orc_set_env_search(void* h, int exact) { ((Scene*)h)->envSearchExact = exact ? 1 : 0; }
void orc_hdri_cdf(void* h, float* cdf, float* radianceSum) {
    Scene& S = *(Scene*)h; if (cdf) memcpy(cdf, S.cdf.data(), S.cdf.size() * 4); *radianceSum = S.radianceSum;
}
void orc_spherical_mapping(const float* dirs, int n, float* uv) {
    for (int i = 0; i < n; i++) { float u, v; sphericalMapping(-1 * v3(dirs + 3 * i), u, v); uv[2 * i] = u; uv[2 * i + 1] = v; }
}
void orc_env_lookup(void* h, const float* dirs, int n, float* rgb) {
    Scene& S = *(Scene*)h;
    for (int i = 0; i < n; i++) { float u, v; sphericalMapping(-1 * v3(dirs + 3 * i), u, v); V3 c = S.hdri.filtered(u, v); rgb[3 * i] = c.x; rgb[3 * i + 1] = c.y; rgb[3 * i + 2] = c.z; }
}
// generateHitData on a given full hit (position unused): in 14 floats like orc_trace's `full` + objectID; out HitData as 27 floats
void orc_hitdata(void* h, const float* full14, int objectID, float* out27) {
    Scene& S = *(Scene*)h; Hit hit = noHit();
    hit.normal = v3(full14 + 3); hit.tangent = v3(full14 + 6); hit.bitangent = v3(full14 + 9); hit.tu = full14[12]; hit.tv = full14[13];
    HitData hd; generateHitData(S, S.mats[S.objMat[objectID]], hd, hit);
    float* o = out27;
    o[0] = hd.metallic; o[1] = hd.roughness; o[2] = hd.clearcoatGloss; o[3] = hd.clearcoat; o[4] = hd.anisotropic; o[5] = hd.eta;
    o[6] = hd.transmission; o[7] = hd.specular; o[8] = hd.specularTint; o[9] = hd.sheenTint; o[10] = hd.subsurface; o[11] = hd.sheen;
    o[12] = hd.emission.x; o[13] = hd.emission.y; o[14] = hd.emission.z; o[15] = hd.albedo.x; o[16] = hd.albedo.y; o[17] = hd.albedo.z;
    o[18] = hd.normal.x; o[19] = hd.normal.y; o[20] = hd.normal.z; o[21] = hd.tangent.x; o[22] = hd.tangent.y; o[23] = hd.tangent.z;
    o[24] = hd.bitangent.x; o[25] = hd.bitangent.y; o[26] = hd.bitangent.z;
}

} // extern "C"
