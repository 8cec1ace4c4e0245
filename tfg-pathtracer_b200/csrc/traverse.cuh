/*
 * traverse.cuh — ray / BVH8 traversal and ray / triangle intersection (device).
 *
 * Replaces BVH::transverse + BVH::intersect + BVH::intersectNode + Tri::hit (S/BVH.hpp:69-175, S/Tri.hpp:38-160):
 *   - 8-wide quantised nodes fetched with 5 x 128-bit loads instead of two 44-byte node copies per binary step;
 *   - the slab test runs on the quantised grid with one FMA per plane (1/dir computed once per ray, not 3 divisions
 *     per box as S/BVH.hpp:75-77);
 *   - culling against the best hit (the reference has none, S/BVH.hpp:88-99) made exact w.r.t. the reference's
 *     ordering key through a per-scene slack (bvh8.h: keySlack);
 *   - triangles read as 3 x 128-bit loads in leaf order (no triIndices indirection to a 152-byte AoS record).
 * The Moeller-Trumbore arithmetic is the reference's, op for op, in uncontracted IEEE float (ex::), so that
 * (triangle, t, u, v) are bit-identical to the reference built with --fmad=false.
 */
#pragma once
#include "common.cuh"

namespace eleven {

struct Ray { F3 o, d; };

// S/Ray.hpp:14-18: the constructor normalises the direction
__device__ __forceinline__ Ray makeRay(F3 o, F3 d) { Ray r; r.o = o; r.d = ex::normalize(d); return r; }

struct HitRec { int tri; float t, u, v, key; };

// ---- S/Tri.hpp:38-68 -------------------------------------------------------------------------------
__device__ __forceinline__ bool mollerTrumbore(const Ray& r, F3 v0, F3 e1, F3 e2, float& t, float& u, float& v) {
    const float EPSILON = 0.0000001f;
    F3 pvec = ex::cross(r.d, e2);
    float det = ex::dot(e1, pvec);
    float inv_det = ex::div(1.0f, det);            // (float)(1.0 / (double)det) == 1.0f / det in IEEE (innocuous double rounding)
    if (det > -EPSILON && det < EPSILON) return false;
    F3 tvec = ex::sub(r.o, v0);
    u = ex::mul(ex::dot(tvec, pvec), inv_det);
    if (u < 0.0f || u > 1.0f) return false;
    F3 qvec = ex::cross(tvec, e1);
    v = ex::mul(ex::dot(r.d, qvec), inv_det);
    if (v < 0.0f || ex::add(u, v) > 1.0f) return false;
    t = ex::mul(ex::dot(e2, qvec), inv_det);
    if (t < 0.0f) return false;
    return true;
}

// The same test in the arithmetic of the reference's SHIPPING build (-use_fast_math: contracted multiply-adds, MUFU reciprocal): the
// render kernels of the fast-math configuration use it; (t, u, v) differ from the exact ones by rounding only.  The closest-hit
// contract (eleven_trace_closest, the parity configuration) always runs mollerTrumbore above.
__device__ __forceinline__ bool mollerTrumboreFast(const Ray& r, F3 v0, F3 e1, F3 e2, float& t, float& u, float& v) {
    const float EPSILON = 0.0000001f;
    const float px = fmaf(r.d.y, e2.z, -(r.d.z * e2.y)), py = fmaf(r.d.z, e2.x, -(r.d.x * e2.z)), pz = fmaf(r.d.x, e2.y, -(r.d.y * e2.x));
    const float det = fmaf(e1.x, px, fmaf(e1.y, py, e1.z * pz));
    if (fabsf(det) < EPSILON) return false;
    const float inv_det = __fdividef(1.0f, det);
    const float tx = r.o.x - v0.x, ty = r.o.y - v0.y, tz = r.o.z - v0.z;
    u = fmaf(tx, px, fmaf(ty, py, tz * pz)) * inv_det;
    if (u < 0.0f || u > 1.0f) return false;
    const float qx = fmaf(ty, e1.z, -(tz * e1.y)), qy = fmaf(tz, e1.x, -(tx * e1.z)), qz = fmaf(tx, e1.y, -(ty * e1.x));
    v = fmaf(r.d.x, qx, fmaf(r.d.y, qy, r.d.z * qz)) * inv_det;
    if (v < 0.0f || u + v > 1.0f) return false;
    t = fmaf(e2.x, qx, fmaf(e2.y, qy, e2.z * qz)) * inv_det;
    return t >= 0.0f;
}

// ---- S/Tri.hpp:70-92: hit position with the shadow-terminator shift ---------------------------------
struct TriGeom { F3 v0, v1, v2, n0, n1, n2; };

__device__ __forceinline__ TriGeom loadTriGeom(const float4* __restrict__ shadeTris, int tri) {
    const float4* p = shadeTris + (size_t)tri * 9;
    float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3), e = __ldg(p + 4);
    TriGeom g;
    g.v0 = f3(a.x, a.y, a.z); g.v1 = f3(a.w, b.x, b.y); g.v2 = f3(b.z, b.w, c.x);
    g.n0 = f3(c.y, c.z, c.w); g.n1 = f3(d.x, d.y, d.z); g.n2 = f3(d.w, e.x, e.y);
    return g;
}
// v0 + (v1 - v0)*u + (v2 - v0)*v   (Vector3 expression order)
__device__ __forceinline__ F3 baryLerp(F3 a, F3 b, F3 c, float u, float v) {
    return ex::add(ex::add(a, ex::mul(ex::sub(b, a), u)), ex::mul(ex::sub(c, a), v));
}
__device__ __forceinline__ F3 projectOnPlane(F3 p, F3 o, F3 n) {       // S/Tri.hpp:34-36: position - dot(position - origin, normal) * normal
    float s = ex::dot(ex::sub(p, o), n);
    return ex::sub(p, f3(ex::mul(n.x, s), ex::mul(n.y, s), ex::mul(n.z, s)));
}
__device__ __forceinline__ F3 hitPosition(const Ray& r, const TriGeom& g, float t, float u, float v, F3& shadingNormal) {
    F3 geom = ex::madd(r.o, r.d, t);
    shadingNormal = baryLerp(g.n0, g.n1, g.n2, u, v);
    F3 p0 = projectOnPlane(geom, g.v0, g.n0), p1 = projectOnPlane(geom, g.v1, g.n1), p2 = projectOnPlane(geom, g.v2, g.n2);
    F3 sp = baryLerp(p0, p1, p2, u, v);
    bool convex = ex::dot(ex::sub(sp, geom), shadingNormal) > 0.0f;
    return convex ? sp : geom;
}
// the reference's ordering key: |hit.position - ray.origin| (S/BVH.hpp:170)
__device__ __forceinline__ float hitKey(const Ray& r, F3 pos) { return ex::length(ex::sub(pos, r.o)); }

// ---- BVH8 traversal -------------------------------------------------------------------------------
enum { TRACE_CLOSEST_KEY = 0, TRACE_CLOSEST_T = 1, TRACE_ANY = 2 };

struct TraceCounters { uint32_t nodes, tris, keys; };

#define EL_STACK 40

__device__ __forceinline__ uint32_t extractByte(uint32_t x, uint32_t i) { return (x >> (i * 8)) & 0xffu; }

} // namespace eleven
