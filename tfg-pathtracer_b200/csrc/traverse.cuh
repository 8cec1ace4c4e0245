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

struct TraceCounters { uint32_t nodes, tris; };

#define EL_STACK 40

__device__ __forceinline__ uint32_t extractByte(uint32_t x, uint32_t i) { return (x >> (i * 8)) & 0xffu; }

/* Traverses the BVH8.  MODE: closest by reference key, closest by t, or any hit with t in [0, tmax).
 * `tmax` bounds t for TRACE_ANY (use INFINITY for "any hit at all").  Returns true when something was hit. */
template <int MODE, bool COUNT>
__device__ __forceinline__ bool traverse(const DevScene& S, const Ray& ray, float tmaxAny, HitRec& best, TraceCounters* cnt) {
    best.tri = -1; best.t = 0.f; best.u = 0.f; best.v = 0.f; best.key = 0.f;
    if (S.nodeCount == 0) return false;

    const F3 o = ray.o, d = ray.d;
    // 1/d with |d| clamped away from zero: keeps every slab distance finite (no inf*0 NaNs), still conservative
    const float dx = fabsf(d.x) > 1e-20f ? d.x : copysignf(1e-20f, d.x);
    const float dy = fabsf(d.y) > 1e-20f ? d.y : copysignf(1e-20f, d.y);
    const float dz = fabsf(d.z) > 1e-20f ? d.z : copysignf(1e-20f, d.z);
    const float idx = 1.0f / dx, idy = 1.0f / dy, idz = 1.0f / dz;
    const uint32_t octinv = (dx >= 0.f ? 4u : 0u) | (dy >= 0.f ? 2u : 0u) | (dz >= 0.f ? 1u : 0u);
    const uint32_t octinv4 = octinv * 0x01010101u;

    // cull bound on t: for the key mode a candidate with MT parameter t has key >= t - slack
    const float epsRay = 4e-6f * (fabsf(o.x) + fabsf(o.y) + fabsf(o.z));
    const float slack = (MODE == TRACE_CLOSEST_KEY) ? S.keySlack + epsRay : 0.f;
    float tcull = (MODE == TRACE_ANY) ? tmaxAny : INFINITY;
    float bestKey = INFINITY;

    uint2 stack[EL_STACK];
    int sp = 0;
    uint2 ngroup = make_uint2(0u, 0x80000000u);     // root: "child" bit 31 of a virtual parent with childBase 0
    uint2 tgroup = make_uint2(0u, 0u);

    for (;;) {
        if (ngroup.y > 0x00ffffffu) {
            // pop the nearest pending child of the current node group (root: virtual parent, imask 0 -> index 0)
            const uint32_t imask = ngroup.y;
            const uint32_t bit = 31u - __clz(ngroup.y);
            ngroup.y &= ~(1u << bit);
            if (ngroup.y > 0x00ffffffu) { stack[sp++] = ngroup; }
            const uint32_t slot = (bit - 24u) ^ octinv;
            const uint32_t rank = __popc(imask & ~(0xffffffffu << slot));
            const uint32_t nodeIndex = ngroup.x + rank;
            const float4* np = S.nodes + (size_t)nodeIndex * 5;
            const float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
            if (COUNT) cnt->nodes++;

            const uint32_t eim = __float_as_uint(n0.w);
            const float ax = __uint_as_float((eim & 0xffu) << 23) * idx;
            const float ay = __uint_as_float(((eim >> 8) & 0xffu) << 23) * idy;
            const float az = __uint_as_float(((eim >> 16) & 0xffu) << 23) * idz;
            const float ox = (n0.x - o.x) * idx, oy = (n0.y - o.y) * idy, oz = (n0.z - o.z) * idz;

            ngroup.x = __float_as_uint(n1.x);
            tgroup.x = __float_as_uint(n1.y);
            tgroup.y = 0;
            uint32_t hitmask = 0;
#pragma unroll
            for (int half = 0; half < 2; half++) {
                const uint32_t meta4 = __float_as_uint(half ? n1.w : n1.z);
                const uint32_t isInner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
                const uint32_t innerMask4 = (isInner4 >> 4) * 0xffu;
                const uint32_t bitIndex4 = (meta4 ^ (octinv4 & innerMask4)) & 0x1f1f1f1fu;
                const uint32_t childBits4 = (meta4 >> 5) & 0x07070707u;
                const uint32_t qlox = __float_as_uint(half ? n2.y : n2.x), qloy = __float_as_uint(half ? n2.w : n2.z);
                const uint32_t qloz = __float_as_uint(half ? n3.y : n3.x), qhix = __float_as_uint(half ? n3.w : n3.z);
                const uint32_t qhiy = __float_as_uint(half ? n4.y : n4.x), qhiz = __float_as_uint(half ? n4.w : n4.z);
                const uint32_t xmin = dx < 0.f ? qhix : qlox, xmax = dx < 0.f ? qlox : qhix;
                const uint32_t ymin = dy < 0.f ? qhiy : qloy, ymax = dy < 0.f ? qloy : qhiy;
                const uint32_t zmin = dz < 0.f ? qhiz : qloz, zmax = dz < 0.f ? qloz : qhiz;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const float tminx = fmaf((float)extractByte(xmin, j), ax, ox), tmaxx = fmaf((float)extractByte(xmax, j), ax, ox);
                    const float tminy = fmaf((float)extractByte(ymin, j), ay, oy), tmaxy = fmaf((float)extractByte(ymax, j), ay, oy);
                    const float tminz = fmaf((float)extractByte(zmin, j), az, oz), tmaxz = fmaf((float)extractByte(zmax, j), az, oz);
                    const float cmin = fmaxf(fmaxf(tminx, tminy), fmaxf(tminz, 0.f));
                    const float cmax = fminf(fminf(tmaxx, tmaxy), fminf(tmaxz, tcull));
                    // relative padding makes the test robust against rounding in the fused slab distances
                    if (cmin * 0.9999995f <= cmax * 1.0000005f) {
                        const uint32_t cb = extractByte(childBits4, j), bi = extractByte(bitIndex4, j);
                        hitmask |= cb << bi;
                    }
                }
            }
            ngroup.y = (hitmask & 0xff000000u) | (eim >> 24);
            tgroup.y = hitmask & 0x00ffffffu;
        } else {
            tgroup = ngroup;
            ngroup = make_uint2(0u, 0u);
        }

        while (tgroup.y != 0) {
            const uint32_t ti = 31u - __clz(tgroup.y);
            tgroup.y &= ~(1u << ti);
            const float4* tp = S.slots + (size_t)(tgroup.x + ti) * 3;
            const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
            if (COUNT) cnt->tris++;
            float t, u, v;
            if (mollerTrumbore(ray, f3(a.x, a.y, a.z), f3(a.w, b.x, b.y), f3(b.z, b.w, c.x), t, u, v)) {
                const int tri = __float_as_int(c.y);
                if (MODE == TRACE_ANY) {
                    if (t < tmaxAny) { best.tri = tri; best.t = t; best.u = u; best.v = v; best.key = t; return true; }
                } else if (MODE == TRACE_CLOSEST_T) {
                    if (best.tri < 0 || t < best.t || (t == best.t && tri < best.tri)) {
                        best.tri = tri; best.t = t; best.u = u; best.v = v; best.key = t; tcull = t;
                    }
                } else {
                    if (t <= tcull) {
                        TriGeom g = loadTriGeom(S.shadeTris, tri);
                        F3 sn;
                        const float key = hitKey(ray, hitPosition(ray, g, t, u, v, sn));
                        if (best.tri < 0 || key < bestKey || (key == bestKey && tri < best.tri)) {
                            best.tri = tri; best.t = t; best.u = u; best.v = v; best.key = key; bestKey = key;
                            tcull = fmaf(key, 1.000004f, slack);
                        }
                    }
                }
            }
        }

        if (ngroup.y <= 0x00ffffffu) {
            if (sp == 0) break;
            ngroup = stack[--sp];
        }
    }
    return best.tri >= 0;
}

} // namespace eleven
