/*
 * trace_engine.cuh — the persistent, warp-synchronous ray queue engine used by every traversal kernel.
 *
 * Why (ncu on the first version, profiles/r1_extend_v1.txt): the closest-hit kernel is ISSUE-bound, not memory-bound
 * (issue slots 70-73 % busy, DRAM 1 %, L2 2-5 %), and on incoherent bounce rays only 4.7 of 32 lanes were active per
 * issued instruction: lanes idled (a) inside per-lane `while (triangles)` loops whose trip count is the warp maximum,
 * and (b) after their ray finished, until the slowest ray of the 32-ray batch finished.  This engine removes both:
 *
 *   - ONE unit of work per lane per iteration: every active lane either intersects one node (8 child boxes) or one
 *     triangle; the two phases are warp-uniformly skipped when no lane needs them (__any_sync), so a lane is never
 *     parked in somebody else's inner loop;
 *   - dynamic ray fetch: a lane whose ray finished writes its result immediately and goes idle; as soon as REFILL lanes
 *     are idle the warp grabs that many rays from the global queue with one atomicAdd (warp-level work fetch) and the
 *     idle lanes start new rays while the others continue theirs;
 *   - child-box bytes are converted to float with PRMT + FADD (full-rate pipes) instead of I2F (quarter-rate
 *     conversion pipe): 48 conversions per node made I2F the busiest pipe of the old kernel.
 *
 * Source supplies rays by queue index, Sink consumes finished rays; both are small functors so that the same engine
 * serves the extension rays, the shadow rays and the plain ray batches of the test hook.
 *
 * Round 2: finished lanes keep their result and the sinks run together at the next queue fetch (EL_DEFER_SINK: the shadow
 * kernels' sink is a 90-instruction MIS combination that used to run with 1.7-3.3 lanes active); the render kernels of the
 * fast-math configuration intersect triangles with mollerTrumboreFast (FAST_TRI); the ray's sign tests come from its octant
 * bits (64 registers: 8 CTAs per SM for k_extend, 9 for the shadow kernels).
 */
#pragma once
#include "traverse.cuh"

namespace eleven {

#ifndef EL_TRIS_PER_ITER
#define EL_TRIS_PER_ITER 2     /* triangle tests per lane per loop iteration */
#endif
/* Round 2 tried a per-warp shared-memory RING of set-up rays (one atomicAdd + one coalesced queue read + the set-up arithmetic for 32
 * rays with all lanes busy; an idle lane then starts a ray with 12 LDS) so that the refill could run at 2-8 idle lanes instead of 12.
 * Bit-identical, and SLOWER on one box back to back: k_extend 17.54 vs 16.70 ms, shadow 6.21 vs 5.80 (refill at 2 / 4 / 6 / 8 idle lanes:
 * 17.69 / 17.54 / 17.56 / 17.62; a 32-slot ring: 17.40) — profiles/r2_variants_session5.json.  The refill threshold is NOT where the
 * lanes are lost, and 15 KB of shared memory per CTA come out of the L1 the node fetches live in.  Removed again. */
#ifndef EL_DEFER_SINK
#define EL_DEFER_SINK 1
#endif
#ifndef EL_REFILL
#define EL_REFILL 12           /* idle lanes that trigger a queue fetch (4 / 8 / 12 / 16 measured on 16-sample waves: k_extend 19.9 / 18.85 / 18.4 / 18.5 ms) */
#endif

struct LaneRay {
    Ray ray;
    float tmaxAny;             /* TRACE_ANY: accept hits with t < tmaxAny */
    uint32_t tag;              /* opaque to the engine (path id / ray index) */
};

// 2^15 + (byte j of q4) as a float with ONE instruction and no conversion pipe: PRMT drops the byte into mantissa bits 8..15
// of 0x47000000 (= 2^15, ulp 2^-8).  The 2^15 is folded into the FMA's addend once per node and axis (oN = o - 2^15 a), so a
// near plane costs PRMT + FFMA.  (Round 1 used 0x4B000000|b = 2^23 + b and an FADD per byte: folding THAT offset loses half a
// cell, because 2^23 a is 2^15 cells away; 2^15 a is 128 node widths away and its rounding, <= 2^-9 cell, is covered by moving
// oN towards the ray origin by 2^-23 |oN|, i.e. the near planes only ever move outwards.)
// `magic` holds 0x47000000 but arrives as scene DATA (DevScene::byteMagic) so that it lives in a register: PRMT takes one
// immediate, and with both the selector and the constant known, ptxas kept the constant as the immediate and re-materialised
// the selector with a MOV in front of every PRMT (24 extra ALU-pipe instructions per node on the busiest pipe).
__device__ __forceinline__ float byteMagic15(uint32_t q4, uint32_t magic, uint32_t selector) {
    return __uint_as_float(__byte_perm(q4, magic, selector));
}

__device__ __forceinline__ float byteI2F(uint32_t q4, int j) { return (float)((q4 >> (8 * j)) & 0xffu); }   // I2F.U8 on the conversion pipe

#ifndef EL_SAT
#define EL_SAT 0               /* 1: near planes through FFMA.SAT on a per-ray power-of-two time scale (see traceQueue).  Exact, 8 FMNMX fewer per
                                * node, but measured SLOWER (k_extend 18.04 vs 17.84 ms, shadow kernels 72 instead of 64 registers: 7.30 vs 6.88 ms,
                                * profiles/r1_variants_session3.json): the ray set-up and the scale register cost more than the clamps save */
#endif
// a * b + c clamped to [0, 1] in the FMA pipe (FFMA.SAT): the `max(tmin, 0)` of the slab test for free
__device__ __forceinline__ float fmaSat(float a, float b, float c) {
#if EL_SAT
    float r; asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r;
#else
    return fmaxf(fmaf(a, b, c), 0.f);
#endif
}

// Child hit mask -> traversal masks by table (shared memory, 3 KB per CTA, filled by the CTA itself):
//   expand[h]  bit 3*i + k (k = 0..2) set iff bit i of h is set: AND with Node8::triMask = the triangles of the hit leaves;
//   perm[o][h] bit (i ^ o) set iff bit i of h is set: the hit internal children in front-to-back order for ray octant o
//              (highest bit = nearest child, as the pop takes the highest bit).
// Round 1 built both masks child by child from an (offset, count) byte: 2 extracts, a shift and an OR per child on the ALU
// pipe, which the node test saturates (ncu: ALU pipe 72 % at 78 % issue; SASS: 150 of the 272 node-phase instructions on ALU).
struct TraceLut { uint32_t expand[256]; uint8_t perm[8][256]; };


__device__ __forceinline__ void traceLutInit(TraceLut& L) {
    for (uint32_t h = threadIdx.x; h < 256u; h += blockDim.x) {
        uint32_t e = 0;
        for (uint32_t i = 0; i < 8u; i++) if ((h >> i) & 1u) e |= 7u << (3u * i);
        L.expand[h] = e;
        for (uint32_t o = 0; o < 8u; o++) {
            uint32_t p = 0;
            for (uint32_t i = 0; i < 8u; i++) if ((h >> i) & 1u) p |= 1u << (i ^ o);
            L.perm[o][h] = (uint8_t)p;
        }
    }
    __syncthreads();
}

template <int MODE, bool COUNT, bool NEED_KEY, bool FAST_TRI, class Source, class Sink>
__device__ __forceinline__ void traceQueue(const DevScene& S, uint32_t n, uint32_t* workCounter, Source& src, Sink& sink, TraceCounters& tc) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t FULL = 0xffffffffu;
    const uint32_t ltMask = (1u << lane) - 1u;
    const uint32_t magic = S.byteMagic;
    __shared__ TraceLut lut;
    traceLutInit(lut);
    // Per-ray time scale 2^-k with 2^k beyond the far end of the scene (root box) as seen from the ray origin: all slab
    // distances of the node test live in [0, 1] then, and FFMA.SAT clamps the near planes at 0 for free.  A power of two,
    // folded into 1/dir: every product and sum is the unscaled one times 2^-k exactly, so the test decides as before (the
    // clamp at 1 can only ADD children, and only beyond the scene).
    float4 root0 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (EL_SAT && S.nodeCount) root0 = __ldg(S.nodes);

    bool active = false, exhausted = (S.nodeCount == 0 && false);
    LaneRay lr; lr.tag = 0; lr.tmaxAny = INFINITY; lr.ray.o = f3(0.f); lr.ray.d = f3(0.f, 0.f, 1.f);
    float idx = 0.f, idy = 0.f, idz = 0.f, tscale = 1.f;
    float dx = 1.f, dy = 1.f, dz = 1.f;
    uint32_t octinv = 0, tvalid = 0, tpostValid = 0;
    float slack = 0.f, epsRay = 0.f, tcull = INFINITY, bestLo = INFINITY, bestHi = INFINITY;   // [bestLo, bestHi] brackets the best key
    bool bestExact = false;
    HitRec best; best.tri = -1; best.t = best.u = best.v = best.key = 0.f;
    uint2 ngroup = make_uint2(0u, 0u), tgroup = make_uint2(0u, 0u), tpost = make_uint2(0u, 0u);   // tpost: one postponed triangle group
    uint2 stack[EL_STACK];
    int sp = 0;
#if EL_DEFER_SINK
    // A finished lane keeps its result in its (now idle) registers and the sinks run TOGETHER at the next queue fetch, i.e. with EL_REFILL or more
    // lanes at once.  Called the moment a lane finished, the shadow kernels' sink (the MIS combination: three dependent record loads, ~90
    // instructions) ran 6.9 M times per launch with 1.7-3.3 of 32 lanes active: 14 % of the kernel's warp instructions and 16 % of its stall samples
    // (ncu source page, profiles/r2_final_trace_ncu_full.txt).
    bool pending = false;
    auto flush = [&]() {
        if (pending) {
            if (MODE == TRACE_CLOSEST_KEY && NEED_KEY && best.tri >= 0 && !bestExact) {
                const TriGeom g = loadTriGeom(S.shadeTris, best.tri);
                F3 sn;
                best.key = hitKey(lr.ray, hitPosition(lr.ray, g, best.t, best.u, best.v, sn));
            }
            sink.done(lr, best); pending = false;
        }
    };
#endif

    for (;;) {
        // ---- dynamic fetch ---------------------------------------------------------------------------------------
        const uint32_t idle = __ballot_sync(FULL, !active);
        if (idle != 0u && !exhausted && (__popc(idle) >= EL_REFILL || idle == FULL)) {
#if EL_DEFER_SINK
            flush();
#endif
            const uint32_t cnt = __popc(idle);
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(workCounter, cnt);
            base = __shfl_sync(FULL, base, 0);
            if (base + cnt >= n) exhausted = true;
            if (!active) {
                const uint32_t qi = base + __popc(idle & ltMask);
                if (qi < n) {
                    src.load(qi, lr);
                    const F3 d = lr.ray.d;
                    dx = fabsf(d.x) > 1e-20f ? d.x : copysignf(1e-20f, d.x);
                    dy = fabsf(d.y) > 1e-20f ? d.y : copysignf(1e-20f, d.y);
                    dz = fabsf(d.z) > 1e-20f ? d.z : copysignf(1e-20f, d.z);
                    const F3 o = lr.ray.o;
                    if (EL_SAT) {
                        const uint32_t re = __float_as_uint(root0.w);      // root box: [p, p + 255 * 2^e] per axis
                        const float far1 = (fabsf(o.x - root0.x) + fabsf(o.y - root0.y) + fabsf(o.z - root0.z)) +
                                           256.f * (__uint_as_float((re & 0xffu) << 23) + __uint_as_float(((re >> 8) & 0xffu) << 23) + __uint_as_float(((re >> 16) & 0xffu) << 23));
                        const uint32_t fe = min(max(__float_as_uint(far1) >> 23, 64u), 190u);   // biased exponent of the bound, kept far from under/overflow
                        tscale = __uint_as_float((253u - fe) << 23);                              // 2^-(e+1) < 1 / far1
                    }
                    idx = (1.0f / dx) * tscale; idy = (1.0f / dy) * tscale; idz = (1.0f / dz) * tscale;
                    octinv = (dx >= 0.f ? 4u : 0u) | (dy >= 0.f ? 2u : 0u) | (dz >= 0.f ? 1u : 0u);
                    epsRay = 4e-6f * (fabsf(o.x) + fabsf(o.y) + fabsf(o.z));
                    slack = (MODE == TRACE_CLOSEST_KEY) ? S.keySlack + epsRay : 0.f;
                    tcull = (MODE == TRACE_ANY) ? lr.tmaxAny : INFINITY;
                    bestLo = bestHi = INFINITY; bestExact = false;
                    best.tri = -1; best.t = best.u = best.v = best.key = 0.f;
                    ngroup = make_uint2(0u, S.nodeCount ? 0x80000000u : 0u);
                    tgroup = make_uint2(0u, 0u); tpost = make_uint2(0u, 0u); tvalid = tpostValid = 0u;
                    sp = 0;
                    active = true;
                }
            }
        }
        if (__ballot_sync(FULL, active) == 0u) {
#if EL_DEFER_SINK
            flush();
#endif
            if (exhausted) break;
            continue;
        }

        // ---- bookkeeping: promote the postponed triangle group, pop a node group, or finish -------------------------------
        if (active) {
            if (tgroup.y == 0u && tpost.y != 0u) { tgroup = tpost; tvalid = tpostValid; tpost.y = 0u; }
            if (ngroup.y <= 0x00ffffffu && sp > 0) ngroup = stack[--sp];
            if (ngroup.y <= 0x00ffffffu && tgroup.y == 0u) {
#if EL_DEFER_SINK
                pending = true; active = false;
#else
                if (MODE == TRACE_CLOSEST_KEY && NEED_KEY && best.tri >= 0 && !bestExact) {
                    const TriGeom g = loadTriGeom(S.shadeTris, best.tri);
                    F3 sn;
                    best.key = hitKey(lr.ray, hitPosition(lr.ray, g, best.t, best.u, best.v, sn));
                }
                sink.done(lr, best); active = false;
#endif
            }
        }

        // ---- node phase: one node (8 quantised child boxes) ------------------------------------------------------------
        // A lane does node work AND triangle work in the same iteration whenever it has both: the triangles found by this
        // node are postponed (tpost) while an older group is still being tested, so neither phase idles the lane.
        const bool doNode = active && ngroup.y > 0x00ffffffu && tpost.y == 0u;
        if (__any_sync(FULL, doNode)) {
            if (doNode) {
                const uint32_t imask = ngroup.y;
                const uint32_t bit = 31u - __clz(ngroup.y);
                ngroup.y &= ~(1u << bit);
                if (ngroup.y > 0x00ffffffu) stack[sp++] = ngroup;
                const uint32_t slot = (bit - 24u) ^ octinv;
                const uint32_t rank = __popc(imask & ~(0xffffffffu << slot));
                const uint32_t nodeIndex = ngroup.x + rank;
                const float4* np = S.nodes + (size_t)nodeIndex * 5;
                const float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
                // KEY mode: a triangle below this node can only win if t - shift - e <= bestHi, and its shift is bounded
                // by this node's subtree maximum (Node8::slack): cull the children with that LOCAL slack instead of the scene maximum
                const float tcullNode = ((MODE == TRACE_CLOSEST_KEY) ? (bestHi + n1.w + epsRay) * 1.00001f : tcull) * tscale;
                if (COUNT) tc.nodes++;
                const F3 o = lr.ray.o;
                const uint32_t eim = __float_as_uint(n0.w);
                const float ax = __uint_as_float((eim & 0xffu) << 23) * idx;
                const float ay = __uint_as_float(((eim >> 8) & 0xffu) << 23) * idy;
                const float az = __uint_as_float(((eim >> 16) & 0xffu) << 23) * idz;
                const float ox = (n0.x - o.x) * idx, oy = (n0.y - o.y) * idy, oz = (n0.z - o.z) * idz;
                // addends of the near planes: o - 2^15 a, nudged towards the ray origin by its own rounding bound (see byteMagic15)
                const float oxN0 = fmaf(-32768.f, ax, ox), oyN0 = fmaf(-32768.f, ay, oy), ozN0 = fmaf(-32768.f, az, oz);
                const float oxN = fmaf(-fabsf(oxN0), 1.1920929e-7f, oxN0), oyN = fmaf(-fabsf(oyN0), 1.1920929e-7f, oyN0), ozN = fmaf(-fabsf(ozN0), 1.1920929e-7f, ozN0);
                uint32_t h8 = 0;                                              // bit i: child box i (= octant slot i) is hit
#pragma unroll
                for (int half = 0; half < 2; half++) {
                    const uint32_t qlox = __float_as_uint(half ? n2.y : n2.x), qloy = __float_as_uint(half ? n2.w : n2.z);
                    const uint32_t qloz = __float_as_uint(half ? n3.y : n3.x), qhix = __float_as_uint(half ? n3.w : n3.z);
                    const uint32_t qhiy = __float_as_uint(half ? n4.y : n4.x), qhiz = __float_as_uint(half ? n4.w : n4.z);
                    // octinv bit set = direction component >= 0 (near plane = low plane)
                    const bool xNeg = !(octinv & 4u), yNeg = !(octinv & 2u), zNeg = !(octinv & 1u);
                    const uint32_t xmin = xNeg ? qhix : qlox, xmax = xNeg ? qlox : qhix;
                    const uint32_t ymin = yNeg ? qhiy : qloy, ymax = yNeg ? qloy : qhiy;
                    const uint32_t zmin = zNeg ? qhiz : qloz, zmax = zNeg ? qloz : qhiz;
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const uint32_t sel = 0x7504u + ((uint32_t)j << 4);       // bytes: 0x00, q_j, 0x00, 0x47
                        // pipe balancing (ncu: ALU pipe 65 % busy, XU 6 %): the near planes are decoded with PRMT (ALU pipe, the
                        // offset folded into the FMA), the far planes with I2F (conversion unit), so neither pipe carries all 48
                        const float tminx = fmaSat(byteMagic15(xmin, magic, sel), ax, oxN), tmaxx = fmaf(byteI2F(xmax, j), ax, ox);
                        const float tminy = fmaSat(byteMagic15(ymin, magic, sel), ay, oyN), tmaxy = fmaf(byteI2F(ymax, j), ay, oy);
                        const float tminz = fmaSat(byteMagic15(zmin, magic, sel), az, ozN), tmaxz = fmaf(byteI2F(zmax, j), az, oz);
                        const float cmin = fmaxf(fmaxf(tminx, tminy), tminz);     // >= 0 already (FFMA.SAT)
                        const float cmax = fminf(fminf(tmaxx, tmaxy), fminf(tmaxz, tcullNode));
                        if (cmin <= cmax * 1.000001f) h8 |= 1u << (4 * half + j);  // relative slack for the rounding of the fused distances
                    }
                }
                // empty slots hold a zero box: whatever the test says about them, imask and triMask drop them
                const uint32_t im8 = eim >> 24, triMask = __float_as_uint(n1.z);
                ngroup.x = __float_as_uint(n1.x);
                ngroup.y = __byte_perm(im8, (uint32_t)lut.perm[octinv][h8 & im8], 0x4210);   // (hit internal children, front to back) << 24 | imask
                const uint32_t newTris = lut.expand[h8] & triMask;
                if (newTris) {
                    if (tgroup.y != 0u) { tpost = make_uint2(__float_as_uint(n1.y), newTris); tpostValid = triMask; }
                    else { tgroup = make_uint2(__float_as_uint(n1.y), newTris); tvalid = triMask; }
                }
            }
        }

        // ---- closest-hit candidate (t, u, v of an accepted Moeller-Trumbore test) --------------------------------------
        auto consider = [&](float t, float u, float v, int tri, float sT) {
            if (MODE == TRACE_CLOSEST_T) {
                if (best.tri < 0 || t < best.t || (t == best.t && tri < best.tri)) {
                    best.tri = tri; best.t = t; best.u = u; best.v = v; best.key = t; tcull = t;
                }
            } else {
                // Closest by the reference KEY = |hit.position - origin| (S/BVH.hpp:170), evaluated LAZILY: the key
                // of a candidate lies in [t - s - e, t + s + e] (s = this triangle's shadow-terminator shift bound
                // from the builder, e = rounding allowance), so the exact key (5 more 128-bit loads + ~120 flops on
                // a divergent path) is only computed when two candidates' brackets overlap.
                const float e = fmaf(t, 4e-6f, epsRay);
                const float cLo = t - sT - e, cHi = t + sT + e;
                if (best.tri < 0 || cHi < bestLo) {            // certainly closer than the current best
                    best.tri = tri; best.t = t; best.u = u; best.v = v; bestLo = cLo; bestHi = cHi; bestExact = false;
                    tcull = (bestHi + slack) * 1.00001f;       // a later triangle with t beyond this cannot have a smaller key
                } else if (cLo <= bestHi) {                      // brackets overlap: decide on exact keys
                    if (COUNT) tc.keys++;
                    if (!bestExact) {
                        const TriGeom gb = loadTriGeom(S.shadeTris, best.tri);
                        F3 snb;
                        best.key = hitKey(lr.ray, hitPosition(lr.ray, gb, best.t, best.u, best.v, snb));
                        bestLo = bestHi = best.key; bestExact = true;
                    }
                    const TriGeom g = loadTriGeom(S.shadeTris, tri);
                    F3 sn;
                    const float key = hitKey(lr.ray, hitPosition(lr.ray, g, t, u, v, sn));
                    if (key < best.key || (key == best.key && tri < best.tri)) {
                        best.tri = tri; best.t = t; best.u = u; best.v = v; best.key = key; bestLo = bestHi = key;
                    }
                    tcull = (bestHi + slack) * 1.00001f;
                }
            }
        };

        // ---- triangle phase, per lane: up to EL_TRIS_PER_ITER triangles (leaf nodes yield ~2.5 triangles per node visit, a
        //      triangle test costs about a third of a node test: this keeps the two phases balanced at the leaf level) ---------
#pragma unroll 1
        for (int round = 0; round < (MODE == TRACE_ANY ? 1 : EL_TRIS_PER_ITER); round++) {   // measured: 2 rounds -10 % on closest hit, +3 % on any hit
        if (active && tgroup.y == 0u && tpost.y != 0u) { tgroup = tpost; tvalid = tpostValid; tpost.y = 0u; }
        const bool doTri = active && tgroup.y != 0u;
        if (!__any_sync(FULL, doTri)) break;
        {
            if (doTri) {
                const uint32_t tb = 1u << (31u - __clz(tgroup.y));
                tgroup.y &= ~tb;
                const float4* tp = S.slots + (size_t)(tgroup.x + __popc(tvalid & (tb - 1u))) * 3;   // Node8::triMask: slots are packed in bit order
                const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
                if (COUNT) tc.tris++;
                float t, u, v;
                const bool triHit = FAST_TRI ? mollerTrumboreFast(lr.ray, f3(a.x, a.y, a.z), f3(a.w, b.x, b.y), f3(b.z, b.w, c.x), t, u, v)
                                             : mollerTrumbore(lr.ray, f3(a.x, a.y, a.z), f3(a.w, b.x, b.y), f3(b.z, b.w, c.x), t, u, v);
                if (triHit) {
                    const int tri = __float_as_int(c.y);
                    if (MODE == TRACE_ANY) {
                        if (t < lr.tmaxAny) {
                            best.tri = tri; best.t = t; best.u = u; best.v = v; best.key = t;
#if EL_DEFER_SINK
                            pending = true; active = false;
#else
                            sink.done(lr, best); active = false;
#endif
                        }
                    } else consider(t, u, v, tri, c.w);
                }
            }
        }
        }   // triangle rounds
    }
}

} // namespace eleven
