/*
 * kernels_trace.cuh — the traversal kernels: thin Source/Sink adapters around trace_engine.cuh.
 *
 *   k_extend        closest hit for every queued path (throwRay, S/kernel.cu:152; BVH::transverse, S/BVH.hpp:120)
 *   k_shadowEnv     environment NEE shadow rays: any valid hit occludes (S/kernel.cu:246-248)
 *   k_shadowLight   point-light NEE shadow rays (S/kernel.cu:192-197)
 *   the last shadow stage of a bounce also performs the balance-heuristic MIS combination and the throughput update
 *   (shade, S/kernel.cu:351-357) in its Sink, so no separate combine pass reads the NEE records again
 *   k_traceBatch    the closest-hit contract on plain ray batches (eleven_trace_*)
 */
#pragma once
#include "kernels.cuh"
#include "trace_engine.cuh"

namespace eleven {

#ifndef EL_SHADOW_MIN_CTAS
#define EL_SHADOW_MIN_CTAS 9      /* 56 registers: shadow kernels 5.93 -> 5.73 ms per 16-spp step; 10 CTAs: 6.10 (profiles/r2_variants_session7.json) */
#endif
#ifndef EL_SHADOW_PREFETCH
#define EL_SHADOW_PREFETCH 1      /* shadow kernels 5.80 -> 5.70 ms per 16-spp step (profiles/r2_variants_session6.json) */
#endif
#ifndef EL_FAST_TRI
#define EL_FAST_TRI 1            /* fast-math configuration: Moeller-Trumbore with contracted multiply-adds and MUFU reciprocal (mollerTrumboreFast) */
#endif
#ifndef EL_EXTEND_MIN_CTAS
#define EL_EXTEND_MIN_CTAS 8      /* 64 registers, 8 CTAs per SM: k_extend 16.69 -> 16.06 ms per 16-spp step since the ray's sign tests come from the octant bits
                                   * (3 registers fewer); 7 CTAs x 72 registers was round 1's point, 6 CTAs: 17.75 ms (profiles/r2_variants_session6.json) */
#endif

// ---- extension rays -----------------------------------------------------------------------------------------------------
struct ExtendSource {
    const WaveState& W;
    __device__ __forceinline__ void load(uint32_t qi, LaneRay& lr) const {
        const uint32_t pid = W.qCur[qi];
        const float4 o = W.ray[2 * (size_t)pid], d = W.ray[2 * (size_t)pid + 1];
        lr.ray.o = f3(o.x, o.y, o.z); lr.ray.d = f3(d.x, d.y, d.z);
        lr.tmaxAny = INFINITY; lr.tag = qi;                              // the sink needs the queue position (hitBucket)
    }
};
struct ExtendSink {
    const WaveState& W; const DevScene& S;
    __device__ __forceinline__ void done(const LaneRay& lr, const HitRec& h) const {
        const uint32_t pid = W.qCur[lr.tag];
        W.hit[pid] = make_float4(__int_as_float(h.tri), h.t, h.u, h.v);
        // shading bucket (material, escaped rays last) at the queue position: k_classify then streams qCur + hitBucket
        // instead of chasing qCur -> hit -> triMaterial (3 dependent loads per entry: 4 % of the frame in round-1 ncu)
        W.hitBucket[lr.tag] = h.tri < 0 ? (uint8_t)EL_MISS_BUCKET : (uint8_t)min((uint32_t)__ldg(S.triMaterial + h.tri), (uint32_t)(EL_MISS_BUCKET - 1));
    }
};
template <int MODE, bool COUNT, bool FM>
__global__ void __launch_bounds__(128, EL_EXTEND_MIN_CTAS) k_extend(const __grid_constant__ WaveState W, const __grid_constant__ DevScene S) {
    TraceCounters tc; tc.nodes = 0; tc.tris = 0; tc.keys = 0;
    ExtendSource src{W}; ExtendSink sink{W, S};
    traceQueue<MODE, COUNT, false, FM && EL_FAST_TRI>(S, W.cnt[CNT_CUR], &W.cnt[CNT_WORK_TRACE], src, sink, tc);
    if (COUNT) {
        atomicAdd(&W.stats[ST_NODES], (unsigned long long)tc.nodes); atomicAdd(&W.stats[ST_TRIS], (unsigned long long)tc.tris); atomicAdd(&W.stats[ST_KEYS], (unsigned long long)tc.keys);
        atomicAdd(&W.stats[ST_NODES_EXT], (unsigned long long)tc.nodes); atomicAdd(&W.stats[ST_TRIS_EXT], (unsigned long long)tc.tris);
    }
}

// p / sum with the bits of the IEEE division.  0 / sum is +0 for any sum > 0, and that case is branched around: the division
// takes its out-of-line slow path for a zero numerator (ncu source page of k_shadowEnv: 8-9 % of the kernel's warp instructions
// in that subroutine at 2 active threads - p_p is 0 in every scene without point lights, p_e whenever the shadow ray is
// occluded).  `asm volatile` because the compiler otherwise computes the quotient speculatively and selects afterwards.
__device__ __forceinline__ float misWeight(float p, float sum) {
    float r;
    if (p == 0.f && sum > 0.f) r = 0.f;
    else asm volatile("div.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(p), "f"(sum));
    return r;
}

// ---- MIS combination (shade, S/kernel.cu:351-357) --------------------------------------------------------------------------
// hdriPdf is DEFINED as 0 when the environment shadow ray is occluded (the reference leaves it uninitialised).
__device__ __forceinline__ void misCombine(const WaveState& W, uint32_t pid, bool lights, bool envOccluded, bool lightOccluded) {
    const float4 ed = neeRec(W, pid, NEE_ENV_DIR), ec = neeRec(W, pid, NEE_ENV_C);
    const float4 bc = W.sceneHasEmission ? neeRec(W, pid, NEE_BRDF_C) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float pB = ec.w;
    const float pE = envOccluded ? 0.f : ed.w;
    const F3 CE = envOccluded ? f3(0.f) : f3(ec.x, ec.y, ec.z);
    float pP = 0.f; F3 CP = f3(0.f);
    if (lights) {
        const float4 lc = neeRec(W, pid, NEE_LIGHT_C);
        pP = lc.w;
        if (!lightOccluded) CP = f3(lc.x, lc.y, lc.z);
    }
    const float sum = pE + pP + pB;
    const float w1 = misWeight(pE, sum), w2 = misWeight(pP, sum), w3 = misWeight(pB, sum);
    const float4 thr4 = W.tr[2 * (size_t)pid];
    const F3 thr = f3(thr4.x, thr4.y, thr4.z);
    const F3 mix = f3(w1 * CE.x + w2 * CP.x + w3 * bc.x, w1 * CE.y + w2 * CP.y + w3 * bc.y, w1 * CE.z + w2 * CP.z + w3 * bc.z);
    float4 r = W.tr[2 * (size_t)pid + 1];
    r.x += thr.x * mix.x; r.y += thr.y * mix.y; r.z += thr.z * mix.z;
    W.tr[2 * (size_t)pid + 1] = r;
    const float4 tm = neeRec(W, pid, NEE_THR_MUL);
    W.tr[2 * (size_t)pid] = make_float4(thr.x * tm.x, thr.y * tm.y, thr.z * tm.z, 0.f);
}

// ---- environment shadow rays --------------------------------------------------------------------------------------------------
struct ShadowEnvSource {
    const WaveState& W;
    __device__ __forceinline__ void load(uint32_t qi, LaneRay& lr) const {
        const uint32_t pid = W.qNee[qi];
        const float4 p = neeRec(W, pid, NEE_POS), e = neeRec(W, pid, NEE_ENV_DIR);
        lr.ray.o = f3(p.x, p.y, p.z); lr.ray.d = f3(e.x, e.y, e.z);       // Ray(point + newDir*0.001, newDir), S/kernel.cu:246: built by k_shade
        lr.tmaxAny = INFINITY; lr.tag = pid;
#if EL_SHADOW_PREFETCH
        // what the sink's MIS combination will read when this ray is done (second sector of the NEE line, throughput / radiance): into
        // L2 now, while the ray is traced (ncu: long_scoreboard 4.7-6.3 warps per issue cycle, a quarter of the stall samples in the sink)
        asm volatile("prefetch.global.L2 [%0];" :: "l"(&neeRec(W, pid, NEE_ENV_C)));
        asm volatile("prefetch.global.L2 [%0];" :: "l"(&W.tr[2 * (size_t)pid]));
#endif
    }
};
template <bool LIGHTS>
struct ShadowEnvSink {
    const WaveState& W;
    __device__ __forceinline__ void done(const LaneRay& lr, const HitRec& h) const {
        const bool occluded = h.tri >= 0;
        if (LIGHTS) { if (occluded) { float4 e = neeRec(W, lr.tag, NEE_ENV_DIR); e.w = -1.f; neeRec(W, lr.tag, NEE_ENV_DIR) = e; } }   // p_e < 0 marks "occluded" for the light stage
        else misCombine(W, lr.tag, false, occluded, false);
    }
};
template <bool LIGHTS, bool COUNT, bool FM>
__global__ void __launch_bounds__(128, EL_SHADOW_MIN_CTAS) k_shadowEnv(const __grid_constant__ WaveState W, const __grid_constant__ DevScene S) {
    TraceCounters tc; tc.nodes = 0; tc.tris = 0; tc.keys = 0;
    ShadowEnvSource src{W}; ShadowEnvSink<LIGHTS> sink{W};
    traceQueue<TRACE_ANY, COUNT, false, FM && EL_FAST_TRI>(S, W.cnt[CNT_NEE], &W.cnt[CNT_WORK_CONNECT], src, sink, tc);
    if (COUNT) { atomicAdd(&W.stats[ST_NODES], (unsigned long long)tc.nodes); atomicAdd(&W.stats[ST_TRIS], (unsigned long long)tc.tris); atomicAdd(&W.stats[ST_KEYS], (unsigned long long)tc.keys); }
}

// ---- point-light shadow rays -------------------------------------------------------------------------------------------------
template <int HITMODE>
struct ShadowLightSource {
    const WaveState& W;
    __device__ __forceinline__ void load(uint32_t qi, LaneRay& lr) const {
        const uint32_t pid = W.qNee[qi];
        const float4 p = neeRec(W, pid, NEE_HITPOS), l = neeRec(W, pid, NEE_LIGHT_DIR);
        const F3 w = f3(l.x, l.y, l.z);
        lr.ray = makeRay(ex::madd(f3(p.x, p.y, p.z), w, 0.001f), w);      // S/kernel.cu:192
        lr.tmaxAny = HITMODE == ELEVEN_HIT_KEY ? INFINITY : l.w - 0.001f;
        lr.tag = pid;
    }
};
template <int HITMODE>
struct ShadowLightSink {
    const WaveState& W; const DevScene& S;
    __device__ __forceinline__ void done(const LaneRay& lr, const HitRec& h) const {
        const uint32_t pid = lr.tag;
        bool occluded = h.tri >= 0;
        if (HITMODE == ELEVEN_HIT_KEY && occluded) {
            // the reference takes the CLOSEST hit and compares |hit.position - point| with the light distance (S/kernel.cu:193-197)
            const float4 p = neeRec(W, pid, NEE_HITPOS);
            const TriGeom g = loadTriGeom(S.shadeTris, h.tri);
            F3 sn;
            const F3 hp = hitPosition(lr.ray, g, h.t, h.u, h.v, sn);
            occluded = length(hp - f3(p.x, p.y, p.z)) < neeRec(W, pid, NEE_LIGHT_DIR).w;
        }
        const bool envOccluded = neeRec(W, pid, NEE_ENV_DIR).w < 0.f;
        misCombine(W, pid, true, envOccluded, occluded);
    }
};
template <int HITMODE, bool COUNT, bool FM>
__global__ void __launch_bounds__(128, (HITMODE == ELEVEN_HIT_KEY ? 7 : EL_SHADOW_MIN_CTAS)) k_shadowLight(   // closest-hit (KEY) flavour: 72 registers
const __grid_constant__ WaveState W, const __grid_constant__ DevScene S) {
    TraceCounters tc; tc.nodes = 0; tc.tris = 0; tc.keys = 0;
    ShadowLightSource<HITMODE> src{W}; ShadowLightSink<HITMODE> sink{W, S};
    traceQueue<(HITMODE == ELEVEN_HIT_KEY ? TRACE_CLOSEST_KEY : TRACE_ANY), COUNT, false, FM && EL_FAST_TRI>(S, W.cnt[CNT_NEE], &W.cnt[CNT_WORK_LIGHT], src, sink, tc);
    if (COUNT) { atomicAdd(&W.stats[ST_NODES], (unsigned long long)tc.nodes); atomicAdd(&W.stats[ST_TRIS], (unsigned long long)tc.tris); atomicAdd(&W.stats[ST_KEYS], (unsigned long long)tc.keys); }
}

// ---- plain ray batches (test hook / bench) -----------------------------------------------------------------------------------------
struct BatchSource {
    const float* rays;
    __device__ __forceinline__ void load(uint32_t qi, LaneRay& lr) const {
        const float* r = rays + 6 * (size_t)qi;
        lr.ray = makeRay(f3(r[0], r[1], r[2]), f3(r[3], r[4], r[5]));    // Ray's constructor normalises (S/Ray.hpp:14-18)
        lr.tmaxAny = INFINITY; lr.tag = qi;
    }
};
struct BatchSink {
    ElevenHit* hits;
    __device__ __forceinline__ void done(const LaneRay& lr, const HitRec& h) const {
        ElevenHit o; o.tri = h.tri; o.t = h.t; o.u = h.u; o.v = h.v; o.key = h.key;
        hits[lr.tag] = o;
    }
};
template <int MODE, bool COUNT>
__global__ void __launch_bounds__(128) k_traceBatch(const float* __restrict__ rays, uint32_t n, ElevenHit* __restrict__ hits,
                                                  const __grid_constant__ DevScene S, uint32_t* workCounter, unsigned long long* stats) {
    TraceCounters tc; tc.nodes = 0; tc.tris = 0; tc.keys = 0;
    BatchSource src{rays}; BatchSink sink{hits};
    traceQueue<MODE, COUNT, true, false>(S, n, workCounter, src, sink, tc);      // the closest-hit contract: always the exact test
    if (COUNT && stats) { atomicAdd(&stats[ST_NODES], (unsigned long long)tc.nodes); atomicAdd(&stats[ST_TRIS], (unsigned long long)tc.tris); atomicAdd(&stats[ST_KEYS], (unsigned long long)tc.keys); }
}

} // namespace eleven
