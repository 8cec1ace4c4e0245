/*
 * kernels.cuh — the wavefront path-tracing pipeline (device kernels).
 *
 * Replaces the reference's megakernel `renderingKernel` (S/kernel.cu:369-481: one thread = one pixel-sample,
 * 188 registers, 2.8 KB of local stack, one launch + one host sync per sample) by a queue-driven pipeline whose
 * stages are separate persistent kernels:
 *
 *   raygen  -> [ extend -> shade -> connect ] x maxBounces -> accumulate
 *
 *   raygen      camera ray with jitter + thin lens (calculateCameraRay, S/kernel.cu:260-337)
 *   extend      closest hit over the BVH8 (throwRay/BVH::transverse, S/kernel.cu:152, S/BVH.hpp:120)
 *   shade       hit attributes, material + texture fetch, DisneySample/Eval/Pdf, NEE set-up (generateHitData,
 *               calculateBounce, shade, hdriLight, pointLight: S/kernel.cu:54-119,175-258,339-367)
 *   connect     shadow rays + the balance-heuristic MIS combination (S/kernel.cu:246-248,192-197,351-357)
 *   accumulate  clamp, NaN rejection, film sums (S/kernel.cu:445-480)
 *
 * K = 2^logK samples of every pixel per wave (K = 1 with the reference's sequential per-pixel RNG), path id =
 * film index * K + k, so film updates need no atomics and the K samples of a pixel are summed in sample order.  Why K > 1:
 * the duration of a traversal kernel has a floor set by its slowest rays (~150-200 us on the benchmark scene, the late
 * bounces of a 2 M-path wave ran at that floor: profiles/r1_launches_final.csv), paid once per kernel whatever the size of
 * the wave.  Queues hold path ids;
 * their counts live on the device and every kernel is a persistent grid that fetches 32-ray batches per warp with
 * one atomicAdd ("warp-level work fetch"), so a whole wave is enqueued without a single host synchronisation.
 * Round 2: the connect (shadow) stage of bounce b runs on an auxiliary stream concurrently with extend + classify of bounce
 * b+1 (eleven_api.cu: eleven_render; k_advance phases 2 / 3 split the counter hand-over accordingly).
 */
#pragma once
#include "shading.cuh"

namespace eleven {

enum { CNT_CUR = 0, CNT_WORK_TRACE = 1, CNT_NEXT = 2, CNT_NEE = 3 /* (NEXT, NEE) = one aligned 64-bit word: k_shade reserves both with ONE atomic */, CNT_WORK_SHADE = 4, CNT_WORK_CONNECT = 5, CNT_WORK_LIGHT = 6, CNT_WORK_CLASSIFY = 7,
       CNT_BUCKET0 = 8, CNT_COUNT = 16 };
enum { EL_BUCKETS = 8, EL_MISS_BUCKET = 7 };   // shading queue buckets: materials 0..5, 6 = all further materials, 7 = escaped rays
enum { ST_RAYS_EXT = 0, ST_RAYS_ENV = 1, ST_RAYS_LIGHT = 2, ST_NODES = 3, ST_TRIS = 4, ST_KEYS = 5, ST_NODES_EXT = 6, ST_TRIS_EXT = 7, ST_COUNT = 8 };   // 3-5: all traversal kernels; 6-7: k_extend alone

struct WaveState {
    // per path (index = film index * K + k)
    float4* ray;                         // 2 x float4 = one 32-byte sector per path: (origin.xyz, depth bits), (normalised direction.xyz, 0);
                                         // depth = number of hit bounces so far ("i")
    float4* tr;                          // 2 x float4 = one sector per path: throughput ("reduction"), accumulated radiance ("light")
    float4* hit;                         // tri (as int bits), t, u, v
    uint8_t* hitBucket;                  // per QUEUE position of the current bounce: shading bucket of the hit (written by k_extend's sink)
    float4* aovN; float4* aovT; float4* aovB;
    Xorwow* rng;                         // per-pixel XORWOW state, persistent across samples (reference mode)
    // NEE records written by shade, consumed by connect
    float4* nee;                         // NEE_STRIDE x float4 = one 128-byte line per path (NEE_*): the records every hit writes come
                                         // first, so a scene without emission and point lights touches two full sectors
    uint32_t sceneHasEmission;           // any material with emission: NEE_BRDF_C is written / read
    // queues
    uint32_t* qCur; uint32_t* qNext; uint32_t* qNee;
    uint32_t* qBucket;                   // EL_BUCKETS x pathCapacity: the shading queue, sorted by material (k_classify)
    uint32_t* cnt;                       // CNT_*
    unsigned long long* stats;           // ST_*
    // film: per-pixel sums of the four device passes, ONE allocation (BEAUTY, NORMAL, TANGENT, BITANGENT back to back, so that a
    // single ncclReduce covers them); .w of every record = number of accepted samples of the pixel as a float (exact below 2^24).
    // Sum and count of a pixel travel in one 16-byte store, so a snapshot taken on another stream while a wave accumulates sees
    // every pixel either before or after its update, never a sum without its count (eleven_get_film during eleven_render).
    float4* filmBeauty; float4* filmNormal; float4* filmTangent; float4* filmBitangent;
    uint32_t* pathCount;
    uint32_t nPixels;
    uint32_t pathCapacity;               // nPixels * largest K: stride of the qBucket rows
};

// slots of a path's NEE line
enum { NEE_ENV_DIR = 0,                  // direction of the environment shadow ray (w_e as Ray's ctor re-normalises it), p_e   } sector 0: all the
       NEE_POS = 1,                      // its origin P + 0.001 w_e                                                               } shadow ray needs
       NEE_ENV_C = 2,                    // C_e.xyz, p_b        } sector 1: the rest of the MIS combination
       NEE_THR_MUL = 3,                  // f*cos/p_b (throughput update factor)
       NEE_BRDF_C = 4,                   // C_b.xyz (only with emission)
       NEE_LIGHT_DIR = 5,                // w_l.xyz, dist (only with point lights)
       NEE_LIGHT_C = 6,                  // C_p.xyz, p_p
       NEE_HITPOS = 7,                   // hit position P (only with point lights: origin of the light shadow ray)
       NEE_STRIDE = 8 };
__device__ __forceinline__ float4& neeRec(const WaveState& W, uint32_t pid, int k) { return W.nee[(size_t)pid * NEE_STRIDE + k]; }

struct RenderParams {
    uint32_t rngMode, envMode, hitMode, maxBounces, flags;
    uint32_t sampleIndex;                // global index of the first sample this wave renders (fast rng)
    uint32_t sampleStride;               // global index step between the K samples of a wave
    uint32_t logK;                       // this wave carries K = 2^logK samples per pixel
    uint32_t seedLo, seedHi;
    CamRot rot;
};

__device__ __forceinline__ uint32_t warpFetch(uint32_t* counter, uint32_t lane) {
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(counter, 32u);
    return __shfl_sync(0xffffffffu, base, 0);
}

// ---- XORWOW per-pixel seeding: curand_init(0, idx, 0) (S/kernel.cu:140) --------------------------------------
// seqMat[k] = step^(2^67 * 2^k) as 160 columns x 5 words (built on the host by repeated squaring over GF(2)).
__global__ void k_rngInit(Xorwow* __restrict__ rng, const uint32_t* __restrict__ seqMat, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t v[5];
    const uint32_t s0 = 0u ^ 0xaad26b49u, s1 = 0u ^ 0xf7dcefddu;          // seed 0 (curand_kernel.h:780-791)
    const uint32_t t0 = 1099087573u * s0, t1 = 2591861531u * s1;
    const uint32_t d = 6615241u + t1 + t0;
    v[0] = 123456789u + t0; v[1] = 362436069u ^ t0; v[2] = 521288629u + t1; v[3] = 88675123u ^ t1; v[4] = 5783321u + t0;
    for (uint32_t k = 0; (i >> k) != 0u; k++) {
        if (!((i >> k) & 1u)) continue;
        const uint32_t* M = seqMat + (size_t)k * 800;
        uint32_t r[5] = {0u, 0u, 0u, 0u, 0u};
        for (int w = 0; w < 5; w++) {
            uint32_t bitsLeft = v[w];
            while (bitsLeft) {
                const int b = __ffs(bitsLeft) - 1; bitsLeft &= bitsLeft - 1;
                const uint32_t* c = M + (w * 32 + b) * 5;
                r[0] ^= __ldg(c); r[1] ^= __ldg(c + 1); r[2] ^= __ldg(c + 2); r[3] ^= __ldg(c + 3); r[4] ^= __ldg(c + 4);
            }
        }
        for (int w = 0; w < 5; w++) v[w] = r[w];
    }
    Xorwow s; s.v0 = v[0]; s.v1 = v[1]; s.v2 = v[2]; s.v3 = v[3]; s.v4 = v[4]; s.d = d;
    rng[i] = s;
}

__global__ void k_filmReset(WaveState W) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= W.nPixels) return;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    W.filmBeauty[i] = z; W.filmNormal[i] = z; W.filmTangent[i] = z; W.filmBitangent[i] = z;
    W.pathCount[i] = 0u;
}

// fast-mode uniforms: 4 per call, keyed by (pixel, sample, dimension block)
__device__ __forceinline__ uint4 fastBits(const RenderParams& P, uint32_t pid, uint32_t block) {
    const uint32_t pixel = pid >> P.logK, k = pid & ((1u << P.logK) - 1u);
    return philox4x32(make_uint4(pixel, P.sampleIndex + k * P.sampleStride, block, 0x11e7e0u), make_uint2(P.seedLo, P.seedHi));
}

// ---- raygen ------------------------------------------------------------------------------------------------
template <bool FM>
__global__ void __launch_bounds__(256) k_raygen(WaveState W, const __grid_constant__ DevScene S, const __grid_constant__ RenderParams P) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;        // path id
    const uint32_t nPaths = W.nPixels << P.logK;
    if (i >= nPaths) return;
    const uint32_t pix = i >> P.logK;
    const uint32_t Wd = S.cam.xRes, Hd = S.cam.yRes;
    const int x = (int)(pix % Wd), y = (int)(Hd - 1u - pix / Wd);    // film index = W*(H-1-y)+x (S/kernel.cu:378)
    float r1, r2, r3, r4, r5;
    if (P.rngMode == ELEVEN_RNG_REFERENCE) {
        Xorwow s = W.rng[i];                                          // logK == 0 in this mode
        r1 = xorwowUniform(s); r2 = xorwowUniform(s); r3 = xorwowUniform(s); r4 = xorwowUniform(s); r5 = xorwowUniform(s);
        W.rng[i] = s;
    } else {
        const uint4 a = fastBits(P, i, 0u);
        r1 = u32ToUniform(a.x); r2 = u32ToUniform(a.y); r3 = u32ToUniform(a.z); r4 = u32ToUniform(a.w);
        r5 = S.cam.bokeh ? u32ToUniform(fastBits(P, i, 1u).x) : 0.5f;
    }
    const Ray ray = cameraRay<FM>(S.cam, P.rot, x, y, r1, r2, r3, r4, r5);
    W.ray[2 * (size_t)i] = make_float4(ray.o.x, ray.o.y, ray.o.z, __uint_as_float(0u));   // depth 0: the first-hit AOVs are written by k_shade at depth 0, k_accumulate reads them only where depth > 0
    W.ray[2 * (size_t)i + 1] = make_float4(ray.d.x, ray.d.y, ray.d.z, 0.f);
    W.tr[2 * (size_t)i] = make_float4(1.f, 1.f, 1.f, 0.f);
    W.tr[2 * (size_t)i + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
    W.qCur[i] = i;
    if (i == 0) {
        W.cnt[CNT_CUR] = nPaths; W.cnt[CNT_NEXT] = 0u; W.cnt[CNT_NEE] = 0u;
        W.cnt[CNT_WORK_TRACE] = 0u; W.cnt[CNT_WORK_SHADE] = 0u; W.cnt[CNT_WORK_CONNECT] = 0u; W.cnt[CNT_WORK_LIGHT] = 0u;
        W.cnt[CNT_WORK_CLASSIFY] = 0u;
        for (int b = 0; b < EL_BUCKETS; b++) W.cnt[CNT_BUCKET0 + b] = 0u;
    }
}


// ---- classify: bucket the traced paths by material (escaped rays last) for a coherent shading queue ------------------
// Two coalesced streams in (k_extend's sink left each path's bucket at its queue position), one scattered stream out.
// A CTA takes chunks of 2 048 entries: ranks inside the chunk come from shared-memory counters (warp-aggregated), the
// chunk then reserves its space in the 8 global buckets with 8 atomics.  (One global atomic per warp and bucket made
// 2.6 M same-address atomics per launch the bottleneck: 0.74 ms for 33 M entries at 6 % issue utilisation and 12 % DRAM.)
enum { EL_CLASSIFY_PER_THREAD = 16 };
__global__ void __launch_bounds__(128) k_classify(WaveState W, const __grid_constant__ DevScene S) {
    __shared__ uint32_t shist[EL_BUCKETS], sbase[EL_BUCKETS];
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n = W.cnt[CNT_CUR];
    const uint32_t chunk = 128u * EL_CLASSIFY_PER_THREAD;
    for (uint32_t c0 = blockIdx.x * chunk; c0 < n; c0 += gridDim.x * chunk) {        // static partition: no work-fetch atomic
        if (threadIdx.x < EL_BUCKETS) shist[threadIdx.x] = 0u;
        __syncthreads();
        uint32_t pid[EL_CLASSIFY_PER_THREAD], key[EL_CLASSIFY_PER_THREAD];            // key = bucket | rank << 3
#pragma unroll
        for (int k = 0; k < EL_CLASSIFY_PER_THREAD; k++) {
            const uint32_t qi = c0 + (uint32_t)k * 128u + threadIdx.x;
            const bool valid = qi < n;
            uint32_t bucket = EL_MISS_BUCKET;
            pid[k] = 0u;
            if (valid) { pid[k] = W.qCur[qi]; bucket = W.hitBucket[qi]; }
            const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
            key[k] = 0xffffffffu;
            if (valid) {
                const uint32_t peers = __match_any_sync(vmask, bucket);
                const uint32_t leader = __ffs(peers) - 1u;
                uint32_t pos = 0;
                if (lane == leader) pos = atomicAdd(&shist[bucket], (uint32_t)__popc(peers));
                pos = __shfl_sync(peers, pos, leader) + __popc(peers & ((1u << lane) - 1u));
                key[k] = bucket | (pos << 3);
            }
        }
        __syncthreads();
        if (threadIdx.x < EL_BUCKETS) sbase[threadIdx.x] = shist[threadIdx.x] ? atomicAdd(&W.cnt[CNT_BUCKET0 + threadIdx.x], shist[threadIdx.x]) : 0u;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < EL_CLASSIFY_PER_THREAD; k++) {
            if (key[k] == 0xffffffffu) continue;
            const uint32_t bucket = key[k] & 7u;
            W.qBucket[(size_t)bucket * W.pathCapacity + sbase[bucket] + (key[k] >> 3)] = pid[k];
        }
        __syncthreads();
    }
}

// ---- shade -----------------------------------------------------------------------------------------------------
__device__ __forceinline__ void appendWarpAggregated(uint32_t* q, uint32_t* counter, bool pred, uint32_t value) {
    const uint32_t mask = __ballot_sync(0xffffffffu, pred);   // called by all 32 lanes of the warp
    if (!pred) return;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t leader = __ffs(mask) - 1u;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(counter, (uint32_t)__popc(mask));
    base = __shfl_sync(mask, base, leader);
    q[base + __popc(mask & ((1u << lane) - 1u))] = value;
}

// Both of k_shade's queue appends with ONE 64-bit atomicAdd on the (CNT_NEXT, CNT_NEE) pair: the ncu source page had 9 % of the kernel's stall samples
// on the two shuffles that wait for the two atomics' results.  k_shade 6.47 -> 6.12 ms per 16-spp step (profiles/r2_variants_session9.json).
__device__ __forceinline__ void appendNextAndNee(const WaveState& W, bool toNext, bool toNee, uint32_t pid) {
    const uint32_t mNext = __ballot_sync(0xffffffffu, toNext), mNee = __ballot_sync(0xffffffffu, toNee);   // called by all 32 lanes
    if ((mNext | mNee) == 0u) return;
    const uint32_t lane = threadIdx.x & 31u;
    unsigned long long old = 0ull;
    if (lane == 0u) old = atomicAdd(reinterpret_cast<unsigned long long*>(&W.cnt[CNT_NEXT]), (unsigned long long)__popc(mNext) | ((unsigned long long)__popc(mNee) << 32));
    old = __shfl_sync(0xffffffffu, old, 0);
    const uint32_t below = (1u << lane) - 1u;
    if (toNext) W.qNext[(uint32_t)old + __popc(mNext & below)] = pid;
    if (toNee) W.qNee[(uint32_t)(old >> 32) + __popc(mNee & below)] = pid;
}

#ifndef EL_SHADE_ONE_ATOMIC
#define EL_SHADE_ONE_ATOMIC 1
#endif
#ifndef EL_SHADE_STATIC
#define EL_SHADE_STATIC 0        /* 1: queue blocks by a static stride over the grid's warps instead of the atomic work fetch (5 % of the stall samples sit on its shuffle).
                                  * Measured SLOWER: k_shade 6.48 vs 6.12 ms (profiles/r2_variants_session9.json): blocks of escaped rays and of hits do not cost the same.
                                  * Issuing the fetch atomic one iteration ahead of its broadcast: 6.32 vs 6.12 ms (session 11), dropped. */
#endif
#ifndef EL_SHADE_PREFETCH
#define EL_SHADE_PREFETCH 1
#endif
#ifndef EL_SHADE_MIN_CTAS
#define EL_SHADE_MIN_CTAS 5      /* <= 102 registers (96 used, no spills), 20 warps/SM.  Without the prefetch 6 CTAs x 80 registers was the best point
                                  * (round 1); with it the queue -> hit part of the gather chain is hidden and the 200 bytes of spills cost more than the
                                  * sixth CTA brings: 7.52 -> 6.46 ms per 16-spp step (profiles/r2_variants_session2.json) */
#endif
template <bool FM>
__global__ void __launch_bounds__(128, EL_SHADE_MIN_CTAS) k_shade(WaveState W, const __grid_constant__ DevScene S, const __grid_constant__ RenderParams P) {
    const uint32_t lane = threadIdx.x & 31u;
    // the shading queue is the concatenation of the material buckets written by k_classify: consecutive entries share a
    // material (same textures, same branches), escaped rays come last
    uint32_t prefix[EL_BUCKETS + 1];
    prefix[0] = 0;
#pragma unroll
    for (int b = 0; b < EL_BUCKETS; b++) prefix[b + 1] = prefix[b] + W.cnt[CNT_BUCKET0 + b];
    const uint32_t n = prefix[EL_BUCKETS];
    // queue position -> path id through the bucket prefix sums
    auto queueEntry = [&](uint32_t qi) -> uint32_t {
        uint32_t b = 0;
#pragma unroll
        for (int k = 1; k < EL_BUCKETS; k++) b += (qi >= prefix[k]) ? 1u : 0u;
        return W.qBucket[(size_t)b * W.pathCapacity + (qi - prefix[b])];
    };
#if EL_SHADE_PREFETCH
    // Software pipeline over the warp's queue blocks: the chain queue entry -> hit record -> triangle -> texel is four dependent
    // gathers (ncu: long_scoreboard 4.2 warps per issue cycle, issue 54 %).  The NEXT block's work fetch, queue entry and hit record
    // are requested while the current block is shaded, so an iteration starts with the triangle index already in registers:
    // 8.65 -> 7.52 ms per 16-spp step at 6 CTAs/SM, 6.40 ms at 5 CTAs/SM without spills (profiles/r2_variants_session{2,3}.json).
    // A two-deep version (hit record of block i+1 at the top, prefetch.global.L2 of its ray / throughput / triangle, queue entry of
    // block i+2) measured SLOWER: 6.69 ms — the extra L2 prefetches compete with the demand gathers of a kernel at 56 % of the HBM roof.
#if EL_SHADE_STATIC
    const uint32_t stride = gridDim.x * (blockDim.x >> 5) * 32u;     // every block of 32 queue entries costs about the same: a static stride balances as well as the atomic
    uint32_t base = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32u;
#else
    uint32_t base = warpFetch(&W.cnt[CNT_WORK_SHADE], lane);
#endif
    uint32_t pid = 0; float4 hv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (base < n && base + lane < n) { pid = queueEntry(base + lane); hv = W.hit[pid]; }
    for (;;) {
        if (base >= n) break;
        const uint32_t qi = base + lane;
#if EL_SHADE_STATIC
        const uint32_t nbase = (n - base > stride) ? base + stride : n;          // no wrap-around for n close to 2^32
#else
        const uint32_t nbase = warpFetch(&W.cnt[CNT_WORK_SHADE], lane);
#endif
        uint32_t npid = 0; float4 nhv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (nbase < n && nbase + lane < n) npid = queueEntry(nbase + lane);
        bool toNee = false, toNext = false;
        if (qi < n) {
#else
    for (;;) {
        const uint32_t base = warpFetch(&W.cnt[CNT_WORK_SHADE], lane);
        if (base >= n) break;
        const uint32_t qi = base + lane;
        bool toNee = false, toNext = false;
        uint32_t pid = 0;
        if (qi < n) {
            pid = queueEntry(qi);
            const float4 hv = W.hit[pid];
#endif
            const int tri = __float_as_int(hv.x);
            const float4 o4 = W.ray[2 * (size_t)pid], d4 = W.ray[2 * (size_t)pid + 1];
            Ray ray; ray.o = f3(o4.x, o4.y, o4.z); ray.d = f3(d4.x, d4.y, d4.z);
            const float4 thr4 = W.tr[2 * (size_t)pid];
            const F3 thr = f3(thr4.x, thr4.y, thr4.z);
            if (tri < 0) {
                // escaped: light += env(dir) * reduction (S/kernel.cu:414-419)
                const F3 e = envLookup(S, ray.d);
                float4 r = W.tr[2 * (size_t)pid + 1];
                r.x += e.x * thr.x; r.y += e.y * thr.y; r.z += e.z * thr.z;
                W.tr[2 * (size_t)pid + 1] = r;
            } else {
                const uint32_t depth = __float_as_uint(o4.w);
                // --- random numbers: 3 for the bounce, 3 for shade of which only the first is used.  Drawn FIRST:
                //     they need only (pid, depth), and the environment sample's table + texel fetches (two dependent gathers into
                //     ~200 MB) can then fly while the triangle and the texture maps are fetched
                float b1, b2, b3, s1; uint32_t aliasBitsA = 0, aliasBitsB = 0, lightBits = 0;
                if (P.rngMode == ELEVEN_RNG_REFERENCE) {
                    Xorwow s = W.rng[pid];
                    b1 = xorwowUniform(s); b2 = xorwowUniform(s); b3 = xorwowUniform(s);
                    s1 = xorwowUniform(s); xorwowNext(s); xorwowNext(s);
                    W.rng[pid] = s;
                } else {
                    const uint4 a = fastBits(P, pid, 2u + 2u * depth);
                    b1 = u32ToUniform(a.x); b2 = u32ToUniform(a.y); b3 = u32ToUniform(a.z); s1 = u32ToUniform(a.w);
                    const uint4 c = fastBits(P, pid, 3u + 2u * depth);
                    aliasBitsA = c.x; aliasBitsB = c.y; lightBits = c.z;
                }
                // --- environment NEE, part 1: pick the texel (hdriLight, S/kernel.cu:236-242) and fetch it
                const int EW = S.hdri.width, EH = S.hdri.height;
                int texel;
                if (P.envMode == ELEVEN_ENV_CDF) texel = cdfSearch(S.cdf, s1, EW * EH);
                else {
                    const uint32_t nT = (uint32_t)(EW * EH);
                    const uint32_t k = P.rngMode == ELEVEN_RNG_REFERENCE ? min(nT - 1u, (uint32_t)(s1 * (float)nT)) : __umulhi(aliasBitsA, nT);
                    const float xi = P.rngMode == ELEVEN_RNG_REFERENCE ? s1 * (float)nT - floorf(s1 * (float)nT) : u32ToUniform(aliasBitsB);
                    const AliasEntry ae = S.alias[k];
                    texel = xi <= ae.prob ? (int)k : (int)ae.alias;
                }
                const float sx = (float)(texel % EW), sy = (float)(texel / EW);
                const float nu = sx / (float)EW, nv = sy / (float)EH;
                float iu, iv; inverseTransformUV(S.hdri, nu, nv, iu, iv);
                const float4 ev = envTexelRaw(S.hdri, (int)(iu * EW), (int)(iv * EH));

                // --- hit attributes: Tri::hit's second half (S/Tri.hpp:70-157), exact arithmetic -----------------
                const float t = hv.y, u = hv.z, v = hv.w;
                const float4* tp = S.shadeTris + (size_t)tri * 9;
                const float4 q4 = __ldg(tp + 4), q5 = __ldg(tp + 5), q6 = __ldg(tp + 6), q7 = __ldg(tp + 7), q8 = __ldg(tp + 8);
                const TriGeom g = loadTriGeom(S.shadeTris, tri);
#if EL_SHADE_PREFETCH
                if (nbase < n && nbase + lane < n) nhv = W.hit[npid];   // the next block's hit record: requested behind this path's own gathers
#endif
                F3 N;
                const F3 Pp = hitPosition(ray, g, t, u, v, N);
                const F3 t0 = f3(q4.z, q4.w, q5.x), t1 = f3(q5.y, q5.z, q5.w), t2 = f3(q6.x, q6.y, q6.z);
                const F3 T = baryLerp(t0, t1, t2, u, v);
                const float sign = q6.w;
                const F3 nxt = ex::cross(N, T);
                const F3 B = f3(ex::mul(nxt.x, sign), ex::mul(nxt.y, sign), ex::mul(nxt.z, sign));
                const F3 uv0 = f3(q7.x, q7.y, 0.f), uv1 = f3(q7.z, q7.w, 0.f), uv2 = f3(q8.x, q8.y, 0.f);
                const F3 tUV = baryLerp(uv0, uv1, uv2, u, v);
                const int objectID = __float_as_int(q8.z);
                const int mid = S.objectMaterial[objectID];
                const DevMaterial& m = S.materials[mid];
                HitData hd;
                generateHitData<FM>(S, m, S.packed[mid], hd, N, T, B, tUV.x, tUV.y);

                const BrdfFrame bf = makeBrdfFrame<FM>(hd, ray.d);
                const F3 L = disneySample<FM>(hd, bf, b1, b2, b3);
                const F3 fB = disneyEval<FM>(hd, bf, L);
                const float pB = disneyPdf<FM>(hd, bf, L);

                // --- environment NEE, part 2 (hdriLight, S/kernel.cu:243-256) -----------------------------------
                const F3 rsm = M<FM>::normalized(reverseSphericalMapping<FM>(iu, iv));
                const F3 wE = f3(-rsm.x, -rsm.y, -rsm.z);
                const float pE = hdriPdf<FM>(S, ev, (int)(iv * EH));
                const F3 fE = disneyEval<FM>(hd, bf, wE);
                const float cE = fabsf(dot(wE, hd.normal));
                const F3 CE = f3(M<FM>::div(fE.x * cE * ev.x, pE), M<FM>::div(fE.y * cE * ev.y, pE), M<FM>::div(fE.z * cE * ev.z, pE));

                // --- point-light NEE set-up (pointLight, S/kernel.cu:175-205) -----------------------------------
                F3 wL = f3(0.f), CP = f3(0.f); float dist = 0.f, pP = 0.f;
                if (S.lightCount > 0) {
                    pP = (float)((double)(float)S.lightCount / (2.0 * (double)EL_PI));
                    int li = P.rngMode == ELEVEN_RNG_REFERENCE ? (int)((float)S.lightCount * s1) : (int)__umulhi(lightBits, S.lightCount);
                    if (li >= (int)S.lightCount) li = (int)S.lightCount - 1;
                    const float* lp = S.lights + 6 * li;
                    const F3 lpos = f3(lp[0], lp[1], lp[2]), lrad = f3(lp[3], lp[4], lp[5]);
                    wL = M<FM>::normalized(lpos - Pp);
                    dist = length(lpos - Pp);
                    const F3 val = M<FM>::div3(lrad, dist * dist);
                    const F3 fL = disneyEval<FM>(hd, bf, wL);
                    const float cL = fabsf(dot(wL, hd.normal));
                    CP = f3(M<FM>::div(val.x * fL.x * cL, pP), M<FM>::div(val.y * fL.y * cL, pP), M<FM>::div(val.z * fL.z * cL, pP));
                }
                // --- emission through the BRDF strategy + throughput factor (shade, S/kernel.cu:349,357) ------------
                const float cB = fabsf(dot(L, hd.normal));
                const F3 mulB = f3(M<FM>::div(fB.x * cB, pB), M<FM>::div(fB.y * cB, pB), M<FM>::div(fB.z * cB, pB));
                const F3 CB = hd.emission * mulB;

                // the environment shadow ray, complete: Ray(point + newDir*0.001, newDir) with the constructor's normalisation
                // (S/kernel.cu:246, S/Ray.hpp:14-18).  Same operations as before, moved here from the shadow kernel's ray set-up, where
                // they were 7-8 % of an issue-bound kernel's instructions; this kernel waits on gathers and has the slots free.
                const Ray er = makeRay(ex::madd(Pp, wE, 0.001f), wE);
                neeRec(W, pid, NEE_ENV_DIR) = make_float4(er.d.x, er.d.y, er.d.z, pE);
                neeRec(W, pid, NEE_ENV_C) = make_float4(CE.x, CE.y, CE.z, pB);
                neeRec(W, pid, NEE_POS) = make_float4(er.o.x, er.o.y, er.o.z, 0.f);
                neeRec(W, pid, NEE_THR_MUL) = make_float4(mulB.x, mulB.y, mulB.z, 0.f);
                if (W.sceneHasEmission) neeRec(W, pid, NEE_BRDF_C) = make_float4(CB.x, CB.y, CB.z, 0.f);
                if (S.lightCount > 0) {
                    neeRec(W, pid, NEE_HITPOS) = make_float4(Pp.x, Pp.y, Pp.z, 0.f);
                    neeRec(W, pid, NEE_LIGHT_DIR) = make_float4(wL.x, wL.y, wL.z, dist);
                    neeRec(W, pid, NEE_LIGHT_C) = make_float4(CP.x, CP.y, CP.z, pP);
                }
                if (depth == 0) {                                   // first-hit AOVs (S/kernel.cu:436-440)
                    W.aovN[pid] = make_float4(N.x, N.y, N.z, 0.f);
                    W.aovT[pid] = make_float4(T.x, T.y, T.z, 0.f);
                    W.aovB[pid] = make_float4(B.x, B.y, B.z, 0.f);
                }
                // next ray: Ray(P + L*0.001, L) (S/kernel.cu:442)
                const Ray nr = makeRay(ex::madd(Pp, L, 0.001f), L);
                W.ray[2 * (size_t)pid] = make_float4(nr.o.x, nr.o.y, nr.o.z, __uint_as_float(depth + 1u));
                W.ray[2 * (size_t)pid + 1] = make_float4(nr.d.x, nr.d.y, nr.d.z, 0.f);
                toNee = true;
                if ((P.flags & ELEVEN_FLAG_SKIP_NULL_NEE) && S.lightCount == 0 && isfinite(pE) &&
                    CE.x == 0.f && CE.y == 0.f && CE.z == 0.f && CB.x == 0.f && CB.y == 0.f && CB.z == 0.f) {
                    // the MIS sum is w1*0 + 0 + w3*0 whatever the shadow ray finds (e.g. the environment sample lies below
                    // the surface): no shadow ray, apply the throughput factor here
                    W.tr[2 * (size_t)pid] = make_float4(thr.x * mulB.x, thr.y * mulB.y, thr.z * mulB.z, 0.f);
                    toNee = false;
                }
                toNext = depth + 1u < P.maxBounces;
                if ((P.flags & ELEVEN_FLAG_TERMINATE_DEAD_PATHS) && toNext) {
                    // throughput after this bounce is thr * mulB: exactly zero means no later bounce can contribute
                    const F3 nt = thr * mulB;
                    if (nt.x == 0.f && nt.y == 0.f && nt.z == 0.f) toNext = false;
                }
            }
        }
#if EL_SHADE_ONE_ATOMIC
        appendNextAndNee(W, toNext, toNee, pid);
#else
        appendWarpAggregated(W.qNee, &W.cnt[CNT_NEE], toNee, pid);
        appendWarpAggregated(W.qNext, &W.cnt[CNT_NEXT], toNext, pid);
#endif
#if EL_SHADE_PREFETCH
        if (nbase < n && nbase + lane < n && !(qi < n && __float_as_int(hv.x) >= 0)) nhv = W.hit[npid];   // lanes that shaded no hit this round (escaped ray / ragged end) have not asked yet
        base = nbase; pid = npid; hv = nhv;
#endif
    }
}


// ---- queue bookkeeping between bounces (single thread) ---------------------------------------------------------------
__global__ void k_advance(WaveState W, uint32_t lights, int phase) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (phase == 0) {            // after extend+shade: account rays, arm connect
        W.stats[ST_RAYS_EXT] += W.cnt[CNT_CUR];
        W.stats[ST_RAYS_ENV] += W.cnt[CNT_NEE];
        if (lights) W.stats[ST_RAYS_LIGHT] += W.cnt[CNT_NEE];
        W.cnt[CNT_WORK_CONNECT] = 0u; W.cnt[CNT_WORK_LIGHT] = 0u;
    } else if (phase == 1) {     // after connect: next bounce
        W.cnt[CNT_CUR] = W.cnt[CNT_NEXT]; W.cnt[CNT_NEXT] = 0u; W.cnt[CNT_NEE] = 0u;
        W.cnt[CNT_WORK_TRACE] = 0u; W.cnt[CNT_WORK_SHADE] = 0u; W.cnt[CNT_WORK_CLASSIFY] = 0u;
        for (int b = 0; b < EL_BUCKETS; b++) W.cnt[CNT_BUCKET0 + b] = 0u;
    } else if (phase == 2) {     // stage overlap: the next bounce's extend + classify start while connect still reads the NEE queue
        W.cnt[CNT_CUR] = W.cnt[CNT_NEXT]; W.cnt[CNT_NEXT] = 0u;
        W.cnt[CNT_WORK_TRACE] = 0u; W.cnt[CNT_WORK_SHADE] = 0u; W.cnt[CNT_WORK_CLASSIFY] = 0u;
        for (int b = 0; b < EL_BUCKETS; b++) W.cnt[CNT_BUCKET0 + b] = 0u;
    } else {                     // stage overlap: connect has finished, shade may refill the NEE queue
        W.cnt[CNT_NEE] = 0u;
    }
}

// ---- accumulate (S/kernel.cu:445-480) with sums instead of running means -------------------------------------------------
// The K samples of a pixel are added in sample order, so the sums are the same floats for every K.  A block owns 256 >> logK
// pixels = 256 consecutive paths: every thread loads ITS path's records (coalesced 16-byte loads; one thread per pixel
// walking its K records strides the warp over K x 16 bytes and ran at a fifth of the HBM rate), stages the clamped values
// in shared memory, and 4 threads per pixel (one per film pass) do the ordered sums.
__global__ void __launch_bounds__(256) k_accumulate(WaveState W, uint32_t logK) {
    __shared__ float sv[4][256][3];
    __shared__ uint32_t sdepth[256];
    __shared__ uint8_t sok[256];
    const uint32_t tid = threadIdx.x;
    const uint32_t nPaths = W.nPixels << logK;
    const uint32_t p = blockIdx.x * 256u + tid;
    if (p < nPaths) {
        float4 r = W.tr[2 * (size_t)p + 1];
        r.x = clampf_(r.x, 0.f, 10.f); r.y = clampf_(r.y, 0.f, 10.f); r.z = clampf_(r.z, 0.f, 10.f);
        const bool ok = !isnan(r.x) && !isnan(r.y) && !isnan(r.z);
        const uint32_t dep = __float_as_uint(W.ray[2 * (size_t)p].w);
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 n = dep ? W.aovN[p] : z, t = dep ? W.aovT[p] : z, bt = dep ? W.aovB[p] : z;     // camera ray escaped: AOVs are 0 (S/kernel.cu:388-391)
        sv[0][tid][0] = r.x; sv[0][tid][1] = r.y; sv[0][tid][2] = r.z;
        sv[1][tid][0] = n.x; sv[1][tid][1] = n.y; sv[1][tid][2] = n.z;
        sv[2][tid][0] = t.x; sv[2][tid][1] = t.y; sv[2][tid][2] = t.z;
        sv[3][tid][0] = bt.x; sv[3][tid][1] = bt.y; sv[3][tid][2] = bt.z;
        sdepth[tid] = dep; sok[tid] = ok ? 1 : 0;
    }
    __syncthreads();
    const uint32_t K = 1u << logK, pixPerBlock = 256u >> logK;
    const uint32_t pass = tid & 3u;                                // film pass; 4 threads per pixel, 64 pixels per sweep
    float4* film = pass == 0 ? W.filmBeauty : pass == 1 ? W.filmNormal : pass == 2 ? W.filmTangent : W.filmBitangent;
    for (uint32_t lp = tid >> 2; lp < pixPerBlock; lp += 64u) {
        const uint32_t pix = blockIdx.x * pixPerBlock + lp;
        if (pix >= W.nPixels) break;
        float4 f = film[pix];
        uint32_t count = 0, paths = 0;
        for (uint32_t k = 0; k < K; k++) {
            const uint32_t q = (lp << logK) + k;
            paths += sdepth[q];
            if (sok[q]) { f.x += sv[pass][q][0]; f.y += sv[pass][q][1]; f.z += sv[pass][q][2]; count++; }
        }
        f.w += (float)count;
        film[pix] = f;
        if (pass == 0) W.pathCount[pix] += paths;
    }
}

// film read-back: mean, alpha = 1 (S/kernel.cu:137,461-463)
// (plain loads, not __ldg: the film may be written by a concurrent wave on the render stream)
__global__ void k_resolve(const float4* sums, float4* __restrict__ out, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 s = sums[i];
    const float inv = s.w > 0.f ? 1.0f / s.w : 0.f;
    out[i] = make_float4(s.x * inv, s.y * inv, s.z * inv, 1.0f);
}
__global__ void k_sampleCounts(const float4* sums, uint32_t* __restrict__ out, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint32_t)sums[i].w;
}
// fused resolve -> 8-bit with the reference's output curve fastPow(clamp01(x), 1/2.2)*255 (S/main.cpp:156-158, S/PostProcessing.cpp:30-33)
__device__ __forceinline__ double fastPowDev(double a, double b) {
    int hi = __double2hiint(a);
    hi = (int)(b * (double)(hi - 1072632447) + 1072632447.0);
    return __hiloint2double(hi, 0);
}
__global__ void k_resolve8(const float4* sums, uchar4* __restrict__ out, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 s = sums[i];
    const float inv = s.w > 0.f ? 1.0f / s.w : 0.f;
    const float v[4] = {s.x * inv, s.y * inv, s.z * inv, 1.0f};
    unsigned char o[4];
    for (int k = 0; k < 4; k++) {
        const float x = clampf_(v[k], 0.f, 1.f);
        o[k] = (unsigned char)(fastPowDev((double)x, 1.0 / 2.2) * 255.0);
    }
    out[i] = make_uchar4(o[0], o[1], o[2], o[3]);
}


} // namespace eleven
