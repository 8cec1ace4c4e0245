/*
 * eleven_api.cu — the C ABI of include/eleven_b200.h: context, scene upload, wave scheduling, film read-back.
 *
 * Replaces the host half of S/kernel.cu (renderSetup :566-661, renderCuda :665-686, getBuffers :688-710,
 * getSamples :712-724; S/ = reference src/tfg-pathtracer) with a context object (no globals, no fixed
 * 1920x1080 symbols, S/kernel.cu:41-47), checked CUDA calls and zero host synchronisations inside a render call.
 */
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>          /* types only: the library is dlopen'ed on first use (eleven_comm_*), so libeleven_b200.so has no link-time NCCL dependency */

#include "../../include/eleven_b200.h"
#include "bvh8.h"
#include "kernels.cuh"
#include "kernels_trace.cuh"
#include "bvh8_build_gpu.cuh"
#include "scene_upload.cuh"
#include "test_hooks.cuh"

using namespace eleven;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }

#define CK(call)                                                                                           \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess) {                                                                           \
            return fail(ELEVEN_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
        }                                                                                                  \
    } while (0)

struct ElevenCtx {
    ElevenConfig cfg;
    cudaStream_t stream = nullptr;               // render stream: every kernel of eleven_render / eleven_trace_* / uploads
    cudaStream_t snapStream = nullptr;           // snapshot stream: film / counter read-backs, concurrent with a running eleven_render
                                                 // (the reference polls getBuffers on its bufferStream while the kernel runs, S/kernel.cu:688-710)
    cudaStream_t auxStream = nullptr;            // the shadow stage of bounce b runs here, concurrently with extend + classify of bounce b+1
    cudaEvent_t evFork = nullptr, evJoin = nullptr;
    bool overlap = true;                         // ELEVEN_OVERLAP=0 serialises the stages on the render stream (A/B knob)
    std::mutex snapMutex;                        // serialises the users of the snapshot stream and of d_resolve
    std::mutex statsMutex;                       // host-side ElevenStats fields written at the end of eleven_render
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evR0 = nullptr, evR1 = nullptr;
    ncclComm_t comm = nullptr;                   // film-reduce communicator (eleven_comm_init_rank / eleven_comm_init_all)
    int commRank = -1, commSize = 0;
    float4* d_reduced = nullptr;                 // root only: sum over ranks of the film passes (eleven_reduce_film)
    float* d_commWarm = nullptr;                 // 8 MB reduced once when the communicator is created (NCCL's lazy channel set-up), then freed
    int reducedPasses = 0;
    std::vector<void*> sceneAllocs, waveAllocs;
    DevScene scene;
    WaveState W;
    RenderParams P;
    bool haveScene = false;
    uint32_t nPixels = 0, width = 0, height = 0;
    uint32_t maxLogK = 0;                        // a wave carries up to 2^maxLogK samples of every pixel
    uint32_t samplesRendered = 0;
    int numSMs = 148;
    uint32_t* d_seqMat = nullptr;
    float4* d_resolve = nullptr;
    uint32_t* d_workCounter = nullptr;
    ElevenCamera* d_camera = nullptr;
    std::vector<cudaEvent_t> evPool;
    ElevenStats stats;
    gpubvh::BuildArena bvhArena;                 // device BVH build scratch, kept for re-builds
};

extern "C" int eleven_abi_version(void) { return ELEVEN_ABI_VERSION; }
extern "C" const char* eleven_last_error(void) { return g_err.c_str(); }

template <typename T>
static int devAlloc(std::vector<void*>& list, T** p, size_t count) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T));
    if (e != cudaSuccess) return fail(ELEVEN_ERR_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    list.push_back(q); *p = (T*)q; return ELEVEN_OK;
}
template <typename T>
static int devUpload(std::vector<void*>& list, const T** p, const T* host, size_t count) {
    T* q = nullptr;
    int rc = devAlloc(list, &q, count); if (rc) return rc;
    if (count) CK(cudaMemcpy(q, host, count * sizeof(T), cudaMemcpyHostToDevice));
    *p = q; return ELEVEN_OK;
}
static void freeAll(std::vector<void*>& list) { for (void* p : list) cudaFree(p); list.clear(); }

static void commDestroy(ElevenCtx* c);

extern "C" int eleven_init(const ElevenConfig* cfg, ElevenCtx** out) {
    if (!cfg || !out) return fail(ELEVEN_ERR_ARG, "eleven_init: null argument");
    if (cfg->rng_mode > 1 || cfg->env_mode > 1 || cfg->hit_mode > 1 || cfg->bvh_builder > 1) return fail(ELEVEN_ERR_ARG, "eleven_init: bad mode");
    if (cfg->wave_spp > 16 || (cfg->wave_spp & (cfg->wave_spp - 1u)) != 0u) return fail(ELEVEN_ERR_ARG, "eleven_init: wave_spp must be 0 (auto) or a power of two <= 16");
    if (cfg->wave_spp > 1 && cfg->rng_mode == ELEVEN_RNG_REFERENCE)
        return fail(ELEVEN_ERR_ARG, "eleven_init: the reference RNG stream is sequential per pixel; wave_spp > 1 needs ELEVEN_RNG_FAST");
    if ((cfg->flags & ELEVEN_FLAG_TERMINATE_DEAD_PATHS) && cfg->rng_mode == ELEVEN_RNG_REFERENCE)
        return fail(ELEVEN_ERR_ARG, "eleven_init: dead-path termination changes the XORWOW stream; use it with ELEVEN_RNG_FAST only");
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(ELEVEN_ERR_ARG, "eleven_init: no such CUDA device");
    CK(cudaSetDevice(cfg->device));
    ElevenCtx* c = new ElevenCtx();
    c->cfg = *cfg;
    if (c->cfg.max_bounces == 0) c->cfg.max_bounces = 5;
    if (c->cfg.sample_stride == 0) c->cfg.sample_stride = 1;
    memset(&c->scene, 0, sizeof c->scene); memset(&c->W, 0, sizeof c->W); memset(&c->stats, 0, sizeof c->stats);
    cudaError_t e = cudaDeviceGetAttribute(&c->numSMs, cudaDevAttrMultiProcessorCount, cfg->device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    int prLo = 0, prHi = 0;
    if (e == cudaSuccess) e = cudaDeviceGetStreamPriorityRange(&prLo, &prHi);
    // highest priority: the persistent render kernels fill every SM, so a snapshot's blocks are placed when the running kernel retires
    if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&c->snapStream, cudaStreamNonBlocking, prHi);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->auxStream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->evFork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->evJoin, cudaEventDisableTiming);
    if (const char* ov = getenv("ELEVEN_OVERLAP")) c->overlap = atoi(ov) != 0;
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev1);
    if (e == cudaSuccess) e = cudaEventCreate(&c->evR0);
    if (e == cudaSuccess) e = cudaEventCreate(&c->evR1);
    if (e != cudaSuccess) {                      // nothing half-built survives a failed init
        eleven_destroy(c);
        return fail(ELEVEN_ERR_CUDA, std::string("eleven_init: ") + cudaGetErrorString(e));
    }
    *out = c;
    return ELEVEN_OK;
}

extern "C" void eleven_destroy(ElevenCtx* c) {
    if (!c) return;
    cudaSetDevice(c->cfg.device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->snapStream) cudaStreamSynchronize(c->snapStream);
    if (c->auxStream) cudaStreamSynchronize(c->auxStream);
    commDestroy(c);
    freeAll(c->sceneAllocs); freeAll(c->waveAllocs);
    if (c->d_reduced) cudaFree(c->d_reduced);
    if (c->d_commWarm) cudaFree(c->d_commWarm);
    if (c->bvhArena.base) cudaFree(c->bvhArena.base);
    for (cudaEvent_t e : {c->ev0, c->ev1, c->evR0, c->evR1, c->evFork, c->evJoin}) if (e) cudaEventDestroy(e);
    if (c->auxStream) cudaStreamDestroy(c->auxStream);
    for (cudaEvent_t e : c->evPool) cudaEventDestroy(e);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->snapStream) cudaStreamDestroy(c->snapStream);
    delete c;
}

// ---- host-side preparation ---------------------------------------------------------------------------------------
// stb_image's patched LDR decode (S/stb_image.h:127-136,1863): published algorithm (Ankerl's approximate pow)
static double fastPowHost(double a, double b) {
    union { double d; int32_t x[2]; } u; u.d = a;
    u.x[1] = (int32_t)(b * (u.x[1] - 1072632447) + 1072632447);
    u.x[0] = 0;
    return u.d;
}
static void buildLut(float* lut512) {
    for (int i = 0; i < 256; i++) {
        lut512[i] = (float)(fastPowHost((float)i / 255.0f, 2.2f) * 1.0f);
        lut512[256 + i] = (float)(fastPowHost((float)i / 255.0f, 1.0f) * 1.0f);
    }
}
// XORWOW sequence-skip matrices: step^(2^67 * 2^k), k = 0..23, by repeated squaring over GF(2) (see k_rngInit)
static void xorwowStepHost(uint32_t v[5]) {
    uint32_t t = v[0] ^ (v[0] >> 2);
    v[0] = v[1]; v[1] = v[2]; v[2] = v[3]; v[3] = v[4];
    v[4] = (v[4] ^ (v[4] << 4)) ^ (t ^ (t << 1));
}
struct BitMat { uint32_t col[160][5]; };
static void matVecHost(const BitMat& m, const uint32_t in[5], uint32_t out[5]) {
    uint32_t r[5] = {0, 0, 0, 0, 0};
    for (int w = 0; w < 5; w++) for (int b = 0; b < 32; b++) if (in[w] & (1u << b)) for (int k = 0; k < 5; k++) r[k] ^= m.col[w * 32 + b][k];
    for (int k = 0; k < 5; k++) out[k] = r[k];
}
static void matMulHost(const BitMat& a, const BitMat& b, BitMat& out) { BitMat r; for (int j = 0; j < 160; j++) matVecHost(a, b.col[j], r.col[j]); out = r; }
static void buildSeqMats(std::vector<uint32_t>& flat, int count) {
    BitMat m;
    for (int j = 0; j < 160; j++) { uint32_t v[5] = {0, 0, 0, 0, 0}; v[j / 32] = 1u << (j % 32); xorwowStepHost(v); for (int k = 0; k < 5; k++) m.col[j][k] = v[k]; }
    for (int i = 0; i < 67; i++) matMulHost(m, m, m);
    flat.resize((size_t)count * 800);
    for (int k = 0; k < count; k++) {
        memcpy(&flat[(size_t)k * 800], m.col, 800 * 4);
        matMulHost(m, m, m);
    }
}

static int uploadTexture(ElevenCtx* c, const ElevenTexture& t, DevTex& out, bool hdri) {
    if (!t.data || t.width <= 0 || t.height <= 0) return fail(ELEVEN_ERR_ARG, "texture: null data or bad size");
    if (t.format > ELEVEN_TEX_U8_LINEAR) return fail(ELEVEN_ERR_ARG, "texture: unknown format");
    if (hdri && t.format != ELEVEN_TEX_F32_RGB) return fail(ELEVEN_ERR_ARG, "hdri must be ELEVEN_TEX_F32_RGB");
    const size_t n = (size_t)t.width * t.height;
    out.width = t.width; out.height = t.height; out.xTile = t.xTile; out.yTile = t.yTile; out.xOffset = t.xOffset; out.yOffset = t.yOffset;
    out.format = t.format; out.filter = t.filter;
    // the caller's texels go up as they are and are expanded on the device (scene_upload.cuh); the staging buffer is stream-ordered
    const size_t rawBytes = n * (t.format == ELEVEN_TEX_F32_RGB ? 12 : 3);
    void* raw = nullptr;
    cudaError_t e = cudaMallocAsync(&raw, rawBytes, c->stream);
    if (e != cudaSuccess) return fail(ELEVEN_ERR_NOMEM, std::string("cudaMallocAsync(texture staging): ") + cudaGetErrorString(e));
    CK(cudaMemcpyAsync(raw, t.data, rawBytes, cudaMemcpyHostToDevice, c->stream));
    const unsigned grid = (unsigned)((n + 255) / 256);
    int rc;
    if (t.format == ELEVEN_TEX_F32_RGB) {
        float4* d = nullptr; if ((rc = devAlloc(c->sceneAllocs, &d, n))) return rc;
        k_expandRgbF32<<<grid, 256, 0, c->stream>>>((const float*)raw, d, n);
        out.data = d;
    } else {
        uchar4* d = nullptr; if ((rc = devAlloc(c->sceneAllocs, &d, n))) return rc;
        k_expandRgb8<<<grid, 256, 0, c->stream>>>((const uint8_t*)raw, d, n);
        out.data = d;
    }
    CK(cudaGetLastError());
    CK(cudaFreeAsync(raw, c->stream));
    return ELEVEN_OK;
}

static int allocWaveK(ElevenCtx* c);

static int allocWave(ElevenCtx* c) {
    // samples per pixel in flight: as many as keep a wave within 2^26 paths (~24 GB of wave state on a 180 GB device: 16 at 1080p,
    // 8 at 3840x2160), at most 16; halved below until the allocation fits
    c->maxLogK = 0;
    if (c->cfg.rng_mode == ELEVEN_RNG_FAST) {
        if (c->cfg.wave_spp) { while ((1u << c->maxLogK) < c->cfg.wave_spp) c->maxLogK++; }
        else while (c->maxLogK < 4 && ((uint64_t)c->nPixels << (c->maxLogK + 1)) <= (1ull << 26)) c->maxLogK++;
    }
    if (((uint64_t)c->nPixels << c->maxLogK) > 0x7fffffffull) return fail(ELEVEN_ERR_ARG, "wave_spp x resolution exceeds 2^31 paths");
    // an automatic width that does not fit the device's free memory is halved until it does (the image does not depend on it)
    for (;;) {
        const int rc = allocWaveK(c);
        if (rc != ELEVEN_ERR_NOMEM || c->cfg.wave_spp != 0 || c->maxLogK == 0) return rc;
        cudaGetLastError();
        c->maxLogK--;
    }
}

static int allocWaveK(ElevenCtx* c) {
    freeAll(c->waveAllocs);
    WaveState& W = c->W; memset(&W, 0, sizeof W);
    W.nPixels = c->nPixels;
    W.pathCapacity = c->nPixels << c->maxLogK;
    const size_t n = W.pathCapacity, npx = c->nPixels;
    int rc = 0;
#define A(field, type) if ((rc = devAlloc(c->waveAllocs, &W.field, n))) return rc;
#define APX(field, type) if ((rc = devAlloc(c->waveAllocs, &W.field, npx))) return rc;
    if ((rc = devAlloc(c->waveAllocs, &W.ray, n * 2))) return rc;
    if ((rc = devAlloc(c->waveAllocs, &W.nee, n * NEE_STRIDE))) return rc;
    if ((rc = devAlloc(c->waveAllocs, &W.tr, n * 2))) return rc;
    A(hit, float4) A(aovN, float4) A(aovT, float4) A(aovB, float4)
    APX(rng, Xorwow) A(hitBucket, uint8_t)
    A(qCur, uint32_t) A(qNext, uint32_t) A(qNee, uint32_t)
    if ((rc = devAlloc(c->waveAllocs, &W.qBucket, n * EL_BUCKETS))) return rc;
    if ((rc = devAlloc(c->waveAllocs, &W.filmBeauty, npx * 4))) return rc;        // the four passes back to back (WaveState)
    W.filmNormal = W.filmBeauty + npx; W.filmTangent = W.filmBeauty + 2 * npx; W.filmBitangent = W.filmBeauty + 3 * npx;
    APX(pathCount, uint32_t)
#undef A
#undef APX
    const size_t npxResolve = npx;
    if ((rc = devAlloc(c->waveAllocs, &W.cnt, CNT_COUNT))) return rc;
    if ((rc = devAlloc(c->waveAllocs, &W.stats, ST_COUNT))) return rc;
    if ((rc = devAlloc(c->waveAllocs, &c->d_resolve, npxResolve))) return rc;
    if ((rc = devAlloc(c->waveAllocs, &c->d_workCounter, 4))) return rc;
    CK(cudaMemsetAsync(W.cnt, 0, CNT_COUNT * sizeof(uint32_t), c->stream));
    CK(cudaMemsetAsync(W.stats, 0, ST_COUNT * sizeof(unsigned long long), c->stream));
    return ELEVEN_OK;
}

static int resetFilm(ElevenCtx* c) {
    const uint32_t n = c->nPixels;
    k_filmReset<<<(n + 255) / 256, 256, 0, c->stream>>>(c->W);
    if (c->cfg.rng_mode == ELEVEN_RNG_REFERENCE)
        k_rngInit<<<(n + 127) / 128, 128, 0, c->stream>>>(c->W.rng, c->d_seqMat, n);
    CK(cudaMemsetAsync(c->W.stats, 0, ST_COUNT * sizeof(unsigned long long), c->stream));
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->stream));
    c->samplesRendered = 0;
    std::lock_guard<std::mutex> lk(c->statsMutex);
    c->stats.render_ms = 0; c->stats.trace_ms = 0; c->stats.kernel_launches = 0; c->stats.pixel_samples = 0; c->stats.reduce_ms = 0; c->stats.reduce_calls = 0;
    c->stats.extend_ms = c->stats.shade_ms = c->stats.connect_ms = c->stats.other_ms = 0; c->stats.extend_launches = 0;
    return ELEVEN_OK;
}

static void setCameraParams(ElevenCtx* c, const ElevenCamera& cam);

extern "C" int eleven_scene_upload(ElevenCtx* c, const ElevenSceneDesc* d) {
    if (!c || !d) return fail(ELEVEN_ERR_ARG, "eleven_scene_upload: null argument");
    if (d->camera.xRes == 0 || d->camera.yRes == 0) return fail(ELEVEN_ERR_ARG, "camera resolution is zero");
    if (d->triCount && !d->tris) return fail(ELEVEN_ERR_ARG, "tris is null");
    if (d->materialCount == 0 || !d->materials) return fail(ELEVEN_ERR_ARG, "at least one material is required");
    if (d->objectCount == 0 || !d->objectMaterial) return fail(ELEVEN_ERR_ARG, "objectMaterial is required");
    CK(cudaSetDevice(c->cfg.device));
    freeAll(c->sceneAllocs);
    c->d_seqMat = nullptr; c->d_camera = nullptr;
    c->haveScene = false;
    DevScene& S = c->scene; memset(&S, 0, sizeof S);
    S.byteMagic = 0x47000000u;
    int rc;
    // ELEVEN_UPLOAD_TRACE=1: wall-clock split of this call on stderr (host stages are synchronous)
    const bool trace = getenv("ELEVEN_UPLOAD_TRACE") != nullptr;
    auto tNow = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double tLast = tNow();
    auto lap = [&](const char* what) { if (!trace) return; cudaDeviceSynchronize(); const double t = tNow(); fprintf(stderr, "[eleven_scene_upload] %-28s %8.1f ms\n", what, t - tLast); tLast = t; };
    // The environment's CDF is a sequential float running sum (it must reproduce the reference's bits, S/HDRI.hpp:107-128) and the alias
    // table a sequential pairing: host work, ~0.2 s for 4096x2048 texels.  It runs on its own thread while this one feeds the GPU
    // (triangles, BVH8 build, textures).
    std::vector<float> envCdf; std::vector<AliasEntry> envAlias; float envSum = 0.f; std::string envErr;
    struct Joiner { std::thread t; void join() { if (t.joinable()) t.join(); } ~Joiner() { join(); } } envJoin;
    if (!d->hdri.data || d->hdri.width <= 0 || d->hdri.height <= 0 || d->hdri.format != ELEVEN_TEX_F32_RGB) return fail(ELEVEN_ERR_ARG, "hdri must be a non-empty ELEVEN_TEX_F32_RGB texture");
    envJoin.t = std::thread([&envCdf, &envAlias, &envSum, &envErr, d]() {
        const int EW = d->hdri.width, EH = d->hdri.height; const size_t n = (size_t)EW * EH;
        const float* src = (const float*)d->hdri.data;
        auto texel = [&](int x, int y) -> size_t {          // Texture::getValueFromCoordinates addressing, S/Texture.hpp:95-108
            x = (int)(d->hdri.xTile * (x + d->hdri.xOffset * EW)) % EW;
            y = (int)(d->hdri.yTile * (y + d->hdri.yOffset * EH)) % EH;
            long long idx = (long long)y * EW + x; if (idx < 0) idx = 0; if (idx >= (long long)n) idx = (long long)n - 1;
            return (size_t)idx;
        };
        float sum = 0;
        for (int j = 0; j < EH; j++) for (int i = 0; i < EW; i++) { const float* p = src + 3 * texel(i, j); sum += p[0] + p[1] + p[2]; }
        if (!(sum > 0)) { envErr = "environment has zero total radiance (the reference hangs on it, SURVEY F10)"; return; }
        std::vector<float>& cdf = envCdf; cdf.resize(n + 1); cdf[0] = 0; size_t k = 0;
        for (int j = 0; j < EH; j++) for (int i = 0; i < EW; i++) { const float* p = src + 3 * texel(i, j); cdf[k + 1] = cdf[k] + (p[0] + p[1] + p[2]) / sum; k++; }
        envSum = sum;
        // Walker/Vose alias table over P_i = cdf[i+1]-cdf[i]: the texel weights the reference's pdf assumes
        std::vector<AliasEntry>& alias = envAlias; alias.resize(n);
        std::vector<double> p(n); double tot = 0;
        for (size_t i = 0; i < n; i++) { p[i] = std::max(0.0, (double)cdf[i + 1] - (double)cdf[i]); tot += p[i]; }
        std::vector<uint32_t> small, large; small.reserve(n); large.reserve(n);
        for (size_t i = 0; i < n; i++) { p[i] = p[i] * (double)n / tot; (p[i] < 1.0 ? small : large).push_back((uint32_t)i); }
        while (!small.empty() && !large.empty()) {
            const uint32_t s = small.back(), l = large.back(); small.pop_back();
            alias[s].prob = (float)p[s]; alias[s].alias = l;
            p[l] = (p[l] + p[s]) - 1.0;
            if (p[l] < 1.0) { large.pop_back(); small.push_back(l); }
        }
        for (uint32_t i : large) { alias[i].prob = 1.0f; alias[i].alias = i; }
        for (uint32_t i : small) { alias[i].prob = 1.0f; alias[i].alias = i; }
    });


    // triangles -> per-triangle material, BVH8, shading records
    std::vector<int32_t> triMat(d->triCount);
    for (uint32_t i = 0; i < d->triCount; i++) {
        const int32_t o = d->tris[i].objectID;
        if (o < 0 || (uint32_t)o >= d->objectCount) return fail(ELEVEN_ERR_ARG, "triangle objectID out of range");
        const int32_t m = d->objectMaterial[o];
        if (m < 0 || (uint32_t)m >= d->materialCount) return fail(ELEVEN_ERR_ARG, "object material out of range");
        triMat[i] = m;
    }
    if ((rc = devUpload(c->sceneAllocs, &S.triMaterial, triMat.data(), triMat.size()))) return rc;
    const bool deviceBuild = c->cfg.bvh_builder == ELEVEN_BVH_DEVICE && d->triCount > 0;
    if (deviceBuild) {
        // triangles go to the device once (152 B each); boxes, the BVH8, the triangle slots and the shading records are all
        // produced there (bvh8_build_gpu.cuh)
        ElevenTri* d_tris = nullptr;
        cudaError_t e = cudaMalloc((void**)&d_tris, (size_t)d->triCount * sizeof(ElevenTri));
        if (e != cudaSuccess) return fail(ELEVEN_ERR_NOMEM, std::string("cudaMalloc(triangles): ") + cudaGetErrorString(e));
        e = cudaMemcpyAsync(d_tris, d->tris, (size_t)d->triCount * sizeof(ElevenTri), cudaMemcpyHostToDevice, c->stream);
        gpubvh::DeviceBvh db; std::string berr;
        std::vector<PresplitPiece> pieces;                  // sliver triangles enter the build as several references (bvh8_build.cpp)
        const char* ps = getenv("ELEVEN_PRESPLIT");
        bool ok = e == cudaSuccess;
        const auto tPs = std::chrono::steady_clock::now();
        if (ok && (!ps || atoi(ps) != 0)) ok = gpubvh::devicePresplit(d_tris, d->tris, d->triCount, c->stream, pieces, berr);
        const double presplitMs = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tPs).count();
        ok = ok && gpubvh::buildBvh8Device(d_tris, S.triMaterial, d->triCount, pieces, c->stream, c->bvhArena, db, berr);
        float4* st = nullptr;
        if (ok && (rc = devAlloc(c->sceneAllocs, &st, (size_t)d->triCount * 9)) == 0) {
            gpubvh::k_shadeTris<<<(d->triCount + 255) / 256, 256, 0, c->stream>>>(d_tris, d->triCount, (float*)st);
            e = cudaStreamSynchronize(c->stream);
        }
        cudaFree(d_tris);
        if (!ok) return fail(ELEVEN_ERR_CUDA, berr.empty() ? std::string("triangle upload: ") + cudaGetErrorString(e) : berr);
        c->sceneAllocs.push_back(db.nodes); c->sceneAllocs.push_back(db.slots); c->sceneAllocs.push_back(db.nodeSlack);
        if (rc) return rc;
        if (e != cudaSuccess) return fail(ELEVEN_ERR_CUDA, std::string("k_shadeTris: ") + cudaGetErrorString(e));
        if (db.maxDepth >= EL_STACK) return fail(ELEVEN_ERR_UNSUPPORTED, "BVH deeper than the traversal stack");
        S.nodes = db.nodes; S.slots = db.slots; S.nodeSlack = db.nodeSlack; S.shadeTris = st;
        S.nodeCount = db.nodeCount; S.triCount = d->triCount; S.keySlack = db.keySlack;
        c->stats.bvh_build_ms = db.buildMs + presplitMs; c->stats.bvh_nodes = db.nodeCount; c->stats.bvh_tri_slots = db.slotCount; c->stats.key_slack = db.keySlack;
    } else {
        Bvh8 bvh;
        buildBvh8(d->tris, d->triCount, triMat.data(), bvh, (int)std::max(1u, std::thread::hardware_concurrency()));
        if (bvh.maxDepth >= EL_STACK) return fail(ELEVEN_ERR_UNSUPPORTED, "BVH deeper than the traversal stack");
        c->stats.bvh_build_ms = bvh.buildMs; c->stats.bvh_nodes = (uint32_t)bvh.nodes.size(); c->stats.bvh_tri_slots = (uint32_t)bvh.slots.size();
        c->stats.key_slack = bvh.keySlack;
        if ((rc = devUpload(c->sceneAllocs, &S.nodes, (const float4*)bvh.nodes.data(), bvh.nodes.size() * 5))) return rc;
        if ((rc = devUpload(c->sceneAllocs, &S.slots, (const float4*)bvh.slots.data(), bvh.slots.size() * 3))) return rc;
        if ((rc = devUpload(c->sceneAllocs, &S.nodeSlack, bvh.nodeSlack.data(), bvh.nodeSlack.size()))) return rc;
        S.nodeCount = d->triCount ? (uint32_t)bvh.nodes.size() : 0u; S.triCount = d->triCount; S.keySlack = bvh.keySlack;
        std::vector<float> st((size_t)d->triCount * 36);
        for (uint32_t i = 0; i < d->triCount; i++) {
            const ElevenTri& T = d->tris[i]; float* o = &st[(size_t)i * 36];
            memcpy(o, T.vertices, 36); memcpy(o + 9, T.normals, 36); memcpy(o + 18, T.tangents, 36);
            o[27] = T.tangentsSign;
            o[28] = T.uv[0][0]; o[29] = T.uv[0][1]; o[30] = T.uv[1][0]; o[31] = T.uv[1][1]; o[32] = T.uv[2][0]; o[33] = T.uv[2][1];
            memcpy(o + 34, &T.objectID, 4); o[35] = 0.f;
        }
        if ((rc = devUpload(c->sceneAllocs, &S.shadeTris, (const float4*)st.data(), (size_t)d->triCount * 9))) return rc;
    }
    lap("triangles + BVH8");
    if ((rc = devUpload(c->sceneAllocs, &S.objectMaterial, d->objectMaterial, d->objectCount))) return rc;
    static_assert(sizeof(DevMaterial) == sizeof(ElevenMaterial), "material layout");
    for (uint32_t i = 0; i < d->materialCount; i++) {
        const ElevenMaterial& m = d->materials[i];
        const int32_t ids[5] = {m.albedoTextureID, m.emissionTextureID, m.roughnessTextureID, m.metallicTextureID, m.normalTextureID};
        for (int k = 0; k < 5; k++) if (ids[k] >= (int32_t)d->textureCount) return fail(ELEVEN_ERR_ARG, "material texture id out of range");
    }
    if ((rc = devUpload(c->sceneAllocs, &S.materials, (const DevMaterial*)d->materials, d->materialCount))) return rc;
    std::vector<DevTex> texs(d->textureCount);
    for (uint32_t i = 0; i < d->textureCount; i++) if ((rc = uploadTexture(c, d->textures[i], texs[i], false))) return rc;
    if ((rc = devUpload(c->sceneAllocs, &S.textures, texs.data(), texs.size()))) return rc;
    lap("material textures");
    {   // packed map records (DevPackedMaps): materials whose four maps are 8-bit, unfiltered and congruent
        std::vector<DevPackedMaps> packed(d->materialCount);
        memset(packed.data(), 0, packed.size() * sizeof(DevPackedMaps));
        for (uint32_t i = 0; i < d->materialCount && !getenv("ELEVEN_NO_PACKED_MAPS"); i++) {
            const ElevenMaterial& m = d->materials[i];
            const int32_t ids[4] = {m.albedoTextureID, m.roughnessTextureID, m.metallicTextureID, m.normalTextureID};
            bool ok = true;
            for (int k = 0; k < 4 && ok; k++) {
                if (ids[k] < 0) { ok = false; break; }
                const ElevenTexture& t = d->textures[ids[k]], &t0 = d->textures[ids[0]];
                ok = t.format != ELEVEN_TEX_F32_RGB && (k == 3 || t.filter == 0) && t.width == t0.width && t.height == t0.height &&
                     t.xTile == t0.xTile && t.yTile == t0.yTile && t.xOffset == t0.xOffset && t.yOffset == t0.yOffset;
            }
            if (!ok) continue;
            const ElevenTexture& t0 = d->textures[ids[0]];
            const size_t n = (size_t)t0.width * t0.height;
            uint2* dp = nullptr;
            if ((rc = devAlloc(c->sceneAllocs, &dp, n))) return rc;
            k_packMaps<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>((const uchar4*)texs[ids[0]].data, (const uchar4*)texs[ids[1]].data,
                                                                             (const uchar4*)texs[ids[2]].data, (const uchar4*)texs[ids[3]].data, dp, n);
            CK(cudaGetLastError());
            DevPackedMaps& P = packed[i];
            P.data = dp; P.width = t0.width; P.height = t0.height; P.xTile = t0.xTile; P.yTile = t0.yTile; P.xOffset = t0.xOffset; P.yOffset = t0.yOffset;
            P.albedoFormat = d->textures[ids[0]].format; P.roughFormat = d->textures[ids[1]].format; P.metalFormat = d->textures[ids[2]].format; P.normalFormat = d->textures[ids[3]].format;
        }
        if ((rc = devUpload(c->sceneAllocs, &S.packed, packed.data(), packed.size()))) return rc;
    }
    lap("packed material maps");
    float lut[512]; buildLut(lut);
    if ((rc = devUpload(c->sceneAllocs, &S.lut, lut, 512))) return rc;

    // environment: float4 texels + the reference's float CDF (S/HDRI.hpp:107-128) + alias table over its increments (built by envWorker)
    if ((rc = uploadTexture(c, d->hdri, S.hdri, true))) return rc;
    envJoin.join();
    if (!envErr.empty()) return fail(ELEVEN_ERR_ARG, envErr);
    S.radianceSum = envSum;
    if ((rc = devUpload(c->sceneAllocs, &S.cdf, envCdf.data(), envCdf.size()))) return rc;
    if ((rc = devUpload(c->sceneAllocs, &S.alias, envAlias.data(), envAlias.size()))) return rc;
    lap("environment + CDF + alias");
    S.lightCount = d->pointLightCount;
    if ((rc = devUpload(c->sceneAllocs, &S.lights, (const float*)d->pointLights, (size_t)d->pointLightCount * 6))) return rc;
    static_assert(sizeof(DevCamera) == sizeof(ElevenCamera), "camera layout");
    memcpy(&S.cam, &d->camera, sizeof(DevCamera));

    c->width = d->camera.xRes; c->height = d->camera.yRes;
    const uint32_t nPix = c->width * c->height;
    if (nPix != c->nPixels || !c->W.ray) { c->nPixels = nPix; if ((rc = allocWave(c))) return rc; }
    lap("wave state allocation");
    c->W.sceneHasEmission = 0u;
    for (uint32_t i = 0; i < d->materialCount; i++) {
        const ElevenMaterial& m = d->materials[i];
        if (m.emissionTextureID >= 0 || m.emission[0] != 0.f || m.emission[1] != 0.f || m.emission[2] != 0.f) c->W.sceneHasEmission = 1u;
    }
    if (c->cfg.rng_mode == ELEVEN_RNG_REFERENCE && !c->d_seqMat) {
        std::vector<uint32_t> flat; buildSeqMats(flat, 32);
        const uint32_t* p = nullptr;
        if ((rc = devUpload(c->sceneAllocs, &p, flat.data(), flat.size()))) return rc;
        c->d_seqMat = (uint32_t*)p;
    }
    // per-render constants
    RenderParams& P = c->P; memset(&P, 0, sizeof P);
    P.rngMode = c->cfg.rng_mode; P.envMode = c->cfg.env_mode; P.hitMode = c->cfg.hit_mode; P.maxBounces = c->cfg.max_bounces; P.flags = c->cfg.flags;
    P.seedLo = (uint32_t)c->cfg.seed; P.seedHi = (uint32_t)(c->cfg.seed >> 32);
    setCameraParams(c, d->camera);
    { const ElevenCamera* dc = nullptr; if ((rc = devUpload(c->sceneAllocs, &dc, &d->camera, 1))) return rc; c->d_camera = (ElevenCamera*)dc; }
    c->haveScene = true;
    return resetFilm(c);
}

static void setCameraParams(ElevenCtx* c, const ElevenCamera& cam) {
    memcpy(&c->scene.cam, &cam, sizeof(DevCamera));
    // rotation *= PI/180.0 in float, then sin/cos (S/kernel.cu:299-306)
    const float k = (float)((double)EL_PI / 180.0);
    const float rx = cam.rotation[0] * k, ry = cam.rotation[1] * k, rz = cam.rotation[2] * k;
    RenderParams& P = c->P;
    P.rot.sx = sinf(rx); P.rot.cx = cosf(rx); P.rot.sy = sinf(ry); P.rot.cy = cosf(ry); P.rot.sz = sinf(rz); P.rot.cz = cosf(rz);
}

extern "C" int eleven_set_camera(ElevenCtx* c, const ElevenCamera* cam) {
    if (!c || !cam) return fail(ELEVEN_ERR_ARG, "eleven_set_camera: null argument");
    if (!c->haveScene) return fail(ELEVEN_ERR_STATE, "eleven_set_camera: no scene uploaded");
    if (cam->xRes != c->width || cam->yRes != c->height) return fail(ELEVEN_ERR_ARG, "eleven_set_camera: resolution change needs eleven_scene_upload");
    CK(cudaSetDevice(c->cfg.device));
    setCameraParams(c, *cam);
    // the camera travels to the device as a kernel argument (__grid_constant__); mirror it in device memory too so that
    // the copy is observable (and so that later device-side consumers can read it)
    CK(cudaMemcpyAsync(c->d_camera, cam, sizeof(ElevenCamera), cudaMemcpyHostToDevice, c->stream));
    return ELEVEN_OK;
}

extern "C" int eleven_film_reset(ElevenCtx* c) {
    if (!c || !c->haveScene) return fail(ELEVEN_ERR_STATE, "eleven_film_reset: no scene uploaded");
    CK(cudaSetDevice(c->cfg.device));
    return resetFilm(c);
}

// Persistent grids are sized to what is RESIDENT: SMs x the CTAs per SM the kernel's registers / shared memory admit (occupancy
// query, cached per kernel).  A larger grid only adds CTAs that wait for a slot, fill their mask tables and find the queue empty.
// ELEVEN_GRID_EXTEND / ELEVEN_GRID_SHADOW override the CTAs per SM (tuning knobs).
template <typename K>
static int persistentGrid(ElevenCtx* c, K kernel, const char* envOverride) {
    static std::mutex m; static std::vector<std::pair<const void*, int>> cache;
    if (envOverride) if (const char* e = getenv(envOverride)) return c->numSMs * std::max(1, atoi(e));
    std::lock_guard<std::mutex> lk(m);
    for (auto& kv : cache) if (kv.first == (const void*)kernel) return c->numSMs * kv.second;
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, 128, 0) != cudaSuccess || n < 1) { cudaGetLastError(); n = 8; }
    cache.push_back({(const void*)kernel, n});
    return c->numSMs * n;
}
#define LAUNCH_PERSISTENT_ON(st, kernel, env, ...) kernel<<<persistentGrid(c, kernel, env), 128, 0, st>>>(__VA_ARGS__)
#define LAUNCH_PERSISTENT(kernel, env, ...) LAUNCH_PERSISTENT_ON(c->stream, kernel, env, __VA_ARGS__)

template <bool COUNT, bool FM>
static void launchExtend(ElevenCtx* c) {
    if (c->cfg.hit_mode == ELEVEN_HIT_KEY) LAUNCH_PERSISTENT((k_extend<TRACE_CLOSEST_KEY, COUNT, FM>), "ELEVEN_GRID_EXTEND", c->W, c->scene);
    else LAUNCH_PERSISTENT((k_extend<TRACE_CLOSEST_T, COUNT, FM>), "ELEVEN_GRID_EXTEND", c->W, c->scene);
}
// shadow stages of one bounce; the last one also performs the MIS combination (kernels_trace.cuh).  Returns launches.
template <bool COUNT, bool FM>
static int launchConnect(ElevenCtx* c, cudaStream_t st) {
    if (c->scene.lightCount == 0) { LAUNCH_PERSISTENT_ON(st, (k_shadowEnv<false, COUNT, FM>), "ELEVEN_GRID_SHADOW", c->W, c->scene); return 1; }
    LAUNCH_PERSISTENT_ON(st, (k_shadowEnv<true, COUNT, FM>), "ELEVEN_GRID_SHADOW", c->W, c->scene);
    if (c->cfg.hit_mode == ELEVEN_HIT_KEY && !(c->cfg.flags & ELEVEN_FLAG_ANYHIT_LIGHT_SHADOWS)) LAUNCH_PERSISTENT_ON(st, (k_shadowLight<ELEVEN_HIT_KEY, COUNT, FM>), "ELEVEN_GRID_SHADOW", c->W, c->scene);
    else LAUNCH_PERSISTENT_ON(st, (k_shadowLight<ELEVEN_HIT_MIN_T, COUNT, FM>), "ELEVEN_GRID_SHADOW", c->W, c->scene);
    return 2;
}
template <bool FM>
static void launchExtendRt(ElevenCtx* c, bool count) { if (count) launchExtend<true, FM>(c); else launchExtend<false, FM>(c); }
template <bool FM>
static int launchConnectRt(ElevenCtx* c, bool count, cudaStream_t st) { return count ? launchConnect<true, FM>(c, st) : launchConnect<false, FM>(c, st); }

extern "C" int eleven_render(ElevenCtx* c, int spp) {
    if (!c) return fail(ELEVEN_ERR_ARG, "eleven_render: null context");
    if (!c->haveScene) return fail(ELEVEN_ERR_STATE, "eleven_render: no scene uploaded");
    if (spp < 0) return fail(ELEVEN_ERR_ARG, "eleven_render: negative sample count");
    CK(cudaSetDevice(c->cfg.device));
    const uint32_t n = c->nPixels;
    const int gridPix = (int)((n + 255) / 256);
    c->P.sampleStride = c->cfg.sample_stride;
    const bool count = (c->cfg.flags & ELEVEN_FLAG_COUNTERS) != 0;
    const bool timeK = (c->cfg.flags & ELEVEN_FLAG_TIME_KERNELS) != 0;
    const bool fastMath = (c->cfg.flags & ELEVEN_FLAG_FAST_MATH) != 0;
    // stage timing: one event after every stage launch; stage i spans events [i, i+1)
    std::vector<int> evKind;                      // 0 extend, 1 shade, 2 connect, 3 other
    size_t evUsed = 0;
    auto mark = [&](int kind) {
        if (!timeK) return;
        if (evUsed == c->evPool.size()) { cudaEvent_t e; cudaEventCreate(&e); c->evPool.push_back(e); }
        cudaEventRecord(c->evPool[evUsed++], c->stream); evKind.push_back(kind);
    };
    CK(cudaEventRecord(c->ev0, c->stream));
    mark(3);
    for (int s = 0; s < spp;) {
        // this wave: the largest power of two of samples that fits both the wave buffers and what is left to render
        uint32_t logK = c->maxLogK;
        while (logK > 0 && (1 << logK) > spp - s) logK--;
        c->P.logK = logK;
        const int gridPaths = (int)((((size_t)n << logK) + 255) / 256);
        c->P.sampleIndex = c->cfg.sample_offset + c->samplesRendered * c->cfg.sample_stride;
        if (fastMath) k_raygen<true><<<gridPaths, 256, 0, c->stream>>>(c->W, c->scene, c->P);
        else k_raygen<false><<<gridPaths, 256, 0, c->stream>>>(c->W, c->scene, c->P);
        mark(3);
        // Stage overlap: the shadow stage of bounce b (reads the NEE records and queue, writes throughput / radiance) and extend +
        // classify of bounce b+1 (read the new rays, write hits and the shading queue) touch disjoint data, so the shadow stage runs
        // on the aux stream while the render stream goes on; they join in front of shade(b+1), which needs the throughput and
        // re-uses the NEE queue.  Every traversal kernel is a persistent grid whose last rays leave most of the machine idle for
        // 100-200 us: with two kernels in flight the one kernel's tail is filled by the other's CTAs.  Off with stage timing
        // (the events would time overlapped stages) and with ELEVEN_OVERLAP=0.
        const bool overlap = c->overlap && !timeK;
        bool joinPending = false;
        for (uint32_t b = 0; b < c->cfg.max_bounces; b++) {
            if (fastMath) launchExtendRt<true>(c, count); else launchExtendRt<false>(c, count);
            mark(0);
            LAUNCH_PERSISTENT(k_classify, nullptr, c->W, c->scene);
            if (joinPending) { CK(cudaStreamWaitEvent(c->stream, c->evJoin, 0)); k_advance<<<1, 32, 0, c->stream>>>(c->W, c->scene.lightCount, 3); joinPending = false; }
            if (fastMath) LAUNCH_PERSISTENT(k_shade<true>, nullptr, c->W, c->scene, c->P);
            else LAUNCH_PERSISTENT(k_shade<false>, nullptr, c->W, c->scene, c->P);
            mark(1);
            k_advance<<<1, 32, 0, c->stream>>>(c->W, c->scene.lightCount, 0);
            mark(3);
            if (overlap) {
                CK(cudaEventRecord(c->evFork, c->stream));
                CK(cudaStreamWaitEvent(c->auxStream, c->evFork, 0));
                c->stats.kernel_launches += fastMath ? launchConnectRt<true>(c, count, c->auxStream) : launchConnectRt<false>(c, count, c->auxStream);
                CK(cudaEventRecord(c->evJoin, c->auxStream));
                joinPending = true;
                k_advance<<<1, 32, 0, c->stream>>>(c->W, c->scene.lightCount, 2);
            } else {
                c->stats.kernel_launches += fastMath ? launchConnectRt<true>(c, count, c->stream) : launchConnectRt<false>(c, count, c->stream);
                mark(2);
                k_advance<<<1, 32, 0, c->stream>>>(c->W, c->scene.lightCount, 1);
                mark(3);
            }
            c->stats.extend_launches += 1;
            std::swap(c->W.qCur, c->W.qNext);      // kernel arguments are captured at launch: the next bounce reads the list just written
            c->stats.kernel_launches += 5;
        }
        if (joinPending) { CK(cudaStreamWaitEvent(c->stream, c->evJoin, 0)); k_advance<<<1, 32, 0, c->stream>>>(c->W, c->scene.lightCount, 3); }
        k_accumulate<<<gridPaths, 256, 0, c->stream>>>(c->W, logK);
        mark(3);
        c->stats.kernel_launches += 2;
        c->samplesRendered += 1u << logK;
        s += 1 << logK;
    }
    CK(cudaEventRecord(c->ev1, c->stream));
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->stream));
    float ms = 0; CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    std::lock_guard<std::mutex> lk(c->statsMutex);
    c->stats.render_ms += ms;
    for (size_t i = 1; i < evUsed; i++) {
        float t = 0; CK(cudaEventElapsedTime(&t, c->evPool[i - 1], c->evPool[i]));
        double* dst[4] = {&c->stats.extend_ms, &c->stats.shade_ms, &c->stats.connect_ms, &c->stats.other_ms};
        *dst[evKind[i]] += t;
    }
    c->stats.pixel_samples += (uint64_t)n * (uint64_t)spp;
    return ELEVEN_OK;
}

static int passIndex(int pass) {               // position of a pass in the contiguous film block
    switch (pass) {
        case ELEVEN_PASS_BEAUTY: return 0;
        case ELEVEN_PASS_NORMAL: return 1;
        case ELEVEN_PASS_TANGENT: return 2;
        case ELEVEN_PASS_BITANGENT: return 3;
        default: return -1;
    }
}
static float4* filmPass(ElevenCtx* c, int pass) {
    const int k = passIndex(pass);
    return k < 0 ? nullptr : c->W.filmBeauty + (size_t)k * c->nPixels;
}

// ---- film read-back: on the SNAPSHOT stream, concurrent with a running eleven_render (see the header) --------------------
enum { SRC_LOCAL = 0, SRC_REDUCED = 1 };
static int filmSource(ElevenCtx* c, int pass, int source, const char* who, float4** src) {
    if (!c->haveScene) return fail(ELEVEN_ERR_STATE, std::string(who) + ": no scene uploaded");
    const int k = passIndex(pass);
    if (k < 0) return fail(ELEVEN_ERR_UNSUPPORTED, std::string(who) + ": pass not produced on the device (DENOISE is a host-side OIDN pass in the reference)");
    if (source == SRC_LOCAL) { *src = c->W.filmBeauty + (size_t)k * c->nPixels; return ELEVEN_OK; }
    if (!c->d_reduced) return fail(ELEVEN_ERR_STATE, std::string(who) + ": no reduced film (call eleven_reduce_film; only the root holds one)");
    if (k >= c->reducedPasses) return fail(ELEVEN_ERR_STATE, std::string(who) + ": this pass was not part of the last eleven_reduce_film (all_passes = 0)");
    *src = c->d_reduced + (size_t)k * c->nPixels; return ELEVEN_OK;
}
static int getFilm(ElevenCtx* c, int pass, float* rgba, size_t nPixels, int source, const char* who) {
    if (!c || !rgba) return fail(ELEVEN_ERR_ARG, std::string(who) + ": null argument");
    float4* src = nullptr;
    if (int rc = filmSource(c, pass, source, who, &src)) return rc;
    if (nPixels != c->nPixels) return fail(ELEVEN_ERR_ARG, std::string(who) + ": n_pixels must be W*H");
    CK(cudaSetDevice(c->cfg.device));
    std::lock_guard<std::mutex> lk(c->snapMutex);
    k_resolve<<<(c->nPixels + 255) / 256, 256, 0, c->snapStream>>>(src, c->d_resolve, c->nPixels);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(rgba, c->d_resolve, (size_t)c->nPixels * 16, cudaMemcpyDeviceToHost, c->snapStream));
    CK(cudaStreamSynchronize(c->snapStream));
    return ELEVEN_OK;
}
static int resolve8(ElevenCtx* c, int pass, uint8_t* rgba8, size_t nPixels, int source, const char* who) {
    if (!c || !rgba8) return fail(ELEVEN_ERR_ARG, std::string(who) + ": null argument");
    float4* src = nullptr;
    if (int rc = filmSource(c, pass, source, who, &src)) return rc;
    if (nPixels != c->nPixels) return fail(ELEVEN_ERR_ARG, std::string(who) + ": n_pixels must be W*H");
    CK(cudaSetDevice(c->cfg.device));
    std::lock_guard<std::mutex> lk(c->snapMutex);
    k_resolve8<<<(c->nPixels + 255) / 256, 256, 0, c->snapStream>>>(src, (uchar4*)c->d_resolve, c->nPixels);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(rgba8, c->d_resolve, (size_t)c->nPixels * 4, cudaMemcpyDeviceToHost, c->snapStream));
    CK(cudaStreamSynchronize(c->snapStream));
    return ELEVEN_OK;
}
static int sampleCounts(ElevenCtx* c, uint32_t* out, size_t nPixels, int source, const char* who) {
    if (!c || !out) return fail(ELEVEN_ERR_ARG, std::string(who) + ": null argument");
    float4* src = nullptr;
    if (int rc = filmSource(c, ELEVEN_PASS_BEAUTY, source, who, &src)) return rc;
    if (nPixels != c->nPixels) return fail(ELEVEN_ERR_ARG, std::string(who) + ": n_pixels must be W*H");
    CK(cudaSetDevice(c->cfg.device));
    std::lock_guard<std::mutex> lk(c->snapMutex);
    k_sampleCounts<<<(c->nPixels + 255) / 256, 256, 0, c->snapStream>>>(src, (uint32_t*)c->d_resolve, c->nPixels);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, c->d_resolve, (size_t)c->nPixels * 4, cudaMemcpyDeviceToHost, c->snapStream));
    CK(cudaStreamSynchronize(c->snapStream));
    return ELEVEN_OK;
}
extern "C" int eleven_get_film(ElevenCtx* c, int pass, float* rgba, size_t n) { return getFilm(c, pass, rgba, n, SRC_LOCAL, "eleven_get_film"); }
extern "C" int eleven_get_film_reduced(ElevenCtx* c, int pass, float* rgba, size_t n) { return getFilm(c, pass, rgba, n, SRC_REDUCED, "eleven_get_film_reduced"); }
extern "C" int eleven_resolve_rgba8(ElevenCtx* c, int pass, uint8_t* o, size_t n) { return resolve8(c, pass, o, n, SRC_LOCAL, "eleven_resolve_rgba8"); }
extern "C" int eleven_resolve_rgba8_reduced(ElevenCtx* c, int pass, uint8_t* o, size_t n) { return resolve8(c, pass, o, n, SRC_REDUCED, "eleven_resolve_rgba8_reduced"); }
extern "C" int eleven_get_sample_counts(ElevenCtx* c, uint32_t* o, size_t n) { return sampleCounts(c, o, n, SRC_LOCAL, "eleven_get_sample_counts"); }
extern "C" int eleven_get_sample_counts_reduced(ElevenCtx* c, uint32_t* o, size_t n) { return sampleCounts(c, o, n, SRC_REDUCED, "eleven_get_sample_counts_reduced"); }

extern "C" int eleven_get_pathcount(ElevenCtx* c, int32_t* out, size_t nPixels) {
    if (!c || !out) return fail(ELEVEN_ERR_ARG, "eleven_get_pathcount: null argument");
    if (!c->haveScene) return fail(ELEVEN_ERR_STATE, "eleven_get_pathcount: no scene uploaded");
    if (nPixels != c->nPixels) return fail(ELEVEN_ERR_ARG, "eleven_get_pathcount: n_pixels must be W*H");
    CK(cudaSetDevice(c->cfg.device));
    std::lock_guard<std::mutex> lk(c->snapMutex);
    CK(cudaMemcpyAsync(out, c->W.pathCount, (size_t)c->nPixels * 4, cudaMemcpyDeviceToHost, c->snapStream));
    CK(cudaStreamSynchronize(c->snapStream));
    return ELEVEN_OK;
}

extern "C" int eleven_get_samples(ElevenCtx* c) {
    if (!c || !c->haveScene) return fail(ELEVEN_ERR_STATE, "eleven_get_samples: no scene uploaded");
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (cudaSetDevice(c->cfg.device) != cudaSuccess) return fail(ELEVEN_ERR_CUDA, "cudaSetDevice");
    std::lock_guard<std::mutex> lk(c->snapMutex);
    if (cudaMemcpyAsync(&v, c->W.filmBeauty, 16, cudaMemcpyDeviceToHost, c->snapStream) != cudaSuccess) return fail(ELEVEN_ERR_CUDA, "cudaMemcpyAsync");
    if (cudaStreamSynchronize(c->snapStream) != cudaSuccess) return fail(ELEVEN_ERR_CUDA, "cudaStreamSynchronize");
    return (int)v.w;
}

extern "C" int eleven_get_stats(ElevenCtx* c, ElevenStats* out) {
    if (!c || !out) return fail(ELEVEN_ERR_ARG, "eleven_get_stats: null argument");
    if (c->haveScene) {
        CK(cudaSetDevice(c->cfg.device));
        unsigned long long st[ST_COUNT];
        std::vector<uint32_t> pc(c->nPixels);
        float4 f0 = make_float4(0.f, 0.f, 0.f, 0.f);
        {
            std::lock_guard<std::mutex> lk(c->snapMutex);
            CK(cudaMemcpyAsync(st, c->W.stats, sizeof st, cudaMemcpyDeviceToHost, c->snapStream));
            CK(cudaMemcpyAsync(pc.data(), c->W.pathCount, (size_t)c->nPixels * 4, cudaMemcpyDeviceToHost, c->snapStream));
            CK(cudaMemcpyAsync(&f0, c->W.filmBeauty, 16, cudaMemcpyDeviceToHost, c->snapStream));
            CK(cudaStreamSynchronize(c->snapStream));
        }
        uint64_t hb = 0; for (uint32_t v : pc) hb += v;
        std::lock_guard<std::mutex> lk(c->statsMutex);
        c->stats.rays_extension = st[ST_RAYS_EXT]; c->stats.rays_shadow_env = st[ST_RAYS_ENV]; c->stats.rays_shadow_light = st[ST_RAYS_LIGHT];
        c->stats.nodes_visited = st[ST_NODES]; c->stats.tris_tested = st[ST_TRIS]; c->stats.key_evals = st[ST_KEYS];
        c->stats.nodes_visited_extend = st[ST_NODES_EXT]; c->stats.tris_tested_extend = st[ST_TRIS_EXT];
        c->stats.hit_bounces = hb; c->stats.samples_done = (uint32_t)f0.w;
        *out = c->stats;
        return ELEVEN_OK;
    }
    std::lock_guard<std::mutex> lk(c->statsMutex);
    *out = c->stats;
    return ELEVEN_OK;
}

// ---- closest-hit test hook / bench ---------------------------------------------------------------------------------
extern "C" int eleven_trace_device(ElevenCtx* c, const float* d_rays, size_t n, ElevenHit* d_hits, int anyHit, float* ms) {
    if (!c || !d_rays || !d_hits) return fail(ELEVEN_ERR_ARG, "eleven_trace_device: null argument");
    if (!c->haveScene) return fail(ELEVEN_ERR_STATE, "eleven_trace_device: no scene uploaded");
    if (n > 0xfffffff0ull) return fail(ELEVEN_ERR_ARG, "eleven_trace_device: too many rays");
    CK(cudaSetDevice(c->cfg.device));
    const bool count = (c->cfg.flags & ELEVEN_FLAG_COUNTERS) != 0;
    CK(cudaMemsetAsync(c->d_workCounter, 0, 4, c->stream));
    CK(cudaEventRecord(c->ev0, c->stream));
    const uint32_t nn = (uint32_t)n;
#define LAUNCH_BATCH(MODE, CNT) LAUNCH_PERSISTENT((k_traceBatch<MODE, CNT>), nullptr, d_rays, nn, d_hits, c->scene, c->d_workCounter, c->W.stats)   /* grid = resident CTAs, like the render kernels */
    if (anyHit) { if (count) LAUNCH_BATCH(TRACE_ANY, true); else LAUNCH_BATCH(TRACE_ANY, false); }
    else if (c->cfg.hit_mode == ELEVEN_HIT_KEY) { if (count) LAUNCH_BATCH(TRACE_CLOSEST_KEY, true); else LAUNCH_BATCH(TRACE_CLOSEST_KEY, false); }
    else { if (count) LAUNCH_BATCH(TRACE_CLOSEST_T, true); else LAUNCH_BATCH(TRACE_CLOSEST_T, false); }
#undef LAUNCH_BATCH
    CK(cudaEventRecord(c->ev1, c->stream));
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->stream));
    float t = 0; CK(cudaEventElapsedTime(&t, c->ev0, c->ev1));
    c->stats.trace_ms += t; c->stats.kernel_launches += 1;
    if (ms) *ms = t;
    return ELEVEN_OK;
}

extern "C" int eleven_trace_closest(ElevenCtx* c, const float* rays, size_t n, ElevenHit* hits) {
    if (!c || (n && (!rays || !hits))) return fail(ELEVEN_ERR_ARG, "eleven_trace_closest: null argument");
    if (!c->haveScene) return fail(ELEVEN_ERR_STATE, "eleven_trace_closest: no scene uploaded");
    if (n == 0) return ELEVEN_OK;
    CK(cudaSetDevice(c->cfg.device));
    float* d_rays = nullptr; ElevenHit* d_hits = nullptr;
    CK(cudaMalloc(&d_rays, n * 24));
    cudaError_t e = cudaMalloc(&d_hits, n * sizeof(ElevenHit));
    if (e != cudaSuccess) { cudaFree(d_rays); return fail(ELEVEN_ERR_NOMEM, "cudaMalloc hits"); }
    int rc = ELEVEN_OK;
    if (cudaMemcpyAsync(d_rays, rays, n * 24, cudaMemcpyHostToDevice, c->stream) != cudaSuccess) rc = fail(ELEVEN_ERR_CUDA, "H2D rays");
    if (!rc) rc = eleven_trace_device(c, d_rays, n, d_hits, 0, nullptr);
    if (!rc && cudaMemcpyAsync(hits, d_hits, n * sizeof(ElevenHit), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) rc = fail(ELEVEN_ERR_CUDA, "D2H hits");
    if (!rc && cudaStreamSynchronize(c->stream) != cudaSuccess) rc = fail(ELEVEN_ERR_CUDA, "sync");
    cudaFree(d_rays); cudaFree(d_hits);
    return rc;
}

// Test hook: the acceleration structure as it lives on the device (Node8 80 B, TriSlot 48 B, per-node slack), so that
// tests can validate the tree of either builder on the host.
extern "C" int eleven_bvh_download(ElevenCtx* c, void* nodes, size_t nodeCap, void* slots, size_t slotCap, float* nodeSlack) {
    if (!c) return fail(ELEVEN_ERR_ARG, "eleven_bvh_download: null context");
    if (!c->haveScene) return fail(ELEVEN_ERR_STATE, "eleven_bvh_download: no scene uploaded");
    if (nodeCap < c->stats.bvh_nodes || slotCap < c->stats.bvh_tri_slots) return fail(ELEVEN_ERR_ARG, "eleven_bvh_download: buffers too small (see ElevenStats bvh_nodes / bvh_tri_slots)");
    CK(cudaSetDevice(c->cfg.device));
    if (nodes) CK(cudaMemcpyAsync(nodes, c->scene.nodes, (size_t)c->stats.bvh_nodes * 80, cudaMemcpyDeviceToHost, c->stream));
    if (slots) CK(cudaMemcpyAsync(slots, c->scene.slots, (size_t)c->stats.bvh_tri_slots * 48, cudaMemcpyDeviceToHost, c->stream));
    if (nodeSlack) CK(cudaMemcpyAsync(nodeSlack, c->scene.nodeSlack, (size_t)c->stats.bvh_nodes * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return ELEVEN_OK;
}

// Test hook: the HOST builder on its own (bvh8_build.cpp; replaces BVH::build, S/BVH.hpp:187-330).  Needs no GPU and no context,
// so the tree layout the kernels walk is checked by the CPU test-suite too.  counts = {nodes, slots}; returns ELEVEN_ERR_ARG with
// the counts filled in when a buffer is too small.
extern "C" int eleven_bvh_build_host(const ElevenTri* tris, uint32_t n, const int32_t* triMaterial, int threads,
                                     void* nodes, size_t nodeCap, void* slots, size_t slotCap, float* nodeSlack,
                                     uint32_t* counts, float* keySlack) {
    if ((!tris && n) || !counts) return fail(ELEVEN_ERR_ARG, "eleven_bvh_build_host: null argument");
    Bvh8 bvh;
    buildBvh8(tris, n, triMaterial, bvh, threads > 0 ? threads : (int)std::max(1u, std::thread::hardware_concurrency()));
    counts[0] = (uint32_t)bvh.nodes.size(); counts[1] = (uint32_t)bvh.slots.size();
    if (keySlack) *keySlack = bvh.keySlack;
    if ((nodes && nodeCap < bvh.nodes.size()) || (slots && slotCap < bvh.slots.size()) || (nodeSlack && nodeCap < bvh.nodes.size()))
        return fail(ELEVEN_ERR_ARG, "eleven_bvh_build_host: buffers too small (counts returned)");
    if (nodes) memcpy(nodes, bvh.nodes.data(), bvh.nodes.size() * sizeof(Node8));
    if (slots) memcpy(slots, bvh.slots.data(), bvh.slots.size() * sizeof(TriSlot));
    if (nodeSlack) memcpy(nodeSlack, bvh.nodeSlack.data(), bvh.nodeSlack.size() * sizeof(float));
    return ELEVEN_OK;
}

// ---- device plumbing for the one-process-per-GPU driver ------------------------------------------------------------------
extern "C" int eleven_film_sums_device(ElevenCtx* c, int pass, void** p, size_t* nFloats) {
    if (!c || !p || !nFloats) return fail(ELEVEN_ERR_ARG, "eleven_film_sums_device: null argument");
    if (!c->haveScene) return fail(ELEVEN_ERR_STATE, "eleven_film_sums_device: no scene uploaded");
    float4* src = filmPass(c, pass);
    if (!src) return fail(ELEVEN_ERR_UNSUPPORTED, "eleven_film_sums_device: pass not available");
    *p = src; *nFloats = (size_t)c->nPixels * 4; return ELEVEN_OK;
}
extern "C" int eleven_device_alloc(ElevenCtx* c, size_t bytes, void** p) {
    if (!c || !p) return fail(ELEVEN_ERR_ARG, "eleven_device_alloc: null argument");
    CK(cudaSetDevice(c->cfg.device));
    cudaError_t e = cudaMalloc(p, std::max<size_t>(bytes, 1));
    if (e != cudaSuccess) return fail(ELEVEN_ERR_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    return ELEVEN_OK;
}
extern "C" int eleven_host_alloc(ElevenCtx* c, size_t bytes, void** p) {
    if (!c || !p) return fail(ELEVEN_ERR_ARG, "eleven_host_alloc: null argument");
    CK(cudaSetDevice(c->cfg.device));
    cudaError_t e = cudaMallocHost(p, std::max<size_t>(bytes, 1));
    if (e != cudaSuccess) return fail(ELEVEN_ERR_NOMEM, std::string("cudaMallocHost: ") + cudaGetErrorString(e));
    return ELEVEN_OK;
}
extern "C" int eleven_host_free(ElevenCtx* c, void* p) {
    if (!c) return fail(ELEVEN_ERR_ARG, "eleven_host_free: null context");
    CK(cudaSetDevice(c->cfg.device));
    CK(cudaFreeHost(p));
    return ELEVEN_OK;
}
extern "C" int eleven_device_free(ElevenCtx* c, void* p) {
    if (!c) return fail(ELEVEN_ERR_ARG, "eleven_device_free: null context");
    CK(cudaSetDevice(c->cfg.device)); CK(cudaFree(p)); return ELEVEN_OK;
}
extern "C" int eleven_device_upload(ElevenCtx* c, void* d, const void* h, size_t bytes) {
    if (!c || !d || !h) return fail(ELEVEN_ERR_ARG, "eleven_device_upload: null argument");
    CK(cudaSetDevice(c->cfg.device));
    CK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, c->stream)); CK(cudaStreamSynchronize(c->stream)); return ELEVEN_OK;
}
extern "C" int eleven_device_download(ElevenCtx* c, void* h, const void* d, size_t bytes) {
    if (!c || !d || !h) return fail(ELEVEN_ERR_ARG, "eleven_device_download: null argument");
    CK(cudaSetDevice(c->cfg.device));
    CK(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, c->stream)); CK(cudaStreamSynchronize(c->stream)); return ELEVEN_OK;
}

// ---- multi-GPU: ONE ncclReduce of the film records to the root ---------------------------------------------------
// The reference is single-GPU (cudaSetDevice(0), S/kernel.cu:604) and keeps running means, which cannot be combined; here the film
// is sums + counts (.w), so the whole exchange step of a job is one sum-reduce over NVLink.  libnccl.so.2 is loaded on first use:
// inside a torch process this resolves to the NCCL torch already loaded, in the CLI to the system library.
namespace {
struct NcclApi {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string err;
};
NcclApi* ncclApi() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {getenv("ELEVEN_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) { if (n && (api.h = dlopen(n, RTLD_NOW | RTLD_LOCAL))) break; }
        if (!api.h) { api.err = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "not found"); return; }
        bool ok = true;
        auto sym = [&](const char* n) { void* p = dlsym(api.h, n); if (!p) { ok = false; api.err = std::string("libnccl: missing symbol ") + n; } return p; };
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.Reduce = (decltype(api.Reduce))sym("ncclReduce");
        api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
        if (!ok) { dlclose(api.h); api.h = nullptr; }
    });
    return &api;
}
int ncclFail(const char* what, ncclResult_t r) { return fail(ELEVEN_ERR_CUDA, std::string(what) + ": " + ncclApi()->GetErrorString(r)); }
} // namespace

static void commDestroy(ElevenCtx* c) {
    if (c->comm) { ncclApi()->CommDestroy(c->comm); c->comm = nullptr; c->commRank = -1; c->commSize = 0; }
}

// NCCL sets up its channels (peer buffers over NVLink, proxy threads) lazily inside the FIRST collective of a communicator: ~0.5 s on a
// B200 box.  A job has exactly one film reduce, i.e. it would pay that at the very end, on the critical path between the last wave and
// the picture.  So communicator creation ends with an 8 MB reduce on the auxiliary stream: the set-up cost moves to where the caller
// can overlap it with the scene upload and the rendering (the CLI creates the communicator on a thread of its own).
static const size_t COMM_WARM_FLOATS = 2u << 20;       // 8 MB: large enough for the protocol and channel buffers a film-sized reduce uses (a 16-byte warm-up left
                                                       // 37-44 ms of set-up in the first 33 MB reduce on 4 / 8 GPUs)
static int commWarmupAlloc(ElevenCtx* c) {             // outside any NCCL group: cudaMalloc may synchronise
    if (!c->d_commWarm) { CK(cudaMalloc((void**)&c->d_commWarm, COMM_WARM_FLOATS * 4)); CK(cudaMemsetAsync(c->d_commWarm, 0, COMM_WARM_FLOATS * 4, c->auxStream)); }
    return ELEVEN_OK;
}
static int commWarmupEnqueue(ElevenCtx* c) {
    ncclResult_t r = ncclApi()->Reduce(c->d_commWarm, c->d_commWarm, COMM_WARM_FLOATS, ncclFloat, ncclSum, 0, c->comm, c->auxStream);
    return r == ncclSuccess ? ELEVEN_OK : ncclFail("ncclReduce (communicator warm-up)", r);
}

extern "C" int eleven_comm_unique_id(void* idOut) {
    static_assert(sizeof(ncclUniqueId) == ELEVEN_COMM_ID_BYTES, "ncclUniqueId size");
    if (!idOut) return fail(ELEVEN_ERR_ARG, "eleven_comm_unique_id: null argument");
    NcclApi* N = ncclApi();
    if (!N->h) return fail(ELEVEN_ERR_UNSUPPORTED, N->err);
    ncclUniqueId id;
    ncclResult_t r = N->GetUniqueId(&id);
    if (r != ncclSuccess) return ncclFail("ncclGetUniqueId", r);
    memcpy(idOut, &id, sizeof id);
    return ELEVEN_OK;
}

extern "C" int eleven_comm_init_rank(ElevenCtx* c, const void* idIn, int nranks, int rank) {
    if (!c || !idIn) return fail(ELEVEN_ERR_ARG, "eleven_comm_init_rank: null argument");
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(ELEVEN_ERR_ARG, "eleven_comm_init_rank: bad rank / nranks");
    NcclApi* N = ncclApi();
    if (!N->h) return fail(ELEVEN_ERR_UNSUPPORTED, N->err);
    CK(cudaSetDevice(c->cfg.device));
    commDestroy(c);
    ncclUniqueId id; memcpy(&id, idIn, sizeof id);
    ncclResult_t r = N->CommInitRank(&c->comm, nranks, id, rank);
    if (r != ncclSuccess) { c->comm = nullptr; return ncclFail("ncclCommInitRank", r); }
    c->commRank = rank; c->commSize = nranks;
    if (int rc = commWarmupAlloc(c)) return rc;
    if (int rc = commWarmupEnqueue(c)) return rc;
    CK(cudaStreamSynchronize(c->auxStream));
    cudaFree(c->d_commWarm); c->d_commWarm = nullptr;
    return ELEVEN_OK;
}

extern "C" int eleven_comm_init_all(ElevenCtx** ctxs, int n) {
    if (!ctxs || n < 1) return fail(ELEVEN_ERR_ARG, "eleven_comm_init_all: bad argument");
    NcclApi* N = ncclApi();
    if (!N->h) return fail(ELEVEN_ERR_UNSUPPORTED, N->err);
    std::vector<int> devs(n); std::vector<ncclComm_t> comms(n);
    for (int i = 0; i < n; i++) {
        if (!ctxs[i]) return fail(ELEVEN_ERR_ARG, "eleven_comm_init_all: null context");
        devs[i] = ctxs[i]->cfg.device;
        for (int j = 0; j < i; j++) if (devs[j] == devs[i]) return fail(ELEVEN_ERR_ARG, "eleven_comm_init_all: two contexts on the same device (NCCL needs one rank per device)");
        commDestroy(ctxs[i]);
    }
    ncclResult_t r = N->CommInitAll(comms.data(), n, devs.data());
    if (r != ncclSuccess) return ncclFail("ncclCommInitAll", r);
    for (int i = 0; i < n; i++) { ctxs[i]->comm = comms[i]; ctxs[i]->commRank = i; ctxs[i]->commSize = n; }
    if (n > 1) {
        int rc = ELEVEN_OK;
        for (int i = 0; i < n; i++) { CK(cudaSetDevice(devs[i])); if ((rc = commWarmupAlloc(ctxs[i]))) return rc; }
        N->GroupStart();                                    // one thread drives all ranks: the collective must be issued as a group
        for (int i = 0; i < n && !rc; i++) { if (cudaSetDevice(devs[i]) != cudaSuccess) rc = fail(ELEVEN_ERR_CUDA, "cudaSetDevice"); else rc = commWarmupEnqueue(ctxs[i]); }
        r = N->GroupEnd();
        if (rc) return rc;
        if (r != ncclSuccess) return ncclFail("ncclGroupEnd (communicator warm-up)", r);
        for (int i = 0; i < n; i++) { CK(cudaSetDevice(devs[i])); CK(cudaStreamSynchronize(ctxs[i]->auxStream)); cudaFree(ctxs[i]->d_commWarm); ctxs[i]->d_commWarm = nullptr; }
    }
    return ELEVEN_OK;
}

extern "C" int eleven_reduce_film(ElevenCtx* c, int root, int allPasses) {
    if (!c) return fail(ELEVEN_ERR_ARG, "eleven_reduce_film: null context");
    if (!c->haveScene) return fail(ELEVEN_ERR_STATE, "eleven_reduce_film: no scene uploaded");
    const int size = c->comm ? c->commSize : 1, rank = c->comm ? c->commRank : 0;
    if (root < 0 || root >= size) return fail(ELEVEN_ERR_ARG, "eleven_reduce_film: root out of range");
    CK(cudaSetDevice(c->cfg.device));
    const int passes = allPasses ? 4 : 1;
    const size_t nFloats = (size_t)c->nPixels * 4 * passes;
    if (rank == root && !c->d_reduced) {
        cudaError_t e = cudaMalloc((void**)&c->d_reduced, (size_t)c->nPixels * 4 * sizeof(float4));
        if (e != cudaSuccess) { c->d_reduced = nullptr; return fail(ELEVEN_ERR_NOMEM, std::string("cudaMalloc(reduced film): ") + cudaGetErrorString(e)); }
    }
    CK(cudaEventRecord(c->evR0, c->stream));
    if (c->comm) {
        // the render stream orders the reduce behind every wave accumulated so far; the snapshot stream never touches d_reduced
        // while this runs because its readers take snapMutex and the root takes it here too
        std::unique_lock<std::mutex> lk(c->snapMutex, std::defer_lock);
        if (rank == root) lk.lock();
        ncclResult_t r = ncclApi()->Reduce(c->W.filmBeauty, rank == root ? (void*)c->d_reduced : (void*)c->W.filmBeauty, nFloats, ncclFloat, ncclSum, root, c->comm, c->stream);
        if (r != ncclSuccess) return ncclFail("ncclReduce", r);
        CK(cudaEventRecord(c->evR1, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    } else {
        std::lock_guard<std::mutex> lk(c->snapMutex);
        CK(cudaMemcpyAsync(c->d_reduced, c->W.filmBeauty, nFloats * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
        CK(cudaEventRecord(c->evR1, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    if (rank == root) c->reducedPasses = passes;
    float ms = 0; CK(cudaEventElapsedTime(&ms, c->evR0, c->evR1));
    std::lock_guard<std::mutex> lk(c->statsMutex);
    c->stats.reduce_ms += ms; c->stats.reduce_calls += 1; c->stats.kernel_launches += 1;
    return ELEVEN_OK;
}

// ---- known-answer test hooks (test_hooks.cuh) ------------------------------------------------------------------------------------
namespace {
struct Scratch {                                  // device buffers of one hook call, freed on every exit path
    std::vector<void*> p;
    ~Scratch() { for (void* q : p) cudaFree(q); }
    template <typename T> T* up(const T* host, size_t count, cudaStream_t s) {
        void* q = nullptr;
        if (cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T)) != cudaSuccess) return nullptr;
        p.push_back(q);
        if (host && count && cudaMemcpyAsync(q, host, count * sizeof(T), cudaMemcpyHostToDevice, s) != cudaSuccess) return nullptr;
        return (T*)q;
    }
};
} // namespace
#define HOOK_DOWN(dst, src, count) CK(cudaMemcpyAsync(dst, src, (size_t)(count) * sizeof(*(dst)), cudaMemcpyDeviceToHost, c->stream))

extern "C" int eleven_test_disney(ElevenCtx* c, const float* records, size_t n, int fastMath, float* evalPdf, float* sample) {
    if (!c || (n && (!records || !evalPdf || !sample))) return fail(ELEVEN_ERR_ARG, "eleven_test_disney: null argument");
    if (n == 0) return ELEVEN_OK;
    CK(cudaSetDevice(c->cfg.device));
    Scratch S;
    float* dr = S.up(records, n * 30, c->stream); float* de = S.up<float>(nullptr, n * 4, c->stream); float* ds = S.up<float>(nullptr, n * 3, c->stream);
    if (!dr || !de || !ds) return fail(ELEVEN_ERR_NOMEM, "eleven_test_disney: device scratch");
    const unsigned g = (unsigned)((n + 127) / 128);
    if (fastMath) k_testDisney<true><<<g, 128, 0, c->stream>>>(dr, (uint32_t)n, de, ds);
    else k_testDisney<false><<<g, 128, 0, c->stream>>>(dr, (uint32_t)n, de, ds);
    CK(cudaGetLastError());
    HOOK_DOWN(evalPdf, de, n * 4); HOOK_DOWN(sample, ds, n * 3);
    CK(cudaStreamSynchronize(c->stream));
    return ELEVEN_OK;
}

extern "C" int eleven_test_hdri(ElevenCtx* c, const float* r, const float* r2, size_t n, int envMode, int fastMath, int32_t* xy, float* dir, float* pdf) {
    if (!c || (n && (!r || !xy || !dir || !pdf))) return fail(ELEVEN_ERR_ARG, "eleven_test_hdri: null argument");
    if (envMode == ELEVEN_ENV_ALIAS && n && !r2) return fail(ELEVEN_ERR_ARG, "eleven_test_hdri: the alias table needs a second uniform");
    if (!c->haveScene) return fail(ELEVEN_ERR_STATE, "eleven_test_hdri: no scene uploaded");
    if (n == 0) return ELEVEN_OK;
    CK(cudaSetDevice(c->cfg.device));
    Scratch S;
    float* dr = S.up(r, n, c->stream); float* dr2 = S.up(r2 ? r2 : r, n, c->stream);
    int32_t* dxy = S.up<int32_t>(nullptr, n * 2, c->stream); float* dd = S.up<float>(nullptr, n * 3, c->stream); float* dp = S.up<float>(nullptr, n, c->stream);
    if (!dr || !dr2 || !dxy || !dd || !dp) return fail(ELEVEN_ERR_NOMEM, "eleven_test_hdri: device scratch");
    const unsigned g = (unsigned)((n + 127) / 128);
    if (fastMath) k_testHdri<true><<<g, 128, 0, c->stream>>>(c->scene, dr, dr2, (uint32_t)n, envMode, dxy, dd, dp);
    else k_testHdri<false><<<g, 128, 0, c->stream>>>(c->scene, dr, dr2, (uint32_t)n, envMode, dxy, dd, dp);
    CK(cudaGetLastError());
    HOOK_DOWN(xy, dxy, n * 2); HOOK_DOWN(dir, dd, n * 3); HOOK_DOWN(pdf, dp, n);
    CK(cudaStreamSynchronize(c->stream));
    return ELEVEN_OK;
}

extern "C" int eleven_test_env_lookup(ElevenCtx* c, const float* dirs, size_t n, float* rgb) {
    if (!c || (n && (!dirs || !rgb))) return fail(ELEVEN_ERR_ARG, "eleven_test_env_lookup: null argument");
    if (!c->haveScene) return fail(ELEVEN_ERR_STATE, "eleven_test_env_lookup: no scene uploaded");
    if (n == 0) return ELEVEN_OK;
    CK(cudaSetDevice(c->cfg.device));
    Scratch S;
    float* dd = S.up(dirs, n * 3, c->stream); float* dc = S.up<float>(nullptr, n * 3, c->stream);
    if (!dd || !dc) return fail(ELEVEN_ERR_NOMEM, "eleven_test_env_lookup: device scratch");
    k_testEnvLookup<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(c->scene, dd, (uint32_t)n, dc);
    CK(cudaGetLastError());
    HOOK_DOWN(rgb, dc, n * 3);
    CK(cudaStreamSynchronize(c->stream));
    return ELEVEN_OK;
}

extern "C" int eleven_test_hitdata(ElevenCtx* c, const float* attrs, const int32_t* objectIds, size_t n, int fastMath, float* out) {
    if (!c || (n && (!attrs || !objectIds || !out))) return fail(ELEVEN_ERR_ARG, "eleven_test_hitdata: null argument");
    if (!c->haveScene) return fail(ELEVEN_ERR_STATE, "eleven_test_hitdata: no scene uploaded");
    if (n == 0) return ELEVEN_OK;
    CK(cudaSetDevice(c->cfg.device));
    Scratch S;
    float* da = S.up(attrs, n * 14, c->stream); int32_t* di = S.up(objectIds, n, c->stream); float* dout = S.up<float>(nullptr, n * 21, c->stream);
    if (!da || !di || !dout) return fail(ELEVEN_ERR_NOMEM, "eleven_test_hitdata: device scratch");
    const unsigned g = (unsigned)((n + 127) / 128);
    if (fastMath) k_testHitData<true><<<g, 128, 0, c->stream>>>(c->scene, da, di, (uint32_t)n, dout);
    else k_testHitData<false><<<g, 128, 0, c->stream>>>(c->scene, da, di, (uint32_t)n, dout);
    CK(cudaGetLastError());
    HOOK_DOWN(out, dout, n * 21);
    CK(cudaStreamSynchronize(c->stream));
    return ELEVEN_OK;
}
