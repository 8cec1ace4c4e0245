/*
 * eleven_b200.h — C ABI of the B200-native path-tracing hot path.
 *
 * This is the drop-in boundary for the four C++ entry points of the reference
 * renderer (S/ = /root/reference/src/tfg-pathtracer/):
 *
 *     cudaError_t renderSetup(Scene*)                          S/kernel.h:80, S/kernel.cu:566
 *     cudaError_t renderCuda(Scene*, int sampleTarget)         S/kernel.h:78, S/kernel.cu:665
 *     cudaError_t getBuffers(RenderData&, int*, int)           S/kernel.h:82, S/kernel.cu:688
 *     int         getSamples()                                 S/kernel.h:84, S/kernel.cu:712
 *
 * Everything crossing the boundary is a plain pointer + size; no C++ types,
 * no torch types, no std::string.  All structs are POD, 4-byte aligned,
 * little-endian.  Every function returns ELEVEN_OK (0) or a negative error
 * code; eleven_last_error() returns the message of the last failure on the
 * calling thread.  Nothing is printed-and-ignored (contrast S/kernel.cu:657).
 */
#ifndef ELEVEN_B200_H
#define ELEVEN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ELEVEN_ABI_VERSION 3

/* ---- status codes ------------------------------------------------------- */
enum {
    ELEVEN_OK            =  0,
    ELEVEN_ERR_ARG       = -1,   /* bad argument / null pointer / bad size      */
    ELEVEN_ERR_CUDA      = -2,   /* a CUDA runtime call failed                   */
    ELEVEN_ERR_STATE     = -3,   /* call order violated (e.g. render before upload) */
    ELEVEN_ERR_NOMEM     = -4,
    ELEVEN_ERR_UNSUPPORTED = -5
};

/* ---- film passes: same numbering as enum Passes, S/kernel.h:7 ----------- */
enum { ELEVEN_PASS_BEAUTY = 0, ELEVEN_PASS_DENOISE = 1, ELEVEN_PASS_NORMAL = 2,
       ELEVEN_PASS_TANGENT = 3, ELEVEN_PASS_BITANGENT = 4, ELEVEN_PASS_COUNT = 5 };

/* ---- configuration ------------------------------------------------------ */
enum { ELEVEN_RNG_REFERENCE = 0,   /* per-pixel XORWOW, curand_init(0, idx, 0) (S/kernel.cu:140); stream-exact, 1 GPU */
       ELEVEN_RNG_FAST      = 1 }; /* counter-based hash keyed by (pixel, sample, dimension); multi-GPU safe       */
enum { ELEVEN_ENV_CDF   = 0,       /* the reference's flat float CDF + its binarySearch (S/HDRI.hpp:107-162)      */
       ELEVEN_ENV_ALIAS = 1 };     /* Walker alias table over the same texel weights                              */
enum { ELEVEN_HIT_KEY   = 0,       /* closest = min |hit.position-origin| with shadow-terminator shift (S/BVH.hpp:170) */
       ELEVEN_HIT_MIN_T = 1 };     /* closest = min t (classic); differs only inside the per-scene shift bound    */

enum { ELEVEN_BVH_HOST   = 0,      /* binned-SAH BVH8 built on the host cores (bvh8_build.cpp)                      */
       ELEVEN_BVH_DEVICE = 1 };    /* binned-SAH BVH8 built on the GPU (bvh8_build_gpu.cuh)                         */

typedef struct ElevenConfig {
    int32_t  device;        /* CUDA device ordinal                                                     */
    uint32_t rng_mode;      /* ELEVEN_RNG_*                                                            */
    uint32_t env_mode;      /* ELEVEN_ENV_*                                                            */
    uint32_t hit_mode;      /* ELEVEN_HIT_*                                                            */
    uint32_t max_bounces;   /* MAXBOUNCES, S/Definitions.h:9 (5)                                       */
    uint32_t sample_offset; /* fast rng: global index of this context's first sample (multi-GPU split) */
    uint32_t sample_stride; /* fast rng: global sample index advances by this per local sample (>=1)   */
    uint32_t flags;         /* ELEVEN_FLAG_*                                                           */
    uint64_t seed;          /* fast rng key; the reference mode always uses seed 0 like the reference  */
    uint32_t wave_spp;      /* samples of every pixel in flight per wave (power of two, <= 16); 0 = auto: as many as keep
                             * a wave <= 2^26 paths (and within the device's free memory).  Only ELEVEN_RNG_FAST can have more than 1 (the reference's per-pixel
                             * XORWOW stream is sequential across samples, S/kernel.cu:380,480).  Same image for any value. */
    uint32_t bvh_builder;   /* ELEVEN_BVH_*                                                            */
} ElevenConfig;

#define ELEVEN_FLAG_TERMINATE_DEAD_PATHS 1u  /* stop paths whose throughput is exactly 0 (only legal with RNG_FAST) */
#define ELEVEN_FLAG_COUNTERS             2u  /* count nodes/triangles visited per ray (slower; for the roofline)    */
#define ELEVEN_FLAG_TIME_KERNELS         4u  /* CUDA-event timing around every pipeline stage (ElevenStats *_ms)   */
#define ELEVEN_FLAG_FAST_MATH           16u  /* shading with MUFU reciprocal/rsqrt/sin/cos/log2/exp2 and float intermediates, like the
                                              * reference's shipping -use_fast_math build; geometry (t,u,v,key) stays bit-exact */
#define ELEVEN_FLAG_ANYHIT_LIGHT_SHADOWS 32u  /* point-light shadow rays stop at the first occluder with t < |L - P| instead of taking the
                                              * reference's closest hit and comparing |hit.position - P| (S/kernel.cu:193-197): differs only
                                              * where an occluder lies within the shadow-terminator shift + 1 mm of the light itself */
#define ELEVEN_FLAG_SKIP_NULL_NEE        8u  /* no env shadow ray when its contribution is 0 whatever it hits (no point
                                              * lights, no emission, BRDF 0 towards the sample): same image, fewer rays */

/* ---- scene description (what renderSetup copies out of Scene, S/kernel.cu:566-661) ---- */

/* Camera.hpp:6-40 */
typedef struct ElevenCamera {
    uint32_t xRes, yRes;
    float    focalLength, sensorWidth, sensorHeight, aperture, focusDistance;
    float    rotation[3];     /* degrees, applied X then Y then Z (S/kernel.cu:299-306) */
    float    position[3];
    uint32_t bokeh;
} ElevenCamera;

/* Tri.hpp:13-19, 152 bytes, identical field order so a std::vector<Tri> can be passed as is. */
typedef struct ElevenTri {
    float   vertices[3][3];
    float   uv[3][3];        /* z unused */
    float   normals[3][3];
    float   tangents[3][3];
    float   tangentsSign;
    int32_t objectID;
} ElevenTri;

/* Material.hpp:8-37 without the std::string name. */
typedef struct ElevenMaterial {
    int32_t albedoTextureID, emissionTextureID, roughnessTextureID,
            metallicTextureID, normalTextureID, opacityTextureID;
    float   albedo[3], emission[3], opacity[3];
    float   roughness, metallic, clearcoatGloss, clearcoat, anisotropic, eta,
            transmission, specular, specularTint, sheenTint, subsurface, sheen;
} ElevenMaterial;

enum { ELEVEN_TEX_F32_RGB   = 0,   /* float RGB, 12 B/texel: Texture::data as the reference holds it (S/Texture.hpp:18) */
       ELEVEN_TEX_U8_SRGB   = 1,   /* 8-bit RGB; decoded through the 256-entry fastPow(x/255, 2.2) table (S/stb_image.h:127-136,1863) */
       ELEVEN_TEX_U8_LINEAR = 2 }; /* 8-bit RGB; decoded through the fastPow(x/255, 1.0) table (not the identity!)      */

/* Texture.hpp:14-34.  Row 0 is the first row of `data` exactly as the reference indexes it
 * (LDR textures are flipped on load, S/Texture.hpp:49; the HDRI is not). */
typedef struct ElevenTexture {
    const void* data;
    uint32_t format;          /* ELEVEN_TEX_* */
    int32_t  width, height;
    float    xTile, yTile, xOffset, yOffset;
    uint32_t filter;          /* 0 = NO_FILTER (nearest), 1 = BILINEAR (S/Texture.hpp:10) */
} ElevenTexture;

/* PointLight.hpp:6-20 */
typedef struct ElevenPointLight { float position[3]; float radiance[3]; } ElevenPointLight;

typedef struct ElevenSceneDesc {
    ElevenCamera            camera;
    uint32_t                triCount;
    const ElevenTri*        tris;              /* Scene::tris, S/Scene.hpp:28 */
    uint32_t                objectCount;
    const int32_t*          objectMaterial;    /* meshObjects[i].materialID, S/MeshObject.hpp:22 */
    uint32_t                materialCount;
    const ElevenMaterial*   materials;
    uint32_t                textureCount;
    const ElevenTexture*    textures;
    ElevenTexture           hdri;              /* must be ELEVEN_TEX_F32_RGB; HDRI::texture, S/HDRI.hpp:14 */
    uint32_t                pointLightCount;
    const ElevenPointLight* pointLights;
} ElevenSceneDesc;

/* ---- results ------------------------------------------------------------ */

/* One closest-hit record.  The reference's Hit (S/Hit.hpp:6-13) carries neither the
 * triangle nor t,u,v; the parity contract (BASELINE.json north_star) is on these. */
typedef struct ElevenHit {
    int32_t tri;      /* index into ElevenSceneDesc::tris, -1 = miss */
    float   t, u, v;  /* Moeller-Trumbore outputs, S/Tri.hpp:40-68    */
    float   key;      /* |hit.position - ray.origin|, the reference's ordering key (S/BVH.hpp:170) */
} ElevenHit;

typedef struct ElevenStats {
    uint64_t pixel_samples;      /* W*H*spp rendered by this context                        */
    uint64_t rays_extension;     /* closest-hit rays traced                                  */
    uint64_t rays_shadow_env;    /* environment NEE shadow rays                              */
    uint64_t rays_shadow_light;  /* point-light NEE shadow rays                              */
    uint64_t hit_bounces;        /* the reference's "paths": sum of dev_pathcount (S/kernel.cu:445) */
    uint64_t nodes_visited;      /* BVH8 nodes fetched   (only with ELEVEN_FLAG_COUNTERS)    */
    uint64_t tris_tested;        /* triangles intersected (only with ELEVEN_FLAG_COUNTERS)   */
    uint64_t kernel_launches;    /* CUDA kernels launched by eleven_render so far            */
    double   render_ms;          /* device time inside eleven_render (CUDA events)           */
    double   trace_ms;           /* device time of eleven_trace_* batches                    */
    double   extend_ms;          /* with ELEVEN_FLAG_TIME_KERNELS: closest-hit kernels       */
    double   shade_ms;           /*   "   : shade kernels                                    */
    double   connect_ms;         /*   "   : shadow-ray + MIS kernels                         */
    double   other_ms;           /*   "   : raygen, queue bookkeeping, accumulate            */
    uint64_t extend_launches;    /* closest-hit kernel launches inside eleven_render         */
    double   bvh_build_ms;       /* wall time of the BVH8 build (host or device builder)     */
    uint32_t bvh_nodes;          /* BVH8 node count                                          */
    uint32_t bvh_tri_slots;      /* triangle slots in leaf order (>= triCount: a pre-split sliver triangle owns several) */
    float    key_slack;          /* per-scene bound on |key - t| used for culling in HIT_KEY */
    uint32_t samples_done;       /* per-pixel sample count of pixel 0 (getSamples)           */
    uint64_t key_evals;          /* exact reference-key evaluations (only with ELEVEN_FLAG_COUNTERS) */
    double   reduce_ms;          /* device time inside eleven_reduce_film (CUDA events on the render stream)    */
    uint64_t reduce_calls;       /* eleven_reduce_film calls so far                                              */
    uint64_t nodes_visited_extend; /* nodes_visited / tris_tested of the closest-hit kernel (k_extend) alone: the per-ray      */
    uint64_t tris_tested_extend;   /* figures of the roofline's algorithmic bytes (only with ELEVEN_FLAG_COUNTERS)            */
} ElevenStats;

typedef struct ElevenCtx ElevenCtx;

/* ---- entry points ------------------------------------------------------- */

int  eleven_abi_version(void);
const char* eleven_last_error(void);

/* Creates a context on cfg->device.  Replaces the implicit global state + cudaSetDevice(0)
 * of renderSetup (S/kernel.cu:46-49,604). */
int  eleven_init(const ElevenConfig* cfg, ElevenCtx** out);
void eleven_destroy(ElevenCtx* ctx);

/* renderSetup (S/kernel.cu:566-661): build the acceleration structure, copy the scene,
 * zero the film, seed the per-pixel RNG (setupKernel, S/kernel.cu:121-150). */
int  eleven_scene_upload(ElevenCtx* ctx, const ElevenSceneDesc* scene);

/* renderCuda (S/kernel.cu:665-686): add `spp` samples to every pixel; blocking. */
int  eleven_render(ElevenCtx* ctx, int spp);

/* getBuffers (S/kernel.cu:688-710): RGBA float, index W*(H-1-y)+x, A = 1; BEAUTY is the mean of
 * per-sample radiance clamped to [0,10] (S/kernel.cu:447-463).  n_pixels must be W*H.
 * Threading (S/main.cpp:132-184: the main thread polls getBuffers/getSamples on bufferStream every 100 ms while the
 * render thread sits in renderCuda): eleven_get_film, eleven_resolve_rgba8, eleven_get_pathcount, eleven_get_samples,
 * eleven_get_sample_counts and eleven_get_stats run on the context's SNAPSHOT stream and may be called from another
 * thread while eleven_render is running; they return the film as of the waves accumulated so far, without waiting for
 * the render to end.  Every pixel is seen either before or after a wave's update (sum and sample count share one
 * 16-byte record); different pixels may be one wave apart, as in the reference.  After eleven_render has returned the
 * film is exact. */
int  eleven_get_film(ElevenCtx* ctx, int pass, float* rgba, size_t n_pixels);
int  eleven_get_pathcount(ElevenCtx* ctx, int32_t* out, size_t n_pixels);   /* dev_pathcount */
int  eleven_get_samples(ElevenCtx* ctx);                                    /* getSamples, S/kernel.cu:712 */
/* dev_samples (S/kernel.cu:44): per-pixel count of accepted samples; NaN samples are dropped without counting
 * (S/kernel.cu:449,477), so pixels can lag the number of samples rendered. */
int  eleven_get_sample_counts(ElevenCtx* ctx, uint32_t* out, size_t n_pixels);
int  eleven_get_stats(ElevenCtx* ctx, ElevenStats* out);

/* Replace the camera (same resolution) without touching the scene: 56 bytes host->device.  The reference has no such
 * call (its camera is baked by renderSetup); it is what an interactive host needs between frames. */
int  eleven_set_camera(ElevenCtx* ctx, const ElevenCamera* camera);

/* Zero film + counters and re-seed, keeping the uploaded scene (the reference has no such call:
 * it is what re-running renderSetup's setupKernel would do). */
int  eleven_film_reset(ElevenCtx* ctx);

/* Test hook for the closest-hit contract: BVH::transverse (S/BVH.hpp:120-157) on a ray batch.
 * rays = n * {ox,oy,oz,dx,dy,dz}; directions are normalised like Ray's ctor (S/Ray.hpp:14-18).
 * Host pointers; copies are inside the call. */
int  eleven_trace_closest(ElevenCtx* ctx, const float* rays, size_t n, ElevenHit* hits);
/* Same, device pointers already resident (bench: inputs in HBM). `any_hit` != 0 traces
 * occlusion only (hits[i].tri = -1 or the first triangle found). Returns device ms in *ms. */
int  eleven_trace_device(ElevenCtx* ctx, const float* d_rays, size_t n, ElevenHit* d_hits,
                         int any_hit, float* ms);

/* Test hook for the acceleration structure (replaces nothing in the reference, whose BVH is a host object,
 * S/BVH.hpp:50-62): copies the device-resident BVH8 to host buffers — nodes (80 B each, bvh8.h Node8), triangle slots in
 * leaf order (48 B each, TriSlot) and the per-node culling slack.  Capacities in elements; counts are in ElevenStats. */
int  eleven_bvh_download(ElevenCtx* ctx, void* nodes, size_t node_cap, void* slots, size_t slot_cap, float* node_slack);

/* Test hook: the host BVH8 builder on its own (replaces BVH::build, S/BVH.hpp:187-330).  No GPU, no context.  Buffers may be
 * NULL to query the sizes: counts[0] = nodes, counts[1] = triangle slots.  threads <= 0: all host cores. */
int  eleven_bvh_build_host(const ElevenTri* tris, uint32_t n, const int32_t* tri_material, int threads,
                           void* nodes, size_t node_cap, void* slots, size_t slot_cap, float* node_slack,
                           uint32_t* counts, float* key_slack);

/* ---- multi-GPU (the reference has none: cudaSetDevice(0), S/kernel.cu:604) -------------------------------
 * The scene is replicated, the samples are split (ElevenConfig.sample_offset/stride), every context accumulates per-pixel
 * SUMS whose .w carries the accepted-sample count (exact below 2^24 samples per pixel), and the one exchange step of a job is
 * ONE ncclReduce(sum, float32) of those records to the root over NVLink, followed by the fused resolve on the root.
 * NCCL (libnccl.so.2) is loaded on first use; without it these calls return ELEVEN_ERR_UNSUPPORTED. */
#define ELEVEN_COMM_ID_BYTES 128
/* ncclGetUniqueId: call on one rank, hand the 128 bytes to every rank (any transport: torch.distributed, MPI, a file). */
int  eleven_comm_unique_id(void* id_out);
/* One process (or thread) per GPU: ncclCommInitRank on the context's device.  Collective: every rank must call it.
 * Both init calls end with an 8 MB reduce on the context's auxiliary stream: NCCL builds its channels inside the first collective
 * of a communicator (~0.5 s), and a job's only reduce would otherwise pay that between the last wave and the picture.  Create the
 * communicator early (the CLI does it on a thread of its own while the scene uploads and renders). */
int  eleven_comm_init_rank(ElevenCtx* ctx, const void* id, int nranks, int rank);
/* One process driving n contexts on n different devices (the `eleven --gpus N` CLI): ncclCommInitAll; rank i = ctxs[i]. */
int  eleven_comm_init_all(ElevenCtx** ctxs, int n);
/* Sums the film records of all ranks into the ROOT context's reduced film (a separate buffer: local films are left as
 * they are, so the call can be repeated as the render progresses).  all_passes = 0: BEAUTY only (W*H float4);
 * 1: BEAUTY + NORMAL + TANGENT + BITANGENT (4*W*H float4) — either way ONE ncclReduce, enqueued on the render stream
 * behind the waves rendered so far; blocks until it has completed.  Collective; must not overlap eleven_render on the
 * same context.  Without a communicator (single GPU) it copies the local film into the reduced film. */
int  eleven_reduce_film(ElevenCtx* ctx, int root, int all_passes);
/* eleven_get_film / eleven_resolve_rgba8 / eleven_get_sample_counts on the REDUCED film (root only, after eleven_reduce_film). */
int  eleven_get_film_reduced(ElevenCtx* ctx, int pass, float* rgba, size_t n_pixels);
int  eleven_resolve_rgba8_reduced(ElevenCtx* ctx, int pass, uint8_t* rgba8, size_t n_pixels);
int  eleven_get_sample_counts_reduced(ElevenCtx* ctx, uint32_t* out, size_t n_pixels);
/* Device pointer to the local film sums of a pass (W*H float4, .w = sample count), for callers that bring their own
 * collective.  The four passes are contiguous in pass order BEAUTY, NORMAL, TANGENT, BITANGENT. */
int  eleven_film_sums_device(ElevenCtx* ctx, int pass, void** d_ptr, size_t* n_floats);
/* Page-locked host memory for film read-backs / ray batches: copies to and from it run at PCIe/C2C rate instead of going
 * through the driver's staging buffer (the reference's host film buffers are pageable `new float[]`, S/main.cpp:44-49). */
int  eleven_host_alloc(ElevenCtx* ctx, size_t bytes, void** h_ptr);
int  eleven_host_free(ElevenCtx* ctx, void* h_ptr);
int  eleven_device_alloc(ElevenCtx* ctx, size_t bytes, void** d_ptr);
int  eleven_device_free(ElevenCtx* ctx, void* d_ptr);
int  eleven_device_upload(ElevenCtx* ctx, void* d_dst, const void* h_src, size_t bytes);
int  eleven_device_download(ElevenCtx* ctx, void* h_dst, const void* d_src, size_t bytes);

/* Fused resolve: mean, alpha, optional 8-bit pack with the reference's output curve
 * fastPow(clamp01(x), 1/2.2)*255 (S/main.cpp:156-158).  rgba8 is a HOST buffer of W*H*4 bytes. */
int  eleven_resolve_rgba8(ElevenCtx* ctx, int pass, uint8_t* rgba8, size_t n_pixels);

/* ---- known-answer test hooks for the shading functions (SURVEY §8c iii-iv; the reference's own debugging pattern:
 * printBRDFMaterial / printHDRISampling, S/kernel.cu:726-794).  Host pointers; the device functions exercised are the ones
 * k_shade calls.  fast_math selects the arithmetic flavour (ELEVEN_FLAG_FAST_MATH). ------------------------------------- */
/* records: n x 30 floats = HitData scalars in S/kernel.h:46-69 order (metallic, roughness, clearcoatGloss, clearcoat,
 * anisotropic, eta, transmission, specular, specularTint, sheenTint, subsurface, sheen), emission[3], albedo[3], normal[3],
 * ray direction[3] (normalised like Ray's ctor), L[3], r1 r2 r3.  Out: DisneyEval rgb + DisneyPdf (n x 4), DisneySample (n x 3)
 * (S/Disney.hpp:108-253).  Needs no scene. */
int  eleven_test_disney(ElevenCtx* ctx, const float* records, size_t n, int fast_math, float* eval_pdf, float* sample);
/* HDRI::sample(r) -> texel (x, y), the NEE direction and HDRI::pdf (S/HDRI.hpp:130-162, S/kernel.cu:236-243) through the
 * reference CDF search (env_mode 0) or the alias table (env_mode 1: r = the table uniform, r2 = the second uniform). */
int  eleven_test_hdri(ElevenCtx* ctx, const float* r, const float* r2, size_t n, int env_mode, int fast_math, int32_t* xy, float* dir, float* pdf);
/* radiance seen by an escaped ray of direction dirs[i] (S/kernel.cu:415-417). */
int  eleven_test_env_lookup(ElevenCtx* ctx, const float* dirs, size_t n, float* rgb);
/* generateHitData (S/kernel.cu:54-119) for given interpolated attributes: in n x 14 floats (position[3] unused, normal[3],
 * tangent[3], bitangent[3], tu, tv) + object ids; out n x 21 floats (HitData scalars, emission, albedo, shading normal). */
int  eleven_test_hitdata(ElevenCtx* ctx, const float* attrs, const int32_t* object_ids, size_t n, int fast_math, float* out);

#ifdef __cplusplus
}
#endif
#endif /* ELEVEN_B200_H */
