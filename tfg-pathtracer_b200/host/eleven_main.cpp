/*
 * eleven_main.cpp — the `eleven <scene_path> <#samples> <output.bmp>` command line of the reference (README.md:12-16,
 * S/main.cpp:111-195) over the C ABI of include/eleven_b200.h.  Headless: no GL preview, no OIDN thread (out of scope,
 * SURVEY §2).  Output is a 24-bit BMP written with the reference's save curve fastPow(clamp01(x), 1/2.2)*255
 * (S/main.cpp:156-158; the reference actually writes PNG bytes whatever the extension).
 *
 *   eleven <scene_dir | scene.flat> <#samples> <output.bmp> [--mode fast|parity] [--gpus N] [--slice K] [--bvh device|host]
 *          [--raw file.f32]        linear float RGBA film
 *          [--preview file.bmp]    8-bit snapshot after every slice (what the reference shows in its preview window)
 *          [--aov prefix]          prefix_{normal,tangent,bitangent}.f32: first-hit passes, RGBA float (the denoiser hand-off of S/main.cpp:72-74)
 *   eleven --dump-flat <scene_dir> <out.flat>          (loader only, needs no GPU)
 */
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <future>
#include <thread>
#include <vector>

#include "../../include/eleven_b200.h"
#include "scene_loader.h"

using namespace eleven_host;

static double nowS() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct DeviceJob {
    int device = 0, spp = 0; ElevenConfig cfg; ElevenSceneDesc* desc = nullptr;
    const char* previewPath = nullptr;        // device 0 only: 8-bit snapshot of the film after every slice (the reference's preview window, S/main.cpp:132-184 + Window.hpp)
    bool allPasses = false;                   // reduce the AOV passes too (--aov)
    ElevenStats stats; std::string err; int rc = 0;
    double uploadS = 0, renderS = 0, reduceS = 0;
    ElevenCtx* ctx = nullptr;
    std::shared_future<int> commReady;        // the NCCL communicator is created (and warmed up) on its own thread while this device uploads and renders
};

// One host thread per device: upload the (replicated) scene, render this device's share of the samples, then take part in the ONE
// ncclReduce that sums the film records of all devices onto device 0 (eleven_reduce_film; a plain copy when there is one device).
static void runDevice(DeviceJob* j, int slice, bool report) {
    double t0 = nowS();
    if ((j->rc = eleven_scene_upload(j->ctx, j->desc))) { j->err = eleven_last_error(); }
    j->uploadS = nowS() - t0;
    // the communicator is being created meanwhile (its cudaMalloc / IPC set-up calls are device-wide synchronisation points that would
    // stall the wavefront pipeline's launches again and again: 690 instead of 249 ms for 125 spp per device on 8 GPUs): render once it is there
    if (j->commReady.valid()) j->commReady.wait();
    t0 = nowS();
    double lastReport = t0;
    for (int done = 0; !j->rc && done < j->spp;) {
        const int k = std::min(slice, j->spp - done);
        if ((j->rc = eleven_render(j->ctx, k))) { j->err = eleven_last_error(); break; }
        done += k;
        if (j->previewPath) {                                 // progressive preview: resolve on the device, 8-bit BMP on disk (atomic rename)
            const size_t np_ = (size_t)j->desc->camera.xRes * j->desc->camera.yRes;
            std::vector<unsigned char> px(np_ * 4);
            std::string perr, tmp = std::string(j->previewPath) + ".tmp";
            if (!eleven_resolve_rgba8(j->ctx, ELEVEN_PASS_BEAUTY, px.data(), np_) &&
                writeBmp24(tmp, (int)j->desc->camera.xRes, (int)j->desc->camera.yRes, px.data(), perr))
                rename(tmp.c_str(), j->previewPath);
        }
        if (report && (done == j->spp || nowS() - lastReport >= 0.1)) {   // the reference's status line, at its 100 ms polling period (S/main.cpp:132,172-179)
            lastReport = nowS();
            ElevenStats st; eleven_get_stats(j->ctx, &st);
            const double ms = (nowS() - t0) * 1e3;
            printf("\rkPaths/s: %.1f, %d/%d samples, %.2f seconds running, %llu total paths", st.hit_bounces / ms, done, j->spp, ms / 1e3,
                   (unsigned long long)st.hit_bounces);
            fflush(stdout);
        }
    }
    j->renderS = nowS() - t0;
    // a device that failed still joins the collective (with whatever film it has) so that the others do not hang in NCCL
    if (j->commReady.valid() && j->commReady.get() != 0) { if (!j->rc) { j->rc = -1; j->err = "communicator creation failed"; } return; }
    t0 = nowS();
    const int rrc = eleven_reduce_film(j->ctx, 0, j->allPasses ? 1 : 0);
    if (!j->rc && rrc) { j->rc = rrc; j->err = eleven_last_error(); }
    j->reduceS = nowS() - t0;
    eleven_get_stats(j->ctx, &j->stats);
}

int main(int argc, char** argv) {
    const double tProcess = nowS();
    std::string err;
    if (argc >= 4 && !strcmp(argv[1], "--dump-flat")) {
        LoadedScene s;
        const double tl = nowS();
        if (!loadScene(argv[2], s, err)) { fprintf(stderr, "eleven: %s\n", err.c_str()); return 1; }
        printf("loaded in %.0f ms\n", (nowS() - tl) * 1e3);
        if (!saveFlat(s, argv[3], err)) { fprintf(stderr, "eleven: %s\n", err.c_str()); return 1; }
        printf("%zu triangles, %zu objects, %zu materials, %zu textures, %zu lights -> %s\n", s.tris.size(), s.objectMaterial.size(),
               s.materials.size(), s.textures.size(), s.lights.size(), argv[3]);
        return 0;
    }
    if (argc < 4) { fprintf(stderr, "usage: eleven <scene_path> <#samples> <output.bmp> [--mode fast|parity] [--gpus N] [--slice K] [--bvh device|host] [--raw file.f32] [--preview file.bmp] [--aov prefix]\n"); return 2; }
    const std::string scenePath = argv[1], outPath = argv[3];
    const int spp = atoi(argv[2]);
    bool fast = true, deviceBvh = true; int gpus = 1, slice = 16; const char* rawPath = nullptr; const char* previewPath = nullptr; const char* aovPrefix = nullptr;
    for (int i = 4; i < argc; i++) {
        if (!strcmp(argv[i], "--mode") && i + 1 < argc) fast = strcmp(argv[++i], "parity") != 0;
        else if (!strcmp(argv[i], "--gpus") && i + 1 < argc) gpus = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--slice") && i + 1 < argc) slice = std::max(1, atoi(argv[++i]));
        else if (!strcmp(argv[i], "--raw") && i + 1 < argc) rawPath = argv[++i];
        else if (!strcmp(argv[i], "--bvh") && i + 1 < argc) deviceBvh = strcmp(argv[++i], "host") != 0;
        else if (!strcmp(argv[i], "--preview") && i + 1 < argc) previewPath = argv[++i];
        else if (!strcmp(argv[i], "--aov") && i + 1 < argc) aovPrefix = argv[++i];
        else { fprintf(stderr, "eleven: unknown option %s\n", argv[i]); return 2; }
    }
    if (spp <= 0 || gpus < 1) { fprintf(stderr, "eleven: bad sample or GPU count\n"); return 2; }
    if (!fast && gpus > 1) { fprintf(stderr, "eleven: the reference RNG stream cannot be split across GPUs; use --mode fast\n"); return 2; }

    // (Narrowing CUDA_VISIBLE_DEVICES to the devices in use was tried: the 6-7 s before the first CUDA call returns on the 8-GPU boxes of
    // this pool — 1.3-2 s on its 1-GPU boxes — do not depend on it.)
    // CUDA context creation does not depend on the scene: one thread per device creates the contexts while this thread reads and
    // parses the scene files
    std::vector<DeviceJob> jobs(gpus);
    std::vector<ElevenCtx*> ctxs(gpus, nullptr);
    std::vector<std::thread> initTh;
    const double tInit = nowS();
    for (int g = 0; g < gpus; g++) {
        DeviceJob& j = jobs[g]; memset(&j.cfg, 0, sizeof j.cfg);
        j.device = g;
        j.cfg.device = g; j.cfg.max_bounces = 5;
        j.cfg.rng_mode = fast ? ELEVEN_RNG_FAST : ELEVEN_RNG_REFERENCE; j.cfg.env_mode = fast ? ELEVEN_ENV_ALIAS : ELEVEN_ENV_CDF;
        j.cfg.hit_mode = ELEVEN_HIT_KEY; j.cfg.flags = fast ? (ELEVEN_FLAG_TERMINATE_DEAD_PATHS | ELEVEN_FLAG_SKIP_NULL_NEE | ELEVEN_FLAG_FAST_MATH | ELEVEN_FLAG_ANYHIT_LIGHT_SHADOWS) : 0u;
        // the job's samples in waves of 16 (what a context renders per wave at full speed), a contiguous run of whole waves per device:
        // nobody renders ragged 8 + 4 + 1-sample waves (1000 spp on 8 devices: 128 x 7 + 104 instead of 125 x 8 = 7 full waves + 3 small
        // ones each).  The counter RNG is keyed by the global sample index: the image is the same for any split.
        const int waveSpp = 16, waves = (spp + waveSpp - 1) / waveSpp;
        const int w0 = g * (waves / gpus) + std::min(g, waves % gpus), w1 = w0 + waves / gpus + (g < waves % gpus ? 1 : 0);
        const int s0 = std::min(spp, w0 * waveSpp), s1 = std::min(spp, w1 * waveSpp);
        j.cfg.sample_offset = fast ? (uint32_t)s0 : 0u; j.cfg.sample_stride = 1u;
        j.cfg.bvh_builder = deviceBvh ? ELEVEN_BVH_DEVICE : ELEVEN_BVH_HOST;
        j.spp = s1 - s0;
        j.allPasses = aovPrefix != nullptr;
        if (g == 0) j.previewPath = previewPath;                 // device 0's share of the samples is an unbiased picture of its own
        initTh.emplace_back([&j]() { if ((j.rc = eleven_init(&j.cfg, &j.ctx))) j.err = eleven_last_error(); });
    }

    printf("Loading scene \n");
    double t0 = nowS();
    LoadedScene scene;
    const bool loaded = loadScene(scenePath, scene, err);
    const double loadS = nowS() - t0;
    for (auto& t : initTh) t.join();
    const double initS = nowS() - tInit;
    for (auto& j : jobs) if (j.rc) { fprintf(stderr, "eleven: device %d: %s\n", j.device, j.err.c_str()); return 1; }
    if (!loaded) { fprintf(stderr, "eleven: %s\n", err.c_str()); for (auto& j : jobs) eleven_destroy(j.ctx); return 1; }
    ElevenSceneDesc desc = scene.desc();
    printf("%s: %zu triangles, %zu textures, %ux%u, loaded in %.0f ms (CUDA contexts ready after %.0f ms, in parallel)\n", scenePath.c_str(), scene.tris.size(),
           scene.textures.size(), desc.camera.xRes, desc.camera.yRes, loadS * 1e3, initS * 1e3);
    for (int g = 0; g < gpus; g++) { jobs[g].desc = &desc; ctxs[g] = jobs[g].ctx; }
    // NCCL communicator + its first (tiny) collective on a thread of its own: ~1-2 s that overlap the upload and the rendering
    std::string commErr; double commS = 0;
    std::shared_future<int> commReady;
    if (gpus > 1) {
        commReady = std::async(std::launch::async, [&]() { const double tc = nowS(); const int rc = eleven_comm_init_all(ctxs.data(), gpus); if (rc) commErr = eleven_last_error(); commS = nowS() - tc; return rc; }).share();
        for (auto& j : jobs) j.commReady = commReady;
    }
    t0 = nowS();
    std::vector<std::thread> th;
    for (int g = 1; g < gpus; g++) th.emplace_back(runDevice, &jobs[g], slice, false);
    runDevice(&jobs[0], slice, true);
    for (auto& t : th) t.join();
    printf("\n");
    if (gpus > 1 && commReady.get() != 0) { fprintf(stderr, "eleven: %s\n", commErr.c_str()); return 1; }
    for (auto& j : jobs) if (j.rc) { fprintf(stderr, "eleven: device %d: %s\n", j.device, j.err.c_str()); return 1; }
    const double wall = nowS() - t0;

    // device 0 holds the film of the whole job (sum over devices of sums and counts): fused resolve -> 8-bit there, one D2H
    const size_t n = (size_t)desc.camera.xRes * desc.camera.yRes;
    std::vector<unsigned char> rgba(n * 4);
    if (eleven_resolve_rgba8_reduced(jobs[0].ctx, ELEVEN_PASS_BEAUTY, rgba.data(), n)) { fprintf(stderr, "eleven: %s\n", eleven_last_error()); return 1; }
    printf("Saving file %s...\n", outPath.c_str());
    if (!writeBmp24(outPath, (int)desc.camera.xRes, (int)desc.camera.yRes, rgba.data(), err)) { fprintf(stderr, "eleven: %s\n", err.c_str()); return 1; }
    if (rawPath) {
        std::vector<float> mean(n * 4);
        if (eleven_get_film_reduced(jobs[0].ctx, ELEVEN_PASS_BEAUTY, mean.data(), n)) { fprintf(stderr, "eleven: %s\n", eleven_last_error()); return 1; }
        FILE* f = fopen(rawPath, "wb"); if (f) { fwrite(mean.data(), 4, mean.size(), f); fclose(f); }
    }
    if (aovPrefix) {                                             // denoiser hand-off: first-hit NORMAL / TANGENT / BITANGENT means over all devices' samples
        const char* names[3] = {"normal", "tangent", "bitangent"};
        const int passes[3] = {ELEVEN_PASS_NORMAL, ELEVEN_PASS_TANGENT, ELEVEN_PASS_BITANGENT};
        std::vector<float> aov(n * 4);
        for (int k = 0; k < 3; k++) {
            if (eleven_get_film_reduced(jobs[0].ctx, passes[k], aov.data(), n)) { fprintf(stderr, "eleven: %s\n", eleven_last_error()); return 1; }
            const std::string path = std::string(aovPrefix) + "_" + names[k] + ".f32";
            FILE* f = fopen(path.c_str(), "wb");
            if (!f) { fprintf(stderr, "eleven: cannot write %s\n", path.c_str()); return 1; }
            fwrite(aov.data(), 4, aov.size(), f); fclose(f);
        }
    }
    printf("Saved!\n");
    double renderMs = 0, reduceMs = 0, uploadS = 0, renderS = 0; uint64_t rays = 0, samples = 0;
    for (auto& j : jobs) {
        renderMs = std::max(renderMs, j.stats.render_ms); reduceMs = std::max(reduceMs, j.stats.reduce_ms);
        uploadS = std::max(uploadS, j.uploadS); renderS = std::max(renderS, j.renderS);
        rays += j.stats.rays_extension + j.stats.rays_shadow_env + j.stats.rays_shadow_light; samples += j.stats.pixel_samples;
    }
    printf("%d spp on %d GPU(s): render %.1f ms (wall incl. upload %.1f ms), %.1f M pixel-samples/s, %.1f Mrays/s, BVH8 %u nodes built in %.1f ms\n",
           spp, gpus, renderMs, wall * 1e3, samples / renderMs / 1e3, rays / renderMs / 1e3, jobs[0].stats.bvh_nodes, jobs[0].stats.bvh_build_ms);
    // the job as a user times it (process start -> picture on disk), phase by phase: what a strong-scaling number must include
    printf("job: {\"gpus\": %d, \"spp\": %d, \"load_s\": %.3f, \"init_s\": %.3f, \"comm_init_s\": %.3f, \"upload_s\": %.3f, \"render_s\": %.3f, \"render_device_ms\": %.1f, \"reduce_device_ms\": %.2f, \"total_s\": %.3f}\n",
           gpus, spp, loadS, initS, commS, uploadS, renderS, renderMs, reduceMs, nowS() - tProcess);
    for (auto& j : jobs) eleven_destroy(j.ctx);
    return 0;
}
