// multi_gpu.cu -- one render job on several GPUs of one node, inside the library: dr_scene_clone, dr_render_multi.
//
// Chains never interact (DRMLTProcess::generateWork hands out seeds, results are only summed, drmlt_proc.cpp:856-883), so the
// path shards into independent units: GPU g bootstraps its own sample range, runs its own chains into a private film.
// Two exchanges, mirroring what the reference does between its initialisation threads and its work units:
//   ONE all-reduce of {sum luminance, sample count} -> the global b  (drmlt.cpp:498-546: the per-thread means are averaged)
//   ONE reduce(sum) of the films onto the first GPU before develop   (DRMLTProcess::processResult, drmlt_proc.cpp:856-867)
// One host thread per GPU, NCCL over NVLink (ncclCommInitAll in-process communicators).  NCCL is resolved with dlopen at
// the first multi-GPU job, so that the library carries no load-time dependency on it (single-GPU hosts, the CPU test image).
#include "scene.h"
#include <dlfcn.h>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace {

// the handful of NCCL entry points used (nccl.h: ncclResult_t = int, 0 = success; ncclDataType_t float32 = 7, float64 = 8; ncclSum = 0)
typedef struct ncclComm *ncclComm_t;
enum { NCCL_SUM = 0, NCCL_F32 = 7, NCCL_F64 = 8 };
struct Nccl {
    int (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Reduce)(const void *, void *, size_t, int, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
};
Nccl &nccl() {
    static Nccl n;
    static std::once_flag once;
    std::call_once(once, [] {
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
        n.CommInitAll = (int (*)(ncclComm_t *, int, const int *)) dlsym(h, "ncclCommInitAll");
        n.CommDestroy = (int (*)(ncclComm_t)) dlsym(h, "ncclCommDestroy");
        n.AllReduce = (int (*)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t)) dlsym(h, "ncclAllReduce");
        n.Reduce = (int (*)(const void *, void *, size_t, int, int, int, ncclComm_t, cudaStream_t)) dlsym(h, "ncclReduce");
        n.GetErrorString = (const char *(*)(int)) dlsym(h, "ncclGetErrorString");
        n.ok = n.CommInitAll && n.CommDestroy && n.AllReduce && n.Reduce;
    });
    return n;
}

struct Rank {
    dr_status status = DR_OK;
    std::string error;
    dr_stats stats;
    double bootSum = 0.0, bootCount = 0.0;
};

} // namespace

// dr_render on `n` GPUs (devices[0] develops the image).  scenes[g] lives on devices[g] (dr_scene_create / dr_scene_clone).
extern "C" dr_status dr_render_multi(const dr_scene *scenes, int32_t n, const dr_config *cfgIn, float *imageRgb, dr_stats *stats) {
    if (!scenes || n <= 0 || !cfgIn || !imageRgb) { dr_set_error("dr_render_multi: bad argument"); return DR_ERR_INVALID_ARG; }
    for (int g = 0; g < n; ++g) if (!scenes[g]) { dr_set_error("dr_render_multi: null scene"); return DR_ERR_INVALID_ARG; }
    if (n == 1) return dr_render(scenes[0], cfgIn, imageRgb, stats);
    if (cfgIn->two_stage && !cfgIn->first_stage && !cfgIn->importance_map) {
        // mltLuminancePass (util.cpp:96-199): the nested low-resolution job, on the first GPU (it is 1 / 16^2 of the pixels);
        // every GPU then renders with the same map
        int32_t W = 0, H = 0;
        dr_status st = dr_film_size(scenes[0], cfgIn, &W, &H);
        if (st) return st;
        std::vector<float> map((size_t) W * H);
        if ((st = dr_importance_map(scenes[0], cfgIn, map.data(), nullptr))) return st;
        dr_config c = *cfgIn;
        c.importance_map = map.data();
        return dr_render_multi(scenes, n, &c, imageRgb, stats);
    }
    Nccl &nc = nccl();
    if (!nc.ok) { dr_set_error("dr_render_multi: NCCL (libnccl.so.2) is not available"); return DR_ERR_UNSUPPORTED; }
    std::vector<int> devs(n);
    for (int g = 0; g < n; ++g) devs[g] = scenes[g]->device;
    std::vector<ncclComm_t> comms(n);
    int rc = nc.CommInitAll(comms.data(), n, devs.data());
    if (rc) { dr_set_error("ncclCommInitAll failed: %s", nc.GetErrorString ? nc.GetErrorString(rc) : "?"); return DR_ERR_CUDA; }

    std::vector<Rank> ranks(n);
    auto work = [&](int g) {
        Rank &R = ranks[g];
        auto good = [&](dr_status s) { if (s == DR_OK) return true; if (R.status == DR_OK) { R.status = s; R.error = dr_last_error(); } return false; };
        cudaSetDevice(devs[g]);
        dr_config cfg = *cfgIn;
        cfg.rank = g; cfg.world_size = n;
        dr_job job = nullptr;
        double *dbuf = nullptr;
        cudaStream_t stream = nullptr;
        // Every rank takes part in both collectives whatever happened to it before (a rank that failed contributes zeros and a
        // raised error flag): nobody is left waiting in a collective.
        bool ok = good(dr_job_create(scenes[g], &cfg, &job));
        if (cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) != cudaSuccess || cudaMalloc((void **) &dbuf, 4 * sizeof(double)) != cudaSuccess) {
            dr_set_error("dr_render_multi: stream / buffer creation failed on device %d", devs[g]);
            ok = good(DR_ERR_CUDA);
        }
        double v[4] = { 0.0, 0.0, 0.0, 0.0 };            // sum, count, error flag
        ok = ok && good(dr_job_bootstrap(job, &v[0], &v[1]));
        if (!ok) { v[0] = v[1] = 0.0; v[2] = 1.0; }
        if (stream && dbuf) {
            cudaMemcpyAsync(dbuf, v, sizeof(v), cudaMemcpyHostToDevice, stream);
            nc.AllReduce(dbuf, dbuf, 4, NCCL_F64, NCCL_SUM, comms[g], stream);
            cudaMemcpyAsync(v, dbuf, sizeof(v), cudaMemcpyDeviceToHost, stream);
            cudaStreamSynchronize(stream);
        }
        R.bootSum = v[0]; R.bootCount = v[1];
        const bool allOk = v[2] == 0.0 && stream && dbuf;
        double b = v[1] > 0.0 ? v[0] / v[1] : 0.0;       // pathsampler.cpp:922-934
        if (cfg.technique == DR_TECH_MMLT) b *= cfg.max_depth;
        float *film = nullptr;
        int64_t nFloats = 0;
        if (ok && allOk) {
            ok = (g != 0 || good(dr_job_direct(job))) && good(dr_job_seed_chains(job, b));
            if (ok) {
                const int64_t per = std::max<int64_t>(1, dr_job_total_mutations(job) / dr_job_num_chains(job));
                ok = good(dr_job_run(job, per)) && good(dr_job_film_device(job, &film, &nFloats));
                if (!ok) film = nullptr;
            }
        }
        if (allOk) {
            // the film reduce: every rank of the all-reduce above enters it (a failed rank with a zero film of the right size)
            float *zero = nullptr;
            if (!film) {
                int32_t W = 0, H = 0;
                dr_film_size(scenes[g], &cfg, &W, &H);
                nFloats = (int64_t) W * H * 4;
                if (cudaMalloc((void **) &zero, (size_t) nFloats * sizeof(float)) == cudaSuccess) cudaMemset(zero, 0, (size_t) nFloats * sizeof(float));
                film = zero;
            }
            if (film) {
                nc.Reduce(film, film, (size_t) nFloats, NCCL_F32, NCCL_SUM, 0, comms[g], stream);
                cudaStreamSynchronize(stream);
            }
            if (ok && g == 0) ok = good(dr_job_develop(job, imageRgb));
            if (zero) cudaFree(zero);
        }
        if (ok) dr_job_stats(job, &R.stats); else memset(&R.stats, 0, sizeof(R.stats));
        if (job) dr_job_destroy(job);
        if (dbuf) cudaFree(dbuf);
        if (stream) cudaStreamDestroy(stream);
    };
    std::vector<std::thread> threads;
    for (int g = 1; g < n; ++g) threads.emplace_back(work, g);
    work(0);
    for (auto &t : threads) t.join();
    for (int g = 0; g < n; ++g) nc.CommDestroy(comms[g]);
    for (int g = 0; g < n; ++g)
        if (ranks[g].status != DR_OK) { dr_set_error("dr_render_multi: GPU %d: %s", devs[g], ranks[g].error.c_str()); return ranks[g].status; }
    if (stats) {
        // counters add up; times are those of the slowest GPU; b is the global one
        dr_stats s = ranks[0].stats;
        uint64_t *acc = reinterpret_cast<uint64_t *>(&s);
        const size_t nCounters = offsetof(dr_stats, luminance) / sizeof(uint64_t);
        for (int g = 1; g < n; ++g) {
            const uint64_t *o = reinterpret_cast<const uint64_t *>(&ranks[g].stats);
            for (size_t i = 0; i < nCounters; ++i) acc[i] += o[i];
            s.bootstrap_ms = std::max(s.bootstrap_ms, ranks[g].stats.bootstrap_ms);
            s.chains_ms = std::max(s.chains_ms, ranks[g].stats.chains_ms);
            s.kernel_launches += ranks[g].stats.kernel_launches;
        }
        *stats = s;
    }
    return DR_OK;
}
