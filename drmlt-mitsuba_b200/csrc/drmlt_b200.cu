// drmlt_b200.cu -- host side of the C ABI (include/drmlt_b200.h): configuration, scene upload,
// bootstrap / chain orchestration, develop and the replay entry points.
//
// Host-side counterparts in the reference: DRMLT::DRMLT / PSSMLT::PSSMLT (parameter parsing,
// src/integrators/drmlt/drmlt.cpp:178-351, src/integrators/pssmlt/pssmlt.cpp:166-308),
// DRMLT::render / PSSMLT::render (work sizing + bootstrap orchestration, drmlt.cpp:393-611) and
// DRMLTProcess::develop (drmlt_proc.cpp:813-854).  There is no CPU fallback: every entry point
// that computes needs a CUDA device and fails with DR_ERR_NO_DEVICE / DR_ERR_CUDA otherwise.
#include "machine.cuh"
#include "util_kernels.h"
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>
#include <chrono>

// ------------------------------------------------------------------ errors
static thread_local char g_error[1024] = "";
void dr_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}
#define CK(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess) {                                                                          \
            dr_set_error("%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);            \
            return DR_ERR_CUDA;                                                                           \
        }                                                                                                 \
    } while (0)
#define CKL() CK(cudaGetLastError())

extern "C" int dr_abi_version(void) { return DR_ABI_VERSION; }
extern "C" const char *dr_last_error(void) { return g_error; }
extern "C" int dr_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// ------------------------------------------------------------------ configuration
extern "C" void dr_config_default(dr_config *c) {
    memset(c, 0, sizeof(*c));
    c->integrator = DR_INTEGRATOR_DRMLT;
    c->technique = -1;                 // required (drmlt.cpp:194-203)
    c->type = -1;                      // required for drmlt (drmlt.cpp:313-322)
    c->max_depth = -1;
    c->rr_depth = 5;
    c->direct_sampling = 1;
    c->direct_samples = 16;
    c->luminance_samples = 100000;
    c->p_large = 0.3f;
    c->work_units = -1;
    c->kelemen_style_weights = 1;
    c->two_stage = 0;
    c->timeout = 0;
    c->average_luminance = -1.0f;
    c->light_image = 1;
    c->acceptance_map = 0;
    c->timid_after_large = 0;
    c->fix_emitter_path = 0;
    c->use_mixture = 0;
    c->sigma = 1.0f / 64.0f;
    c->scale_second = 0.1f;
    c->kelemen_style_mutation = 1;
    c->mutation_size_low = 1.0f / 1024.0f;
    c->mutation_size_high = 1.0f / 64.0f;
    c->sample_count = 4;               // IndependentSampler default sampleCount (samplers/independent.cpp)
    c->rfilter = DR_FILTER_GAUSSIAN;
    c->n_chains = 0;
    c->seed = 0x5eed5eedull;
    c->rank = 0;
    c->world_size = 1;
    c->ray_epsilon = 0.f;
    c->shadow_epsilon = 0.f;
    c->first_stage = 0;
    c->first_stage_size_reduction = 16;
    c->film_width = c->film_height = 0;
    c->crop_offset_x = c->crop_offset_y = 0;
    c->crop_width = c->crop_height = 0;
    c->importance_map = nullptr;
    c->n_lanes = 0;
    c->depth_balance = 1;
}

static bool parse_bool(const char *v, int *out) {
    if (!strcmp(v, "true") || !strcmp(v, "1")) { *out = 1; return true; }
    if (!strcmp(v, "false") || !strcmp(v, "0")) { *out = 0; return true; }
    return false;
}
static bool parse_int(const char *v, long long *out) {
    char *end = nullptr;
    long long x = strtoll(v, &end, 10);
    if (end == v || *end) return false;
    *out = x;
    return true;
}
static bool parse_float(const char *v, float *out) {
    char *end = nullptr;
    double x = strtod(v, &end);
    if (end == v || *end) return false;
    *out = (float) x;
    return true;
}

extern "C" dr_status dr_config_set(dr_config *c, const char *key, const char *value) {
    if (!c || !key || !value) { dr_set_error("dr_config_set: null argument"); return DR_ERR_INVALID_ARG; }
    struct BoolKey { const char *name; int32_t dr_config::*field; };
    struct IntKey { const char *name; int32_t dr_config::*field; };
    struct FloatKey { const char *name; float dr_config::*field; };
    static const BoolKey bools[] = {
        { "directSampling", &dr_config::direct_sampling }, { "kelemenStyleWeights", &dr_config::kelemen_style_weights },
        { "twoStage", &dr_config::two_stage }, { "lightImage", &dr_config::light_image },
        { "acceptanceMap", &dr_config::acceptance_map }, { "timidAfterLarge", &dr_config::timid_after_large },
        { "fixEmitterPath", &dr_config::fix_emitter_path }, { "useMixture", &dr_config::use_mixture },
        { "kelemenStyleMutation", &dr_config::kelemen_style_mutation }, { "firstStage", &dr_config::first_stage },
        { "depthBalance", &dr_config::depth_balance },
    };
    static const IntKey ints[] = {
        { "maxDepth", &dr_config::max_depth }, { "rrDepth", &dr_config::rr_depth }, { "directSamples", &dr_config::direct_samples },
        { "luminanceSamples", &dr_config::luminance_samples }, { "workUnits", &dr_config::work_units },
        { "timeout", &dr_config::timeout }, { "sampleCount", &dr_config::sample_count }, { "chains", &dr_config::n_chains },
        { "rank", &dr_config::rank }, { "worldSize", &dr_config::world_size },
        { "firstStageSizeReduction", &dr_config::first_stage_size_reduction }, { "lanes", &dr_config::n_lanes },
        // film plugin parameters (src/librender/film.cpp:30-48)
        { "width", &dr_config::film_width }, { "height", &dr_config::film_height },
        { "cropOffsetX", &dr_config::crop_offset_x }, { "cropOffsetY", &dr_config::crop_offset_y },
        { "cropWidth", &dr_config::crop_width }, { "cropHeight", &dr_config::crop_height },
    };
    static const FloatKey floats[] = {
        { "pLarge", &dr_config::p_large }, { "averageLuminance", &dr_config::average_luminance }, { "sigma", &dr_config::sigma },
        { "scaleSecond", &dr_config::scale_second }, { "mutationSizeLow", &dr_config::mutation_size_low },
        { "mutationSizeHigh", &dr_config::mutation_size_high }, { "rayEpsilon", &dr_config::ray_epsilon },
        { "shadowEpsilon", &dr_config::shadow_epsilon },
    };
    if (!strcmp(key, "integrator")) {
        if (!strcmp(value, "pssmlt")) c->integrator = DR_INTEGRATOR_PSSMLT;
        else if (!strcmp(value, "drmlt")) c->integrator = DR_INTEGRATOR_DRMLT;
        else { dr_set_error("Unknown integrator \"%s\" (pssmlt|drmlt)", value); return DR_ERR_INVALID_ARG; }
        return DR_OK;
    }
    if (!strcmp(key, "technique")) {   // drmlt.cpp:194-203
        if (!strcmp(value, "path")) c->technique = DR_TECH_PATH;
        else if (!strcmp(value, "bdpt")) c->technique = DR_TECH_BDPT;
        else if (!strcmp(value, "mmlt")) c->technique = DR_TECH_MMLT;
        else { dr_set_error("Unknown technique type"); return DR_ERR_INVALID_ARG; }
        return DR_OK;
    }
    if (!strcmp(key, "type")) {        // drmlt.cpp:313-322 ("mirasym" selects the orbital sampler)
        if (!strcmp(value, "green")) c->type = DR_TYPE_GREEN;
        else if (!strcmp(value, "mira")) c->type = DR_TYPE_MIRA;
        else if (!strcmp(value, "orbital") || !strcmp(value, "mirasym")) c->type = DR_TYPE_ORBITAL;
        else { dr_set_error("Unknown implementation type"); return DR_ERR_INVALID_ARG; }
        return DR_OK;
    }
    if (!strcmp(key, "rfilter")) {
        if (!strcmp(value, "gaussian")) c->rfilter = DR_FILTER_GAUSSIAN;
        else if (!strcmp(value, "box")) c->rfilter = DR_FILTER_BOX;
        else if (!strcmp(value, "tent")) c->rfilter = DR_FILTER_TENT;
        else if (!strcmp(value, "mitchell")) c->rfilter = DR_FILTER_MITCHELL;
        else if (!strcmp(value, "catmullrom")) c->rfilter = DR_FILTER_CATMULLROM;
        else if (!strcmp(value, "lanczos")) c->rfilter = DR_FILTER_LANCZOS;
        else { dr_set_error("Unsupported reconstruction filter \"%s\" (gaussian|box|tent|mitchell|catmullrom|lanczos)", value); return DR_ERR_UNSUPPORTED; }
        return DR_OK;
    }
    if (!strcmp(key, "seed")) {
        char *end = nullptr;
        unsigned long long x = strtoull(value, &end, 0);
        if (end == value || *end) { dr_set_error("seed: not an integer: \"%s\"", value); return DR_ERR_INVALID_ARG; }
        c->seed = x;
        return DR_OK;
    }
    for (const BoolKey &k : bools)
        if (!strcmp(key, k.name)) {
            int v;
            if (!parse_bool(value, &v)) { dr_set_error("%s: not a boolean: \"%s\"", key, value); return DR_ERR_INVALID_ARG; }
            c->*(k.field) = v;
            return DR_OK;
        }
    for (const IntKey &k : ints)
        if (!strcmp(key, k.name)) {
            long long v;
            if (!parse_int(value, &v)) { dr_set_error("%s: not an integer: \"%s\"", key, value); return DR_ERR_INVALID_ARG; }
            c->*(k.field) = (int32_t) v;
            return DR_OK;
        }
    for (const FloatKey &k : floats)
        if (!strcmp(key, k.name)) {
            float v;
            if (!parse_float(value, &v)) { dr_set_error("%s: not a number: \"%s\"", key, value); return DR_ERR_INVALID_ARG; }
            c->*(k.field) = v;
            return DR_OK;
        }
    dr_set_error("Unknown parameter \"%s\"", key);
    return DR_ERR_INVALID_ARG;
}

extern "C" dr_status dr_config_validate(dr_config *c) {
    if (!c) { dr_set_error("dr_config_validate: null"); return DR_ERR_INVALID_ARG; }
    if (c->integrator != DR_INTEGRATOR_PSSMLT && c->integrator != DR_INTEGRATOR_DRMLT) { dr_set_error("Unknown integrator"); return DR_ERR_INVALID_ARG; }
    if (c->technique < DR_TECH_PATH || c->technique > DR_TECH_MMLT) { dr_set_error("Unknown technique type"); return DR_ERR_INVALID_ARG; }
    if (c->integrator == DR_INTEGRATOR_DRMLT && (c->type < DR_TYPE_GREEN || c->type > DR_TYPE_ORBITAL)) {
        dr_set_error("Unknown implementation type"); return DR_ERR_INVALID_ARG;
    }
    if (c->technique == DR_TECH_MMLT && c->max_depth == -1) { dr_set_error("Impossible to use MMLT with no max depth"); return DR_ERR_INVALID_ARG; }
    if (c->technique == DR_TECH_MMLT) { c->direct_sampling = 0; c->kelemen_style_weights = 0; }   // drmlt.cpp:229-231, 266-268
    if (c->fix_emitter_path && c->technique != DR_TECH_MMLT) { dr_set_error("Impossible to use fixEmitterPath without MMLT"); return DR_ERR_INVALID_ARG; }
    if (c->integrator == DR_INTEGRATOR_DRMLT && c->scale_second > 1.0f) { dr_set_error("scaleSecond is bigger than the first stage"); return DR_ERR_INVALID_ARG; }
    if (!(c->p_large >= 0.f && c->p_large <= 1.f)) { dr_set_error("pLarge must be in [0, 1]"); return DR_ERR_INVALID_ARG; }
    if (c->sample_count <= 0) { dr_set_error("sampleCount must be positive"); return DR_ERR_INVALID_ARG; }
    if (c->world_size < 1 || c->rank < 0 || c->rank >= c->world_size) { dr_set_error("rank/worldSize out of range"); return DR_ERR_INVALID_ARG; }
    // limits of the GPU path
    if (c->max_depth <= 0 || c->max_depth + 3 > DR_MAXK) {
        dr_set_error("maxDepth must be in [1, %d] on the GPU path (got %d)", DR_MAXK - 3, c->max_depth); return DR_ERR_UNSUPPORTED;
    }
    if (c->two_stage && c->first_stage_size_reduction <= 0) { dr_set_error("firstStageSizeReduction must be positive"); return DR_ERR_INVALID_ARG; }   // Assert, drmlt.cpp:409
    if (c->film_width < 0 || c->film_height < 0 || c->crop_offset_x < 0 || c->crop_offset_y < 0 || c->crop_width < 0 || c->crop_height < 0) {
        dr_set_error("Invalid crop window specification!"); return DR_ERR_INVALID_ARG;       // film.cpp:44-48
    }
    if (c->technique == DR_TECH_BDPT && c->direct_sampling) {
        // the reference overflows its direct sampler in this mode
        dr_set_error("technique=bdpt requires directSampling=false on the GPU path"); return DR_ERR_UNSUPPORTED;
    }
    if (c->rfilter < DR_FILTER_GAUSSIAN || c->rfilter > DR_FILTER_LANCZOS) { dr_set_error("Unsupported rfilter"); return DR_ERR_UNSUPPORTED; }
    if (c->rfilter == DR_FILTER_TABLE && !(c->filter_radius > 0.0 && std::isfinite(c->filter_radius))) {
        dr_set_error("rfilter: a filter table needs a positive radius"); return DR_ERR_INVALID_ARG;          // Assert(m_radius > 0), rfilter.cpp:38
    }
    // "Hack to detect box filter" (drmlt_proc.cpp:75-79): radius - 0.500010f > 1e-6
    if (c->acceptance_map && c->integrator == DR_INTEGRATOR_DRMLT &&
        !(c->rfilter == DR_FILTER_BOX || (c->rfilter == DR_FILTER_TABLE && !(c->filter_radius - (double) 0.500010f > 1e-6)))) {
        dr_set_error("Box filter required for acceptance map!"); return DR_ERR_INVALID_ARG;
    }
    return DR_OK;
}

// hasRoughDielectric: some shape's BSDF draws a third number per sample (offsetRoughDielectric, pssmlt_utils.h:35-52)
static void max_dimensions(const dr_config *c, int depth, int *se, int *em, int *di, bool hasRoughDielectric = false) {   // pssmlt_utils.h:27-77
    const int offsetRR = (c->rr_depth < c->max_depth ? 1 : 0) + (hasRoughDielectric ? 1 : 0);
    int m;
    if (c->technique == DR_TECH_MMLT) { m = (depth + 2) * 3; if (m & 1) m++; *se = m; *em = m; *di = 1; }
    else if (c->technique == DR_TECH_PATH) { m = (c->max_depth + 2) * (4 + offsetRR); if (m & 1) m++; *se = m; *em = 0; *di = 0; }
    else { m = (c->max_depth + 2) * (2 + offsetRR); if (m & 1) m++; *se = m; *em = m; *di = c->direct_sampling ? c->max_depth : 0; }
}
extern "C" void dr_max_dimensions(const dr_config *cfg, int depth, int *sensor, int *emitter, int *direct) {
    max_dimensions(cfg, depth, sensor, emitter, direct);
}

// ------------------------------------------------------------------ scene
struct HostUpload { void *dev; char *host; size_t bytes; };   // host: pinned staging copy (kept for dr_scene_reupload)

struct SceneImpl : dr_scene_t {
    std::vector<HostUpload> uploads;     // host staging copies (kept for dr_scene_reupload)
    unsigned int *dOrder = nullptr;      // leaf order -> caller's triangle index
    uint32_t nTextures = 0;
    size_t uploadBytes = 0;
};

template <class T>
static dr_status upload(SceneImpl *s, const std::vector<T> &v, const T **devOut) {
    HostUpload u;
    const size_t bytes = std::max<size_t>(v.size() * sizeof(T), 16);
    u.bytes = bytes; u.host = nullptr;                       // (the pinned staging copy is made by the first dr_scene_reupload)
    CK(cudaMalloc(&u.dev, bytes));
    s->allocations.push_back(u.dev);
    CK(cudaMemset(u.dev, 0, bytes));
    if (!v.empty()) CK(cudaMemcpy(u.dev, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    s->uploads.push_back(u);
    s->bytes += bytes;
    s->uploadBytes += bytes;
    *devOut = (const T *) u.dev;
    return DR_OK;
}
// an array that was produced on the device (build_scene_gpu): the scene takes ownership
template <class T>
static void adopt(SceneImpl *s, T *dev, size_t count, const T **devOut, bool own = true) {
    HostUpload u;
    u.bytes = std::max<size_t>(count * sizeof(T), 16); u.host = nullptr; u.dev = dev;
    if (own) s->allocations.push_back(u.dev);
    s->uploads.push_back(u);
    s->bytes += u.bytes;
    s->uploadBytes += u.bytes;
    *devOut = dev;
}
// pinned host copies of the scene arrays, for dr_scene_reupload (made on first use: a one-shot render never needs them)
static dr_status ensure_staging(SceneImpl *s) {
    for (HostUpload &u : s->uploads)
        if (!u.host) {
            CK(cudaMallocHost((void **) &u.host, u.bytes));
            CK(cudaMemcpy(u.host, u.dev, u.bytes, cudaMemcpyDeviceToHost));
        }
    return DR_OK;
}

static float as_float_bits(int i) { float f; memcpy(&f, &i, 4); return f; }
// texture-coordinate bits of a triangle record (scene.h): the mesh has texture coordinates | it carries UV tangents
static uint32_t tri_uv_bits(const dr_scene_desc *d, size_t prim) {
    const uint32_t f = d->tri_flags ? d->tri_flags[prim] : 0u;
    if (!d->texcoords || (f & DR_TRI_NO_TEXCOORDS)) return 0u;
    return DR_MF_HAS_UV | ((f & DR_TRI_UV_TANGENTS) ? DR_MF_UV_TANGENTS : 0u);
}

// fresnelDiffuseReflectance(eta, fast = false) (src/libcore/util.cpp:815-867): integral of the unpolarised Fresnel
// reflectance F(sqrt(xi), eta) over xi in [0, 1].  The reference integrates with an adaptive Gauss-Lobatto rule to a relative
// error of 1e-5; here: xi = x^2, composite Simpson in x (the integrand's kink at the critical angle is resolved by the 2^16
// intervals), error < 1e-8.
static double fresnel_dielectric_host(double cosThetaI, double eta) {        // util.cpp:659-693
    if (eta == 1.0) return 0.0;
    const double scale = cosThetaI > 0 ? 1.0 / eta : eta;
    const double cosThetaTSqr = 1.0 - (1.0 - cosThetaI * cosThetaI) * (scale * scale);
    if (cosThetaTSqr <= 0.0) return 1.0;
    const double ci = std::fabs(cosThetaI), ct = std::sqrt(cosThetaTSqr);
    const double Rs = (ci - eta * ct) / (ci + eta * ct), Rp = (eta * ci - ct) / (eta * ci + ct);
    return 0.5 * (Rs * Rs + Rp * Rp);
}
static double fresnel_diffuse_reflectance(double eta) {
    const int n = 1 << 16;
    const double h = 1.0 / n;
    auto f = [&](double x) { return fresnel_dielectric_host(x, eta) * 2.0 * x; };
    double s = f(0.0) + f(1.0);
    for (int i = 1; i < n; ++i) s += f(i * h) * ((i & 1) ? 4.0 : 2.0);
    return s * h / 3.0;
}

extern "C" void dr_scene_destroy(dr_scene scene) {
    if (!scene) return;
    SceneImpl *s = static_cast<SceneImpl *>(scene);
    cudaSetDevice(s->device);
    for (void *p : s->allocations) cudaFree(p);
    for (HostUpload &u : s->uploads) if (u.host) cudaFreeHost(u.host);
    delete s;
}

extern "C" dr_status dr_scene_create(const dr_scene_desc *d, int device, dr_scene *out) {
    // (DRMLT_BVH=gpu|host overrides the default builder of this entry point: a tuning aid)
    const char *env = getenv("DRMLT_BVH");
    return dr_scene_create_ex(d, device, env && !strcmp(env, "gpu") ? DR_SCENE_BVH_GPU : DR_SCENE_BVH_HOST, out);
}
extern "C" void dr_scene_bvh_info(dr_scene scene, int32_t *builder, int32_t *nNodes, int32_t *stackBound, double *buildMs) {
    if (!scene) return;
    if (builder) *builder = scene->bvhBuilder;
    if (nNodes) *nNodes = (int32_t) scene->nNodes;
    if (stackBound) *stackBound = scene->bvhDepth;
    if (buildMs) *buildMs = scene->bvhBuildMs;
}
extern "C" dr_status dr_scene_create_ex(const dr_scene_desc *d, int device, uint32_t flags, dr_scene *out) {
    if (!d || !out) { dr_set_error("dr_scene_create: null argument"); return DR_ERR_INVALID_ARG; }
    *out = nullptr;
    if (!d->positions || !d->indices || !d->tri_material || !d->tri_emitter || !d->materials || d->n_triangles == 0 ||
        d->n_vertices == 0 || d->n_materials == 0) {
        dr_set_error("dr_scene_create: empty scene or missing buffers"); return DR_ERR_INVALID_ARG;
    }
    if (d->n_triangles >= (1u << 29)) { dr_set_error("dr_scene_create: too many triangles"); return DR_ERR_UNSUPPORTED; }
    if (d->n_materials >= (1u << 24)) { dr_set_error("dr_scene_create: too many materials"); return DR_ERR_UNSUPPORTED; }
    if (d->n_emitters && !d->emitters) { dr_set_error("dr_scene_create: emitters missing"); return DR_ERR_INVALID_ARG; }
    if (d->camera.film_width <= 0 || d->camera.film_height <= 0) { dr_set_error("dr_scene_create: bad film size"); return DR_ERR_INVALID_ARG; }
    for (uint32_t i = 0; i < d->n_triangles; ++i) {
        if (d->tri_material[i] >= d->n_materials) { dr_set_error("triangle %u: material index out of range", i); return DR_ERR_INVALID_ARG; }
        if (d->tri_emitter[i] >= (int32_t) d->n_emitters) { dr_set_error("triangle %u: emitter index out of range", i); return DR_ERR_INVALID_ARG; }
        for (int v = 0; v < 3; ++v)
            if (d->indices[3 * (size_t) i + v] >= d->n_vertices) { dr_set_error("triangle %u: vertex index out of range", i); return DR_ERR_INVALID_ARG; }
    }
    for (uint32_t m = 0; m < d->n_materials; ++m)
    {
        const dr_material &mat = d->materials[m];
        if (mat.type < DR_BSDF_DIFFUSE || mat.type > DR_BSDF_ROUGHPLASTIC) {
            dr_set_error("material %u: unsupported BSDF type %d", m, mat.type); return DR_ERR_UNSUPPORTED;
        }
        if (mat.type == DR_BSDF_ROUGHPLASTIC) {
            if (!d->rough_tables || mat.table >= d->n_rough_tables) {
                dr_set_error("material %u: roughplastic needs its rough-transmittance table (dr_scene_desc.rough_tables[%u])", m, mat.table);
                return DR_ERR_INVALID_ARG;
            }
            // MicrofacetDistribution + RoughPlastic (roughplastic.cpp:210-230): eta != 1; the Phong distribution is not on this path
            if (!(mat.eta[0] > 0.f) || mat.eta[0] == 1.f) { dr_set_error("The interior and exterior indices of refraction must be positive and differ!"); return DR_ERR_INVALID_ARG; }
        }
    }
    // bitmap textures (ABI 6)
    if (d->n_textures > DR_MAX_TEXTURES || (d->n_textures && !d->textures)) { dr_set_error("dr_scene_create: bad texture table"); return DR_ERR_INVALID_ARG; }
    for (uint32_t t = 0; t < d->n_textures; ++t) {
        const dr_texture &tx = d->textures[t];
        if (!tx.texels || tx.width == 0 || tx.height == 0 || tx.width > (1u << 15) || tx.height > (1u << 15) || tx.wrap_u > DR_WRAP_ONE || tx.wrap_v > DR_WRAP_ONE) {
            dr_set_error("texture %u: missing texels, bad size or wrap mode", t); return DR_ERR_INVALID_ARG;
        }
    }
    for (uint32_t m = 0; m < d->n_materials; ++m) {
        const uint32_t tr = (d->materials[m].flags >> 8) & 0xfffu, tt = d->materials[m].flags >> 20;
        if (tr > d->n_textures || tt > d->n_textures) { dr_set_error("material %u: texture index out of range", m); return DR_ERR_INVALID_ARG; }
    }
    if (!d->texcoords && d->tri_flags)
        for (uint32_t i = 0; i < d->n_triangles; ++i)
            if (d->tri_flags[i] & DR_TRI_UV_TANGENTS) { dr_set_error("triangle %u: DR_TRI_UV_TANGENTS without texcoords", i); return DR_ERR_INVALID_ARG; }
    if (d->tri_flags)
        for (uint32_t i = 0; i < d->n_triangles; ++i)
            if ((d->tri_flags[i] & DR_TRI_UV_TANGENTS) && (d->tri_flags[i] & DR_TRI_NO_TEXCOORDS)) { dr_set_error("triangle %u: DR_TRI_UV_TANGENTS on a mesh without texcoords", i); return DR_ERR_INVALID_ARG; }
    for (uint32_t e = 0; e < d->n_emitters; ++e) {
        const dr_emitter &em = d->emitters[e];
        if (em.n_tris == 0 || (uint64_t) em.first_tri + em.n_tris > d->n_triangles) { dr_set_error("emitter %u: triangle range out of bounds", e); return DR_ERR_INVALID_ARG; }
        for (uint32_t t = 0; t < em.n_tris; ++t)
            if (d->tri_emitter[em.first_tri + t] != (int32_t) e) { dr_set_error("emitter %u: tri_emitter mismatch at triangle %u", e, em.first_tri + t); return DR_ERR_INVALID_ARG; }
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        dr_set_error("no CUDA device available (there is no CPU fallback)");
        return DR_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= ndev) { dr_set_error("device %d out of range (%d devices)", device, ndev); return DR_ERR_INVALID_ARG; }
    CK(cudaSetDevice(device));

    SceneImpl *s = new SceneImpl();
    s->device = device;
    s->filmW = d->camera.film_width; s->filmH = d->camera.film_height;
    s->nTris = d->n_triangles;
    for (uint32_t i = 0; i < d->n_triangles; ++i) s->typeMask |= 1u << d->materials[d->tri_material[i]].type;
    dr_status st = DR_OK;
    auto fail = [&](dr_status code) { dr_scene_destroy(s); return code; };

    const size_t nT = d->n_triangles;
    bool anySmooth = false;
    for (size_t i = 0; i < nT; ++i)
        if (d->normals && d->tri_flags && (d->tri_flags[i] & DR_TRI_SMOOTH)) { anySmooth = true; break; }
    BuiltBVH bvh;
    GpuScene gpu;
    struct GpuGuard { GpuScene &g; bool adopted = false; ~GpuGuard() { if (!adopted) g.release(); } } gpuGuard{ gpu };
    const auto tBuild = std::chrono::steady_clock::now();
    bool gpuBuilt = false;
#ifdef DR_BVH4
    if ((flags & DR_SCENE_BVH_GPU) && d->n_triangles >= 1024) {
        bool tooDeep = false;
        gpuBuilt = build_scene_gpu(d, anySmooth, DR_STACK, gpu, &tooDeep);
        if (!gpuBuilt && !tooDeep) return fail(DR_ERR_CUDA);
        if (gpuBuilt) { bvh.maxDepth = gpu.stackBound; s->nNodes = gpu.nNodes; }
    }
#endif
    if (!gpuBuilt) {
        build_bvh(d->positions, d->indices, d->n_triangles, bvh);
#ifdef DR_BVH4
        collapse_bvh4(bvh);
#endif
    }
    s->bvhBuilder = gpuBuilt ? (int) DR_SCENE_BVH_GPU : (int) DR_SCENE_BVH_HOST;
    s->bvhDepth = bvh.maxDepth;
    s->bvhBuildMs = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tBuild).count();
    if (bvh.maxDepth >= DR_STACK) { dr_set_error("dr_scene_create: BVH depth %d exceeds the traversal stack (%d)", bvh.maxDepth, DR_STACK); return fail(DR_ERR_UNSUPPORTED); }
    if (!gpuBuilt) s->nNodes = (uint32_t) (bvh.nodes.size() / DR_NODE_F4);

    // triangles and shading normals in leaf order (device build: packed by build_scene_gpu)
    std::vector<float4> tris(gpuBuilt ? 0 : 3 * nT), normals;
    if (anySmooth && !gpuBuilt) normals.assign(3 * nT, make_float4(0, 0, 0, 0));
    auto P = [&](uint32_t v) { return f3(d->positions[3 * (size_t) v], d->positions[3 * (size_t) v + 1], d->positions[3 * (size_t) v + 2]); };
    auto N = [&](uint32_t v) { return f3(d->normals[3 * (size_t) v], d->normals[3 * (size_t) v + 1], d->normals[3 * (size_t) v + 2]); };
    for (size_t slot = 0; slot < (gpuBuilt ? 0 : nT); ++slot) {
        const uint32_t prim = bvh.order[slot];
        const uint32_t i0 = d->indices[3 * (size_t) prim], i1 = d->indices[3 * (size_t) prim + 1], i2 = d->indices[3 * (size_t) prim + 2];
        const float3 p0 = P(i0), p1 = P(i1), p2 = P(i2);
        const bool smooth = anySmooth && (d->tri_flags[prim] & DR_TRI_SMOOTH);
        // material index | BSDF model << 24 (routes the hit to its walk queue) | smooth << 31
        const uint32_t mf = d->tri_material[prim] | ((uint32_t) d->materials[d->tri_material[prim]].type << 24) | (smooth ? 0x80000000u : 0u) |
                            tri_uv_bits(d, prim);
        tris[3 * slot] = make_float4(p0.x, p0.y, p0.z, p1.x);
        tris[3 * slot + 1] = make_float4(p1.y, p1.z, p2.x, p2.y);
        tris[3 * slot + 2] = make_float4(p2.z, as_float_bits((int) prim), as_float_bits((int) mf), as_float_bits(d->tri_emitter[prim]));
        if (smooth) {
            const float3 n0 = N(i0), n1 = N(i1), n2 = N(i2);
            normals[3 * slot] = make_float4(n0.x, n0.y, n0.z, n1.x);
            normals[3 * slot + 1] = make_float4(n1.y, n1.z, n2.x, n2.y);
            normals[3 * slot + 2] = make_float4(n2.z, 0, 0, 0);
        }
    }
    // emitters: triangles in emitter order + area CDFs (trimesh.cpp:405-420, pmf.h)
    std::vector<float4> emTris;
    std::vector<double> emCdf, emitterCdf;
    std::vector<DevEmitter> emitters(d->n_emitters);
    double weightSum = 0.0;
    for (uint32_t e = 0; e < d->n_emitters; ++e) weightSum += (double) d->emitters[e].sampling_weight;
    emitterCdf.push_back(0.0);
    for (uint32_t e = 0; e < d->n_emitters; ++e) {
        const dr_emitter &em = d->emitters[e];
        DevEmitter &de = emitters[e];
        memset(&de, 0, sizeof(de));
        de.firstEmTri = (uint32_t) (emTris.size() / 6);
        de.nTris = em.n_tris;
        de.cdfOffset = (uint32_t) emCdf.size();
        const size_t cdfStart = emCdf.size();
        emCdf.push_back(0.0);
        for (uint32_t t = 0; t < em.n_tris; ++t) {
            const uint32_t prim = em.first_tri + t;
            const uint32_t i0 = d->indices[3 * (size_t) prim], i1 = d->indices[3 * (size_t) prim + 1], i2 = d->indices[3 * (size_t) prim + 2];
            const float3 p0 = P(i0), p1 = P(i1), p2 = P(i2);
            // area in double from the float positions (Triangle::surfaceArea, triangle.cpp:62-68)
            const double ax = (double) p1.x - p0.x, ay = (double) p1.y - p0.y, az = (double) p1.z - p0.z;
            const double bx = (double) p2.x - p0.x, by = (double) p2.y - p0.y, bz = (double) p2.z - p0.z;
            const double cx = ay * bz - az * by, cy = az * bx - ax * bz, cz = ax * by - ay * bx;
            emCdf.push_back(emCdf.back() + 0.5 * std::sqrt(cx * cx + cy * cy + cz * cz));
            const bool smooth = anySmooth && (d->tri_flags[prim] & DR_TRI_SMOOTH);
            emTris.push_back(make_float4(p0.x, p0.y, p0.z, p1.x));
            emTris.push_back(make_float4(p1.y, p1.z, p2.x, p2.y));
            emTris.push_back(make_float4(p2.z, as_float_bits(smooth ? 1 : 0), 0, 0));
            if (smooth) {
                const float3 n0 = N(i0), n1 = N(i1), n2 = N(i2);
                emTris.push_back(make_float4(n0.x, n0.y, n0.z, n1.x));
                emTris.push_back(make_float4(n1.y, n1.z, n2.x, n2.y));
                emTris.push_back(make_float4(n2.z, 0, 0, 0));
            } else {
                for (int k = 0; k < 3; ++k) emTris.push_back(make_float4(0, 0, 0, 0));
            }
        }
        const double area = emCdf.back();
        if (!(area > 0.0)) { dr_set_error("emitter %u has zero area", e); return fail(DR_ERR_INVALID_ARG); }
        for (size_t k = cdfStart; k < emCdf.size(); ++k) emCdf[k] *= 1.0 / area;   // DiscreteDistribution::normalize
        emCdf.back() = 1.0;
        de.radiance[0] = em.radiance[0]; de.radiance[1] = em.radiance[1]; de.radiance[2] = em.radiance[2];
        de.area = area; de.invArea = 1.0 / area;
        de.pdfDiscrete = weightSum > 0.0 ? (double) em.sampling_weight * (1.0 / weightSum) : 0.0;
        emitterCdf.push_back(emitterCdf.back() + (double) em.sampling_weight);
    }
    if (d->n_emitters && weightSum > 0.0) {
        for (double &v : emitterCdf) v *= 1.0 / weightSum;
        emitterCdf.back() = 1.0;
    }
    std::vector<DevMaterial> mats(d->n_materials);
    static_assert(sizeof(DevMaterial) == sizeof(dr_material), "material layout");
    memcpy(mats.data(), d->materials, sizeof(dr_material) * d->n_materials);
    for (DevMaterial &m : mats)
        if (m.type == DR_BSDF_PLASTIC) {                    // SmoothPlastic::configure (plastic.cpp:188-205)
            m.k[0] = (float) fresnel_diffuse_reflectance(1.0 / (double) m.eta[0]);
            const double Y[3] = { 0.212671f, 0.715160f, 0.072169f };   // float literals, as spectrum.h:734-736
            double dAvg = 0, sAvg = 0;
            for (int c = 0; c < 3; ++c) { dAvg += Y[c] * m.reflectance[c]; sAvg += Y[c] * m.transmittance[c]; }
            m.k[1] = (float) (sAvg / (dAvg + sAvg));
        }
    // roughplastic: the caller's tables + m_specularSamplingWeight (roughplastic.cpp:273-277) in slot [102], in double
    std::vector<double> roughTables(d->rough_tables ? d->rough_tables : nullptr,
                                    d->rough_tables ? d->rough_tables + (size_t) d->n_rough_tables * DR_ROUGH_TABLE_DOUBLES : nullptr);
    {
        std::vector<int> owner(d->n_rough_tables, -1);
        for (uint32_t i = 0; i < d->n_materials; ++i) {
            const DevMaterial &m = mats[i];
            if (m.type != DR_BSDF_ROUGHPLASTIC) continue;
            if (owner[m.table] >= 0) { dr_set_error("materials %d and %u share rough table %u (one table per roughplastic material)", owner[m.table], i, m.table); return fail(DR_ERR_INVALID_ARG); }
            owner[m.table] = (int) i;
            const double Y[3] = { 0.212671f, 0.715160f, 0.072169f };
            double dAvg = 0, sAvg = 0;
            for (int c = 0; c < 3; ++c) { dAvg += Y[c] * m.reflectance[c]; sAvg += Y[c] * m.transmittance[c]; }
            roughTables[(size_t) m.table * DR_ROUGH_TABLE_DOUBLES + 102] = sAvg / (dAvg + sAvg);
        }
    }

    DevScene &ds = s->dev;
    memset(&ds, 0, sizeof(ds));
    const unsigned int *dOrder = nullptr;
    if (gpuBuilt) {                          // (the same order of arrays as the host build: dr_scene_clone walks them)
        gpuGuard.adopted = true;
        s->allocations.push_back(gpu.order);               // (owned from here on; listed with the uploads below)
        adopt(s, gpu.nodes, 8 * (size_t) gpu.nNodes, &ds.nodes);
        adopt(s, gpu.tris, 3 * nT, &ds.tris);
        adopt(s, gpu.normals, gpu.normalsCount, &ds.normals);
    } else if ((st = upload(s, bvh.nodes, &ds.nodes)) || (st = upload(s, tris, &ds.tris)) || (st = upload(s, normals, &ds.normals)))
        return fail(st);
    if ((st = upload(s, emTris, &ds.emTris)) || (st = upload(s, emCdf, &ds.emCdf)) || (st = upload(s, emitterCdf, &ds.emitterCdf)) ||
        (st = upload(s, emitters, &ds.emitters)) || (st = upload(s, mats, &ds.materials)) || (st = upload(s, roughTables, &ds.roughTables)))
        return fail(st);
    if (gpuBuilt) adopt(s, (unsigned int *) gpu.order, nT, &dOrder, false);
    else if ((st = upload(s, bvh.order, &dOrder))) return fail(st);
    s->dOrder = const_cast<unsigned int *>(dOrder);
    {   // texture coordinates in leaf order, texel pool, texture table (16-byte placeholders when the scene has none)
        std::vector<float4> uvs, texels;
        std::vector<DevTexture> textures(d->n_textures);
        if (d->texcoords) {
            std::vector<uint32_t> orderHost;
            if (gpuBuilt) {
                orderHost.resize(nT);
                if (cudaMemcpy(orderHost.data(), gpu.order, nT * sizeof(uint32_t), cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); dr_set_error("dr_scene_create: reading the leaf order failed"); return fail(DR_ERR_CUDA); }
            }
            const uint32_t *order = gpuBuilt ? orderHost.data() : bvh.order.data();
            uvs.resize(2 * nT);
            for (size_t slot = 0; slot < nT; ++slot) {
                const uint32_t prim = order[slot];
                const float *t0 = d->texcoords + 2 * (size_t) d->indices[3 * (size_t) prim], *t1 = d->texcoords + 2 * (size_t) d->indices[3 * (size_t) prim + 1],
                            *t2 = d->texcoords + 2 * (size_t) d->indices[3 * (size_t) prim + 2];
                uvs[2 * slot] = make_float4(t0[0], t0[1], t1[0], t1[1]);
                uvs[2 * slot + 1] = make_float4(t2[0], t2[1], 0.f, 0.f);
            }
        }
        for (uint32_t t = 0; t < d->n_textures; ++t) {
            const dr_texture &tx = d->textures[t];
            DevTexture &dt = textures[t];
            memset(&dt, 0, sizeof(dt));
            dt.w = tx.width; dt.h = tx.height; dt.wrapU = tx.wrap_u; dt.wrapV = tx.wrap_v; dt.nearest = tx.nearest ? 1u : 0u;
            dt.first = texels.size();
            dt.scaleU = tx.uv_scale[0]; dt.scaleV = tx.uv_scale[1]; dt.offU = tx.uv_offset[0]; dt.offV = tx.uv_offset[1];
            const size_t n = (size_t) tx.width * tx.height;
            texels.resize(dt.first + n);
            for (size_t i = 0; i < n; ++i) texels[dt.first + i] = make_float4(tx.texels[3 * i], tx.texels[3 * i + 1], tx.texels[3 * i + 2], 0.f);
        }
        if ((st = upload(s, uvs, &ds.uvs)) || (st = upload(s, texels, &ds.texels)) || (st = upload(s, textures, &ds.textures))) return fail(st);
        s->nTextures = d->n_textures;
    }
    ds.nEmitters = (int) d->n_emitters; ds.nTris = (int) d->n_triangles; ds.nNodes = (int) s->nNodes; ds.rootIsLeaf = 0;
    ds.epsilon = 1e-4f; ds.shadowEpsilon = 1e-3f;     // constants.h:29-30 (single precision)

    // pinhole camera (perspective.cpp:126-173)
    const dr_camera &c = d->camera;
    DevCamera &dc = ds.cam;
    for (int r = 0; r < 3; ++r) for (int k = 0; k < 4; ++k) dc.m[4 * r + k] = (double) c.to_world[4 * r + k];
    {   // world -> camera directions: Transform::inverse (transform.h) of the caller's float matrix, which is a rotation only to ~6e-8
        const double *a = dc.m;
        const double c00 = a[5] * a[10] - a[6] * a[9], c01 = a[6] * a[8] - a[4] * a[10], c02 = a[4] * a[9] - a[5] * a[8];
        const double det = a[0] * c00 + a[1] * c01 + a[2] * c02;
        if (!(std::fabs(det) > 1e-12)) { dr_set_error("dr_scene_create: camera to_world is singular"); return fail(DR_ERR_INVALID_ARG); }
        const double id = 1.0 / det;
        dc.inv[0] = c00 * id; dc.inv[1] = (a[2] * a[9] - a[1] * a[10]) * id; dc.inv[2] = (a[1] * a[6] - a[2] * a[5]) * id;
        dc.inv[3] = c01 * id; dc.inv[4] = (a[0] * a[10] - a[2] * a[8]) * id; dc.inv[5] = (a[2] * a[4] - a[0] * a[6]) * id;
        dc.inv[6] = c02 * id; dc.inv[7] = (a[1] * a[8] - a[0] * a[9]) * id; dc.inv[8] = (a[0] * a[5] - a[1] * a[4]) * id;
    }
    dc.pos[0] = c.to_world[3]; dc.pos[1] = c.to_world[7]; dc.pos[2] = c.to_world[11];
    dc.dir[0] = c.to_world[2]; dc.dir[1] = c.to_world[6]; dc.dir[2] = c.to_world[10];
    const double tanHalf = std::tan(0.5 * (double) c.xfov_deg * 3.14159265358979323846 / 180.0);
    dc.tanHalf = tanHalf;
    dc.nearClip = c.near_clip; dc.farClip = c.far_clip;
    camera_set_window(dc, c.film_width, c.film_height, 0, 0, c.film_width, c.film_height);
    *out = s;
    return DR_OK;
}

// A replica of a scene on another GPU (multi-GPU jobs hold one replica per device): the device buffers are filled from the
// pinned staging copies the original keeps -- no second BVH build, no second flattening.
extern "C" dr_status dr_scene_clone(dr_scene scene, int device, dr_scene *out) {
    if (!scene || !out) { dr_set_error("dr_scene_clone: null argument"); return DR_ERR_INVALID_ARG; }
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); dr_set_error("no CUDA device available (there is no CPU fallback)"); return DR_ERR_NO_DEVICE; }
    if (device < 0 || device >= ndev) { dr_set_error("device %d out of range (%d devices)", device, ndev); return DR_ERR_INVALID_ARG; }
    const SceneImpl *src = static_cast<const SceneImpl *>(scene);
    CK(cudaSetDevice(device));
    SceneImpl *s = new SceneImpl();
    s->device = device; s->filmW = src->filmW; s->filmH = src->filmH; s->nTris = src->nTris; s->nNodes = src->nNodes; s->typeMask = src->typeMask; s->nTextures = src->nTextures;
    s->dev = src->dev;
    auto fail = [&](dr_status code) { dr_scene_destroy(s); return code; };
    // the device pointers of DevScene, in the order dr_scene_create uploaded them
    const void **slots[13] = { (const void **) &s->dev.nodes, (const void **) &s->dev.tris, (const void **) &s->dev.normals, (const void **) &s->dev.emTris,
                              (const void **) &s->dev.emCdf, (const void **) &s->dev.emitterCdf, (const void **) &s->dev.emitters, (const void **) &s->dev.materials,
                              (const void **) &s->dev.roughTables, (const void **) &s->dOrder,
                              (const void **) &s->dev.uvs, (const void **) &s->dev.texels, (const void **) &s->dev.textures };
    if (src->uploads.size() != 13) { dr_set_error("dr_scene_clone: unexpected scene layout"); return fail(DR_ERR_INVALID_ARG); }
    for (size_t i = 0; i < src->uploads.size(); ++i) {
        HostUpload u = src->uploads[i];
        u.host = nullptr;                                   // the staging copy stays with the original
        if (cudaMalloc(&u.dev, u.bytes) != cudaSuccess) { dr_set_error("dr_scene_clone: cudaMalloc of %zu bytes failed", u.bytes); cudaGetLastError(); return fail(DR_ERR_CUDA); }
        s->allocations.push_back(u.dev);
        if (cudaMemcpyPeer(u.dev, device, src->uploads[i].dev, src->device, u.bytes) != cudaSuccess) { dr_set_error("dr_scene_clone: copy from device %d failed", src->device); cudaGetLastError(); return fail(DR_ERR_CUDA); }
        *slots[i] = u.dev;
        s->bytes += u.bytes;
    }
    *out = s;
    return DR_OK;
}

// Repeat the host->device copies of the flattened scene (what a plugin pays per render job).
extern "C" dr_status dr_scene_reupload(dr_scene scene, int64_t *bytes) {
    if (!scene) { dr_set_error("dr_scene_reupload: null scene"); return DR_ERR_INVALID_ARG; }
    SceneImpl *s = static_cast<SceneImpl *>(scene);
    CK(cudaSetDevice(s->device));
    dr_status st = ensure_staging(s);
    if (st) return st;
    for (HostUpload &u : s->uploads) CK(cudaMemcpyAsync(u.dev, u.host, u.bytes, cudaMemcpyHostToDevice, 0));
    CK(cudaStreamSynchronize(0));
    if (bytes) *bytes = (int64_t) s->uploadBytes;
    return DR_OK;
}

extern "C" void dr_cancel(dr_scene scene) { if (scene) scene->cancel = 1; }

// ------------------------------------------------------------------ launch parameters
// Fills the constant part of a Machine from the configuration.  `evalDims` (JOB_EVAL): the coordinate
// buffers are laid out for the replayed vectors' own dimensions.
static void make_params(const dr_config &c, int W, int H, double b, const int *evalDims, Machine &p, unsigned typeMask = 0) {
    p.pc.technique = c.technique; p.pc.maxDepth = c.max_depth; p.pc.rrDepth = c.rr_depth;
    p.pc.hasRoughDielectric = (typeMask >> DR_BSDF_ROUGHDIELECTRIC) & 1u;
    p.pc.excludeDirect = c.direct_samples >= 0;        // separateDirect (drmlt.cpp:242)
    p.pc.lightImage = c.light_image != 0;
    p.pc.bdBatch = 0;                                  // (alloc_lanes decides)
    PssParams &pp = p.pp;
    memset(&pp, 0, sizeof(pp));
    pp.seed = c.seed; pp.integrator = c.integrator; pp.type = c.type;
    const double s1 = 1.0 / 1024.0, s2 = 1.0 / 64.0;   // drmlt_sampler.h:201-202
    const double scale = (c.integrator == DR_INTEGRATOR_DRMLT && c.type == DR_TYPE_ORBITAL) ? (double) 1.9f : 1.0;   // :203-205
    pp.kel_s2 = s2 * scale;
    pp.kel_logRatio = -std::log((s2 * scale) / (s1 * scale));
    pp.sigma2 = (double) c.scale_second * (double) c.sigma;
    const double rho = std::exp(-0.25f);   // a FLOAT exponential, as m_rho (drmlt_sampler.h:204)
    pp.cauchy_disp = 2.0 * rho / (1.0 + rho * rho);
    pp.pss_kelemen = c.kelemen_style_mutation;
    pp.pss_s2 = c.mutation_size_high;
    pp.pss_logRatio = -std::log((double) c.mutation_size_high / (double) c.mutation_size_low);
    pp.pss_sigma = c.sigma;
    pp.identity1 = pp.identity2 = 0;
    if (c.integrator == DR_INTEGRATOR_DRMLT && c.technique == DR_TECH_MMLT) {
        pp.identity1 = 1u << SMP_DIRECT;                // setStagesToIdentity (drmlt_proc.cpp:133-136)
        if (c.fix_emitter_path) pp.identity2 = 1u << SMP_EMITTER;   // handleLightTracing (:137-140)
    }
    // subset mode (pss.cuh): exact whenever the strategy coordinate can only change in a large step and no
    // stage reads the current state after such a change
    pp.subset = c.integrator == DR_INTEGRATOR_DRMLT && c.technique == DR_TECH_MMLT && !(c.type == DR_TYPE_GREEN && c.timid_after_large);
    ChainParams &cp = p.cp;
    memset(&cp, 0, sizeof(cp));
    cp.pLarge = c.p_large; cp.b = b;
    cp.acceptanceMap = c.integrator == DR_INTEGRATOR_DRMLT && c.acceptance_map;
    cp.timidAfterLarge = c.timid_after_large; cp.fixEmitterPath = c.fix_emitter_path; cp.useMixture = c.use_mixture;
    cp.kelemenWeights = c.kelemen_style_weights;
    cp.kel_s1 = s1; cp.kel_s2 = s2; cp.kel_logRatio = -std::log(s2 / s1);
    max_dimensions(&c, c.max_depth, &cp.dimS, &cp.dimE, &cp.dimD, p.pc.hasRoughDielectric != 0);   // worst case over the MMLT depths
    int lay[3] = { cp.dimS, cp.dimE, cp.dimD };
    if (evalDims) for (int s = 0; s < 3; ++s) lay[s] = std::max(lay[s], evalDims[s]);
    int off = 0;
    for (int s = 0; s < 3; ++s) { pp.off[s] = off; off += (lay[s] + 1) & ~1; }
    pp.nU = std::max(off, 2);
    // reconstruction filter: radius + the 32-entry table of ReconstructionFilter::configure (rfilter.cpp:37-55) -- from the caller
    // (DR_FILTER_TABLE) or from the plugins' eval functions with their default parameters (src/rfilters/{gaussian,box,tent,
    // mitchell,catmullrom,lanczos}.cpp)
    FilmParams &fp = p.fp;
    fp.w = W; fp.h = H;
    double vals[32], radius;
    if (c.rfilter == DR_FILTER_TABLE) {
        radius = c.filter_radius;
        for (int i = 0; i < 32; ++i) vals[i] = c.filter_table[i];
    } else {
        const double stddev = 0.5;
        switch (c.rfilter) {
            case DR_FILTER_BOX: radius = 0.5 + (double) 1e-5f; break;             // Float 0.5 + a float literal (box.cpp:38)
            case DR_FILTER_TENT: radius = 1.0; break;
            case DR_FILTER_LANCZOS: radius = 3.0; break;                             // lobes = 3
            default: radius = 2.0; break;                                            // gaussian 4 stddev, mitchell, catmullrom
        }
        auto cubic = [](double x, double B, double C) {                              // mitchell.cpp:55-69, catmullrom.cpp:43-58
            x = std::fabs(x);
            const double x2 = x * x, x3 = x2 * x;
            if (x < 1) return (double) (1.0f / 6.0f) * ((12 - 9 * B - 6 * C) * x3 + (-18 + 12 * B + 6 * C) * x2 + (6 - 2 * B));
            if (x < 2) return (double) (1.0f / 6.0f) * ((-B - 6 * C) * x3 + (6 * B + 30 * C) * x2 + (-12 * B - 48 * C) * x + (8 * B + 24 * C));
            return 0.0;
        };
        double sum = 0.0;
        for (int i = 0; i < 31; ++i) {
            const double x = (radius * i) / 31;
            double v;
            switch (c.rfilter) {
                case DR_FILTER_BOX: v = std::fabs(x) <= radius ? 1.0 : 0.0; break;
                case DR_FILTER_TENT: v = std::max(0.0, 1.0 - std::fabs(x / radius)); break;
                case DR_FILTER_MITCHELL: v = cubic(x, (double) (1.0f / 3.0f), (double) (1.0f / 3.0f)); break;   // props.getFloat("B", 1.0f / 3.0f)
                case DR_FILTER_CATMULLROM: v = cubic(x, 0.0, 0.5); break;
                case DR_FILTER_LANCZOS: {                                          // lanczos.cpp:44-57 (Epsilon of the double build)
                    const double ax = std::fabs(x);
                    if (ax < 1e-7) v = 1.0;
                    else if (ax > radius) v = 0.0;
                    else { const double x1 = M_PI * ax, x2 = x1 / radius; v = (std::sin(x1) * std::sin(x2)) / (x1 * x2); }
                    break;
                }
                default: { const double alpha = -1.0 / (2.0 * stddev * stddev); v = std::max(0.0, std::exp(alpha * x * x) - std::exp(alpha * radius * radius)); }
            }
            vals[i] = v; sum += v;
        }
        vals[31] = 0.0;
        sum *= 2 * radius / 31;
        const double normalization = 1.0 / sum;      // multiplied in, as rfilter.cpp:52-54
        for (int i = 0; i < 31; ++i) vals[i] *= normalization;
    }
    for (int i = 0; i < 32; ++i) fp.values[i] = (float) vals[i];
    fp.radius = (float) radius; fp.scaleFactor = (float) (31 / radius);
}

// Film and crop window of a job (Film::Film, src/librender/film.cpp:30-48): the configuration may override the
// camera's film size (the nested first-stage pass does) and select a crop window; every image buffer has the crop size.
struct FilmWindow { int filmW, filmH, cropX, cropY, W, H; };
static dr_status film_window(const dr_scene scene, const dr_config &c, FilmWindow &fw) {
    fw.filmW = c.film_width > 0 ? c.film_width : scene->filmW;
    fw.filmH = c.film_height > 0 ? c.film_height : scene->filmH;
    fw.cropX = c.crop_offset_x; fw.cropY = c.crop_offset_y;
    fw.W = c.crop_width > 0 ? c.crop_width : fw.filmW;
    fw.H = c.crop_height > 0 ? c.crop_height : fw.filmH;
    if (fw.cropX < 0 || fw.cropY < 0 || fw.W <= 0 || fw.H <= 0 || fw.cropX + fw.W > fw.filmW || fw.cropY + fw.H > fw.filmH) {
        dr_set_error("Invalid crop window specification!"); return DR_ERR_INVALID_ARG;
    }
    return DR_OK;
}

static DevScene scene_for(const dr_scene scene, const dr_config &c, const FilmWindow &fw) {
    DevScene ds = scene->dev;
    if (c.ray_epsilon > 0.f) ds.epsilon = c.ray_epsilon;
    if (c.shadow_epsilon > 0.f) ds.shadowEpsilon = c.shadow_epsilon;
    camera_set_window(ds.cam, fw.filmW, fw.filmH, fw.cropX, fw.cropY, fw.W, fw.H);
    return ds;
}

static dr_status check_technique(const dr_config &c) {
    if (c.technique == DR_TECH_BDPT && c.max_depth + 2 > BD_MAXV) { dr_set_error("technique=bdpt: maxDepth must be <= %d", BD_MAXV - 2); return DR_ERR_UNSUPPORTED; }
    return DR_OK;
}

// tuning knobs of the launch geometry (read per call, so a sweep can change them between jobs of one process)
int stage_ctas_per_sm() { const char *e = getenv("DRMLT_STAGE_CTAS"); return e ? std::max(1, atoi(e)) : 16; }
int trace_ctas_per_sm() { const char *e = getenv("DRMLT_TRACE_CTAS"); return e ? std::max(0, atoi(e)) : 0; }

// ------------------------------------------------------------------ job
enum { STAGE_TRACE = 0, STAGE_WALK, STAGE_CHAIN, STAGE_COUNT };
struct dr_job_t;
static void resample_chains(dr_job_t *j, unsigned long long firstChain);

struct dr_job_t {
    dr_scene scene = nullptr;
    int device = 0;
    dr_config cfg;
    int W = 0, H = 0;                           // size of the rendered image = crop window of the film
    float *importance = nullptr;                // [W*H] two-stage importance map on the device (null: none)
    Machine M;                                  // constant part: scene, parameters, lane memory, queues
    int *depth = nullptr;                       // [n] MMLT depth of every chain (or -1)
    unsigned long long *chainId = nullptr, *seedIdx = nullptr;

    float4 *film = nullptr;
    unsigned long long *counters = nullptr;     // [0, ST_COUNT): chain phase, [ST_COUNT, 2 ST_COUNT): bootstrap
    // Wavefront groups: the lanes are split into independent groups, each with its own work queues and stream, so
    // that the latency-bound tail of one group's stage kernels overlaps with the other groups' kernels.
    struct Group { Queues q; cudaStream_t stream = nullptr; cudaEvent_t evJoin = nullptr; uint32_t *countsHost = nullptr; int begin = 0, end = 0; };
    cudaEvent_t evFork = nullptr;
    cudaGraphExec_t graphExec = nullptr;        // the captured rounds (run_machine), updated in place by later runs
    std::vector<Group> groups;
    float *bootLum = nullptr;
    float *bootLumTarget = nullptr;             // two-stage MLT: importance-re-weighted luminances, the seed CDF is built on these
    double *cdf = nullptr, *blockSums = nullptr, *red = nullptr, *redScratch = nullptr;
    float *devImage = nullptr;
    float4 *directFilm = nullptr;               // weighted film of the separate direct-illumination pass
    float *directImage = nullptr;               // ... developed (null until dr_job_direct ran)
    bool haveDirect = false;
    // timeout (drmlt.cpp:296, drmlt_proc.cpp:519-521, 870-877): the chain phase stops issuing rounds after `timeout` seconds
    std::chrono::steady_clock::time_point deadline;
    bool hasDeadline = false, timedOut = false;
    long long nBoot = 0;
    unsigned long long bootFirst = 0;
    const double *replay = nullptr;             // dr_chain_replay: the chains' uniform tables on the device
    long long replayStride = 0;
    int replayDim = 0;
    int nChains = 0;                            // Markov chains of the job (the reference's work units)
    int nLanes = 0;                             // lanes of the wavefront machine; < nChains: the lanes pull chains from a queue
    unsigned int *chainCursor = nullptr;        // work-unit queue: next chain index
    uint32_t epoch = 0;                         // work-unit queue: batches run so far (chain ids of batch e start at e * nChains * worldSize)
    long long totalMutations = 0, mutationsDone = 0;
    uint32_t mutTarget = 0;
    // depth-balanced MMLT chains (dr_config.depth_balance): per-depth mutation factors, seed weights and bootstrap sums
    uint32_t *mutScale = nullptr;               // [256] 16.16 fixed point, indexed by path depth
    float *depthWeight = nullptr;               // [maxDepth] seed weight of depth d + 1
    double *depthSums = nullptr;                // [32]
    bool balanced = false;
    double b = 0.0;
    bool bootstrapped = false, seeded = false;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double bootstrapMs = 0.0, chainsMs = 0.0, totalMs = 0.0, directMs = 0.0;
    uint64_t launches = 0, rounds = 0;
    std::vector<void *> allocations;
    int roundsPerPoll = 16;
    // optional per-stage device timing (dr_job_profile): events around every stage of every round
    bool profile = false;
    std::vector<cudaEvent_t> profEvents;
    double stageMs[STAGE_COUNT] = { 0, 0, 0 };
    uint64_t stageLaunches[STAGE_COUNT] = { 0, 0, 0 };
};

// Device memory of a job comes from the device's stream-ordered pool (cudaMallocAsync) with an unbounded release
// threshold: a long-lived process (the plugin renders frame after frame) pays for the ~7 GB of lane memory once.
// Only what is read before it is written gets cleared (`zero`).
static void pool_init(int device) {
    static bool done[64] = { false };
    if (device < 0 || device >= 64 || done[device]) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long threshold = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
    }
    cudaGetLastError();
    done[device] = true;
}
template <class T>
static dr_status job_alloc(dr_job j, T **p, size_t count, bool zero = false) {
    const cudaError_t e = cudaMallocAsync((void **) p, std::max<size_t>(count * sizeof(T), 16), j->stream);
    if (e != cudaSuccess) {
        size_t fr = 0, tot = 0;
        cudaMemGetInfo(&fr, &tot);
        dr_set_error("cudaMalloc of %zu bytes failed: %s (%zu of %zu bytes free)", count * sizeof(T), cudaGetErrorString(e), fr, tot);
        cudaGetLastError();
        return DR_ERR_CUDA;
    }
    j->allocations.push_back(*p);
    if (zero) CK(cudaMemsetAsync(*p, 0, std::max<size_t>(count * sizeof(T), 16), j->stream));
    return DR_OK;
}

#ifdef DR_FILM_MATCH_STATS
void film_match_stats_print();
#endif
extern "C" void dr_job_destroy(dr_job j) {
    if (!j) return;
#ifdef DR_FILM_MATCH_STATS
    if (j->stream) cudaStreamSynchronize(j->stream);
    film_match_stats_print();
#endif
    cudaSetDevice(j->device);                   // (not j->scene->device: a garbage collector may already have destroyed the scene)
    if (j->stream) cudaStreamSynchronize(j->stream);
    for (void *p : j->allocations) cudaFreeAsync(p, j->stream);
    if (j->stream) cudaStreamSynchronize(j->stream);
    for (auto &g : j->groups) { if (g.countsHost) cudaFreeHost(g.countsHost); if (g.evJoin) cudaEventDestroy(g.evJoin); if (g.stream) cudaStreamDestroy(g.stream); }
    if (j->graphExec) cudaGraphExecDestroy(j->graphExec);
    if (j->evFork) cudaEventDestroy(j->evFork);
    for (cudaEvent_t e : j->profEvents) cudaEventDestroy(e);
    if (j->ev0) cudaEventDestroy(j->ev0);
    if (j->ev1) cudaEventDestroy(j->ev1);
    if (j->stream) cudaStreamDestroy(j->stream);
    cudaGetLastError();
    delete j;
}

static int auto_chains(long long totalMutations) {
    // enough lanes to keep every stage kernel's queue several waves deep on 148 SMs, but >= 64 mutations per chain:
    // at equal mutation count, chains shorter than that lose statistical efficiency (profiles/r01_g_chain_length_study.json)
    long long n = totalMutations / 64;
    n = std::max<long long>(4096, std::min<long long>(n, 1 << 22));
    n = std::min<long long>(n, std::max<long long>(128, totalMutations));
    return (int) ((n + 127) / 128 * 128);
}

// lane memory and work queues of the wavefront machine for `n` lanes
static dr_status alloc_lanes(dr_job j, int n) {
    LaneMem &lm = j->M.lm;
    memset(&lm, 0, sizeof(lm));
    lm.n = n; lm.nU = j->M.pp.nU;
    // coordinate buffers: X and Y always, Z for drmlt's second stage, R for Green's reverse state; MIS records: one per walk step
    lm.ubCount = j->cfg.integrator != DR_INTEGRATOR_DRMLT ? 2 : (j->cfg.type == DR_TYPE_GREEN ? 4 : 3);
    lm.mrSlots = std::min((int) MR_MAXV, j->cfg.max_depth + 3);
    dr_status st;
    if ((st = job_alloc(j, &lm.core, (size_t) n)) || (st = job_alloc(j, &lm.vt, (size_t) n)) || (st = job_alloc(j, &lm.vs, (size_t) n)) ||
        (st = job_alloc(j, &lm.geo, (size_t) 4 * n)) || (st = job_alloc(j, &lm.chain, (size_t) n, true)) ||
        (st = job_alloc(j, &lm.misrec, (size_t) 2 * lm.mrSlots * MR_WORDS * n)) || (st = job_alloc(j, &lm.conn, (size_t) 4 * n)) || (st = job_alloc(j, &lm.ubuf, (size_t) lm.ubCount * lm.nU * n)) ||
        (st = job_alloc(j, &lm.rayd, (size_t) 8 * n)))
        return st;
    if (j->cfg.technique == DR_TECH_PATH && ((st = job_alloc(j, &lm.rayd2, (size_t) 8 * n)) || (st = job_alloc(j, &lm.neeOcc, (size_t) n, true)))) return st;
    if (j->cfg.technique == DR_TECH_BDPT &&      // both subpaths and the splat lists are kept per lane
        ((st = job_alloc(j, &lm.bv, (size_t) 2 * BD_MAXV * n)) || (st = job_alloc(j, &lm.bx, (size_t) 2 * BD_MAXV * n)) ||
         (st = job_alloc(j, &lm.bacc, (size_t) n)) || (st = job_alloc(j, &lm.bsplat, (size_t) 4 * BD_MAXS * 2 * n))))
        return st;
    // batched connections: all (s, t) pairs of a path in one round (k_bdpt.cu); DRMLT_BD_BATCH=0 keeps one pair per round
    lm.bdStride = (j->cfg.max_depth + 1) * (j->cfg.max_depth + 2) / 2;
    const bool bdBatch = j->cfg.technique == DR_TECH_BDPT && lm.bdStride <= BD_MAXC && !(getenv("DRMLT_BD_BATCH") && atoi(getenv("DRMLT_BD_BATCH")) == 0);
    j->M.pc.bdBatch = bdBatch ? 1 : 0;
    if (bdBatch &&
        ((st = job_alloc(j, &lm.bconn, (size_t) lm.bdStride * n)) || (st = job_alloc(j, &lm.brayd, (size_t) 8 * lm.bdStride * n)) ||
         (st = job_alloc(j, &lm.bvis, (size_t) n, true)) || (st = job_alloc(j, &lm.bpend, (size_t) n, true)) || (st = job_alloc(j, &lm.bcount, (size_t) n, true))))
        return st;
    // groups of about 1 M lanes -- the persistent traversal kernel needs several rays per resident thread to amortise
    // its latency-bound tail, while >= 3 groups are needed for the stages of different groups to overlap (measured at
    // 4 M lanes: 1 group 110 M mutations/s, 2: 136, 3: 144, 4: 146, 6: 143, 8: 140; at 8 M lanes 8 groups: 148) --
    // but never fewer than 64K lanes per group, at most 8 groups (DRMLT_GROUPS overrides)
    int G = std::max(3, std::min(8, n >> 20));
    G = std::max(1, std::min(G, n / 65536));
    if (getenv("DRMLT_GROUPS")) G = std::max(1, std::min(64, atoi(getenv("DRMLT_GROUPS"))));
    G = std::min(G, std::max(1, n / 32));
    j->groups.resize(G);
    for (int g = 0; g < G; ++g) {
        dr_job_t::Group &gr = j->groups[g];
        gr.begin = (int) ((long long) n * g / G); gr.end = (int) ((long long) n * (g + 1) / G);
        gr.q.n = gr.end - gr.begin;
        if ((st = job_alloc(j, &gr.q.items, (size_t) Q_COUNT * gr.q.n)) || (st = job_alloc(j, &gr.q.rays, (size_t) 4 * 2 * gr.q.n)) || (st = job_alloc(j, &gr.q.aux, (size_t) Q_COUNT * gr.q.n, true)) || (st = job_alloc(j, &gr.q.count, (size_t) Q_COUNT + 2, true)))   // + head counters of the two ray queues
            return st;
        gr.q.bn = bdBatch ? gr.q.n * lm.bdStride : 0;
        gr.q.bitems = nullptr; gr.q.brays = nullptr;
        if (bdBatch && ((st = job_alloc(j, &gr.q.bitems, (size_t) 2 * gr.q.bn)) || (st = job_alloc(j, &gr.q.brays, (size_t) 2 * 2 * gr.q.bn)))) return st;
        if (cudaStreamCreateWithFlags(&gr.stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreateWithFlags(&gr.evJoin, cudaEventDisableTiming) != cudaSuccess || cudaMallocHost((void **) &gr.countsHost, sizeof(uint32_t) * Q_COUNT) != cudaSuccess) {
            dr_set_error("group stream creation failed: %s", cudaGetErrorString(cudaGetLastError())); return DR_ERR_CUDA;
        }
    }
    trace_init();
    if (cudaEventCreateWithFlags(&j->evFork, cudaEventDisableTiming) != cudaSuccess) { dr_set_error("event creation failed"); return DR_ERR_CUDA; }
    return DR_OK;
}

static dr_status job_create_common(dr_scene scene, const dr_config *cfgIn, int nLanes, bool chains, const int *evalDims, dr_job *out) {
    if (!scene || !cfgIn || !out) { dr_set_error("dr_job_create: null argument"); return DR_ERR_INVALID_ARG; }
    *out = nullptr;
    dr_config cfg = *cfgIn;
    dr_status st = dr_config_validate(&cfg);
    if (st) return st;
    if ((st = check_technique(cfg))) return st;
    CK(cudaSetDevice(scene->device));
    dr_job j = new dr_job_t();
    j->scene = scene; j->device = scene->device; j->cfg = cfg;
    memset(&j->M, 0, sizeof(j->M));
    FilmWindow fw;
    if ((st = film_window(scene, cfg, fw))) { delete j; return st; }
    j->M.sc = scene_for(scene, cfg, fw);
    const int W = fw.W, H = fw.H;
    j->W = W; j->H = H;
    make_params(cfg, W, H, 1.0, evalDims, j->M, static_cast<SceneImpl *>(scene)->typeMask);
    j->M.traceRefill = getenv("DRMLT_TRACE_REFILL") ? atoi(getenv("DRMLT_TRACE_REFILL")) : 12;   // measured with the 4-wide BVH: 8 / 12 / 16 / 24 / 28 -> k_trace 924 / 910 / 910 / 932 / 935 ms
    j->M.traceDescend = getenv("DRMLT_TRACE_DESCEND") ? atoi(getenv("DRMLT_TRACE_DESCEND")) : 8;
    auto fail = [&](dr_status code) { dr_job_destroy(j); return code; };
    if (cudaStreamCreateWithFlags(&j->stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreate(&j->ev0) != cudaSuccess ||
        cudaEventCreate(&j->ev1) != cudaSuccess) {
        dr_set_error("stream/event creation failed: %s", cudaGetErrorString(cudaGetLastError())); return fail(DR_ERR_CUDA);
    }
    pool_init(scene->device);
    j->profile = getenv("DRMLT_PROFILE_STAGES") != nullptr;
    // this rank's share of W*H*sampleCount mutations (drmlt.cpp:475-476)
    const long long total = (long long) W * H * cfg.sample_count;
    j->totalMutations = total / cfg.world_size + (cfg.rank < total % cfg.world_size ? 1 : 0);
    j->nChains = nLanes > 0 ? nLanes : (cfg.n_chains > 0 ? cfg.n_chains : auto_chains(j->totalMutations));
    // lanes: explicit (`lanes`), else one per chain.  With fewer lanes than chains the lanes pull chains from a queue.
    j->nLanes = j->nChains;
    if (nLanes <= 0 && cfg.n_lanes > 0 && cfg.n_lanes < j->nChains) j->nLanes = std::max(128, (cfg.n_lanes + 127) / 128 * 128);
    if (j->nLanes > j->nChains) j->nLanes = j->nChains;
    const size_t n = (size_t) j->nChains;
    if (j->nLanes < j->nChains && (st = job_alloc(j, &j->chainCursor, 1, true))) return fail(st);
    if ((st = alloc_lanes(j, j->nLanes)) || (st = job_alloc(j, &j->counters, (size_t) 2 * ST_COUNT, true)) || (st = job_alloc(j, &j->red, 4, true)) ||
        (st = job_alloc(j, &j->redScratch, (size_t) lum_reduce_scratch_doubles())))
        return fail(st);
    if (chains) {
        if ((st = job_alloc(j, &j->depth, n)) || (st = job_alloc(j, &j->chainId, n)) || (st = job_alloc(j, &j->seedIdx, n)) ||
            (st = job_alloc(j, &j->film, (size_t) W * H, true)) || (st = job_alloc(j, &j->devImage, (size_t) W * H * 3)))
            return fail(st);
        if (cfg.importance_map && !cfg.first_stage) {      // m_config.importanceMap: splats are divided by it (pathsampler.cpp:1001-1020)
            if ((st = job_alloc(j, &j->importance, (size_t) W * H))) return fail(st);
            if (cudaMemcpyAsync(j->importance, cfg.importance_map, sizeof(float) * (size_t) W * H, cudaMemcpyHostToDevice, j->stream) != cudaSuccess) {
                dr_set_error("importance map upload failed: %s", cudaGetErrorString(cudaGetLastError())); return fail(DR_ERR_CUDA);
            }
            j->M.cp.importance = j->importance;
            if (cfg.integrator == DR_INTEGRATOR_PSSMLT) j->M.cp.kelemenWeights = 0;    // pssmlt_proc.cpp:205
        }
    }
    j->cfg.importance_map = nullptr;                       // the host buffer is not referenced after creation
    if (cudaStreamSynchronize(j->stream) != cudaSuccess) { dr_set_error("dr_job_create: %s", cudaGetErrorString(cudaGetLastError())); return fail(DR_ERR_CUDA); }
    *out = j;
    return DR_OK;
}

extern "C" dr_status dr_job_create(dr_scene scene, const dr_config *cfg, dr_job *out) { return job_create_common(scene, cfg, 0, true, nullptr, out); }

static Machine machine_for(dr_job j, int group, const JobParams &job, unsigned long long *counters, bool withFilm) {
    Machine M = j->M;
    const dr_job_t::Group &gr = j->groups[group];
    M.q = gr.q; M.laneBegin = gr.begin; M.laneEnd = gr.end;
    M.job = job; M.film = withFilm ? j->film : nullptr; M.counters = counters; M.parity = (int) (j->rounds & 1);
    return M;
}
// the job's main stream waits for everything queued on the group streams, and vice versa
static dr_status join_groups(dr_job j) {
    for (auto &g : j->groups) CK(cudaStreamSynchronize(g.stream));
    return DR_OK;
}

// Run the wavefront machine until no lane has work left.  One round = trace (closest, shadow) -> walk per BSDF
// model + connect (MMLT) | path tracer (technique=path) -> chain.  The queue counters are polled every
// `roundsPerPoll` rounds through a pinned mirror.
// One round of one group on stream `st`.
static void launch_round(dr_job j, Machine &M, int g, cudaStream_t st, bool mmlt, unsigned typeMask, bool marks) {
    LaunchCfg lc; lc.stream = st; lc.nLanes = j->groups[g].end - j->groups[g].begin;
    auto mark = [&]() { if (marks) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); j->profEvents.push_back(e); } };
    mark();
    launch_trace(M, lc);
    mark();
    if (mmlt) launch_walk(M, lc, typeMask); else if (j->cfg.technique == DR_TECH_BDPT) launch_bdpt(M, lc); else launch_pt(M, lc);
    mark();
    launch_chain(M, lc);
    mark();
}

// Run the wavefront machine until no lane has work left.  One round = trace (closest, shadow) -> walk per BSDF
// model + connect (MMLT) | path tracer (technique=path) -> chain, per group.  `roundsPerPoll` rounds of all groups
// are captured ONCE into a CUDA graph (one branch per group, so the groups' kernels overlap) and replayed until the
// queue counters, copied to a pinned mirror at the end of every replay, show that every queue is empty.
static dr_status run_machine(dr_job j, const JobParams &job, unsigned long long *counters, bool withFilm) {
    const int G = (int) j->groups.size();
    std::vector<Machine> Ms;
    for (int g = 0; g < G; ++g) Ms.push_back(machine_for(j, g, job, counters, withFilm));
    const bool mmlt = j->cfg.technique == DR_TECH_MMLT;
    const unsigned typeMask = static_cast<SceneImpl *>(j->scene)->typeMask;
    const int walkLaunches = mmlt || j->M.pc.bdBatch ? 2 : 1;
    if (const char *e = getenv("DRMLT_ROUNDS_PER_POLL")) {      // tuning aid: rounds per graph replay = per host poll of the queue counters
        const int r = atoi(e);
        if (r >= 2 && r <= 512) j->roundsPerPoll = r & ~1;
    }
    const int R = j->roundsPerPoll;                             // even: a replay starts at the parity it was captured with
    CK(cudaStreamSynchronize(j->stream));                       // everything queued on the main stream is visible to the groups
    cudaStream_t s0 = j->groups[0].stream;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    const bool useGraph = !j->profile && !getenv("DRMLT_NO_GRAPH");
    // the executable graph lives with the job: every later run (bootstrap, seed replay, each dr_job_run, every progressive
    // slice) captures the same topology with other kernel parameters and UPDATES it in place instead of instantiating anew
    auto cleanup = [&]() { if (graph) cudaGraphDestroy(graph); };
    if (useGraph) {
        const uint64_t round0 = j->rounds;
        CK(cudaStreamBeginCapture(s0, cudaStreamCaptureModeThreadLocal));
        cudaEventRecord(j->evFork, s0);
        for (int g = 1; g < G; ++g) cudaStreamWaitEvent(j->groups[g].stream, j->evFork, 0);
        for (int r = 0; r < R; ++r)
            for (int g = 0; g < G; ++g) {
                Ms[g].parity = (int) ((round0 + r) & 1);
                launch_round(j, Ms[g], g, j->groups[g].stream, mmlt, typeMask, false);
            }
        for (int g = 0; g < G; ++g)
            cudaMemcpyAsync(j->groups[g].countsHost, j->groups[g].q.count, sizeof(uint32_t) * Q_COUNT, cudaMemcpyDeviceToHost, j->groups[g].stream);
        for (int g = 1; g < G; ++g) { cudaEventRecord(j->groups[g].evJoin, j->groups[g].stream); cudaStreamWaitEvent(s0, j->groups[g].evJoin, 0); }
        bool ready = cudaStreamEndCapture(s0, &graph) == cudaSuccess;
        if (ready && j->graphExec) {
            cudaGraphExecUpdateResultInfo info;
            if (cudaGraphExecUpdate(j->graphExec, graph, &info) != cudaSuccess) {      // (another topology: start over)
                cudaGetLastError();
                cudaGraphExecDestroy(j->graphExec);
                j->graphExec = nullptr;
            }
        }
        if (ready && !j->graphExec) ready = cudaGraphInstantiate(&j->graphExec, graph, 0) == cudaSuccess;
        if (!ready) {
            dr_set_error("CUDA graph capture of the wavefront rounds failed: %s", cudaGetErrorString(cudaGetLastError()));
            cleanup();
            return DR_ERR_CUDA;
        }
        exec = j->graphExec;
    }
    for (;;) {
        if (j->scene->cancel) { join_groups(j); cleanup(); dr_set_error("cancelled"); return DR_ERR_CANCELLED; }
        if (job.type == JOB_CHAIN && j->hasDeadline && std::chrono::steady_clock::now() >= j->deadline) { j->timedOut = true; break; }
        if (useGraph) {
            if (cudaGraphLaunch(exec, s0) != cudaSuccess || cudaStreamSynchronize(s0) != cudaSuccess) {
                dr_set_error("wavefront rounds failed: %s", cudaGetErrorString(cudaGetLastError()));
                cleanup();
                return DR_ERR_CUDA;
            }
            j->rounds += R;
        } else {
            for (int r = 0; r < R; ++r) {
                for (int g = 0; g < G; ++g) {
                    Ms[g].parity = (int) (j->rounds & 1);
                    // stage profiling serialises the groups on one stream so that the event intervals are kernel times
                    launch_round(j, Ms[g], g, j->profile ? s0 : j->groups[g].stream, mmlt, typeMask, j->profile);
                }
                ++j->rounds;
            }
            CKL();
            for (int g = 0; g < G; ++g)
                CK(cudaMemcpyAsync(j->groups[g].countsHost, j->groups[g].q.count, sizeof(uint32_t) * Q_COUNT, cudaMemcpyDeviceToHost,
                                   j->profile ? s0 : j->groups[g].stream));
            dr_status st = join_groups(j);
            if (st) return st;
        }
        for (int g = 0; g < G; ++g) j->launches += (uint64_t) R * (1 + walkLaunches + (chain_begin_fused(j->groups[g].end - j->groups[g].begin, j->cfg.integrator) ? 1 : 2));
        if (!j->profEvents.empty()) {
            for (size_t i = 0; i + 3 < j->profEvents.size(); i += 4)
                for (int s = 0; s < STAGE_COUNT; ++s) {
                    float ms = 0.f;
                    cudaEventElapsedTime(&ms, j->profEvents[i + s], j->profEvents[i + s + 1]);
                    j->stageMs[s] += ms;
                }
            j->stageLaunches[STAGE_TRACE] += 1ull * R * G; j->stageLaunches[STAGE_WALK] += (uint64_t) walkLaunches * R * G;
            j->stageLaunches[STAGE_CHAIN] += (uint64_t) R * G;
            for (cudaEvent_t e : j->profEvents) cudaEventDestroy(e);
            j->profEvents.clear();
        }
        const int p = (int) (j->rounds & 1);                      // queues the next round would consume
        bool busy = false;
        for (int g = 0; g < G; ++g) {
            const uint32_t *c = j->groups[g].countsHost;
            busy |= c[Q_RAYC + p] != 0 || c[Q_RAYS + p] != 0 || c[Q_CHAIN + p] != 0 || c[Q_BDS + p] != 0;
        }
        if (!busy) break;
    }
    cleanup();
    return DR_OK;
}

static dr_status setup_lanes(dr_job j, const JobParams &job) {
    CK(cudaStreamSynchronize(j->stream));
    for (int g = 0; g < (int) j->groups.size(); ++g) {
        Machine M = machine_for(j, g, job, j->counters, false);
        LaunchCfg lc; lc.stream = j->groups[g].stream; lc.nLanes = M.laneEnd - M.laneBegin;
        CK(cudaMemsetAsync(M.q.count, 0, sizeof(uint32_t) * (Q_COUNT + 2), lc.stream));
        launch_setup(M, lc, j->depth, j->chainId, j->seedIdx);
        ++j->launches;
    }
    CKL();
    return DR_OK;
}

// luminanceSamples sizing of DRMLT::render (drmlt.cpp:446-473), with the reference's CPU work-unit count standing in
// for "workUnits", and a floor of one bootstrap sample per four resident chains (of all ranks) so that the seed pool
// does not degenerate when hundreds of thousands of chains are resampled from it, and of 1/8 of the mutation budget.
static long long bootstrap_samples(const dr_job j) {
    const dr_config &c = j->cfg;
    const long long desired = c.technique == DR_TECH_PATH ? 200000 : 100000;
    const long long total = (long long) j->W * j->H * c.sample_count;
    long long workUnits = c.work_units > 0 ? c.work_units : std::max<long long>(1, (desired - 1 + total) / desired);
    long long n = c.luminance_samples;
    const long long times = c.technique == DR_TECH_MMLT ? 50 : 10;
    n = std::max(n, workUnits * times);
    n = std::max(n, (long long) j->nChains * c.world_size / 4);
    if (c.technique == DR_TECH_MMLT) n *= c.max_depth;
    // The image is scaled by b, so the relative error of b is a floor of the image error.  On the GPU a bootstrap path
    // costs about as much as a mutation: spend 1/8 of the mutation budget on it (profiles/r01_g_chain_length_study.json:
    // 16 M instead of 0.8 M bootstrap paths lower relMSE 4-10x at 59 M mutations for +0.2 s).
    const long long div = getenv("DRMLT_BOOT_DIV") ? std::max(1, atoi(getenv("DRMLT_BOOT_DIV"))) : 8;   // (tuning aid)
    n = std::max(n, std::min<long long>(total / div, 1ll << 26));   // (capped: 64 M samples = 0.8 GB of luminances + CDF per GPU)
    return n;
}

extern "C" dr_status dr_job_bootstrap(dr_job j, double *sumOut, double *countOut) {
    if (!j) { dr_set_error("dr_job_bootstrap: null job"); return DR_ERR_INVALID_ARG; }
    CK(cudaSetDevice(j->scene->device));
    const dr_config &c = j->cfg;
    const long long all = bootstrap_samples(j);
    long long per = (all + c.world_size - 1) / c.world_size;
    if (c.technique == DR_TECH_MMLT) per = (per + c.max_depth - 1) / c.max_depth * c.max_depth;   // whole depth cycles per rank
    j->nBoot = per;
    j->bootFirst = (unsigned long long) per * (unsigned long long) c.rank;
    dr_status st;
    const long long nb = scan_blocks(per);
    if (!j->bootLum) {
        if (j->importance && (st = job_alloc(j, &j->bootLumTarget, (size_t) per))) return st;
        if ((st = job_alloc(j, &j->bootLum, (size_t) per)) || (st = job_alloc(j, &j->cdf, (size_t) per + 1)) ||
            (st = job_alloc(j, &j->blockSums, (size_t) nb + 1)))
            return st;
    }
    CK(cudaMemsetAsync(j->red, 0, 4 * sizeof(double), j->stream));
    CK(cudaEventRecord(j->ev0, j->stream));
    // the bootstrap paths run through the same wavefront machine as the chains (JOB_BOOT)
    JobParams job;
    memset(&job, 0, sizeof(job));
    job.type = JOB_BOOT; job.nItems = per; job.first = j->bootFirst; job.lumOut = j->bootLum;
    job.lumTargetOut = j->bootLumTarget;
    if ((st = setup_lanes(j, job)) || (st = run_machine(j, job, j->counters + ST_COUNT, false))) return st;
    launch_lum_reduce(j->bootLum, per, j->redScratch, j->red, j->stream);
    // b comes from the un-weighted luminances (generateSeeds copies the luminance before normalize(importanceMap),
    // pathsampler.cpp:899-901).  The seed CDF is built on the chains' TARGET: with an importance map that is the re-weighted
    // luminance.  The reference seeds ~ the un-weighted luminance and lets its ~100 000-mutation work units forget the start;
    // chains of ~64 mutations must start in their stationary distribution (seeding ~ L under a target ~ L / importance
    // biased the cornell box by 2-14 %, measured).
    launch_scan(j->bootLumTarget ? j->bootLumTarget : j->bootLum, per, j->cdf, j->blockSums, j->stream);
    CKL();
    j->launches += 5;
    CK(cudaEventRecord(j->ev1, j->stream));
    double red[2];
    CK(cudaMemcpyAsync(red, j->red, sizeof(red), cudaMemcpyDeviceToHost, j->stream));
    CK(cudaStreamSynchronize(j->stream));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, j->ev0, j->ev1));
    j->bootstrapMs += ms;
    j->bootstrapped = true;
    if (sumOut) *sumOut = red[0];
    if (countOut) *countOut = red[1];
    return DR_OK;
}

// Depth-balanced chain lengths (MMLT, resident chains; GPU execution knob dr_config.depth_balance, no reference equivalent).
// A mutation of a depth-d path costs ~d rays and ~d rounds of the wavefront machine, so with equal chain lengths the
// rounds of a job are set by its deepest chains while the lanes of the shallow ones idle.  Here a chain of depth d runs
// m_d ~ per * dbar / d mutations (dbar = luminance-weighted mean depth of the bootstrap) and the seeds are resampled
// ~ L / m_d: the expected number of mutations spent at depth d stays ~ B_d (the reference's allocation: chains ~ B_d,
// equal lengths), every chain still starts in its stationary distribution, and the film's normalisation (b / mean film
// luminance, drmlt_proc.cpp:813-854) does not depend on the chain count.  The m_d are integers; the seed weights use the
// very same integers, so the per-depth energy is exact in expectation whatever the rounding.
static dr_status balance_depths(dr_job j, long long per) {
    const dr_config &c = j->cfg;
    const int D = c.max_depth;
    dr_status st;
    if (!j->mutScale && ((st = job_alloc(j, &j->mutScale, 256)) || (st = job_alloc(j, &j->depthWeight, 32)) || (st = job_alloc(j, &j->depthSums, 32 + (size_t) depth_sums_scratch_doubles()))))
        return st;
    const float *src = j->bootLumTarget ? j->bootLumTarget : j->bootLum;
    launch_depth_sums(src, j->nBoot, j->bootFirst, D, j->depthSums + 32, j->depthSums, j->stream);
    double B[32];
    CK(cudaMemcpyAsync(B, j->depthSums, 32 * sizeof(double), cudaMemcpyDeviceToHost, j->stream));
    CK(cudaStreamSynchronize(j->stream));
    // cost model of a depth-d mutation: (d + c0)^e rays; c0 = 0, e = 1 measured best (DRMLT_DEPTH_COST0 / DRMLT_DEPTH_COST_EXP: tuning aids)
    const double cost0 = getenv("DRMLT_DEPTH_COST0") ? atof(getenv("DRMLT_DEPTH_COST0")) : 0.0;
    const double costE = getenv("DRMLT_DEPTH_COST_EXP") ? atof(getenv("DRMLT_DEPTH_COST_EXP")) : 1.0;
    double cost[32];
    for (int d = 0; d < 32; ++d) cost[d] = std::pow(std::max(0.25, d + 1 + cost0), costE);
    double sumB = 0.0, sumBd = 0.0;
    for (int d = 0; d < D; ++d) { sumB += B[d]; sumBd += B[d] * cost[d]; }
    if (!(sumB > 0.0)) return DR_OK;
    const double dbar = sumBd / sumB;
    uint32_t scale[256];
    long long m[32];
    // smallest common factor for which the expected total N * sum B / sum (B_d / m_d) reaches the budget N * per
    auto total = [&](double lambda) {
        double q = 0.0;
        for (int d = 0; d < D; ++d) {
            scale[d + 1] = (uint32_t) std::min(4294967295.0, std::floor(65536.0 * lambda * dbar / cost[d] + 0.5));
            m[d] = ((long long) per * scale[d + 1]) >> 16;
            if (m[d] < 1) return -1.0;
            q += B[d] / (double) m[d];
        }
        return sumB / q;
    };
    double lo = 0.25, hi = 4.0;
    if (!(total(hi) >= (double) per)) return DR_OK;          // (degenerate budget: keep equal lengths)
    for (int it = 0; it < 48; ++it) { const double mid = 0.5 * (lo + hi); if (total(mid) >= (double) per) hi = mid; else lo = mid; }
    if (total(hi) < 0.0) return DR_OK;
    float w[32] = { 0 };
    for (int d = 0; d < D; ++d) w[d] = (float) ((double) per / (double) m[d]);
    scale[0] = 65536u;
    for (int d = D + 1; d < 256; ++d) scale[d] = 65536u;
    CK(cudaMemcpyAsync(j->mutScale, scale, sizeof(scale), cudaMemcpyHostToDevice, j->stream));
    CK(cudaMemcpyAsync(j->depthWeight, w, sizeof(w), cudaMemcpyHostToDevice, j->stream));
    launch_scan(src, j->nBoot, j->cdf, j->blockSums, j->stream, j->depthWeight, j->bootFirst, D);
    CK(cudaStreamSynchronize(j->stream));                     // (scale / w are stack arrays)
    CKL();
    j->launches += 5;
    j->balanced = true;
    return DR_OK;
}

// seedPDF.sample per chain (pathsampler.cpp:946-954)
static void resample_chains(dr_job_t *j, unsigned long long firstChain) {
    const dr_config &c = j->cfg;
    launch_resample(j->cdf, j->nBoot, c.seed, firstChain, j->nChains, j->bootFirst, c.max_depth, c.technique, j->seedIdx, j->chainId, j->depth, j->stream);
    ++j->launches;
}

extern "C" dr_status dr_job_seed_chains(dr_job j, double b) {
    if (!j) { dr_set_error("dr_job_seed_chains: null job"); return DR_ERR_INVALID_ARG; }
    if (!j->bootstrapped) { dr_set_error("dr_job_seed_chains: call dr_job_bootstrap first"); return DR_ERR_INVALID_ARG; }
    CK(cudaSetDevice(j->scene->device));
    const dr_config &c = j->cfg;
    if (c.average_luminance != -1.0f) b = c.average_luminance;        // drmlt.cpp:555-558
    if (c.integrator == DR_INTEGRATOR_DRMLT && c.acceptance_map) b = 1.0;   // :550-552
    if (!(b > 0.0) || !std::isfinite(b)) {
        dr_set_error("The average image luminance appears to be zero! This could indicate a problem with the scene setup.");
        return DR_ERR_ZERO_LUMINANCE;                                   // pathsampler.cpp:939-941
    }
    double total = 0.0;
    CK(cudaMemcpyAsync(&total, j->cdf + j->nBoot, sizeof(double), cudaMemcpyDeviceToHost, j->stream));
    CK(cudaStreamSynchronize(j->stream));
    if (!(total > 0.0)) { dr_set_error("bootstrap found no path with non-zero luminance on rank %d", c.rank); return DR_ERR_ZERO_LUMINANCE; }
    j->b = b;
    j->M.cp.b = b;
    const int n = j->nChains;
    // chain ids key the mutation / coin streams: ranks (and work-unit batches) get disjoint 2^32-id ranges, whatever their chain
    // counts (a stride of the rank-local count would overlap when W*H*sampleCount does not divide by the world size)
    const unsigned long long firstChain = (unsigned long long) c.rank << 32;
    CK(cudaEventRecord(j->ev0, j->stream));
    j->balanced = false;
    if (c.depth_balance && c.technique == DR_TECH_MMLT && !j->chainCursor && !j->replay && c.max_depth <= 32) {
        const long long per = std::max<long long>(1, j->totalMutations / n);
        dr_status st;
        if (per >= 8 && (st = balance_depths(j, per))) return st;
    }
    resample_chains(j, firstChain);
    CKL();
    j->epoch = 0;
    if (!j->chainCursor) {
        // resident chains: the lanes evaluate their seed vector (PH_INIT); mutTarget = 0 parks them afterwards
        JobParams job;
        memset(&job, 0, sizeof(job));
        job.type = JOB_CHAIN; job.mutTarget = 0;
        dr_status st;
        if ((st = setup_lanes(j, job)) || (st = run_machine(j, job, j->counters, true))) return st;
    }                                          // (work-unit queue: a chain replays its seed when a lane takes it up)
    j->mutTarget = 0;
    CK(cudaEventRecord(j->ev1, j->stream));
    CK(cudaStreamSynchronize(j->stream));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, j->ev0, j->ev1));
    j->bootstrapMs += ms;
    j->seeded = true;
    return DR_OK;
}

// Work-unit queue: chains [first, first + count) of the current batch, `steps` mutations each, through the job's lanes.
static dr_status run_chain_range(dr_job j, long long first, long long count, long long steps) {
    JobParams job;
    memset(&job, 0, sizeof(job));
    job.type = JOB_CHAIN;
    job.mut0 = 0; job.mutTarget = (uint32_t) steps;
    job.chainCursor = j->chainCursor; job.chainFirst = (unsigned int) first; job.chainEnd = (unsigned int) (first + count);
    job.qDepth = j->depth; job.qChainId = j->chainId; job.qSeedIdx = j->seedIdx;
    const unsigned int cursor0 = (unsigned int) (first + std::min<long long>(count, j->nLanes));   // the lanes start with the first chains
    CK(cudaMemcpyAsync(j->chainCursor, &cursor0, sizeof(cursor0), cudaMemcpyHostToDevice, j->stream));
    dr_status st;
    if ((st = setup_lanes(j, job))) return st;
    return run_machine(j, job, j->counters, true);
}

// a fresh batch of chains for the work-unit queue: batch e uses chain ids (and resampling uniforms) e * nChains * worldSize + ...
static dr_status next_batch(dr_job j) {
    if (j->epoch > 0) {
        const dr_config &c = j->cfg;
        const unsigned long long firstChain = ((unsigned long long) j->epoch * c.world_size + c.rank) << 32;
        resample_chains(j, firstChain);
        CKL();
    }
    ++j->epoch;
    return DR_OK;
}

static dr_status run_chains(dr_job j, long long steps, dr_step_record *records, int recordStride, bool withFilm) {
    if (j->chainCursor) {                                          // work-unit queue: one batch = every chain once
        dr_status st = next_batch(j);
        return st ? st : run_chain_range(j, 0, j->nChains, steps);
    }
    JobParams job;
    memset(&job, 0, sizeof(job));
    job.type = JOB_CHAIN;
    job.mut0 = j->mutTarget;
    j->mutTarget += (uint32_t) steps;
    job.mutTarget = j->mutTarget;
    job.records = records; job.recordStride = recordStride;
    job.replay = j->replay; job.replayStride = j->replayStride; job.replayDim = j->replayDim;
    job.mutScale = j->balanced && !records ? j->mutScale : nullptr;
    CK(cudaStreamSynchronize(j->stream));
    for (int g = 0; g < (int) j->groups.size(); ++g) {             // parked chains start their next mutation
        Machine M = machine_for(j, g, job, j->counters, withFilm);
        LaunchCfg lc; lc.stream = j->groups[g].stream; lc.nLanes = M.laneEnd - M.laneBegin;
        launch_resume(M, lc);
        ++j->launches;
    }
    CKL();
    return run_machine(j, job, j->counters, withFilm);
}

static dr_status job_run_impl(dr_job j, int64_t mutationsPerChain, long long rangeFirst, long long rangeCount);
extern "C" dr_status dr_job_run(dr_job j, int64_t mutationsPerChain) { return job_run_impl(j, mutationsPerChain, 0, -1); }

// rangeCount < 0: resident chains advance, or (work-unit queue) a whole new batch runs; else chains
// [rangeFirst, rangeFirst + rangeCount) of the current batch run (work-unit queue only; used by progressive renders)
static dr_status job_run_impl(dr_job j, int64_t mutationsPerChain, long long rangeFirst, long long rangeCount) {
    if (!j) { dr_set_error("dr_job_run: null job"); return DR_ERR_INVALID_ARG; }
    if (!j->seeded) { dr_set_error("dr_job_run: call dr_job_seed_chains first"); return DR_ERR_INVALID_ARG; }
    if (j->cfg.timeout > 0 && !j->hasDeadline) {               // the reference's timer starts with the chain phase
        j->deadline = std::chrono::steady_clock::now() + std::chrono::seconds(j->cfg.timeout);
        j->hasDeadline = true;
    }
    if (j->timedOut) return DR_OK;
    if (mutationsPerChain < 0 || mutationsPerChain > (1ll << 30)) { dr_set_error("dr_job_run: mutation count out of range"); return DR_ERR_INVALID_ARG; }
    CK(cudaSetDevice(j->scene->device));
    CK(cudaEventRecord(j->ev0, j->stream));
    dr_status st = rangeCount < 0 ? run_chains(j, mutationsPerChain, nullptr, 0, true) : run_chain_range(j, rangeFirst, rangeCount, mutationsPerChain);
    if (st) { cudaStreamSynchronize(j->stream); return st; }
    CK(cudaEventRecord(j->ev1, j->stream));
    CK(cudaStreamSynchronize(j->stream));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, j->ev0, j->ev1));
    j->chainsMs += ms;
    j->mutationsDone += mutationsPerChain * (rangeCount < 0 ? (long long) j->nChains : rangeCount);
    return DR_OK;
}

static dr_status flush_pssmlt(dr_job j);
extern "C" dr_status dr_job_film_device(dr_job j, float **filmDev, int64_t *nFloats) {
    if (!j || !filmDev || !nFloats) { dr_set_error("dr_job_film_device: null argument"); return DR_ERR_INVALID_ARG; }
    if (j->seeded) {                                            // the film handed to a reduce must hold PSSMLT's pending last splats
        CK(cudaSetDevice(j->scene->device));
        dr_status st = flush_pssmlt(j);
        if (st) return st;
    }
    *filmDev = reinterpret_cast<float *>(j->film);
    *nFloats = (int64_t) j->W * j->H * 4;     // RGBA, A unused (16-byte vector atomics)
    return DR_OK;
}

// PSSMLT keeps the weight of the current state and splats it when the state is replaced; the
// last state is splatted when the chain ends (pssmlt_proc.cpp:274-279).
static dr_status flush_pssmlt(dr_job j) {
    if (j->cfg.integrator != DR_INTEGRATOR_PSSMLT) return DR_OK;
    JobParams job;
    memset(&job, 0, sizeof(job));
    CK(cudaStreamSynchronize(j->stream));
    for (int g = 0; g < (int) j->groups.size(); ++g) {
        Machine M = machine_for(j, g, job, j->counters, true);
        LaunchCfg lc; lc.stream = j->groups[g].stream; lc.nLanes = M.laneEnd - M.laneBegin;
        launch_flush_pssmlt(M, lc);
        ++j->launches;
    }
    CKL();
    return join_groups(j);
}

extern "C" dr_status dr_job_develop(dr_job j, float *imageRgb) {
    if (!j || !imageRgb) { dr_set_error("dr_job_develop: null argument"); return DR_ERR_INVALID_ARG; }
    CK(cudaSetDevice(j->scene->device));
    dr_status st = flush_pssmlt(j);
    if (st) return st;
    const long long n = (long long) j->W * j->H;
    CK(cudaMemsetAsync(j->red + 2, 0, sizeof(double), j->stream));
    launch_film_luminance(j->film, j->importance, n, j->red + 2, j->stream);
    CKL();
    double lumSum = 0.0;
    CK(cudaMemcpyAsync(&lumSum, j->red + 2, sizeof(double), cudaMemcpyDeviceToHost, j->stream));
    CK(cudaStreamSynchronize(j->stream));
    const bool accMap = j->M.cp.acceptanceMap;
    const double avg = lumSum / (double) n;
    const float factor = accMap ? 1.0f : (avg > 0.0 ? (float) (j->b / avg) : 0.f);
    launch_develop(j->film, j->importance, n, factor, (j->haveDirect && !accMap) ? j->directImage : nullptr, j->devImage, j->stream);
    CKL();
    j->launches += 2;
    CK(cudaMemcpyAsync(imageRgb, j->devImage, (size_t) n * 3 * sizeof(float), cudaMemcpyDeviceToHost, j->stream));
    CK(cudaStreamSynchronize(j->stream));
    return DR_OK;
}

// pixelSamples x shadingSamples split of directSamples (util.cpp:44-54)
static void direct_split(int directSamples, int *pixelSamples, int *shadingSamples) {
    int p = std::max(directSamples, 1), s = 1;
    while (p > 8) { p /= 2; s *= 2; }
    *pixelSamples = p; *shadingSamples = s;
}

// The separate direct-illumination image (renderDirectComponent, src/libbidir/util.cpp:30-94); develop adds it.
extern "C" dr_status dr_job_direct(dr_job j) {
    if (!j) { dr_set_error("dr_job_direct: null job"); return DR_ERR_INVALID_ARG; }
    if (j->cfg.direct_samples <= 0) return DR_OK;               // directSamples = 0: direct light is excluded and not rendered (drmlt.cpp:479)
    if (j->cfg.two_stage && j->cfg.first_stage) return DR_OK;   // the nested pass renders no direct image (`!nested`, drmlt.cpp:478)
    if (!j->film) { dr_set_error("dr_job_direct: not a render job"); return DR_ERR_INVALID_ARG; }
    CK(cudaSetDevice(j->scene->device));
    const long long n = (long long) j->W * j->H;
    dr_status st;
    if (!j->directFilm && ((st = job_alloc(j, &j->directFilm, (size_t) n)) || (st = job_alloc(j, &j->directImage, (size_t) 3 * n)))) return st;
    CK(cudaMemsetAsync(j->directFilm, 0, sizeof(float4) * n, j->stream));
    int ps, ss;
    direct_split(j->cfg.direct_samples, &ps, &ss);
    CK(cudaEventRecord(j->ev0, j->stream));
    launch_direct(j->M.sc, j->M.fp, j->cfg.seed, ps, ss, j->directFilm, j->directImage, nullptr, j->stream);
    CKL();
    j->launches += 2;
    CK(cudaEventRecord(j->ev1, j->stream));
    CK(cudaStreamSynchronize(j->stream));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, j->ev0, j->ev1));
    j->directMs += ms;
    j->haveDirect = true;
    return DR_OK;
}

extern "C" dr_status dr_job_stats(dr_job j, dr_stats *s) {
    if (!j || !s) { dr_set_error("dr_job_stats: null argument"); return DR_ERR_INVALID_ARG; }
    CK(cudaSetDevice(j->scene->device));
    unsigned long long c[2 * ST_COUNT];
    CK(cudaMemcpyAsync(c, j->counters, sizeof(c), cudaMemcpyDeviceToHost, j->stream));
    CK(cudaStreamSynchronize(j->stream));
    memset(s, 0, sizeof(*s));
    s->mutations = c[ST_MUT];
    s->first_accept = c[ST_FIRST_A]; s->first_base = c[ST_FIRST_B];
    s->large_accept = c[ST_LARGE_A]; s->large_base = c[ST_LARGE_B];
    s->bold_accept = c[ST_BOLD_A]; s->bold_base = c[ST_BOLD_B];
    s->second_accept = c[ST_SECOND_A]; s->second_base = c[ST_SECOND_B];
    s->second_large_accept = c[ST_SECOND_LARGE_A]; s->second_large_base = c[ST_SECOND_LARGE_B];
    s->second_bold_accept = c[ST_SECOND_BOLD_A]; s->second_bold_base = c[ST_SECOND_BOLD_B];
    s->accept = c[ST_ACC_A]; s->accept_base = c[ST_ACC_B];
    s->paths = c[ST_PATHS]; s->rays = c[ST_RAYS];
    s->bootstrap_paths = c[ST_COUNT + ST_PATHS]; s->bootstrap_rays = c[ST_COUNT + ST_RAYS];
    s->luminance = j->b;
    s->bootstrap_ms = j->bootstrapMs; s->chains_ms = j->chainsMs; s->total_ms = j->totalMs;
    s->kernel_launches = j->launches;
    s->rounds = j->rounds;
    s->direct_ms = j->directMs;
    s->trace_ms = j->stageMs[STAGE_TRACE]; s->walk_ms = j->stageMs[STAGE_WALK]; s->chain_ms = j->stageMs[STAGE_CHAIN];
    s->trace_launches = j->stageLaunches[STAGE_TRACE]; s->walk_launches = j->stageLaunches[STAGE_WALK]; s->chain_launches = j->stageLaunches[STAGE_CHAIN];
    return DR_OK;
}

extern "C" void dr_job_profile(dr_job j, int on) { if (j) j->profile = on != 0; }
extern "C" int64_t dr_job_num_chains(dr_job j) { return j ? j->nChains : 0; }
extern "C" int64_t dr_job_total_mutations(dr_job j) { return j ? j->totalMutations : 0; }

struct DevBuf {
    void *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    dr_status alloc(size_t bytes) { CK(cudaMalloc(&p, std::max<size_t>(bytes, 16))); return DR_OK; }
    template <class T> T *as() { return (T *) p; }
};

// ------------------------------------------------------------------ two-stage MLT (src/libbidir/util.cpp:96-199)
extern "C" dr_status dr_film_size(dr_scene scene, const dr_config *cfg, int32_t *width, int32_t *height) {
    if (!scene || !cfg || !width || !height) { dr_set_error("dr_film_size: null argument"); return DR_ERR_INVALID_ARG; }
    FilmWindow fw;
    dr_status st = film_window(scene, *cfg, fw);
    if (st) return st;
    *width = fw.W; *height = fw.H;
    return DR_OK;
}

extern "C" dr_status dr_first_stage_config(dr_scene scene, const dr_config *cfgIn, dr_config *nested) {
    if (!scene || !cfgIn || !nested) { dr_set_error("dr_first_stage_config: null argument"); return DR_ERR_INVALID_ARG; }
    dr_config c = *cfgIn;
    dr_status st = dr_config_validate(&c);
    if (st) return st;
    const int f = c.first_stage_size_reduction;
    if (f <= 0) { dr_set_error("firstStageSizeReduction must be positive"); return DR_ERR_INVALID_ARG; }
    FilmWindow fw;
    if ((st = film_window(scene, c, fw))) return st;
    c.two_stage = 1; c.first_stage = 1;                                   // integratorProps + firstStage=true (util.cpp:151)
    c.film_width = std::max(1, fw.filmW / f); c.film_height = std::max(1, fw.filmH / f);   // :104-110
    c.crop_width = std::max(1, fw.W / f); c.crop_height = std::max(1, fw.H / f);
    c.crop_offset_x = fw.cropX / f;
    c.crop_offset_y = fw.cropX / f;                                       // sic: the reference passes reducedCropOffset.x for Y too (:128)
    const long long spp = (long long) c.sample_count * f;                 // "higher number of mutations/pixel" (:134-136)
    if (spp > 0x7fffffffll) { dr_set_error("dr_first_stage_config: sampleCount * firstStageSizeReduction overflows"); return DR_ERR_INVALID_ARG; }
    c.sample_count = (int32_t) spp;
    c.importance_map = nullptr;
    c.n_chains = 0;                                                       // sized for the small nested job
    *nested = c;
    return DR_OK;
}

// Resampler::Resampler (include/mitsuba/core/rfilter.h:123-177), resampling mode, gaussian filter (stddev 0.5, radius 2)
struct ResampleTable { int taps = 0; std::vector<int> start; std::vector<double> weights; };
static void resample_table(int sourceRes, int targetRes, ResampleTable &t) {
    const double stddev = 0.5, radius = 4 * stddev, alpha = -1.0 / (2.0 * stddev * stddev);
    auto eval = [&](double x) { return std::max(0.0, std::exp(alpha * x * x) - std::exp(alpha * radius * radius)); };   // gaussian.cpp:52-57
    double filterRadius = radius, scale = 1.0, invScale = 1.0;
    if (targetRes < sourceRes) { scale = (double) sourceRes / (double) targetRes; invScale = 1 / scale; filterRadius *= scale; }
    t.taps = (int) std::ceil(filterRadius * 2);
    t.start.resize(targetRes); t.weights.resize((size_t) t.taps * targetRes);
    for (int i = 0; i < targetRes; i++) {
        const double center = (i + 0.5) / targetRes * sourceRes;
        t.start[i] = (int) std::floor(center - filterRadius + 0.5);
        double sum = 0;
        for (int j = 0; j < t.taps; j++) {
            const double pos = t.start[i] + j + 0.5 - center;
            const double w = eval(pos * invScale);
            t.weights[(size_t) i * t.taps + j] = w;
            sum += w;
        }
        const double normalization = 1.0 / sum;
        for (int j = 0; j < t.taps; j++) t.weights[(size_t) i * t.taps + j] *= normalization;
    }
}

extern "C" dr_status dr_resample_luminance(dr_scene scene, const float *imageRgb, int32_t w, int32_t h, int32_t W, int32_t H, float *map) {
    if (!scene || !imageRgb || !map || w <= 0 || h <= 0 || W <= 0 || H <= 0) { dr_set_error("dr_resample_luminance: bad argument"); return DR_ERR_INVALID_ARG; }
    CK(cudaSetDevice(scene->device));
    DevBuf rgb, lum, tmp, out, outF, dStart, dWeights;
    dr_status st;
    if ((st = rgb.alloc(sizeof(float) * 3 * (size_t) w * h)) || (st = lum.alloc(sizeof(double) * (size_t) w * h)) || (st = outF.alloc(sizeof(float) * (size_t) W * H))) return st;
    CK(cudaMemcpy(rgb.p, imageRgb, sizeof(float) * 3 * (size_t) w * h, cudaMemcpyHostToDevice));
    launch_rgb_luminance(rgb.as<float>(), (long long) w * h, lum.as<double>(), 0);     // nestedFilm->develop into an ELuminance bitmap (util.cpp:184-188)
    const double *cur = lum.as<double>();
    int curW = w;
    auto pass = [&](int srcRes, int tgtRes, int other, int alongX, DevBuf &dst) -> dr_status {
        ResampleTable t;
        resample_table(srcRes, tgtRes, t);
        DevBuf ds_, dw_;
        dr_status s2;
        if ((s2 = ds_.alloc(sizeof(int) * t.start.size())) || (s2 = dw_.alloc(sizeof(double) * t.weights.size())) || (s2 = dst.alloc(sizeof(double) * (size_t) tgtRes * other))) return s2;
        CK(cudaMemcpy(ds_.p, t.start.data(), sizeof(int) * t.start.size(), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dw_.p, t.weights.data(), sizeof(double) * t.weights.size(), cudaMemcpyHostToDevice));
        launch_resample_axis(cur, srcRes, tgtRes, other, alongX, ds_.as<int>(), dw_.as<double>(), t.taps, dst.as<double>(), 0);
        CKL();
        CK(cudaDeviceSynchronize());                               // the tables are freed when this scope ends
        return DR_OK;
    };
    // mitsuba::resample (src/libcore/bitmap.cpp:2230-2329): along x first (into a [h][W] temporary), then along y
    if (w != W) { if ((st = pass(w, W, h, 1, tmp))) return st; cur = tmp.as<double>(); curW = W; }
    if (h != H) { if ((st = pass(h, H, curW, 0, out))) return st; cur = out.as<double>(); }
    launch_double_to_float(cur, (long long) W * H, outF.as<float>(), 0);
    CKL();
    CK(cudaMemcpy(map, outF.p, sizeof(float) * (size_t) W * H, cudaMemcpyDeviceToHost));
    return DR_OK;
}

extern "C" dr_status dr_importance_map(dr_scene scene, const dr_config *cfg, float *map, dr_stats *nestedStats) {
    if (!scene || !cfg || !map) { dr_set_error("dr_importance_map: null argument"); return DR_ERR_INVALID_ARG; }
    dr_config nested;
    dr_status st = dr_first_stage_config(scene, cfg, &nested);
    if (st) return st;
    int32_t w, h, W, H;
    if ((st = dr_film_size(scene, &nested, &w, &h)) || (st = dr_film_size(scene, cfg, &W, &H))) return st;
    std::vector<float> img((size_t) 3 * w * h);
    if ((st = dr_render(scene, &nested, img.data(), nestedStats))) return st;      // "Executing first MLT stage" (drmlt.cpp:406-418)
    return dr_resample_luminance(scene, img.data(), w, h, W, H, map);
}

extern "C" dr_status dr_render(dr_scene scene, const dr_config *cfgIn, float *imageRgb, dr_stats *stats) {
    return dr_render_progressive(scene, cfgIn, imageRgb, stats, 0.0, nullptr, nullptr);
}

// dr_render with periodic develops of the partial result: what DRMLTProcess::processResult does for interactive jobs
// (develop + signalRefresh every <= 2 s, drmlt_proc.cpp:856-867) and what `mitsuba -r <sec>` dumps through Scene::flush
// (src/librender/scene.cpp:468-511).  The chain phase runs in slices of about `refreshSeconds`; after each slice the
// film is developed into `imageRgb` and `fn` is called; a non-zero return cancels the job like Integrator::cancel.
extern "C" dr_status dr_render_progressive(dr_scene scene, const dr_config *cfgIn, float *imageRgb, dr_stats *stats,
                                           double refreshSeconds, dr_refresh_fn fn, void *user) {
    if (!scene || !cfgIn || !imageRgb) { dr_set_error("dr_render: null argument"); return DR_ERR_INVALID_ARG; }
    scene->cancel = 0;
    dr_config cfgLocal = *cfgIn;
    const dr_config *cfg = &cfgLocal;
    std::vector<float> importance;
    double firstStageMs = 0.0;
    if (cfgLocal.two_stage && !cfgLocal.first_stage && !cfgLocal.importance_map) {
        int32_t W, H;
        dr_status s0 = dr_film_size(scene, cfg, &W, &H);
        if (s0) return s0;
        importance.resize((size_t) W * H);
        dr_stats ns;
        const auto t0 = std::chrono::steady_clock::now();
        if ((s0 = dr_importance_map(scene, cfg, importance.data(), &ns))) return s0;
        firstStageMs = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        cfgLocal.importance_map = importance.data();
    }
    dr_job j = nullptr;
    dr_status st = dr_job_create(scene, cfg, &j);
    if (st) return st;
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0); cudaEventCreate(&t1);
    cudaEventRecord(t0, j->stream);
    double sum = 0.0, count = 0.0;
    if (!(st = dr_job_bootstrap(j, &sum, &count))) {
        double b = count > 0.0 ? sum / count : 0.0;                              // pathsampler.cpp:922-934
        if (j->cfg.technique == DR_TECH_MMLT) b *= j->cfg.max_depth;
        // the direct image is rendered before the chains start (drmlt.cpp:478-488), so partial develops include it
        if (!(st = dr_job_direct(j)) && !(st = dr_job_seed_chains(j, b))) {
            const long long per = std::max<long long>(1, j->totalMutations / j->nChains);   // nMutations (drmlt.cpp:475-476)
            if (!fn || !(refreshSeconds > 0.0)) st = dr_job_run(j, per);
            else {
                // slices: mutations of the resident chains, or (work-unit queue) ranges of the batch's chains
                const bool queue = j->chainCursor != nullptr;
                const long long whole = queue ? j->nChains : per;
                if (queue) st = next_batch(j);
                const auto start = std::chrono::steady_clock::now();
                long long done = 0, slice = queue ? std::min<long long>(whole, std::max<long long>(j->nLanes, whole / 64)) : std::min<long long>(per, 4);
                while (!st && done < whole && !j->timedOut) {
                    const auto s0 = std::chrono::steady_clock::now();
                    slice = std::min(slice, whole - done);
                    if ((st = queue ? job_run_impl(j, per, done, slice) : dr_job_run(j, slice))) break;
                    done += slice;
                    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - s0).count();
                    // next slice: about refreshSeconds of work at the measured rate (at most 4x growth per step)
                    const double want = sec > 0.0 ? (double) slice * refreshSeconds / sec : (double) slice * 4.0;
                    slice = (long long) std::max(queue ? (double) j->nLanes : 1.0, std::min(want, (double) slice * 4.0));
                    if (done < whole && !j->timedOut) {
                        dr_stats ps;
                        if ((st = dr_job_develop(j, imageRgb)) || (st = dr_job_stats(j, &ps))) break;
                        const double elapsed = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
                        if (fn(imageRgb, j->W, j->H, elapsed, &ps, user)) { dr_set_error("cancelled"); st = DR_ERR_CANCELLED; }
                    }
                }
            }
            if (!st) st = dr_job_develop(j, imageRgb);
        }
    }
    cudaEventRecord(t1, j->stream);
    cudaEventSynchronize(t1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, t0, t1);
    j->totalMs = ms;
    cudaEventDestroy(t0); cudaEventDestroy(t1);
    if (!st && stats) { st = dr_job_stats(j, stats); stats->first_stage_ms = firstStageMs; }
    dr_job_destroy(j);
    return st;
}

// ------------------------------------------------------------------ replay / parity entry points

extern "C" dr_status dr_trace_rays(dr_scene scene, const dr_ray *rays, int64_t n, int shadow, dr_hit *hits) {
    if (!scene || (n > 0 && (!rays || !hits)) || n < 0) { dr_set_error("dr_trace_rays: bad argument"); return DR_ERR_INVALID_ARG; }
    if (n == 0) return DR_OK;
    SceneImpl *s = static_cast<SceneImpl *>(scene);
    CK(cudaSetDevice(s->device));
    DevBuf dr, dh;
    dr_status st;
    if ((st = dr.alloc(n * sizeof(dr_ray))) || (st = dh.alloc(n * sizeof(dr_hit)))) return st;
    CK(cudaMemcpy(dr.p, rays, n * sizeof(dr_ray), cudaMemcpyHostToDevice));
    launch_trace_rays(s->dev, dr.as<dr_ray>(), n, shadow, s->dOrder, dh.as<dr_hit>(), 0);
    CKL();
    CK(cudaMemcpy(hits, dh.p, n * sizeof(dr_hit), cudaMemcpyDeviceToHost));
    return DR_OK;
}

extern "C" dr_status dr_texture_eval(dr_scene scene, uint32_t texture, const double *uv, int64_t n, double *rgb) {
    if (!scene || (n > 0 && (!uv || !rgb)) || n < 0) { dr_set_error("dr_texture_eval: bad argument"); return DR_ERR_INVALID_ARG; }
    SceneImpl *s = static_cast<SceneImpl *>(scene);
    if (texture >= s->nTextures) { dr_set_error("dr_texture_eval: texture index %u out of range (%u textures)", texture, s->nTextures); return DR_ERR_INVALID_ARG; }
    if (n == 0) return DR_OK;
    CK(cudaSetDevice(s->device));
    DevBuf du, dc;
    dr_status st;
    if ((st = du.alloc(n * 2 * sizeof(double))) || (st = dc.alloc(n * 3 * sizeof(double)))) return st;
    CK(cudaMemcpy(du.p, uv, n * 2 * sizeof(double), cudaMemcpyHostToDevice));
    launch_texture_eval(s->dev, texture, du.as<double>(), n, dc.as<double>(), 0);
    CKL();
    CK(cudaMemcpy(rgb, dc.p, n * 3 * sizeof(double), cudaMemcpyDeviceToHost));
    return DR_OK;
}

extern "C" dr_status dr_direct_image(dr_scene scene, const dr_config *cfgIn, float *imageRgb, double *li) {
    if (!scene || !cfgIn || !imageRgb) { dr_set_error("dr_direct_image: bad argument"); return DR_ERR_INVALID_ARG; }
    dr_config cfg = *cfgIn;
    dr_status st = dr_config_validate(&cfg);
    if (st) return st;
    if (cfg.direct_samples <= 0) { dr_set_error("dr_direct_image: directSamples must be positive"); return DR_ERR_INVALID_ARG; }
    CK(cudaSetDevice(scene->device));
    FilmWindow fw;
    if ((st = film_window(scene, cfg, fw))) return st;
    const long long n = (long long) fw.W * fw.H;
    int ps, ss;
    direct_split(cfg.direct_samples, &ps, &ss);
    Machine M;
    memset(&M, 0, sizeof(M));
    M.sc = scene_for(scene, cfg, fw);
    make_params(cfg, fw.W, fw.H, 1.0, nullptr, M, static_cast<SceneImpl *>(scene)->typeMask);
    DevBuf film, rgb, dli;
    if ((st = film.alloc(sizeof(float4) * n)) || (st = rgb.alloc(sizeof(float) * 3 * n)) || (li && (st = dli.alloc(sizeof(double) * 3 * n * ps)))) return st;
    CK(cudaMemset(film.p, 0, sizeof(float4) * n));
    launch_direct(M.sc, M.fp, cfg.seed, ps, ss, film.as<float4>(), rgb.as<float>(), li ? dli.as<double>() : nullptr, 0);
    CKL();
    CK(cudaMemcpy(imageRgb, rgb.p, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost));
    if (li) CK(cudaMemcpy(li, dli.p, sizeof(double) * 3 * n * ps, cudaMemcpyDeviceToHost));
    return DR_OK;
}

// The film by itself: `n` splats through the reconstruction filter of `cfg` into a w x h film (ImageBlock::put, imageblock.h:149-196).
extern "C" dr_status dr_splat_points(int device, const dr_config *cfgIn, int32_t w, int32_t h, const float *pos, const float *rgb, int64_t n, float *filmRgb) {
    if (!cfgIn || w <= 0 || h <= 0 || n < 0 || !filmRgb || (n > 0 && (!pos || !rgb))) { dr_set_error("dr_splat_points: bad argument"); return DR_ERR_INVALID_ARG; }
    if (cfgIn->rfilter < DR_FILTER_GAUSSIAN || cfgIn->rfilter > DR_FILTER_LANCZOS || (cfgIn->rfilter == DR_FILTER_TABLE && !(cfgIn->filter_radius > 0.0))) {
        dr_set_error("dr_splat_points: unsupported rfilter"); return DR_ERR_UNSUPPORTED;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); dr_set_error("no CUDA device available (there is no CPU fallback)"); return DR_ERR_NO_DEVICE; }
    if (device < 0 || device >= ndev) { dr_set_error("device %d out of range (%d devices)", device, ndev); return DR_ERR_INVALID_ARG; }
    CK(cudaSetDevice(device));
    Machine M;
    memset(&M, 0, sizeof(M));
    dr_config c = *cfgIn;
    if (c.max_depth <= 0) c.max_depth = 1;
    make_params(c, w, h, 1.0, nullptr, M);
    const size_t np = (size_t) w * h;
    DevBuf film, dpos, drgb;
    dr_status st;
    if ((st = film.alloc(np * sizeof(float4))) || (st = dpos.alloc((size_t) n * 8)) || (st = drgb.alloc((size_t) n * 12))) return st;
    CK(cudaMemset(film.p, 0, np * sizeof(float4)));
    if (n) {
        CK(cudaMemcpy(dpos.p, pos, (size_t) n * 8, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(drgb.p, rgb, (size_t) n * 12, cudaMemcpyHostToDevice));
        launch_splat_points(M.fp, film.as<float4>(), dpos.as<float>(), drgb.as<float>(), n, 0);
        CKL();
    }
    std::vector<float4> f4(np);
    CK(cudaMemcpy(f4.data(), film.p, np * sizeof(float4), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < np; ++i) { filmRgb[3 * i] = f4[i].x; filmRgb[3 * i + 1] = f4[i].y; filmRgb[3 * i + 2] = f4[i].z; }
    return DR_OK;
}

static int replay_lanes(int64_t n) { return (int) std::min<int64_t>((n + 127) / 128 * 128, 1 << 18); }

extern "C" dr_status dr_eval_paths(dr_scene scene, const dr_config *cfgIn, const float *us, int ds_, const float *ue, int de, const float *ud, int dd,
                                   const int32_t *depth, int64_t n, dr_path_result *out) {
    if (!scene || !cfgIn || n < 0 || (n > 0 && !out)) { dr_set_error("dr_eval_paths: bad argument"); return DR_ERR_INVALID_ARG; }
    if ((ds_ > 0 && !us) || (de > 0 && !ue) || (dd > 0 && !ud) || ds_ < 0 || de < 0 || dd < 0) { dr_set_error("dr_eval_paths: missing primary-sample buffer"); return DR_ERR_INVALID_ARG; }
    if (ds_ > 255 || de > 255 || dd > 255) { dr_set_error("dr_eval_paths: at most 255 coordinates per sampler"); return DR_ERR_INVALID_ARG; }
    if (n == 0) { dr_config c = *cfgIn; return dr_config_validate(&c); }
    dr_job j = nullptr;
    const int evalDims[3] = { ds_, de, dd };
    dr_status st = job_create_common(scene, cfgIn, replay_lanes(n), false, evalDims, &j);
    if (st) return st;
    auto done = [&](dr_status code) { cudaStreamSynchronize(j->stream); dr_job_destroy(j); return code; };
    if (j->cfg.technique == DR_TECH_MMLT) {
        if (!depth) return done((dr_set_error("dr_eval_paths: MMLT needs a depth per path"), DR_ERR_INVALID_ARG));
        for (int64_t i = 0; i < n; ++i)
            if (depth[i] < 1 || depth[i] > j->cfg.max_depth) return done((dr_set_error("dr_eval_paths: depth out of range"), DR_ERR_INVALID_ARG));
    }
    DevBuf bs, be, bd, bdep, bout;
    if ((st = bs.alloc((size_t) n * ds_ * 4)) || (st = be.alloc((size_t) n * de * 4)) || (st = bd.alloc((size_t) n * dd * 4)) ||
        (st = bdep.alloc((size_t) n * 4)) || (st = bout.alloc((size_t) n * sizeof(dr_path_result))))
        return done(st);
    cudaError_t e = cudaSuccess;
    if (ds_) e = cudaMemcpy(bs.p, us, (size_t) n * ds_ * 4, cudaMemcpyHostToDevice);
    if (de && e == cudaSuccess) e = cudaMemcpy(be.p, ue, (size_t) n * de * 4, cudaMemcpyHostToDevice);
    if (dd && e == cudaSuccess) e = cudaMemcpy(bd.p, ud, (size_t) n * dd * 4, cudaMemcpyHostToDevice);
    if (depth && e == cudaSuccess) e = cudaMemcpy(bdep.p, depth, (size_t) n * 4, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return done((dr_set_error("dr_eval_paths: upload failed: %s", cudaGetErrorString(e)), DR_ERR_CUDA));
    JobParams job;
    memset(&job, 0, sizeof(job));
    job.type = JOB_EVAL; job.nItems = n;
    job.us = bs.as<float>(); job.ue = be.as<float>(); job.ud = bd.as<float>(); job.ds = ds_; job.de = de; job.dd = dd;
    job.depthIn = depth ? bdep.as<int>() : nullptr; job.out = bout.as<dr_path_result>();
    if ((st = setup_lanes(j, job)) || (st = run_machine(j, job, j->counters, false))) return done(st);
    if (cudaMemcpy(out, bout.p, (size_t) n * sizeof(dr_path_result), cudaMemcpyDeviceToHost) != cudaSuccess)
        return done((dr_set_error("dr_eval_paths: download failed"), DR_ERR_CUDA));
    return done(DR_OK);
}

extern "C" dr_status dr_bootstrap_luminance(dr_scene scene, const dr_config *cfgIn, uint64_t first, int64_t n, float *luminance, int32_t *depth) {
    if (!scene || !cfgIn || n < 0 || (n > 0 && !luminance)) { dr_set_error("dr_bootstrap_luminance: bad argument"); return DR_ERR_INVALID_ARG; }
    if (n == 0) { dr_config c = *cfgIn; return dr_config_validate(&c); }
    dr_job j = nullptr;
    dr_status st = job_create_common(scene, cfgIn, replay_lanes(n), false, nullptr, &j);
    if (st) return st;
    auto done = [&](dr_status code) { cudaStreamSynchronize(j->stream); dr_job_destroy(j); return code; };
    DevBuf bl;
    if ((st = bl.alloc((size_t) n * 4))) return done(st);
    JobParams job;
    memset(&job, 0, sizeof(job));
    job.type = JOB_BOOT; job.nItems = n; job.first = first; job.lumOut = bl.as<float>();
    if ((st = setup_lanes(j, job)) || (st = run_machine(j, job, j->counters + ST_COUNT, false))) return done(st);
    if (cudaMemcpy(luminance, bl.p, (size_t) n * 4, cudaMemcpyDeviceToHost) != cudaSuccess)
        return done((dr_set_error("dr_bootstrap_luminance: download failed"), DR_ERR_CUDA));
    if (depth)
        for (int64_t i = 0; i < n; ++i)
            depth[i] = j->cfg.technique == DR_TECH_MMLT ? (int32_t) ((first + (uint64_t) i) % (uint64_t) j->cfg.max_depth) + 1 : -1;
    return done(DR_OK);
}

static dr_status chain_steps_impl(dr_scene scene, const dr_config *cfgIn, double b, const uint64_t *seedIndex, const int32_t *depth,
                                  const uint64_t *chainId, int64_t nChains, int64_t steps, dr_step_record *records, float *film,
                                  const double *uniforms, int uniformDim);

extern "C" dr_status dr_chain_steps(dr_scene scene, const dr_config *cfgIn, double b, const uint64_t *seedIndex, const int32_t *depth,
                                    const uint64_t *chainId, int64_t nChains, int64_t steps, dr_step_record *records, float *film) {
    if (nChains > 0 && (!seedIndex || !chainId)) { dr_set_error("dr_chain_steps: bad argument"); return DR_ERR_INVALID_ARG; }
    return chain_steps_impl(scene, cfgIn, b, seedIndex, depth, chainId, nChains, steps, records, film, nullptr, 0);
}

// Chains of the reference replayed: every uniform comes from the caller's table (layout: include/drmlt_b200.h) instead of Philox.
extern "C" dr_status dr_chain_replay(dr_scene scene, const dr_config *cfgIn, double b, const int32_t *depth, int64_t nChains, int64_t steps,
                                     const double *uniforms, int32_t uniformDim, dr_step_record *records, float *film) {
    if (nChains > 0 && (!uniforms || uniformDim <= 0 || (uniformDim & 1))) { dr_set_error("dr_chain_replay: a table of an even number of coordinates per sampler is required"); return DR_ERR_INVALID_ARG; }
    if (nChains < 0 || nChains > (1 << 20)) { dr_set_error("dr_chain_replay: chain count out of range"); return DR_ERR_INVALID_ARG; }
    std::vector<uint64_t> ids((size_t) nChains);
    for (int64_t i = 0; i < nChains; ++i) ids[i] = (uint64_t) i;
    return chain_steps_impl(scene, cfgIn, b, ids.data(), depth, ids.data(), nChains, steps, records, film, uniforms, uniformDim);
}

static dr_status chain_steps_impl(dr_scene scene, const dr_config *cfgIn, double b, const uint64_t *seedIndex, const int32_t *depth,
                                  const uint64_t *chainId, int64_t nChains, int64_t steps, dr_step_record *records, float *film,
                                  const double *uniforms, int uniformDim) {
    if (!scene || !cfgIn || nChains < 0 || steps < 0 || (nChains > 0 && (!seedIndex || !chainId))) { dr_set_error("dr_chain_steps: bad argument"); return DR_ERR_INVALID_ARG; }
    if (nChains == 0) { dr_config c = *cfgIn; return dr_config_validate(&c); }
    if (nChains > (1 << 24) || steps > (1 << 24)) { dr_set_error("dr_chain_steps: too large"); return DR_ERR_INVALID_ARG; }
    dr_job j = nullptr;
    dr_status st = job_create_common(scene, cfgIn, (int) nChains, true, nullptr, &j);
    if (st) return st;
    auto done = [&](dr_status code) { cudaStreamSynchronize(j->stream); dr_job_destroy(j); return code; };
    if (j->cfg.technique == DR_TECH_MMLT && !depth) return done((dr_set_error("dr_chain_steps: MMLT needs a depth per chain"), DR_ERR_INVALID_ARG));
    if (j->cfg.average_luminance != -1.0f) b = j->cfg.average_luminance;
    if (j->cfg.integrator == DR_INTEGRATOR_DRMLT && j->cfg.acceptance_map) b = 1.0;
    j->b = b; j->M.cp.b = b;
    const int n = (int) nChains;
    std::vector<int> dep(n, -1);
    if (j->cfg.technique == DR_TECH_MMLT)
        for (int i = 0; i < n; ++i) {
            dep[i] = depth[i];
            if (dep[i] < 1 || dep[i] > j->cfg.max_depth) return done((dr_set_error("dr_chain_steps: depth out of range"), DR_ERR_INVALID_ARG));
        }
    static_assert(sizeof(unsigned long long) == sizeof(uint64_t), "u64");
    if (cudaMemcpyAsync(j->seedIdx, seedIndex, n * sizeof(uint64_t), cudaMemcpyHostToDevice, j->stream) != cudaSuccess ||
        cudaMemcpyAsync(j->chainId, chainId, n * sizeof(uint64_t), cudaMemcpyHostToDevice, j->stream) != cudaSuccess ||
        cudaMemcpyAsync(j->depth, dep.data(), n * sizeof(int), cudaMemcpyHostToDevice, j->stream) != cudaSuccess ||
        cudaStreamSynchronize(j->stream) != cudaSuccess) {
        dr_set_error("dr_chain_steps: upload failed: %s", cudaGetErrorString(cudaGetLastError()));
        return done(DR_ERR_CUDA);
    }
    DevBuf drec;
    if (records && steps > 0) {
        if ((st = drec.alloc((size_t) n * steps * sizeof(dr_step_record)))) return done(st);
        cudaMemsetAsync(drec.p, 0, (size_t) n * steps * sizeof(dr_step_record), j->stream);
    }
    DevBuf dtab;
    const long long tabStride = 3ll * uniformDim + (long long) steps * (4 + 12ll * uniformDim);
    if (uniforms) {
        if ((st = dtab.alloc((size_t) n * tabStride * sizeof(double)))) return done(st);
        if (cudaMemcpy(dtab.p, uniforms, (size_t) n * tabStride * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) { dr_set_error("dr_chain_replay: table upload failed"); return done(DR_ERR_CUDA); }
        j->replay = dtab.as<double>(); j->replayStride = tabStride; j->replayDim = uniformDim;
    }
    JobParams job;
    memset(&job, 0, sizeof(job));
    job.type = JOB_CHAIN; job.mutTarget = 0;
    job.replay = j->replay; job.replayStride = j->replayStride; job.replayDim = j->replayDim;
    if ((st = setup_lanes(j, job)) || (st = run_machine(j, job, j->counters, film != nullptr))) return done(st);   // seed replay
    j->seeded = true;
    if (steps > 0 && (st = run_chains(j, steps, records ? drec.as<dr_step_record>() : nullptr, (int) steps, film != nullptr))) return done(st);
    if (film && (st = flush_pssmlt(j))) return done(st);
    if (cudaStreamSynchronize(j->stream) != cudaSuccess) { dr_set_error("dr_chain_steps: kernel failed: %s", cudaGetErrorString(cudaGetLastError())); return done(DR_ERR_CUDA); }
    if (records && steps > 0 && cudaMemcpy(records, drec.p, (size_t) n * steps * sizeof(dr_step_record), cudaMemcpyDeviceToHost) != cudaSuccess) {
        dr_set_error("dr_chain_steps: record download failed"); return done(DR_ERR_CUDA);
    }
    if (film) {
        const size_t np = (size_t) j->W * j->H;
        std::vector<float4> f4(np);
        if (cudaMemcpy(f4.data(), j->film, np * sizeof(float4), cudaMemcpyDeviceToHost) != cudaSuccess) { dr_set_error("dr_chain_steps: film download failed"); return done(DR_ERR_CUDA); }
        for (size_t i = 0; i < np; ++i) { film[3 * i] = f4[i].x; film[3 * i + 1] = f4[i].y; film[3 * i + 2] = f4[i].z; }
    }
    return done(DR_OK);
}
