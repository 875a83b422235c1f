/*
 * drmlt_b200.h -- C ABI of the B200-native Markov-chain light-transport hot path.
 *
 * This is the drop-in boundary for the reference's `pssmlt` / `drmlt` integrator
 * plugins (reference: src/integrators/drmlt/drmlt.cpp:176-618,
 * src/integrators/pssmlt/pssmlt.cpp:164-552).  A thin Mitsuba-side plugin
 * (drmlt-mitsuba_b200/shim/) flattens `Scene` into the POD buffers below and calls
 * `dr_render()`; everything behind this header runs on the GPU (sm_100a).
 *
 * Conventions: plain pointers and sizes, no C++ or torch types; the caller owns all
 * host buffers, the library owns device memory; every entry point returns a
 * dr_status (0 = ok) and never throws; `dr_last_error()` gives the message of the
 * last failure on the calling thread.  There is NO CPU fallback: when no CUDA
 * device is usable the calls fail with DR_ERR_NO_DEVICE.
 */
#ifndef DRMLT_B200_H
#define DRMLT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DR_ABI_VERSION 6     /* 6: bitmap textures (dr_texture, dr_scene_desc.texcoords / textures, DR_MAT_TEX_*, DR_TRI_UV_TANGENTS); 5: dr_scene_desc.rough_tables, dr_material.table, dr_config.depth_balance, dr_scene_create_ex */

/* ------------------------------------------------------------------ status */
typedef enum dr_status {
    DR_OK = 0,
    DR_ERR_INVALID_ARG = 1,   /* bad parameter combination (mirrors the reference's Log(EError,..)) */
    DR_ERR_NO_DEVICE = 2,     /* CUDA runtime / device missing: no CPU fallback exists */
    DR_ERR_CUDA = 3,          /* a CUDA call failed */
    DR_ERR_ZERO_LUMINANCE = 4,/* bootstrap mean luminance is zero (pathsampler.cpp:939-941) */
    DR_ERR_CANCELLED = 5,     /* dr_cancel() was called (Integrator::cancel, drmlt.cpp:386-391) */
    DR_ERR_UNSUPPORTED = 6
} dr_status;

/* ------------------------------------------------------------------- scene */

/* BSDF models on the hot path */
typedef enum dr_bsdf_type {
    DR_BSDF_DIFFUSE = 0,        /* src/bsdfs/diffuse.cpp:109-150 */
    DR_BSDF_DIELECTRIC = 1,     /* src/bsdfs/dielectric.cpp:227-400 */
    DR_BSDF_CONDUCTOR = 2,      /* src/bsdfs/conductor.cpp:223-285 */
    DR_BSDF_ROUGHCONDUCTOR = 3, /* src/bsdfs/roughconductor.cpp:250-412 */
    DR_BSDF_ROUGHDIELECTRIC = 4,/* src/bsdfs/roughdielectric.cpp:270-611 (isotropic alpha; uses one extra primary sample per BSDF sample) */
    DR_BSDF_PLASTIC = 5,        /* src/bsdfs/plastic.cpp:240-420: delta specular coating over a diffuse base */
    DR_BSDF_ROUGHPLASTIC = 6    /* src/bsdfs/roughplastic.cpp:325-470: microfacet coating over a diffuse base (constant alpha; needs a rough-
                                   transmittance table, dr_scene_desc.rough_tables) */
} dr_bsdf_type;

#define DR_MAT_TWOSIDED       1u  /* wrapped in <bsdf type="twosided"> (src/bsdfs/twosided.cpp) */
#define DR_MAT_GGX            2u  /* roughconductor / roughdielectric distribution=ggx (else beckmann) */
#define DR_MAT_SAMPLE_VISIBLE 4u  /* roughconductor / roughdielectric sampleVisible=true */
#define DR_MAT_NONLINEAR      8u  /* plastic nonlinear=true (plastic.cpp:162, 268-271) */
/* Bitmap textures on the two colour parameters (src/textures/bitmap.cpp): bits 8-19 of `flags` hold 1 + the index in
 * dr_scene_desc.textures of the texture bound to `reflectance`, bits 20-31 the same for `transmittance` (0 = constant).
 * A textured parameter is looked up at every BSDF evaluation / sample; the constant in the struct is then the texture's
 * average (Texture::getAverage), which only the sampling weights of plastic / roughplastic use (plastic.cpp:188-205). */
#define DR_MAT_TEX_REFLECTANCE(i)   ((((uint32_t) (i) + 1u) & 0xfffu) << 8)
#define DR_MAT_TEX_TRANSMITTANCE(i) ((((uint32_t) (i) + 1u) & 0xfffu) << 20)
#define DR_MAX_TEXTURES 4095

typedef struct dr_material {
    int32_t  type;              /* dr_bsdf_type */
    uint32_t flags;             /* DR_MAT_* */
    float    reflectance[3];    /* diffuse reflectance | specularReflectance | plastic: diffuseReflectance */
    float    transmittance[3];  /* dielectric specularTransmittance | plastic: specularReflectance */
    float    eta[3];            /* conductor eta (RGB) | dielectric, roughdielectric, plastic: intIOR/extIOR in eta[0] */
    float    k[3];              /* conductor k (RGB) (plastic: ignored; the library derives its constants here) */
    float    alpha;             /* roughconductor / roughdielectric / roughplastic alpha (isotropic) */
    uint32_t table;             /* roughplastic: index of its table in dr_scene_desc.rough_tables (others: 0) */
} dr_material;                  /* 64 bytes */

/* Rough Fresnel transmittance of a roughplastic material (src/bsdfs/rtrans.h), reduced to the material's (eta, alpha) as
 * RoughPlastic::configure does (roughplastic.cpp:283-301): DR_ROUGH_TABLE_DOUBLES doubles per table --
 *   [0, 100)  m_externalRoughTransmittance after setEta(eta) + setAlpha(alpha): the 100 cos-theta samples that
 *             RoughTransmittance::eval interpolates (rtrans.h:136-146, evalCubicInterp1D, spline.cpp:23-60)
 *   [100]     m_internalRoughTransmittance->evalDiffuse(alpha) (after setEta(1 / eta))       (roughplastic.cpp:369)
 *   [101]     m_externalRoughTransmittance->evalDiffuse(alpha)                               (:307, informative)
 * The reference reads these from data/microfacet/{beckmann,ggx}.dat; a Mitsuba host gets them from its own RoughTransmittance
 * (shim/mts_plugin.cpp), other hosts from the .dat files (drmlt_mitsuba_b200/rough_tables.py). */
#define DR_ROUGH_TABLE_THETA   100
#define DR_ROUGH_TABLE_DOUBLES 104

/* A bitmap texture as the bidirectional path code sees it (src/textures/bitmap.cpp:432-455, include/mitsuba/render/mipmap.h:503-596):
 * without ray differentials (its.hasUVPartials is never set on this path) `eval` is a bilinear -- or, for filterType=nearest, a
 * box -- lookup in MIP level 0 at uv' = uv * uv_scale + uv_offset (Texture2D::eval, texture.cpp:112-121), out-of-range texels
 * resolved per axis by the wrap mode.  The reference keeps its MIP levels in half precision (bitmap.cpp:175-177), so a Mitsuba
 * host passes exactly what its texture holds.  Texels are linear RGB float32 (the host has already applied gamma / sRGB decoding; a
 * luminance bitmap repeats its channel), row y = 0 first.  The arithmetic of the lookup is double, like the reference's. */
typedef enum dr_wrap { DR_WRAP_REPEAT = 0, DR_WRAP_CLAMP = 1, DR_WRAP_MIRROR = 2, DR_WRAP_ZERO = 3, DR_WRAP_ONE = 4 } dr_wrap;
typedef struct dr_texture {
    uint32_t width, height;
    const float *texels;        /* 3 * width * height */
    uint32_t wrap_u, wrap_v;    /* dr_wrap (wrapModeU / wrapModeV) */
    uint32_t nearest;           /* filterType=nearest: MIPMap::evalBox instead of evalBilinear */
    uint32_t pad;
    double   uv_scale[2];       /* uscale, vscale (default 1; double: the reference parses them into Float = double) */
    double   uv_offset[2];      /* uoffset, voffset (default 0) */
} dr_texture;                   /* 64 bytes */

/* One area emitter = one emissive triangle mesh (src/emitters/area.cpp:67-215).
 * Its triangles are the contiguous range [first_tri, first_tri + n_tris). */
typedef struct dr_emitter {
    uint32_t first_tri;
    uint32_t n_tris;
    float    radiance[3];
    float    sampling_weight;   /* Emitter::getSamplingWeight(), default 1 (scene.cpp:380-383) */
} dr_emitter;

#define DR_TRI_SMOOTH 1u        /* interpolate vertex normals (mesh has normals) */
#define DR_TRI_UV_TANGENTS 2u  /* the mesh carries UV tangents (TriMesh::computeUVTangents, trimesh.cpp: requested by BSDFs whose textures
                                  use ray differentials, i.e. any bitmap texture that is not filterType=nearest): the shading frame is built on
                                  dp/du of the texture parameterisation instead of the edge p1 - p0 (skdtree.h:373-380); needs texcoords.
                                  In this reference EVERY mesh with texture coordinates has them (TriMesh::configure, trimesh.cpp:400-402) */
#define DR_TRI_NO_TEXCOORDS 4u /* this triangle's mesh has no texture coordinates although the scene passes `texcoords`: its.uv is the
                                  barycentric pair, as for a scene without texcoords */

/* Pinhole camera (src/sensors/perspective.cpp:107-460) */
typedef struct dr_camera {
    float   to_world[16];       /* row-major 4x4 camera-to-world, no scale */
    float   xfov_deg;           /* horizontal field of view in degrees */
    float   near_clip, far_clip;
    int32_t film_width, film_height;   /* film size == crop size (no crop window) */
} dr_camera;

typedef struct dr_scene_desc {
    uint32_t n_vertices, n_triangles, n_materials, n_emitters;
    const float    *positions;     /* 3 * n_vertices */
    const float    *normals;       /* 3 * n_vertices, or NULL (then no triangle may be DR_TRI_SMOOTH) */
    const uint32_t *indices;       /* 3 * n_triangles */
    const uint32_t *tri_material;  /* n_triangles, index into materials */
    const int32_t  *tri_emitter;   /* n_triangles, index into emitters or -1 */
    const uint32_t *tri_flags;     /* n_triangles, DR_TRI_* (may be NULL = 0) */
    const dr_material *materials;
    const dr_emitter  *emitters;
    dr_camera camera;
    const double *rough_tables;    /* n_rough_tables * DR_ROUGH_TABLE_DOUBLES, or NULL (no roughplastic material) */
    uint32_t n_rough_tables;
    uint32_t n_textures;           /* <= DR_MAX_TEXTURES */
    const float *texcoords;        /* 2 * n_vertices, or NULL: its.uv is then the barycentric pair (b1, b2) (skdtree.h:399-405) */
    const dr_texture *textures;    /* n_textures, referenced by DR_MAT_TEX_* */
} dr_scene_desc;

/* ------------------------------------------------------------------ config */

typedef enum dr_integrator { DR_INTEGRATOR_PSSMLT = 0, DR_INTEGRATOR_DRMLT = 1 } dr_integrator;
typedef enum dr_technique  { DR_TECH_PATH = 0, DR_TECH_BDPT = 1, DR_TECH_MMLT = 2 } dr_technique;
typedef enum dr_type       { DR_TYPE_GREEN = 0, DR_TYPE_MIRA = 1, DR_TYPE_ORBITAL = 2 } dr_type;
/* reconstruction filter of the film (the plugins under src/rfilters with their default parameters); DR_FILTER_TABLE: the caller supplies
 * what every ReconstructionFilter boils down to after configure() -- radius + the 32-entry discretisation (rfilter.cpp:37-55,
 * include/mitsuba/core/rfilter.h:76-77) -- which covers non-default parameters (gaussian stddev, Mitchell B / C, lanczos lobes) */
typedef enum dr_filter     { DR_FILTER_GAUSSIAN = 0, DR_FILTER_BOX = 1, DR_FILTER_TABLE = 2, DR_FILTER_TENT = 3, DR_FILTER_MITCHELL = 4,
                             DR_FILTER_CATMULLROM = 5, DR_FILTER_LANCZOS = 6 } dr_filter;

/* Parameter names, meaning and defaults follow DRMLT::DRMLT(props) (drmlt.cpp:178-351)
 * and PSSMLT::PSSMLT(props) (pssmlt.cpp:166-308).  Fill with dr_config_default()
 * and then override, or build from "-D key=value" strings with dr_config_set(). */
typedef struct dr_config {
    int32_t integrator;        /* dr_integrator : -D integrator=pssmlt|drmlt */
    int32_t technique;         /* dr_technique  : technique=path|bdpt|mmlt (required) */
    int32_t type;              /* dr_type       : type=green|mira|orbital|mirasym (drmlt, required) */
    int32_t max_depth;         /* maxDepth (-1 = unset; MMLT/path/bdpt here require > 0) */
    int32_t rr_depth;          /* rrDepth = 5 */
    int32_t direct_sampling;   /* directSampling = true (forced false for MMLT) */
    int32_t direct_samples;    /* directSamples = 16; separateDirect = directSamples >= 0 */
    int32_t luminance_samples; /* luminanceSamples = 100000 */
    float   p_large;           /* pLarge = 0.3 */
    int32_t work_units;        /* workUnits = -1 (auto) */
    int32_t kelemen_style_weights;  /* kelemenStyleWeights = true (forced false for MMLT) */
    int32_t two_stage;         /* twoStage = false; true: low-resolution first pass -> importance map (util.cpp:96-199) */
    int32_t timeout;           /* timeout = 0 seconds; > 0: the chain phase stops after that many seconds (drmlt_proc.cpp:519-521) */
    float   average_luminance; /* averageLuminance = -1 (use bootstrap estimate) */
    int32_t light_image;       /* lightImage = true */
    int32_t acceptance_map;    /* acceptanceMap = false (drmlt) */
    int32_t timid_after_large; /* timidAfterLarge = false (drmlt) */
    int32_t fix_emitter_path;  /* fixEmitterPath = false (drmlt, MMLT only) */
    int32_t use_mixture;       /* useMixture = false (drmlt) */
    float   sigma;             /* sigma = 1/64 */
    float   scale_second;      /* scaleSecond = 0.1 (<= 1) */
    int32_t kelemen_style_mutation; /* kelemenStyleMutation = true (pssmlt) */
    float   mutation_size_low;      /* mutationSizeLow = 1/1024 (pssmlt) */
    float   mutation_size_high;     /* mutationSizeHigh = 1/64 (pssmlt) */
    /* carried by other scene objects in the reference */
    int32_t sample_count;      /* sensor sampler's sampleCount = mutations per pixel (drmlt.cpp:400) */
    int32_t rfilter;           /* dr_filter: rfilter=gaussian (stddev .5) | box | tent | mitchell | catmullrom | lanczos; DR_FILTER_TABLE: see below */
    /* GPU execution knobs (no reference equivalent) */
    int32_t n_chains;          /* Markov chains (the reference's work units) per GPU; 0 = auto */
    uint64_t seed;             /* counter-based RNG key (reference: /dev/urandom, random.cpp:473-489) */
    int32_t rank, world_size;  /* chain / bootstrap shard of this process (1 GPU: 0, 1) */
    float   ray_epsilon;       /* 0 = 1e-4f (Mitsuba single precision, constants.h:29) */
    float   shadow_epsilon;    /* 0 = 1e-3f (constants.h:30) */
    /* two-stage MLT (drmlt.cpp:278-293, 401-418; BidirectionalUtils::mltLuminancePass, src/libbidir/util.cpp:96-199) */
    int32_t first_stage;                 /* firstStage = false (internal in the reference: this job IS the nested pass) */
    int32_t first_stage_size_reduction;  /* firstStageSizeReduction = 16 */
    /* film window, carried by the film plugin in the reference (src/librender/film.cpp:30-67): 0 = use dr_camera's film */
    int32_t film_width, film_height;     /* full film size override (the nested pass renders at size / firstStageSizeReduction) */
    int32_t crop_offset_x, crop_offset_y;/* cropOffsetX/Y = 0 */
    int32_t crop_width, crop_height;     /* cropWidth/Height = film size; the rendered image and all buffers have the CROP size */
    /* GPU execution knob: lanes of the wavefront machine.  0 or >= n_chains: one lane per chain (chains stay resident and
     * dr_job_run advances all of them).  Fewer lanes than chains: the lanes pull chains from a work-unit queue (the
     * reference's DRMLTProcess::generateWork, drmlt_proc.cpp:869-883); every dr_job_run then runs a fresh batch of
     * n_chains chains, `mutations_per_chain` mutations each, from newly resampled seeds. */
    int32_t n_lanes;
    /* GPU execution knob (MMLT, resident chains): a chain of path depth d runs ~ mutations_per_chain * dbar / d mutations
     * and the seeds are resampled ~ L * d, so that every lane finishes after about the same number of rays; the mutations
     * spent per depth stay ~ the bootstrap's per-depth luminance as in the reference.  depthBalance = true. */
    int32_t depth_balance;
    /* m_config.importanceMap (drmlt.h:57, internal): host pointer to crop_width*crop_height floats, or NULL.
     * With twoStage=true and NULL here, dr_render computes it with dr_importance_map first. */
    const float *importance_map;
    /* rfilter = DR_FILTER_TABLE: ReconstructionFilter::getRadius() and m_values[0..31] (m_values[i] = evalDiscretized((i + .5) * radius / 31)) */
    double filter_radius;
    double filter_table[32];
} dr_config;

void      dr_config_default(dr_config *cfg);
/* "technique"="mmlt", "sigma"="0.015625", "acceptanceMap"="true", ... (names as in the reference XML). */
dr_status dr_config_set(dr_config *cfg, const char *key, const char *value);
/* Applies the reference's constructor-time rules (MMLT forces directSampling=false and
 * kelemenStyleWeights=false; fixEmitterPath needs MMLT; scaleSecond <= 1; ...). */
dr_status dr_config_validate(dr_config *cfg);

/* ------------------------------------------------------------------- stats */

/* Numerators / denominators of the reference's StatsCounters
 * (drmlt_proc.cpp:34-49,715-768; pssmlt_proc.cpp:33-40). */
typedef struct dr_stats {
    uint64_t mutations;                 /* ++mutationCtr */
    uint64_t first_accept,  first_base;        /* firstLevelRatio */
    uint64_t large_accept,  large_base;        /* largeStepRatio */
    uint64_t bold_accept,   bold_base;         /* boldStepRatio (pssmlt: smallStepRatio) */
    uint64_t second_accept, second_base;       /* secondLevelRatio */
    uint64_t second_large_accept, second_large_base;
    uint64_t second_bold_accept,  second_bold_base;
    uint64_t accept, accept_base;              /* acceptanceRate */
    uint64_t paths, rays;               /* path evaluations / rays cast in the chain phase */
    uint64_t bootstrap_paths, bootstrap_rays;
    double   luminance;                 /* normalization b actually used */
    double   bootstrap_ms, chains_ms, total_ms;  /* device-timed phases */
    uint64_t kernel_launches;           /* launches of this library's kernels */
    uint64_t rounds;                    /* wavefront rounds (one ray per active chain and round) */
    /* per-stage device time, only filled while stage profiling is on (dr_job_profile) */
    double   trace_ms, walk_ms, chain_ms;
    uint64_t trace_launches, walk_launches, chain_launches;
    double   direct_ms;                 /* separate direct-illumination pass (directSamples > 0) */
    double   first_stage_ms;            /* two-stage MLT: wall time of the nested first-stage pass (dr_render only) */
} dr_stats;

/* ------------------------------------------------------------- entry points */

typedef struct dr_scene_t *dr_scene;
typedef struct dr_job_t   *dr_job;

int         dr_abi_version(void);
const char *dr_last_error(void);
int         dr_device_count(void);

/* Flatten -> BVH build (host) -> upload.  Replaces Scene::initialize + ShapeKDTree build
 * (scene.cpp:332-394, skdtree.cpp) for this path. */
dr_status dr_scene_create(const dr_scene_desc *desc, int device, dr_scene *out);
/* dr_scene_create with options.  DR_SCENE_BVH_GPU: build the BVH on the device (Morton-code LBVH, ~100x faster than the host's binned-SAH
 * build and ~15 % more expensive to traverse): for one-shot renders, where the build is on the critical path (the reference builds its
 * kd-tree in Scene::initialize, src/librender/scene.cpp:289-356, before every render).  Scenes of < 1024 triangles, or whose LBVH would
 * exceed the traversal stack, use the host build regardless.  The images do not depend on the builder. */
#define DR_SCENE_BVH_HOST 0u
#define DR_SCENE_BVH_GPU  1u
dr_status dr_scene_create_ex(const dr_scene_desc *desc, int device, uint32_t flags, dr_scene *out);
/* which builder made the tree (DR_SCENE_BVH_*), its node count, the traversal-stack bound and the build time (host clock, ms) */
void      dr_scene_bvh_info(dr_scene scene, int32_t *builder, int32_t *n_nodes, int32_t *stack_bound, double *build_ms);
void      dr_scene_destroy(dr_scene scene);
/* A replica of `scene` on another GPU, filled from the staging copies the original keeps (no second BVH build). */
dr_status dr_scene_clone(dr_scene scene, int device, dr_scene *out);
/* Repeat the host->device copies of the flattened scene buffers (BVH nodes, triangles, normals,
 * emitter tables, materials) from the staging copies kept by dr_scene_create; `bytes` receives
 * the number of bytes copied.  This is the per-job upload a plugin pays when the scene changed. */
dr_status dr_scene_reupload(dr_scene scene, int64_t *bytes);

/* Whole job on one GPU, HOST buffers: DRMLT::render / PSSMLT::render (drmlt.cpp:393-611) +
 * develop (drmlt_proc.cpp:813-854).  `image_rgb` receives the developed W*H*3 float image
 * (what the reference hands to film->setBitmap); in acceptanceMap mode it holds the R/G counts. */
dr_status dr_render(dr_scene scene, const dr_config *cfg, float *image_rgb, dr_stats *stats);
void      dr_cancel(dr_scene scene);
/* dr_render with periodic develops of the partial result -- DRMLTProcess::processResult's develop + signalRefresh for
 * interactive jobs (drmlt_proc.cpp:856-867) and the images `mitsuba -r <sec>` dumps through Scene::flush
 * (src/librender/scene.cpp:468-511).  The chain phase runs in slices of about `refresh_seconds`; after each slice the
 * film is developed into `image_rgb` and `fn(image_rgb, width, height, seconds since the chains started, stats, user)`
 * is called from the calling thread; a non-zero return cancels the job (DR_ERR_CANCELLED). */
typedef int (*dr_refresh_fn)(const float *image_rgb, int32_t width, int32_t height, double seconds, const dr_stats *stats, void *user);
dr_status dr_render_progressive(dr_scene scene, const dr_config *cfg, float *image_rgb, dr_stats *stats,
                                double refresh_seconds, dr_refresh_fn fn, void *user);

/* dr_render on `n` GPUs of one node: scenes[g] is the replica on its GPU (dr_scene_create + dr_scene_clone), scenes[0]'s GPU
 * develops the image.  One host thread per GPU; chains, bootstrap samples and mutation budget are sharded by rank; ONE NCCL
 * all-reduce of {sum luminance, count} gives the global b (the reference averages its initialisation threads' means,
 * drmlt.cpp:498-546), ONE NCCL reduce sums the films onto the first GPU (DRMLTProcess::processResult, drmlt_proc.cpp:856-867).
 * `stats`: counters summed over the GPUs, phase times of the slowest.  NCCL is loaded on first use (libnccl.so.2). */
dr_status dr_render_multi(const dr_scene *scenes, int32_t n, const dr_config *cfg, float *image_rgb, dr_stats *stats);

/* ---- staged API (multi-GPU drivers, tests, benchmarks) -------------------
 * One job = one rank's share of a render.  Sequence:
 *   dr_job_create -> dr_job_bootstrap -> [all-reduce sum_lum/count across ranks] ->
 *   dr_job_seed_chains(b) -> dr_job_run(...)* -> dr_job_film / dr_job_develop.        */
dr_status dr_job_create(dr_scene scene, const dr_config *cfg, dr_job *out);
void      dr_job_destroy(dr_job job);
/* Bootstrap on this rank: generateSeeds (pathsampler.cpp:859-960).  Returns the local
 * luminance sum and the number of (non-NaN) samples; b = max_depth_factor * sum / count. */
dr_status dr_job_bootstrap(dr_job job, double *sum_luminance, double *count);
/* Resample chain seeds from the local seed pool (luminance CDF) and replay them. */
dr_status dr_job_seed_chains(dr_job job, double b);
/* Advance every chain by `mutations_per_chain` iterations of the MLT loop. */
dr_status dr_job_run(dr_job job, int64_t mutations_per_chain);
/* The separate direct-illumination image of the reference's default configuration (directSamples > 0:
 * BidirectionalUtils::renderDirectComponent, src/libbidir/util.cpp:30-94, with the `direct` integrator,
 * src/integrators/direct/direct.cpp:144-305); dr_job_develop adds it (drmlt_proc.cpp:846-847).  No-op for directSamples <= 0. */
dr_status dr_job_direct(dr_job job);
/* Device pointer of the accumulation film: W*H*4 floats (RGB + one unused float per pixel: a splat is one 16-byte vector
 * atomic), un-normalised; *n_floats = W*H*4 -- the buffer and count a multi-GPU driver hands to ncclReduce. */
dr_status dr_job_film_device(dr_job job, float **film_dev, int64_t *n_floats);
/* Develop: image = film * (b / mean pixel luminance)  (drmlt_proc.cpp:823-849).  Host out. */
dr_status dr_job_develop(dr_job job, float *image_rgb);
dr_status dr_job_stats(dr_job job, dr_stats *stats);
/* Stage profiling: CUDA events around every stage of every wavefront round (trace | walk+connect or
 * path tracer | chain); the sums appear in dr_stats.  Off by default (env DRMLT_PROFILE_STAGES=1 turns it on). */
void      dr_job_profile(dr_job job, int on);
int64_t   dr_job_num_chains(dr_job job);
int64_t   dr_job_total_mutations(dr_job job);   /* W*H*sampleCount share of this rank */

/* ---- two-stage MLT ------------------------------------------------------- */

/* Size of the image a job with this configuration renders (the crop window of the film). */
dr_status dr_film_size(dr_scene scene, const dr_config *cfg, int32_t *width, int32_t *height);
/* Configuration of the nested first-stage job (util.cpp:100-159): firstStage = true, film and crop window divided by
 * firstStageSizeReduction (at least 1 pixel), sampleCount multiplied by it, no importance map. */
dr_status dr_first_stage_config(dr_scene scene, const dr_config *cfg, dr_config *nested);
/* Developed first-stage image (w*h*3 floats, host) -> luminance (spectrum.h:640-650) -> up-sampled to W*H with the
 * gaussian reconstruction filter, clamped boundary, values clamped to [0, inf) (util.cpp:180-196,
 * Resampler include/mitsuba/core/rfilter.h:107-324).  Runs on the GPU of `scene`. */
dr_status dr_resample_luminance(dr_scene scene, const float *image_rgb, int32_t w, int32_t h, int32_t W, int32_t H, float *map);
/* BidirectionalUtils::mltLuminancePass on this GPU: nested render with dr_first_stage_config, then
 * dr_resample_luminance.  `map` receives crop_width*crop_height floats of the job described by `cfg`. */
dr_status dr_importance_map(dr_scene scene, const dr_config *cfg, float *map, dr_stats *nested_stats);

/* ---- replay / parity entry points ---------------------------------------- */

typedef struct dr_ray { float o[3]; float mint; float d[3]; float maxt; } dr_ray;
typedef struct dr_hit { float t, u, v; int32_t prim; } dr_hit;   /* prim = -1: miss */

/* Closest-hit over `n` rays (Scene::rayIntersect, skdtree.cpp:111-138).  Host buffers.
 * `shadow`!=0: any-hit only (t/u/v undefined, prim >= 0 means occluded). */
dr_status dr_trace_rays(dr_scene scene, const dr_ray *rays, int64_t n, int shadow, dr_hit *hits);
/* Replay entry point of the texture stage: texture `texture` of the scene evaluated on the device at n intersection uv pairs
 * (uv[2i], uv[2i+1]) exactly as the BSDF stage does -- Texture2D::eval without ray differentials (texture.cpp:112-121) ->
 * TMIPMap::evalBilinear / evalBox (mipmap.h:566-596); rgb gets 3 doubles per lookup. */
dr_status dr_texture_eval(dr_scene scene, uint32_t texture, const double *uv, int64_t n, double *rgb);

#define DR_MAX_SPLATS 12
typedef struct dr_path_result {
    float   luminance;          /* SplatList::luminance (un-normalised) */
    int32_t n_splats;
    int32_t s, t;               /* MMLT strategy (pathsampler.cpp:104-129) | BDPT: last vertex index of the emitter / sensor subpath | else -1 */
    float   mis_weight;         /* MMLT: Path::miWeight of the (s,t) strategy; else 0 */
    float   pos[DR_MAX_SPLATS][2];
    float   value[DR_MAX_SPLATS][3];   /* un-normalised RGB contributions */
    int32_t n_rays;
} dr_path_result;

/* PathSampler::sampleSplats (pathsampler.cpp:79-571) on `n` replayed primary-sample
 * vectors.  u_* are [n][dim_*] row-major; depth[i] is the MMLT depth (ignored otherwise). */
dr_status dr_eval_paths(dr_scene scene, const dr_config *cfg,
                        const float *u_sensor, int dim_sensor,
                        const float *u_emitter, int dim_emitter,
                        const float *u_direct, int dim_direct,
                        const int32_t *depth, int64_t n, dr_path_result *out);

/* Per-mutation record of one chain (accept/reject parity under identical uniforms) */
typedef struct dr_step_record {
    float   L_x, L_y, L_z;      /* current / stage-1 / stage-2 luminance (L_z = 0 if no stage 2) */
    float   a1, a2;
    uint8_t large_step, accept1, did_second, accept2;
} dr_step_record;

/* Run `n_chains` chains for `steps` mutations each from bootstrap indices `seed_index[i]`
 * (MMLT depth `depth[i]`), recording every decision.  Film contributions are discarded
 * unless `film` (host, W*H*3) is given.  Uses the same counter-based uniforms as dr_job_run. */
dr_status dr_chain_steps(dr_scene scene, const dr_config *cfg, double b,
                         const uint64_t *seed_index, const int32_t *depth, const uint64_t *chain_id,
                         int64_t n_chains, int64_t steps, dr_step_record *records, float *film);

/* Chains of the REFERENCE replayed on a recorded uniform stream (SURVEY 8b `uniforms`; DRMLTRenderer::process
 * src/integrators/drmlt/drmlt_proc.cpp:386-771, processMixture :161-380, PSSMLTRenderer::process
 * src/integrators/pssmlt/pssmlt_proc.cpp:110-285): like dr_chain_steps, but every uniform -- the seed state, the large-step and
 * acceptance coins, the samplers' proposal draws -- comes from the caller's table instead of the counter-based generator.
 * A sequential stream does not say WHICH uniform a value is (the reference fills a sampler's proposal lazily, when the path first
 * touches it); the table does: it is the stream sorted into the keyed address space.  Per chain, D = uniform_dim (even)
 * coordinates per sampler (0 sensor, 1 emitter, 2 direct), doubles, NaN = never drawn by the recorded chain (the coordinate
 * then keeps its value in that step):
 *     [3][D]                    seed state: the replayed + padded current vector of each sampler (drmlt_proc.cpp:467-504)
 *     per mutation m < steps:   [4] coins (0 large step, 1 accept 1, 2 accept 2, 3 mixture),
 *                               [3][2 D] stage-1 draws, index 2 * coordinate + draw,  [3][2 D] stage-2 draws
 * i.e. 3 D + steps (4 + 12 D) doubles per chain, chains back to back.  tests/ builds the table with the oracle
 * (orc_chain_stream) from the streams of tests/golden/ref_chain.npz. */
dr_status dr_chain_replay(dr_scene scene, const dr_config *cfg, double b, const int32_t *depth, int64_t n_chains, int64_t steps,
                          const double *uniforms, int32_t uniform_dim, dr_step_record *records, float *film);

/* The film by itself (parity entry point): `n` splats (pos [n][2] in pixel coordinates, rgb [n][3]) through the reconstruction
 * filter of `cfg` into a w x h film (host, w*h*3) -- ImageBlock::put (include/mitsuba/render/imageblock.h:149-196) with the
 * 32-entry table of ReconstructionFilter::configure (src/libcore/rfilter.cpp:37-55); non-finite or negative values are rejected. */
dr_status dr_splat_points(int device, const dr_config *cfg, int32_t w, int32_t h, const float *pos, const float *rgb, int64_t n, float *film_rgb);

/* Bootstrap luminances of samples [first, first+n) (before the x maxDepth MMLT scaling). */
dr_status dr_bootstrap_luminance(dr_scene scene, const dr_config *cfg,
                                 uint64_t first, int64_t n, float *luminance, int32_t *depth);

/* The direct-illumination image alone (parity entry point).  `li` (optional): un-filtered radiance of every pixel
 * sample, [H][W][pixelSamples][3] doubles, pixelSamples x shadingSamples being the split of directSamples (util.cpp:44-54). */
dr_status dr_direct_image(dr_scene scene, const dr_config *cfg, float *image_rgb, double *li);

/* Primary-sample dimensions per sampler (findMaxDimensions, pssmlt_utils.h:27-77). */
void dr_max_dimensions(const dr_config *cfg, int depth, int *sensor, int *emitter, int *direct);

#ifdef __cplusplus
}
#endif
#endif /* DRMLT_B200_H */
