// ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_math.hpp header for what pins it: the reference's own
// sources compiled into oracle/_ref, tests/test_ref_pins.py).
//
// C entry points of the CPU restatement, loaded with ctypes by tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference leg.  Mirrors the replay entry points of
// include/drmlt_b200.h so that parity tests call both sides with the same buffers.
#include "orc_mlt.hpp"
#include <thread>
#include <atomic>
#include <chrono>
#include <cstdio>

using namespace orc;

struct OrcScene { Scene sc; };

static void applyEps(Scene &sc, const dr_config *cfg) {
    // The oracle defaults to the double-precision constants of the reference's default build
    // (constants.h:25-27); parity tests pass the float-build values the GPU uses.
    sc.epsilon = cfg && cfg->ray_epsilon > 0 ? (Float) cfg->ray_epsilon : 1e-7;
    sc.shadowEpsilon = cfg && cfg->shadow_epsilon > 0 ? (Float) cfg->shadow_epsilon : 1e-5;
    // film / crop window of the job (Film::Film, src/librender/film.cpp:30-48): 0 = the camera's film, no crop
    int fW = cfg && cfg->film_width > 0 ? cfg->film_width : sc.cam.filmW, fH = cfg && cfg->film_height > 0 ? cfg->film_height : sc.cam.filmH;
    int cX = cfg ? cfg->crop_offset_x : 0, cY = cfg ? cfg->crop_offset_y : 0;
    int cW = cfg && cfg->crop_width > 0 ? cfg->crop_width : fW, cH = cfg && cfg->crop_height > 0 ? cfg->crop_height : fH;
    sc.cam.setWindow(fW, fH, cX, cY, cW, cH);
}

extern "C" {

void *orc_scene_create(const dr_scene_desc *desc) {
    OrcScene *s = new OrcScene();
    s->sc.load(*desc);
    return s;
}
void orc_scene_destroy(void *h) { delete (OrcScene *) h; }

void orc_philox(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t *out) {
    uint32_t c[4] = { c0, c1, c2, c3 };
    Philox::gen(c, seed);
    for (int i = 0; i < 4; ++i) out[i] = c[i];
}
float orc_uniform(uint64_t seed, uint32_t stream, uint64_t a, uint32_t b, uint32_t j) { return keyedUniform(seed, stream, a, b, j); }

void orc_max_dimensions(const dr_config *cfg, int depth, int *sensor, int *emitter, int *direct) {
    MaxDim md = findMaxDimensions(cfg->max_depth, cfg->rr_depth, depth, cfg->technique, cfg->direct_sampling != 0, false);
    *sensor = md.sensor; *emitter = md.emitter; *direct = md.direct;
}
// ... with the scene's own hasRoughDielectric (pssmlt_utils.h:35-45); held against the reference by tests/test_ref_pins.py
void orc_max_dimensions_scene(void *h, const dr_config *cfg, int depth, int *out3) {
    MaxDim md = findMaxDimensions(cfg->max_depth, cfg->rr_depth, depth, cfg->technique, cfg->direct_sampling != 0, ((OrcScene *) h)->sc.hasRoughDielectric);
    out3[0] = md.sensor; out3[1] = md.emitter; out3[2] = md.direct;
}

int orc_trace_rays(void *h, const dr_ray *rays, int64_t n, int shadow, float ray_epsilon, dr_hit *hits) {
    Scene &sc = ((OrcScene *) h)->sc;
    dr_config tmp; memset(&tmp, 0, sizeof(tmp)); tmp.ray_epsilon = ray_epsilon;
    applyEps(sc, &tmp);
    for (int64_t i = 0; i < n; ++i) {
        Ray r; r.o = Vec3(rays[i].o[0], rays[i].o[1], rays[i].o[2]); r.d = Vec3(rays[i].d[0], rays[i].d[1], rays[i].d[2]);
        r.mint = rays[i].mint; r.maxt = rays[i].maxt;
        Float mint = r.mint, maxt = r.maxt;
        Float t = 0, u = 0, v = 0; int prim = -1;
        bool hit = maxt > mint && sc.traverse(r, mint, maxt, shadow != 0, t, u, v, prim);
        hits[i].prim = hit ? prim : -1;
        hits[i].t = hit ? (float) t : 0; hits[i].u = hit ? (float) u : 0; hits[i].v = hit ? (float) v : 0;
    }
    return 0;
}

static void fillResult(const SplatList &list, uint64_t rays, dr_path_result *out) {
    memset(out, 0, sizeof(*out));
    out->luminance = (float) list.luminance;
    out->n_splats = (int32_t) std::min(list.size(), (size_t) DR_MAX_SPLATS);
    out->s = list.s; out->t = list.t;
    out->mis_weight = (float) list.misWeight;
    for (int k = 0; k < out->n_splats; ++k) {
        out->pos[k][0] = (float) list.splats[k].first.x; out->pos[k][1] = (float) list.splats[k].first.y;
        out->value[k][0] = (float) list.splats[k].second.r; out->value[k][1] = (float) list.splats[k].second.g;
        out->value[k][2] = (float) list.splats[k].second.b;
    }
    out->n_rays = (int32_t) rays;
}

// double-precision outputs for tolerance analysis: lum[n], rgb of splat 0 [n][3], mis[n]
int orc_eval_paths(void *h, const dr_config *cfg, const float *u_sensor, int dim_sensor, const float *u_emitter, int dim_emitter,
                   const float *u_direct, int dim_direct, const int32_t *depth, int64_t n, dr_path_result *out, double *lum_out) {
    Scene &sc = ((OrcScene *) h)->sc;
    applyEps(sc, cfg);
    unsigned nt = std::max(1u, std::thread::hardware_concurrency());
    std::atomic<int64_t> next(0);
    auto work = [&]() {
        SplatList list;
        for (;;) {
            int64_t i = next.fetch_add(256);
            if (i >= n) break;
            for (int64_t j = i; j < std::min(n, i + 256); ++j) {
                ArraySampler se(u_sensor + j * dim_sensor, dim_sensor), em(u_emitter + j * dim_emitter, dim_emitter),
                    di(u_direct + j * dim_direct, dim_direct);
                PathSampler ps(&sc, pathConfigOf(*cfg), &em, &se, &di);
                ps.sampleSplats(list, depth ? depth[j] : -1);
                fillResult(list, ps.ctx.rays, &out[j]);
                if (lum_out) lum_out[j] = list.luminance;
            }
        }
    };
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t) th.emplace_back(work);
    for (auto &t : th) t.join();
    return 0;
}

int orc_bootstrap_luminance(void *h, const dr_config *cfg, uint64_t first, int64_t n, float *luminance, int32_t *depth, double *lum64) {
    Scene &sc = ((OrcScene *) h)->sc;
    applyEps(sc, cfg);
    unsigned nt = std::max(1u, std::thread::hardware_concurrency());
    std::atomic<int64_t> next(0);
    auto work = [&]() {
        SplatList list;
        for (;;) {
            int64_t i = next.fetch_add(256);
            if (i >= n) break;
            for (int64_t j = i; j < std::min(n, i + 256); ++j) {
                bootstrapSample(sc, *cfg, first + (uint64_t) j, list);
                luminance[j] = (float) list.luminance;
                if (lum64) lum64[j] = list.luminance;
                if (depth) depth[j] = bootstrapDepth(*cfg, first + (uint64_t) j);
            }
        }
    };
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t) th.emplace_back(work);
    for (auto &t : th) t.join();
    return 0;
}

// Chains with recorded decisions (mirror of dr_chain_steps).  film (optional) is W*H*3 doubles->floats.
int orc_chain_steps(void *h, const dr_config *cfg, double b, const uint64_t *seed_index, const int32_t *depth,
                    const uint64_t *chain_id, int64_t n_chains, int64_t steps, dr_step_record *records, float *film_out,
                    dr_stats *stats_out, int threads) {
    Scene &sc = ((OrcScene *) h)->sc;
    applyEps(sc, cfg);
    unsigned nt = threads > 0 ? (unsigned) threads : std::max(1u, std::thread::hardware_concurrency());
    nt = (unsigned) std::min<int64_t>(nt, std::max<int64_t>(1, n_chains));
    std::vector<Film> films(film_out ? nt : 0);
    for (auto &f : films) f.init(sc.cam.resX, sc.cam.resY, *cfg);
    std::vector<ChainStats> tstats(nt);
    std::atomic<int64_t> next(0);
    auto work = [&](unsigned tid) {
        ChainRunner runner(sc, *cfg, b, film_out ? &films[tid] : nullptr);
        std::vector<StepRecord> recs(records ? steps : 0);
        for (;;) {
            int64_t i = next.fetch_add(1);
            if (i >= n_chains) break;
            runner.run(chain_id[i], seed_index[i], depth ? depth[i] : -1, (uint64_t) steps, records ? recs.data() : nullptr);
            if (records)
                for (int64_t m = 0; m < steps; ++m) {
                    dr_step_record &o = records[i * steps + m];
                    const StepRecord &r = recs[m];
                    o.L_x = (float) r.L_x; o.L_y = (float) r.L_y; o.L_z = (float) r.L_z; o.a1 = (float) r.a1; o.a2 = (float) r.a2;
                    o.large_step = r.large; o.accept1 = r.acc1; o.did_second = r.did2; o.accept2 = r.acc2;
                }
        }
        tstats[tid] = runner.stats;
    };
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t) th.emplace_back(work, t);
    for (auto &t : th) t.join();
    ChainStats total;
    for (auto &s : tstats) total.add(s);
    if (film_out) {
        size_t n = (size_t) sc.cam.resX * sc.cam.resY * 3;
        for (size_t i = 0; i < n; ++i) {
            double acc = 0;
            for (auto &f : films) acc += f.data[i];
            film_out[i] = (float) acc;
        }
    }
    if (stats_out) {
        memset(stats_out, 0, sizeof(*stats_out));
        stats_out->mutations = total.mutations;
        stats_out->first_accept = total.first_accept; stats_out->first_base = total.first_base;
        stats_out->large_accept = total.large_accept; stats_out->large_base = total.large_base;
        stats_out->bold_accept = total.bold_accept; stats_out->bold_base = total.bold_base;
        stats_out->second_accept = total.second_accept; stats_out->second_base = total.second_base;
        stats_out->second_large_accept = total.second_large_accept; stats_out->second_large_base = total.second_large_base;
        stats_out->second_bold_accept = total.second_bold_accept; stats_out->second_bold_base = total.second_bold_base;
        stats_out->accept = total.accept; stats_out->accept_base = total.accept_base;
        stats_out->paths = total.paths; stats_out->rays = total.rays;
        stats_out->luminance = b;
    }
    return 0;
}

// ONE chain on recorded uniform streams of the reference (oracle/ref/ref_sampler.cpp: ref_drmlt_chain / ref_pssmlt_chain): the seed is
// replayed from `boot_stream` (the ReplayableSampler's stream from the seed's sample index on), everything else -- fillReplay,
// the large-step coins, the samplers' lazy fills in touch order, the acceptance coins -- is consumed, in the reference's call
// order, from `worker_stream`.  film_out is W*H*3 DOUBLES (the work unit's ImageBlock); returns the number of worker uniforms
// consumed (< 0: a stream was too short).
// table_out / table_in (optional): the replay table in keyed address space (orc_mlt.hpp KeyedSource), table_dim coordinates
// per sampler, 3 * dim + steps * (4 + 12 * dim) doubles; with table_in the streams are not used (pass null).
long long orc_chain_stream(void *h, const dr_config *cfg, double b, int depth, const double *boot_stream, long long n_boot,
                           const double *worker_stream, long long n_worker, long long steps, dr_step_record *records,
                           double *film_out, dr_stats *stats_out, double *table_out, const double *table_in, int table_dim) {
    Scene &sc = ((OrcScene *) h)->sc;
    applyEps(sc, cfg);
    Film film;
    film.init(sc.cam.resX, sc.cam.resY, *cfg);
    ChainRunner runner(sc, *cfg, b, film_out ? &film : nullptr);
    // the streams are read through padded copies, so that an (erroneous) over-read is detected instead of crashing
    std::vector<double> boot, work;
    if (!table_in) {
        boot.assign(boot_stream, boot_stream + n_boot); work.assign(worker_stream, worker_stream + n_worker);
        const size_t pad = 4096;
        boot.resize(boot.size() + pad, 0.5); work.resize(work.size() + pad, 0.5);
        runner.bootStream = boot.data(); runner.workerStream = work.data();
    }
    if (table_out) {
        const size_t n = 3 * (size_t) table_dim + (size_t) steps * (4 + 12 * (size_t) table_dim);
        for (size_t i = 0; i < n; ++i) table_out[i] = std::numeric_limits<double>::quiet_NaN();
    }
    runner.tableOut = table_in ? nullptr : table_out; runner.tableIn = table_in; runner.tableDim = table_dim;
    std::vector<StepRecord> recs(records ? steps : 0);
    runner.run(0, 0, depth, (uint64_t) steps, records ? recs.data() : nullptr);
    if (records)
        for (long long m = 0; m < steps; ++m) {
            dr_step_record &o = records[m];
            const StepRecord &r = recs[m];
            o.L_x = (float) r.L_x; o.L_y = (float) r.L_y; o.L_z = (float) r.L_z; o.a1 = (float) r.a1; o.a2 = (float) r.a2;
            o.large_step = r.large; o.accept1 = r.acc1; o.did_second = r.did2; o.accept2 = r.acc2;
        }
    if (film_out) for (size_t i = 0; i < film.data.size(); ++i) film_out[i] = film.data[i];
    if (stats_out) {
        const ChainStats &t = runner.stats;
        memset(stats_out, 0, sizeof(*stats_out));
        stats_out->mutations = t.mutations;
        stats_out->first_accept = t.first_accept; stats_out->first_base = t.first_base;
        stats_out->large_accept = t.large_accept; stats_out->large_base = t.large_base;
        stats_out->bold_accept = t.bold_accept; stats_out->bold_base = t.bold_base;
        stats_out->second_accept = t.second_accept; stats_out->second_base = t.second_base;
        stats_out->second_large_accept = t.second_large_accept; stats_out->second_large_base = t.second_large_base;
        stats_out->second_bold_accept = t.second_bold_accept; stats_out->second_bold_base = t.second_bold_base;
        stats_out->accept = t.accept; stats_out->accept_base = t.accept_base;
        stats_out->paths = t.paths; stats_out->rays = t.rays; stats_out->luminance = b;
    }
    if (!table_in && runner.streamUsed > (size_t) n_worker) return -1;
    return (long long) runner.streamUsed;
}

// Whole render on the CPU (the reported CPU baseline; "port" of DRMLT::render / PSSMLT::render,
// drmlt.cpp:393-611).  n_boot bootstrap samples -> b and the seed CDF -> n_chains chains of
// `steps` mutations each, one chain per work item, `threads` host threads.
int orc_direct_image(void *h, const dr_config *cfg, float *image_rgb, double *li_out);
void orc_first_stage_config(void *h, const dr_config *cfg, dr_config *nested);
void orc_resample_luminance(const float *image_rgb, int w, int h, int W, int H, float *map);
int orc_render(void *h, const dr_config *cfgIn, int64_t n_boot, int64_t n_chains, int64_t steps, int threads,
               float *image_rgb, dr_stats *stats_out, double *seconds_chains) {
    Scene &sc = ((OrcScene *) h)->sc;
    dr_config cfgLocal = *cfgIn;
    const dr_config *cfg = &cfgLocal;
    std::vector<float> importance;
    if (cfgLocal.two_stage && !cfgLocal.first_stage && !cfgLocal.importance_map) {
        // BidirectionalUtils::mltLuminancePass (util.cpp:96-199): nested render at reduced size with sizeFactor x the
        // mutations per pixel (= total mutations / sizeFactor), developed, luminance, up-sampled
        dr_config nested;
        orc_first_stage_config(h, cfg, &nested);
        applyEps(sc, &nested);
        const int w = (int) sc.cam.resX, hh = (int) sc.cam.resY;
        std::vector<float> img((size_t) 3 * w * hh);
        int rc = orc_render(h, &nested, n_boot, n_chains, std::max<int64_t>(1, steps / std::max(1, cfgLocal.first_stage_size_reduction)), threads,
                            img.data(), nullptr, nullptr);
        if (rc) return rc;
        applyEps(sc, cfg);
        const int W = (int) sc.cam.resX, H = (int) sc.cam.resY;
        importance.resize((size_t) W * H);
        orc_resample_luminance(img.data(), w, hh, W, H, importance.data());
        cfgLocal.importance_map = importance.data();
    }
    applyEps(sc, cfg);
    std::vector<float> lum(n_boot);
    std::vector<int32_t> dep(n_boot);
    std::vector<double> lum64(n_boot);
    orc_bootstrap_luminance(h, cfg, 0, n_boot, lum.data(), dep.data(), lum64.data());
    // generateSeeds: running mean over non-NaN samples, x maxDepth for MMLT (pathsampler.cpp:922-934)
    double sum = 0, tok = 0;
    DiscreteDistribution seedPDF;
    std::vector<int64_t> pool;
    // Two-stage MLT: b from the un-weighted luminances (pathsampler.cpp:899-901); the seed pool is weighted by the chains'
    // TARGET, the importance-re-weighted luminance -- the short chains of the GPU design must start in their stationary
    // distribution (the reference seeds ~ the un-weighted luminance and relies on ~100 000-mutation work units); mirrors dr_job_bootstrap.
    std::vector<double> target(lum64);
    if (cfg->importance_map && !cfg->first_stage) {
        const int W = (int) sc.cam.resX, H = (int) sc.cam.resY;
        SplatList list;
        for (int64_t i = 0; i < n_boot; ++i) {
            if (!(lum64[i] > 0)) continue;
            bootstrapSample(sc, *cfg, (uint64_t) i, list);
            list.normalize(cfg->importance_map, W, H);
            target[i] = list.luminance;
        }
    }
    for (int64_t i = 0; i < n_boot; ++i) {
        if (std::isnan(lum64[i])) continue;
        tok += 1; sum += lum64[i];
        if (lum64[i] != 0 && target[i] > 0 && std::isfinite(target[i])) { pool.push_back(i); seedPDF.append(target[i]); }
    }
    double b = tok > 0 ? sum / tok : 0;
    if (cfg->technique == DR_TECH_MMLT) b *= cfg->max_depth;
    if (b == 0 || pool.empty()) return DR_ERR_ZERO_LUMINANCE;
    seedPDF.normalize();
    if (cfg->average_luminance != -1.0f) b = cfg->average_luminance;
    if (cfg->acceptance_map) b = 1.0;
    std::vector<uint64_t> seedIdx(n_chains), chainId(n_chains);
    std::vector<int32_t> depth(n_chains);
    for (int64_t c = 0; c < n_chains; ++c) {
        // one stratum of the seed CDF per chain (the reference draws independently, pathsampler.cpp:946-954); mirrors k_resample
        Float u = ((Float) c + (Float) keyedUniform(cfg->seed, S_RESAMPLE, (uint64_t) c, 0, 0)) / (Float) n_chains;
        int64_t s = pool[seedPDF.sample(u)];
        seedIdx[c] = (uint64_t) s; chainId[c] = (uint64_t) c; depth[c] = dep[s];
    }
    std::vector<float> film((size_t) sc.cam.resX * sc.cam.resY * 3);
    auto t0 = std::chrono::steady_clock::now();
    orc_chain_steps(h, cfg, b, seedIdx.data(), depth.data(), chainId.data(), n_chains, steps, nullptr, film.data(), stats_out, threads);
    auto t1 = std::chrono::steady_clock::now();
    if (seconds_chains) *seconds_chains = std::chrono::duration<double>(t1 - t0).count();
    if (image_rgb) {
        Film f; f.init(sc.cam.resX, sc.cam.resY, *cfg);
        for (size_t i = 0; i < film.size(); ++i) f.data[i] = film[i];
        develop(f, b, cfg->acceptance_map != 0, image_rgb, cfg->first_stage ? nullptr : cfg->importance_map);
        if (cfg->direct_samples > 0 && !cfg->acceptance_map && !(cfg->two_stage && cfg->first_stage)) {   // `!nested` (drmlt.cpp:478)   // value += direct[i] (drmlt_proc.cpp:846-847)
            std::vector<float> direct(film.size());
            orc_direct_image(h, cfg, direct.data(), nullptr);
            for (size_t i = 0; i < film.size(); ++i) image_rgb[i] += direct[i];
        }
    }
    if (stats_out) stats_out->luminance = b;
    return 0;
}

// Nested first-stage configuration (util.cpp:100-159); mirrors dr_first_stage_config.
void orc_first_stage_config(void *h, const dr_config *cfg, dr_config *nested) {
    Scene &sc = ((OrcScene *) h)->sc;
    applyEps(sc, cfg);
    const int f = std::max(1, cfg->first_stage_size_reduction);
    const int fW = cfg->film_width > 0 ? cfg->film_width : sc.cam.filmW, fH = cfg->film_height > 0 ? cfg->film_height : sc.cam.filmH;
    dr_config c = *cfg;
    c.two_stage = 1; c.first_stage = 1;
    c.film_width = std::max(1, fW / f); c.film_height = std::max(1, fH / f);
    c.crop_width = std::max(1, (int) sc.cam.resX / f); c.crop_height = std::max(1, (int) sc.cam.resY / f);
    c.crop_offset_x = cfg->crop_offset_x / f;
    c.crop_offset_y = cfg->crop_offset_x / f;            // sic (util.cpp:128)
    c.sample_count = cfg->sample_count * f;
    c.importance_map = nullptr;
    c.n_chains = 0;
    *nested = c;
}
void orc_resample_luminance(const float *image_rgb, int w, int h, int W, int H, float *map) { resampleLuminance(image_rgb, w, h, W, H, map); }
// the resampling alone, double in / double out: held bit for bit against the reference's Bitmap::resample (tests/test_ref_pins.py)
void orc_resample_map_f64(const double *lum, int w, int h, int W, int H, double *out) {
    const std::vector<Float> r = resampleMap(std::vector<Float>(lum, lum + (size_t) w * h), w, h, W, H);
    for (size_t i = 0; i < r.size(); ++i) out[i] = r[i];
}
// develop of an accumulated film (W*H*3 floats) with an optional importance map (drmlt_proc.cpp:813-854)
void orc_develop(const float *film_rgb, int w, int h, double b, int acceptance_map, const float *importance, float *image_rgb) {
    Film f; f.init(w, h, DR_FILTER_BOX);
    for (size_t i = 0; i < f.data.size(); ++i) f.data[i] = film_rgb[i];
    develop(f, b, acceptance_map != 0, image_rgb, importance);
}

// Film splat of explicit (pos, rgb) pairs -- parity of ImageBlock::put + the filter table.
int orc_splat(int w, int h, int rfilter, const float *pos, const float *rgb, int64_t n, float *film_out) {
    Film f; f.init(w, h, rfilter);
    for (int64_t i = 0; i < n; ++i) f.put(Vec2(pos[2 * i], pos[2 * i + 1]), RGB(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]));
    for (size_t i = 0; i < f.data.size(); ++i) film_out[i] = (float) f.data[i];
    return 0;
}

// The same in double, with the per-splat verdict of put (false = rejected as non-finite / negative): held against the
// reference's own ImageBlock::put by tests/test_ref_pins.py.
int orc_splat_f64(int w, int h, int rfilter, const float *pos, const float *rgb, int64_t n, double *film_out, int *accepted) {
    Film f; f.init(w, h, rfilter);
    for (int64_t i = 0; i < n; ++i)
        accepted[i] = f.put(Vec2(pos[2 * i], pos[2 * i + 1]), RGB(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2])) ? 1 : 0;
    for (size_t i = 0; i < f.data.size(); ++i) film_out[i] = f.data[i];
    return 0;
}

// BSDF leaf access for chi-square / consistency tests (modelled on src/tests/test_chisquare.cpp)
/* roughplastic at leaf level: registers the material's table (DR_ROUGH_TABLE_DOUBLES doubles, as in dr_scene_desc.rough_tables) and
 * returns the value to store in dr_material.table for the orc_bsdf_* calls below */
/* RoughTransmittance::eval(cosTheta) of a reduced table (rtrans.h:136-146) */
double orc_rough_transmittance(const double *table, double cosTheta) {
    dr_material m{};
    m.type = DR_BSDF_ROUGHPLASTIC; m.eta[0] = 1.5f; m.reflectance[0] = m.transmittance[0] = 0.5f;
    m.table = 0;
    detail::prepareRoughPlastic(m, table);
    return detail::RoughPlastic(m).T(cosTheta);
}
uint32_t orc_rough_table_register(const dr_material *m, const double *table) {
    dr_material mm = *m;
    mm.table = 0;
    detail::prepareRoughPlastic(mm, table);
    return mm.table;
}
/* Texture2D::eval of a dr_texture at n intersection uv pairs (no ray differentials): orc_scene.hpp Texture::eval */
void orc_texture_eval(const dr_texture *t, const double *uv, int n, double *rgb) {
    dr_scene_desc d{};
    d.n_textures = 1; d.textures = t;
    Texture tex;
    tex.w = (int) t->width; tex.h = (int) t->height; tex.wrapU = (int) t->wrap_u; tex.wrapV = (int) t->wrap_v; tex.nearest = t->nearest != 0;
    tex.scaleU = t->uv_scale[0]; tex.scaleV = t->uv_scale[1]; tex.offU = t->uv_offset[0]; tex.offV = t->uv_offset[1];
    tex.texels.resize((size_t) tex.w * tex.h);
    for (size_t i = 0; i < tex.texels.size(); ++i) tex.texels[i] = RGB(t->texels[3 * i], t->texels[3 * i + 1], t->texels[3 * i + 2]);
    for (int i = 0; i < n; ++i) {
        const RGB v = tex.eval(Vec2(uv[2 * i], uv[2 * i + 1]));
        rgb[3 * i] = v.r; rgb[3 * i + 1] = v.g; rgb[3 * i + 2] = v.b;
    }
}
void orc_bsdf_sample(const dr_material *m, const double *wi, int mode, double u1, double u2, double *wo, double *weight, double *pdf, int *sampledType) {
    BSDFRecord b(Vec3(wi[0], wi[1], wi[2]), mode);
    Float p = 0;
    dr_material mm = *m; detail::preparePlastic(mm); m = &mm;
    RGB w = bsdfSample(*m, b, p, Vec2(u1, u2), 1e-7);
    wo[0] = b.wo.x; wo[1] = b.wo.y; wo[2] = b.wo.z;
    weight[0] = w.r; weight[1] = w.g; weight[2] = w.b;
    *pdf = p; *sampledType = b.sampledType;
}
// ... with the extra number EUsesSampler BSDFs draw from bRec.sampler (roughdielectric.cpp:555)
void orc_bsdf_sample3(const dr_material *m, const double *wi, int mode, double u1, double u2, double u3, double *wo, double *weight, double *pdf, int *sampledType, double *eta) {
    BSDFRecord b(Vec3(wi[0], wi[1], wi[2]), mode);
    Float p = 0;
    dr_material mm = *m; detail::preparePlastic(mm); m = &mm;
    RGB w = bsdfSample(*m, b, p, Vec2(u1, u2), 1e-7, u3);
    wo[0] = b.wo.x; wo[1] = b.wo.y; wo[2] = b.wo.z;
    weight[0] = w.r; weight[1] = w.g; weight[2] = w.b;
    *pdf = p; *sampledType = b.sampledType; *eta = b.eta;
}
void orc_bsdf_eval(const dr_material *m, const double *wi, const double *wo, int mode, int measure, double *value, double *pdf) {
    BSDFRecord b(Vec3(wi[0], wi[1], wi[2]), Vec3(wo[0], wo[1], wo[2]), mode);
    dr_material mm = *m; detail::preparePlastic(mm); m = &mm;
    RGB v = bsdfEval(*m, b, measure);
    value[0] = v.r; value[1] = v.g; value[2] = v.b;
    *pdf = bsdfPdf(*m, b, measure);
}


// ---- numerical leaves, exported one by one so that tests/test_ref_pins.py can hold them against the reference's own
// code (oracle/_ref/libref_leaf.so, built by oracle/ref/Makefile from the sources under /root/reference) and against
// the fixture that library produced (tests/golden/ref_leaf.npz).
void orc_squareToCosineHemisphere(double u, double v, double *out) { Vec3 d = squareToCosineHemisphere(Vec2(u, v)); out[0] = d.x; out[1] = d.y; out[2] = d.z; }
void orc_squareToUniformDiskConcentric(double u, double v, double *out) { Vec2 p = squareToUniformDiskConcentric(Vec2(u, v)); out[0] = p.x; out[1] = p.y; }
void orc_squareToUniformTriangle(double u, double v, double *out) { Vec2 p = squareToUniformTriangle(Vec2(u, v)); out[0] = p.x; out[1] = p.y; }
double orc_fresnelDielectricExt(double cosThetaI, double eta, double *cosThetaT) { Float ct; Float r = fresnelDielectricExt(cosThetaI, ct, eta); *cosThetaT = ct; return r; }
void orc_fresnelConductorExact(double cosThetaI, const double *eta, const double *k, double *out) {
    RGB r = fresnelConductorExact(cosThetaI, RGB(eta[0], eta[1], eta[2]), RGB(k[0], k[1], k[2]));
    out[0] = r.r; out[1] = r.g; out[2] = r.b;
}
double orc_fresnelDiffuseReflectance(double eta) { return detail::fresnelDiffuseReflectance(eta); }
void orc_coordinateSystem(const double *a, double *b, double *c) {
    Vec3 B, Cv; coordinateSystem(Vec3(a[0], a[1], a[2]), B, Cv);
    b[0] = B.x; b[1] = B.y; b[2] = B.z; c[0] = Cv.x; c[1] = Cv.y; c[2] = Cv.z;
}
double orc_luminance(const double *rgb) { return RGB(rgb[0], rgb[1], rgb[2]).luminance(); }
int orc_triAccel(const double *p0, const double *p1, const double *p2, const double *o, const double *d, double mint, double maxt, double *uvt) {
    TriAccel ta;
    if (triLoad(ta, Vec3(p0[0], p0[1], p0[2]), Vec3(p1[0], p1[1], p1[2]), Vec3(p2[0], p2[1], p2[2])) != 0) return -1;
    Ray ray; ray.o = Vec3(o[0], o[1], o[2]); ray.d = Vec3(d[0], d[1], d[2]); ray.mint = mint; ray.maxt = maxt;
    Float u = 0, v = 0, t = 0;
    bool hit = triIntersect(ta, ray, mint, maxt, u, v, t);
    uvt[0] = u; uvt[1] = v; uvt[2] = t;
    return hit ? 1 : 0;
}
double orc_pmf(const double *weights, int n, const double *xi, int m, int32_t *index, double *reused, double *pmf_out) {
    DiscreteDistribution dist;
    for (int i = 0; i < n; ++i) dist.append(weights[i]);
    Float sum = dist.normalize();
    for (int i = 0; i < n; ++i) pmf_out[i] = dist[i];
    for (int j = 0; j < m; ++j) { Float s = xi[j]; index[j] = (int32_t) dist.sampleReuse(s); reused[j] = s; }
    return sum;
}
void orc_microfacet(int type, double alpha, int sampleVisible, const double *wi, const double *m_in, double u, double v, double *out) {
    dr_material mat{}; mat.alpha = (float) alpha; mat.flags = (type ? DR_MAT_GGX : 0) | (sampleVisible ? DR_MAT_SAMPLE_VISIBLE : 0);
    Microfacet distr(mat);
    distr.alpha = std::max((Float) alpha, (Float) 1e-4f);   // the material record stores alpha as float; the leaf test passes a double
    Vec3 Wi(wi[0], wi[1], wi[2]), M(m_in[0], m_in[1], m_in[2]);
    out[0] = distr.eval(M);
    out[1] = distr.pdf(Wi, M);
    Vec3 Wo = M * (2 * dot(Wi, M)) - Wi;
    out[2] = distr.G(Wi, Wo, M);
    out[3] = distr.smithG1(Wi, M);
    Float pdf;
    Vec3 s = distr.sample(Wi, Vec2(u, v), pdf, 1e-7);
    out[4] = s.x; out[5] = s.y; out[6] = s.z; out[7] = pdf;
}
double orc_kelemen_logpdf(double s1, double s2, double du) { return KelemenKernel(s1, s2).logPdf(du); }


// DRMLTSampler (drmlt_sampler.cpp:189-414) on an explicit current state and an explicit stream of uniforms consumed in call
// order: stage-1 proposal, stage-2 proposal (after setLargeStep(false), as timidAfterLarge does), Green's reverse state
// y* = z - (y - x), Mira's transition ratio.  Held against the reference's own sampler by tests/test_ref_pins.py.
int orc_drmlt_sampler(int type, int maxDim, double sigma, double scaleSecond, int largeStep, const double *uCurrent,
                      const double *stream, double *prop1, double *prop2, double *reverse, double *ratio) {
    KeyedSource src; src.seq = stream; src.seqPos = 0;
    DRMLTSampler s; s.type = type; s.samplerId = 0; s.src = &src; s.maxDim = (size_t) maxDim; s.sigma = sigma; s.scaleSecond = scaleSecond;
    s.uCurrent.assign(uCurrent, uCurrent + maxDim);
    s.setLargeStep(largeStep != 0);
    for (int k = 0; k < maxDim; ++k) prop1[k] = s.primarySample((size_t) k);
    s.nextStage(); s.setLargeStep(false);
    for (int k = 0; k < maxDim; ++k) prop2[k] = s.primarySample((size_t) k);
    if (type == DR_TYPE_GREEN) { s.setReverse(true); for (int k = 0; k < maxDim; ++k) reverse[k] = s.primarySample((size_t) k); }
    *ratio = type == DR_TYPE_MIRA ? s.getTransitionRatio() : 1.0;
    return (int) src.seqPos;
}

// The same sampler over a SEQUENCE of mutations, driven as DRMLTRenderer::process drives it (drmlt_proc.cpp:541-760): outcome 0 =
// accept the first stage, 1 = second stage accepted, 2 = second stage rejected; mode 1 = handleLightTracing() (emitter sampler
// under fixEmitterPath), mode 2 = setStagesToIdentity() (MMLT direct sampler).  NaN marks what a mutation did not produce.
int orc_drmlt_sampler_seq(int type, int mode, int maxDim, double sigma, double scaleSecond, int nMut, const int *large, const int *outcome,
                          const int *lightTracing, const double *uCurrent, const double *stream,
                          double *prop1, double *prop2, double *reverse, double *ratio) {
    KeyedSource src; src.seq = stream; src.seqPos = 0;
    DRMLTSampler s; s.type = type; s.samplerId = 0; s.src = &src; s.maxDim = (size_t) maxDim; s.sigma = sigma; s.scaleSecond = scaleSecond;
    s.stage2Identity = mode == 1; s.identityAll = mode == 2;
    s.uCurrent.assign(uCurrent, uCurrent + maxDim);
    const double nan = std::numeric_limits<double>::quiet_NaN();
    for (int m = 0; m < nMut; ++m) {
        double *p1 = prop1 + (size_t) m * maxDim, *p2 = prop2 + (size_t) m * maxDim, *rv = reverse + (size_t) m * maxDim;
        for (int k = 0; k < maxDim; ++k) p1[k] = p2[k] = rv[k] = nan;
        ratio[m] = nan;
        s.setLargeStep(large[m] != 0);
        for (int k = 0; k < maxDim; ++k) p1[k] = s.primarySample((size_t) k);
        if (outcome[m] == 0) { s.accept(true); continue; }
        s.nextStage(lightTracing[m] != 0); s.setLargeStep(false);
        for (int k = 0; k < maxDim; ++k) p2[k] = s.primarySample((size_t) k);
        if (type == DR_TYPE_GREEN) {
            s.setReverse(true);
            for (int k = 0; k < maxDim; ++k) rv[k] = s.primarySample((size_t) k);
            s.setReverse(false);
        }
        ratio[m] = type == DR_TYPE_MIRA ? s.getTransitionRatio() : 1.0;
        if (outcome[m] == 1) s.accept(false); else s.reject();
    }
    return (int) src.seqPos;
}

// PSSMLTSampler (pssmlt_sampler.cpp:93-166, pssmlt_sampler.h:117-147) on an explicit current state and an explicit stream of
// uniforms consumed in call order, over a SEQUENCE of mutations: per mutation setLargeStep -> primarySample(0..maxDim-1) ->
// accept / reject (eager fill, Kelemen / Gaussian mutation, backup / restore).  Held against the reference's own sampler
// by tests/test_ref_pins.py.
int orc_pssmlt_sampler(int kelemen, int maxDim, double s1, double s2, double sigma, int nMut, const int *large, const int *accepted,
                       const double *uCurrent, const double *stream, double *proposals) {
    KeyedSource src; src.seq = stream; src.seqPos = 0;
    PSSMLTSampler s; s.samplerId = 0; s.src = &src; s.maxDim = (size_t) maxDim; s.useKelemen = kelemen != 0;
    s.configure(s1, s2, sigma);
    s.u.assign(uCurrent, uCurrent + maxDim);
    for (int m = 0; m < nMut; ++m) {
        s.setLargeStep(large[m] != 0);
        for (int k = 0; k < maxDim; ++k) proposals[(size_t) m * maxDim + k] = s.primarySample((size_t) k);
        if (accepted[m]) s.accept(); else s.reject();
    }
    return (int) src.seqPos;
}

// transition kernels (transition.h) for known-answer tests
double orc_kelemen_sample(double s1, double s2, double xi) { return KelemenKernel(s1, s2).sample(xi); }
double orc_kelemen_pdf(double s1, double s2, double du) { return KelemenKernel(s1, s2).pdf(du); }
double orc_gaussian_sample(double sigma, double xi1, double xi2) { return GaussianKernel{ sigma }.sample(xi1, xi2); }
double orc_cauchy_sample(double rho, double xi) { return WrappedCauchyKernel(rho).sample(xi); }
double orc_wrap(double y) { return wrapReflect(y); }

// The separate direct-illumination image (BidirectionalUtils::renderDirectComponent, src/libbidir/util.cpp:30-94):
// pixelSamples x shadingSamples split of directSamples, SamplingIntegrator::renderBlock (src/librender/integrator.cpp)
// with the film's reconstruction filter and weight normalisation.  The reference draws its samples from an `ldsampler`
// seeded from /dev/urandom; here sample j of pixel p uses the keyed uniforms (S_DIRECT, p, j, .): same estimator.
// li_out (optional): the un-filtered radiance of every pixel sample [h][w][pixelSamples][3].
int orc_direct_image(void *h, const dr_config *cfg, float *image_rgb, double *li_out) {
    Scene &sc = ((OrcScene *) h)->sc;
    applyEps(sc, cfg);
    int pixelSamples = std::max(cfg->direct_samples, 1), shadingSamples = 1;
    while (pixelSamples > 8) { pixelSamples /= 2; shadingSamples *= 2; }
    const int W = (int) sc.cam.resX, H = (int) sc.cam.resY;
    Film f; f.init(W, H, *cfg);
    std::vector<Float> weight((size_t) W * H, 0.0);
    PathCtx ctx; ctx.scene = &sc;
    std::vector<Vec2> u(2 * shadingSamples);
    std::vector<Float> extra(shadingSamples);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const uint64_t p = (uint64_t) y * W + x;
            for (int j = 0; j < pixelSamples; ++j) {
                const Vec2 jitter(keyedUniform(cfg->seed, S_DIRECT, p, (uint32_t) j, 0), keyedUniform(cfg->seed, S_DIRECT, p, (uint32_t) j, 1));
                for (int i = 0; i < 2 * shadingSamples; ++i)
                    u[i] = Vec2(keyedUniform(cfg->seed, S_DIRECT, p, (uint32_t) j, 2 + 2 * i), keyedUniform(cfg->seed, S_DIRECT, p, (uint32_t) j, 3 + 2 * i));
                const Vec2 samplePos(x + jitter.x, y + jitter.y);
                Vec3 dl = sc.cam.sampleToDir(samplePos.x / sc.cam.resX, samplePos.y / sc.cam.resY);
                Float invZ = 1.0 / dl.z;
                Ray ray; ray.o = sc.cam.pos; ray.d = sc.cam.xformDir(dl);
                ray.mint = sc.cam.nearClip * invZ; ray.maxt = sc.cam.farClip * invZ;
                for (int i = 0; i < shadingSamples; ++i) extra[i] = keyedUniform(cfg->seed, S_DIRECT, p, (uint32_t) j, 2 + 4 * shadingSamples + i);
                const RGB Li = directLi(ctx, ray, shadingSamples, u.data(), extra.data());
                if (li_out) { double *o = li_out + (((size_t) p) * pixelSamples + j) * 3; o[0] = Li.r; o[1] = Li.g; o[2] = Li.b; }
                // ImageBlock::put(pos, spec, alpha): weighted value + weight channel (imageblock.h:149-196)
                if (!Li.isValid()) continue;
                const Float px = samplePos.x - 0.5, py = samplePos.y - 0.5;
                const int minx = std::max((int) std::ceil(px - f.radius), 0), miny = std::max((int) std::ceil(py - f.radius), 0);
                const int maxx = std::min((int) std::floor(px + f.radius), W - 1), maxy = std::min((int) std::floor(py + f.radius), H - 1);
                for (int yy = miny; yy <= maxy; ++yy) {
                    const Float wy = f.evalDiscretized(yy - py);
                    for (int xx = minx; xx <= maxx; ++xx) {
                        const Float wgt = f.evalDiscretized(xx - px) * wy;
                        Float *dest = &f.data[((size_t) yy * W + xx) * 3];
                        dest[0] += wgt * Li.r; dest[1] += wgt * Li.g; dest[2] += wgt * Li.b;
                        weight[(size_t) yy * W + xx] += wgt;
                    }
                }
            }
        }
    for (size_t i = 0; i < (size_t) W * H; ++i) {             // HDRFilm::develop: divide by the accumulated filter weight
        const Float inv = weight[i] > 0 ? 1.0 / weight[i] : 0.0;
        for (int c = 0; c < 3; ++c) image_rgb[3 * i + c] = (float) (f.data[3 * i + c] * inv);
    }
    return 0;
}

} // extern "C"
