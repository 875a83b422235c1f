/* oracle/_ref path driver -- TEST INFRASTRUCTURE ONLY (never linked into or called by the product).
 *
 * C entry points around the reference's OWN path code, compiled from where it lies under /root/reference by the
 * Makefile next to this file into oracle/_ref/libref_path.so:
 *   libbidir  pathsampler.cpp (PathSampler::sampleSplats: MMLT / BDPT / PT), path.cpp (randomWalk, miWeight),
 *             vertex.cpp, edge.cpp, common.cpp, rsampler.cpp, ...
 *   librender scene.cpp, skdtree.cpp (SAH kd-tree), trimesh.cpp, shape.cpp, emitter.cpp, sensor.cpp, bsdf.cpp, ...
 *   libcore   everything the above needs
 *   plugins   bsdfs/{diffuse,dielectric,conductor,roughconductor,roughdielectric,plastic,twosided}.cpp,
 *             emitters/area.cpp, sensors/perspective.cpp, rfilters/{gaussian,box}.cpp, samplers/independent.cpp,
 *             integrators/path/path.cpp   (each with its CreateInstance renamed CreateInstance_<plugin>)
 * What is NOT the reference's: the arithmetic-free Boost/Eigen stand-ins in stubs/, the TLS stand-in in
 * ref_runtime.cpp, the plugin table and film below, and aborting stubs for symbols of files that cannot be compiled
 * here and are never reached (Bitmap, SpecularManifold, hardware renderer).  The XML loader is not used: the scene
 * is assembled from the same flattened description (include/drmlt_b200.h: dr_scene_desc) the product and the oracle
 * take, through the reference's own object interfaces (Properties, addChild, configure, Scene::initialize).
 *
 * Uniforms are REPLAYED: the three samplers PathSampler draws from are Sampler subclasses that hand out the
 * caller's primary-sample vectors in order (what PSSMLTSampler / DRMLTSampler::next1D/next2D do with
 * primarySample(i), drmlt_sampler.cpp:416-425).
 */
#include <mitsuba/render/scene.h>
#include <mitsuba/render/trimesh.h>
#include <mitsuba/render/film.h>
#include <mitsuba/render/sampler.h>
#include <mitsuba/render/bsdf.h>
#include <mitsuba/render/emitter.h>
#include <mitsuba/render/sensor.h>
#include <mitsuba/core/plugin.h>
#include <mitsuba/core/statistics.h>
#include <mitsuba/core/fstream.h>
#include <mitsuba/core/sched.h>
#include <mitsuba/core/appender.h>
#include <mitsuba/bidir/pathsampler.h>
#include <mitsuba/bidir/util.h>
#include <mitsuba/render/renderjob.h>
#include <mitsuba/render/renderqueue.h>
#include <mitsuba/core/bitmap.h>
#include <mitsuba/core/timer.h>
#include <mitsuba/render/texture.h>
#include <mitsuba/render/mipmap.h>
#include <mitsuba/core/half.h>
#include <map>
#include <tuple>
#include <execinfo.h>
#include <dlfcn.h>
#include <unistd.h>
#include <signal.h>
#include "../../include/drmlt_b200.h"
#include "src/bsdfs/rtrans.h"                   // RoughTransmittance, as roughplastic.cpp includes it
#include "src/integrators/pssmlt_utils.h"     // findMaxDimensions, as the integrators include it (drmlt.cpp, pssmlt.cpp)

using namespace mitsuba;
#define TR(msg) do { if (getenv("REF_TRACE")) fprintf(stderr, "[ref] %s\n", msg); } while (0)

#define REF_PLUGINS(X) X(diffuse) X(dielectric) X(conductor) X(roughconductor) X(roughdielectric) X(plastic) X(roughplastic) X(twosided) \
    X(area) X(rectangle) X(sphere) X(perspective) X(gaussian) X(box) X(tent) X(mitchell) X(catmullrom) X(lanczos) X(independent) X(ldsampler) X(path) X(direct) X(drmlt) X(pssmlt)
#define X(name) extern "C" void *CreateInstance_##name(const Properties &props);
REF_PLUGINS(X)
#undef X

/* ---- plugin table instead of dlopen("plugins/<name>.so") (libcore/plugin.cpp:180-196) */
MTS_NAMESPACE_BEGIN
ref<PluginManager> PluginManager::m_instance = NULL;
PluginManager::PluginManager() {}
PluginManager::~PluginManager() {}
void PluginManager::staticInitialization() { m_instance = new PluginManager(); }
void PluginManager::staticShutdown() { m_instance = NULL; }
void PluginManager::ensurePluginLoaded(const std::string &) {}
std::vector<std::string> PluginManager::getLoadedPlugins() const { return std::vector<std::string>(); }
static ConfigurableObject *ref_make_film(const Properties &props);
/* REF_PLUGIN_DIR=<dir>: a plugin file <dir>/<name>.so takes precedence over the table and is loaded exactly as the reference
 * loads plugins/<name>.so -- dlopen(RTLD_LAZY | RTLD_LOCAL) + dlsym("CreateInstance") (libcore/plugin.cpp:62-96, 222-248).
 * This is how the drop-in plugins built from drmlt-mitsuba_b200/shim/mts_plugin.cpp are exercised by the reference's own
 * RenderJob -> Scene::render -> Integrator::render. */
typedef void *(*CreateInstanceFn)(const Properties &);
static CreateInstanceFn pluginFromDir(const std::string &name) {
    static std::map<std::string, CreateInstanceFn> cache;
    const char *dir = getenv("REF_PLUGIN_DIR");
    if (!dir) return NULL;
    const std::string path = std::string(dir) + "/" + name + ".so";
    std::map<std::string, CreateInstanceFn>::iterator it = cache.find(path);
    if (it != cache.end()) return it->second;
    CreateInstanceFn fn = NULL;
    if (access(path.c_str(), R_OK) == 0) {
        void *handle = dlopen(path.c_str(), RTLD_LAZY | RTLD_LOCAL);
        if (!handle) SLog(EError, "Error while loading plugin \"%s\": %s", path.c_str(), dlerror());
        fn = (CreateInstanceFn) dlsym(handle, "CreateInstance");
        if (!fn || !dlsym(handle, "GetDescription")) SLog(EError, "Could not resolve symbol \"CreateInstance\" / \"GetDescription\" in \"%s\"", path.c_str());
        Class::staticInitialization();     /* "New classes must be registered within the class hierarchy" (plugin.cpp:100-101) */
    }
    cache[path] = fn;
    return fn;
}
ConfigurableObject *PluginManager::createObject(const Properties &props) {
    const std::string name = props.getPluginName();
    if (CreateInstanceFn fn = pluginFromDir(name)) return (ConfigurableObject *) fn(props);
    if (name == "hdrfilm") return ref_make_film(props);      /* the nested film of mltLuminancePass (util.cpp:120-130); hdrfilm itself needs OpenEXR */
#define X(n) if (name == #n) return (ConfigurableObject *) CreateInstance_##n(props);
    REF_PLUGINS(X)
#undef X
    SLog(EError, "oracle/_ref: plugin \"%s\" is not linked in", name.c_str());
    return NULL;
}
ConfigurableObject *PluginManager::createObject(const Class *classType, const Properties &props) {
    ConfigurableObject *o = createObject(props);
    if (!o->getClass()->derivesFrom(classType))
        SLog(EError, "oracle/_ref: plugin \"%s\" has the wrong class", props.getPluginName().c_str());
    return o;
}
MTS_IMPLEMENT_CLASS(PluginManager, false, Object)

/* ---- a film that only carries size, crop window and reconstruction filter (the hdrfilm plugin needs OpenEXR) */
class PinFilm : public Film {
public:
    PinFilm(const Properties &props) : Film(props) {}
    /* SamplingIntegrator::render accumulates weighted blocks into the film (hdrfilm.cpp:352, 388-393): the same storage here */
    void configure() { m_storage = new ImageBlock(Bitmap::ESpectrumAlphaWeight, m_cropSize); m_storage->clear(); }
    void clear() { if (m_storage) m_storage->clear(); }
    void put(const ImageBlock *block) { m_storage->put(block); }
    /* what DRMLTProcess::develop / PSSMLTProcess::develop hand over (drmlt_proc.cpp:850-853): spectrum float pixels */
    void setBitmap(const Bitmap *bitmap, Float multiplier) {
        const Vector2i sz = bitmap->getSize();
        image.resize((size_t) sz.x * sz.y * 3);
        const Float *src = bitmap->getFloatData();
        for (size_t i = 0; i < image.size(); ++i) image[i] = (float) (src[i] * multiplier);
    }
    std::vector<float> image;
    void addBitmap(const Bitmap *, Float) {}
    void setDestinationFile(const fs::path &, uint32_t) {}
    void develop(const Scene *, Float) {}
    /* hdrfilm.cpp:425-470 for the two targets this path asks for: the weighted storage divided by its weight -> ESpectrum
     * (renderDirectComponent, util.cpp:87-91) or ELuminance (mltLuminancePass, util.cpp:184-188) float pixels.  A film that was
     * handed a developed bitmap (setBitmap: the nested MLT job) develops that instead. */
    bool develop(const Point2i &so, const Vector2i &size, const Point2i &to, Bitmap *target) const {
        if (target->getComponentFormat() != Bitmap::EFloat) return false;
        const bool lum = target->getPixelFormat() == Bitmap::ELuminance;
        if (!lum && target->getPixelFormat() != Bitmap::ESpectrum) return false;
        Float *dst = target->getFloatData();
        const int tw = target->getWidth(), tc = target->getChannelCount();
        const Bitmap *src = m_storage->getBitmap();
        const int border = m_storage->getBorderSize(), sw = src->getWidth(), sc = src->getChannelCount();
        const Float *sdata = src->getFloatData();
        for (int y = 0; y < size.y; ++y)
            for (int x = 0; x < size.x; ++x) {
                Spectrum value;
                if (!image.empty()) {
                    const float *p = &image[((size_t) (y + so.y) * m_cropSize.x + (x + so.x)) * 3];
                    for (int q = 0; q < SPECTRUM_SAMPLES; ++q) value[q] = p[q];
                } else {
                    const Float *p = sdata + ((size_t) (y + so.y + border) * sw + (x + so.x + border)) * sc;
                    const Float weight = p[sc - 1], inv = weight != 0 ? (Float) 1 / weight : (Float) 0;
                    for (int q = 0; q < SPECTRUM_SAMPLES; ++q) value[q] = p[q] * inv;
                }
                Float *o = dst + ((size_t) (y + to.y) * tw + (x + to.x)) * tc;
                if (lum) o[0] = value.getLuminance();
                else for (int q = 0; q < SPECTRUM_SAMPLES; ++q) o[q] = value[q];
            }
        if (getenv("REF_TRACE") && lum) {
            Float lo = 1e30, hi = -1e30; int zeros = 0;
            for (int i = 0; i < size.x * size.y; ++i) { lo = std::min(lo, dst[i]); hi = std::max(hi, dst[i]); zeros += dst[i] == 0; }
            fprintf(stderr, "[ref] PinFilm::develop luminance %dx%d: min %g max %g zeros %d (from %s)\n", size.x, size.y, lo, hi, zeros, image.empty() ? "storage" : "bitmap");
        }
        return true;
    }
    bool destinationExists(const fs::path &) const { return false; }
    bool hasAlpha() const { return false; }
    MTS_DECLARE_CLASS()
private:
    ref<ImageBlock> m_storage;
};
MTS_IMPLEMENT_CLASS(PinFilm, false, Film)

static ConfigurableObject *ref_make_film(const Properties &props) {
    Properties fp(props);
    ref<ConfigurableObject> rf = PluginManager::getInstance()->createObject(MTS_CLASS(ReconstructionFilter), Properties("gaussian"));   // hdrfilm's default
    rf->configure();
    PinFilm *film = new PinFilm(fp);
    film->addChild(rf);
    rf->setParent(film);
    return film;
}

/* ---- replayed uniforms */
class ReplaySampler : public Sampler {
public:
    ReplaySampler() : Sampler(Properties()), m_u(NULL), m_n(0), m_i(0) {}
    void set(const float *u, int n) { m_u = u; m_n = n; m_i = 0; }
    ref<Sampler> clone() { return new ReplaySampler(); }
    void generate(const Point2i &) { m_i = 0; }
    void advance() { m_i = 0; }
    void setSampleIndex(size_t) { m_i = 0; }
    Float next1D() { if (m_i >= m_n) { overflow = true; return 0.5; } return (Float) m_u[m_i++]; }
    Point2 next2D() { Float a = next1D(); Float b = next1D(); return Point2(a, b); }
    void request1DArray(size_t) { SLog(EError, "ReplaySampler: sample arrays are not used on this path"); }
    void request2DArray(size_t) { SLog(EError, "ReplaySampler: sample arrays are not used on this path"); }
    std::string toString() const { return "ReplaySampler[]"; }
    bool overflow = false;
    int consumed() const { return m_i; }
    MTS_DECLARE_CLASS()
private:
    const float *m_u; int m_n, m_i;
};
MTS_IMPLEMENT_CLASS(ReplaySampler, false, Sampler)
MTS_NAMESPACE_END

namespace {

struct RefScene {
    ref<Scene> scene;
    ref<PinFilm> film;
    std::vector<ref<BSDF> > bsdfs;      // one per dr_material
    int filmW = 0, filmH = 0;
};

void onSegv(int) { void *bt[64]; int n = backtrace(bt, 64); backtrace_symbols_fd(bt, n, 2); _exit(139); }
bool g_init = false;
void initOnce() {
    if (g_init) return;
    g_init = true;
    if (getenv("REF_TRACE")) signal(SIGSEGV, onSegv);
    /* the start-up sequence of src/mitsuba/mitsuba.cpp (main): class table, threads, logger, spectra, scheduler */
    Class::staticInitialization();
    Object::staticInitialization();
    PluginManager::staticInitialization();
    Statistics::staticInitialization();
    Thread::staticInitialization();
    Logger::staticInitialization();
    FileStream::staticInitialization();
    Spectrum::staticInitialization();
    Scheduler::staticInitialization();
    Thread::getThread()->getLogger()->setLogLevel(getenv("REF_LOG") ? EDebug : EWarn);
    /* data/microfacet/*.dat (roughplastic, rtrans.h:50-51) is resolved through the thread's FileResolver, as mitsuba.cpp sets it up */
    Thread::getThread()->getFileResolver()->appendPath(getenv("REF_DATA_ROOT") ? getenv("REF_DATA_ROOT") : "/root/reference");
}

Spectrum rgbSpectrum(const float *v) { Spectrum s; s.fromLinearRGB(v[0], v[1], v[2]); return s; }

/* A bitmap texture from the texels of a dr_texture.  The reference's own BitmapTexture (src/textures/bitmap.cpp) can only be
 * made from an image FILE (and its format conversion needs Boost.MPL); everything below its file handling is the reference's:
 * this class holds the same TMIPMap<Color3, Color3h> (include/mitsuba/render/mipmap.h: half-precision storage, evalTexel's
 * boundary conditions, evalBilinear / evalBox), derives from the reference's Texture2D (texture.cpp:81-121: uv scale / offset,
 * dispatch on its.hasUVPartials) and restates only the ten lines of BitmapTexture::eval(uv) (bitmap.cpp:432-455). */
class PinTexture : public Texture2D {
public:
    /* bitmap.cpp:175-177 stores TSpectrum<half, 3>; this compiler rejects the explicit TSpectrum<half> -> Color3 conversion inside
     * evalTexel, so the half triple gets the two conversions spelled out (same half class, include/mitsuba/core/half.h) */
    struct Color3h {
        typedef half Scalar;
        half s[3];
        Color3h() {}
        Color3h(const Color3 &c) { for (int i = 0; i < 3; ++i) s[i] = half((float) c[i]); }
        operator Color3() const { return Color3((Float) (float) s[0], (Float) (float) s[1], (Float) (float) s[2]); }
    };
    typedef TMIPMap<Color3, Color3h> MIPMap3;
    using Texture2D::eval;                           // eval(its, filter): texture.cpp:112-121
    static Properties props(const dr_texture &t) {
        Properties p("pintexture");
        p.setFloat("uscale", t.uv_scale[0]); p.setFloat("vscale", t.uv_scale[1]);
        p.setFloat("uoffset", t.uv_offset[0]); p.setFloat("voffset", t.uv_offset[1]);
        return p;
    }
    static ReconstructionFilter::EBoundaryCondition bc(uint32_t w) {
        switch (w) {
            case DR_WRAP_CLAMP: return ReconstructionFilter::EClamp;
            case DR_WRAP_MIRROR: return ReconstructionFilter::EMirror;
            case DR_WRAP_ZERO: return ReconstructionFilter::EZero;
            case DR_WRAP_ONE: return ReconstructionFilter::EOne;
            default: return ReconstructionFilter::ERepeat;
        }
    }
    PinTexture(const dr_texture &t) : Texture2D(props(t)) {
        ref<Bitmap> bitmap = new Bitmap(Bitmap::ERGB, Bitmap::EFloat, Vector2i((int) t.width, (int) t.height));
        bitmap->setGamma(1.0f);
        Float *dst = (Float *) bitmap->getData();
        for (size_t i = 0; i < 3 * (size_t) t.width * t.height; ++i) dst[i] = (Float) t.texels[i];
        m_mipmap = new MIPMap3(bitmap, Bitmap::ERGB, Bitmap::EFloat, NULL, bc(t.wrap_u), bc(t.wrap_v), t.nearest ? ENearest : EBilinear, 20.0f);
    }
    /* serializable like BitmapTexture (the drop-in plugin enumerates a BSDF's textures by serializing it); never unserialized */
    PinTexture(Stream *stream, InstanceManager *manager) : Texture2D(stream, manager) { SLog(EError, "PinTexture cannot be unserialized"); }
    Spectrum eval(const Point2 &uv) const {          // bitmap.cpp:432-455 (the RGB branch)
        Spectrum result;
        Color3 value;
        if (m_mipmap->getFilterType() != ENearest) value = m_mipmap->evalBilinear(0, uv);
        else value = m_mipmap->evalBox(0, uv);
        result.fromLinearRGB(value[0], value[1], value[2]);
        return result;
    }
    Spectrum eval(const Point2 &uv, const Vector2 &, const Vector2 &) const { return eval(uv); }
    Spectrum getAverage() const { Spectrum r; Color3 a = m_mipmap->getAverage(); r.fromLinearRGB(a[0], a[1], a[2]); return r; }   // bitmap.cpp getAverage
    Spectrum getMaximum() const { Spectrum r; Color3 a = m_mipmap->getMaximum(); r.fromLinearRGB(a[0], a[1], a[2]); return r; }
    Spectrum getMinimum() const { Spectrum r; Color3 a = m_mipmap->getMinimum(); r.fromLinearRGB(a[0], a[1], a[2]); return r; }
    bool isConstant() const { return false; }
    bool isMonochromatic() const { return false; }
    bool usesRayDifferentials() const { return true; }   // bitmap.cpp:542-544
    Vector3i getResolution() const { return Vector3i(m_mipmap->getWidth(), m_mipmap->getHeight(), 1); }
    ref<Bitmap> getBitmap(const Vector2i &) const { return m_mipmap->toBitmap(); }      // bitmap.cpp:483-485
    std::string toString() const {                                                      // bitmap.cpp:566-578
        std::ostringstream oss;
        oss << "BitmapTexture[" << endl << "  filename = \"\"," << endl << "  mipmap = " << indent(m_mipmap.toString()) << endl << "]";
        return oss.str();
    }
    MTS_DECLARE_CLASS()
private:
    ref<MIPMap3> m_mipmap;
};
MTS_IMPLEMENT_CLASS_S(PinTexture, false, Texture2D)

ref<BSDF> makeBSDF(const dr_material &m, const std::vector<ref<Texture> > *textures = NULL) {
    const uint32_t texR = (m.flags >> 8) & 0xfffu, texT = m.flags >> 20;       // DR_MAT_TEX_*: 1 + texture index
    PluginManager *pm = PluginManager::getInstance();
    const char *names[] = { "diffuse", "dielectric", "conductor", "roughconductor", "roughdielectric", "plastic", "roughplastic" };
    Properties p(names[m.type]);
    switch (m.type) {
        case DR_BSDF_DIFFUSE: p.setSpectrum("reflectance", rgbSpectrum(m.reflectance)); break;
        case DR_BSDF_DIELECTRIC:
        case DR_BSDF_ROUGHDIELECTRIC:
            p.setFloat("intIOR", (Float) m.eta[0]); p.setFloat("extIOR", 1.0);
            p.setSpectrum("specularReflectance", rgbSpectrum(m.reflectance));
            p.setSpectrum("specularTransmittance", rgbSpectrum(m.transmittance));
            break;
        case DR_BSDF_CONDUCTOR:
        case DR_BSDF_ROUGHCONDUCTOR:
            p.setString("material", "none");      // eta / k are given explicitly (no data/ior lookup)
            p.setSpectrum("eta", rgbSpectrum(m.eta)); p.setSpectrum("k", rgbSpectrum(m.k)); p.setFloat("extEta", 1.0);
            p.setSpectrum("specularReflectance", rgbSpectrum(m.reflectance));
            break;
        case DR_BSDF_PLASTIC:
        case DR_BSDF_ROUGHPLASTIC:
            p.setFloat("intIOR", (Float) m.eta[0]); p.setFloat("extIOR", 1.0);
            p.setSpectrum("diffuseReflectance", rgbSpectrum(m.reflectance));
            p.setSpectrum("specularReflectance", rgbSpectrum(m.transmittance));
            p.setBoolean("nonlinear", (m.flags & DR_MAT_NONLINEAR) != 0);
            break;
    }
    if (m.type == DR_BSDF_ROUGHCONDUCTOR || m.type == DR_BSDF_ROUGHDIELECTRIC || m.type == DR_BSDF_ROUGHPLASTIC) {
        p.setString("distribution", (m.flags & DR_MAT_GGX) ? "ggx" : "beckmann");
        p.setFloat("alpha", (Float) m.alpha);
        p.setBoolean("sampleVisible", (m.flags & DR_MAT_SAMPLE_VISIBLE) != 0);
    }
    ref<BSDF> bsdf = static_cast<BSDF *>(pm->createObject(MTS_CLASS(BSDF), p));
    if (textures && (texR || texT)) {               // <texture name="..."> children replace the constant spectra (e.g. diffuse.cpp addChild)
        const bool plastic = m.type == DR_BSDF_PLASTIC || m.type == DR_BSDF_ROUGHPLASTIC;
        const char *nameR = m.type == DR_BSDF_DIFFUSE ? "reflectance" : plastic ? "diffuseReflectance" : "specularReflectance";
        const char *nameT = plastic ? "specularReflectance" : "specularTransmittance";
        if (texR) bsdf->addChild(nameR, const_cast<Texture *>((*textures)[texR - 1].get()));
        if (texT) bsdf->addChild(nameT, const_cast<Texture *>((*textures)[texT - 1].get()));
    }
    bsdf->configure();
    if (m.flags & DR_MAT_TWOSIDED) {
        ref<BSDF> two = static_cast<BSDF *>(pm->createObject(MTS_CLASS(BSDF), Properties("twosided")));
        two->addChild(bsdf);
        two->configure();
        return two;
    }
    return bsdf;
}

}  // namespace

extern "C" {

void ref_init() { initOnce(); }

/* The reference's own RoughTransmittance (src/bsdfs/rtrans.h), reduced as RoughPlastic::configure reduces it (roughplastic.cpp:283-301):
 * the DR_ROUGH_TABLE_DOUBLES doubles of include/drmlt_b200.h.  `probe`/`probe_out` (optional): eval(cosTheta) of the reduced
 * external table at n_probe angles, for the pins of the 1-D interpolation. */
int ref_rough_table(int ggx, double eta, double alpha, double *table, const double *probe, int n_probe, double *probe_out) {
    initOnce();
    try {
        ref<RoughTransmittance> ext = new RoughTransmittance(ggx ? MicrofacetDistribution::EGGX : MicrofacetDistribution::EBeckmann);
        ext->checkEta((Float) eta); ext->checkAlpha((Float) alpha);
        ref<RoughTransmittance> in = ext->clone();
        ext->setEta((Float) eta);
        in->setEta(1 / (Float) eta);
        for (int i = 0; i < DR_ROUGH_TABLE_DOUBLES; ++i) table[i] = 0.0;
        table[100] = in->evalDiffuse((Float) alpha);
        ext->setAlpha((Float) alpha);
        table[101] = ext->evalDiffuse((Float) alpha);
        /* after setEta + setAlpha the object holds exactly the 100 theta samples (rtrans.h:331-347): read the protected array itself */
        struct Peek : public RoughTransmittance {
            static const Float *trans(const RoughTransmittance *r) { return r->*(&Peek::m_trans); }
            static size_t thetaSamples(const RoughTransmittance *r) { return r->*(&Peek::m_thetaSamples); }
        };
        if (Peek::thetaSamples(ext.get()) != DR_ROUGH_TABLE_THETA) throw std::runtime_error("unexpected number of theta samples");
        for (int k = 0; k < DR_ROUGH_TABLE_THETA; ++k) table[k] = Peek::trans(ext.get())[k];
        for (int i = 0; i < n_probe; ++i) probe_out[i] = ext->eval((Float) probe[i], (Float) alpha);
        return 0;
    } catch (const std::exception &e) { fprintf(stderr, "oracle/_ref: %s\n", e.what()); return 1; }
}

static void *scene_create(const dr_scene_desc *d, int rfilter, const Properties *integrator = NULL, int sampleCount = 1, bool analytic = false);
void *ref_scene_create(const dr_scene_desc *d, int rfilter) {
    try { return scene_create(d, rfilter); }
    catch (const std::exception &e) { fprintf(stderr, "oracle/_ref: %s\n", e.what()); return NULL; }
}
/* an analytic shape plugin (rectangle / sphere) with a diffuse BSDF and, optionally, an area emitter */
static void addAnalytic(Scene *scene, const char *plugin, const Transform &toWorld, const Spectrum &reflectance, const Spectrum *radiance,
                        const Point *center = NULL, Float radius = 0) {
    PluginManager *pm = PluginManager::getInstance();
    Properties sp(plugin);
    if (center) { sp.setPoint("center", *center); sp.setFloat("radius", radius); } else sp.setTransform("toWorld", toWorld);
    ref<Shape> shape = static_cast<Shape *>(pm->createObject(MTS_CLASS(Shape), sp));
    Properties bp("diffuse");
    bp.setSpectrum("reflectance", reflectance);
    ref<ConfigurableObject> bsdf = pm->createObject(MTS_CLASS(BSDF), bp);
    bsdf->configure();
    shape->addChild(bsdf);
    bsdf->setParent(shape);
    if (radiance) {
        Properties ep("area");
        ep.setSpectrum("radiance", *radiance);
        ref<ConfigurableObject> em = pm->createObject(MTS_CLASS(Emitter), ep);
        shape->addChild(em);
        em->setParent(shape);
    }
    shape->configure();
    scene->addChild(shape);
}

static void *scene_create(const dr_scene_desc *d, int rfilter, const Properties *integrator, int sampleCount, bool analytic) {
    initOnce();
    TR("init done");
    PluginManager *pm = PluginManager::getInstance();
    RefScene *rs = new RefScene();
    rs->scene = new Scene(Properties("scene"));

    /* sensor: perspective, fov along x, film + reconstruction filter (perspective.cpp:125-187, film.cpp:30-48) */
    const dr_camera &c = d->camera;
    Matrix4x4 mtx;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) mtx(i, j) = (Float) c.to_world[4 * i + j];
    Properties sp("perspective");
    sp.setTransform("toWorld", Transform(mtx));
    sp.setFloat("fov", (Float) c.xfov_deg);
    sp.setString("fovAxis", "x");
    sp.setFloat("nearClip", (Float) c.near_clip);
    sp.setFloat("farClip", (Float) c.far_clip);
    TR("scene object");
    ref<Sensor> sensor = static_cast<Sensor *>(pm->createObject(MTS_CLASS(Sensor), sp));
    TR("sensor created");
    Properties fp("pinfilm");
    fp.setInteger("width", c.film_width); fp.setInteger("height", c.film_height);
    ref<PinFilm> film = new PinFilm(fp);
    rs->film = film;
    ref<ConfigurableObject> rf = pm->createObject(MTS_CLASS(ReconstructionFilter), Properties(rfilter == DR_FILTER_BOX ? "box" : "gaussian"));
    rf->configure();
    film->addChild(rf);
    rf->setParent(film);
    film->configure();
    sensor->addChild(film);
    film->setParent(sensor);
    Properties smpProps("independent");
    smpProps.setInteger("sampleCount", sampleCount);     // = mutations per pixel (drmlt.cpp:400)
    ref<ConfigurableObject> smp = pm->createObject(MTS_CLASS(Sampler), smpProps);
    smp->configure();
    sensor->addChild(smp);
    smp->setParent(sensor);
    TR("sensor children");
    sensor->configure();
    TR("sensor configured");
    rs->scene->addChild(sensor);
    rs->filmW = c.film_width; rs->filmH = c.film_height;

    std::vector<ref<Texture> > textures;
    for (uint32_t i = 0; i < d->n_textures; ++i) { textures.push_back(new PinTexture(d->textures[i])); textures.back()->configure(); }
    for (uint32_t i = 0; i < d->n_materials; ++i) rs->bsdfs.push_back(makeBSDF(d->materials[i], &textures));

    if (analytic) {
        /* the camera of `d`, but a room of ANALYTIC shapes instead of its triangles: floor, back wall and two side walls
         * (rectangles), a sphere, and a rectangular area light under the ceiling */
        const Float h = 1.0;
        const Spectrum white(0.7f), red = rgbSpectrum((const float[3]) { 0.63f, 0.06f, 0.05f }), green = rgbSpectrum((const float[3]) { 0.12f, 0.45f, 0.1f });
        const Spectrum light(15.0f);
        addAnalytic(rs->scene, "rectangle", Transform::translate(Vector(0, -h, 0)) * Transform::rotate(Vector(1, 0, 0), -90), white, NULL);       // floor, normal +y
        addAnalytic(rs->scene, "rectangle", Transform::translate(Vector(0, h, 0)) * Transform::rotate(Vector(1, 0, 0), 90), white, NULL);        // ceiling, normal -y
        addAnalytic(rs->scene, "rectangle", Transform::translate(Vector(0, 0, -h)), white, NULL);                                                // back wall, normal +z
        addAnalytic(rs->scene, "rectangle", Transform::translate(Vector(-h, 0, 0)) * Transform::rotate(Vector(0, 1, 0), 90), red, NULL);          // left, normal +x
        addAnalytic(rs->scene, "rectangle", Transform::translate(Vector(h, 0, 0)) * Transform::rotate(Vector(0, 1, 0), -90), green, NULL);        // right, normal -x
        const Point c0(0.2f, -0.55f, -0.1f);
        addAnalytic(rs->scene, "sphere", Transform(), white, NULL, &c0, 0.45f);
        addAnalytic(rs->scene, "rectangle", Transform::translate(Vector(0, h - 0.01f, 0)) * Transform::rotate(Vector(1, 0, 0), 90) * Transform::scale(Vector(0.25f, 0.25f, 1)), white, &light);
    }
    /* one TriMesh per (material, emitter, smooth) group, in triangle order (trimesh.h:71-77) */
    typedef std::tuple<uint32_t, int32_t, bool, bool> Key;
    std::map<Key, std::vector<uint32_t> > groups;
    for (uint32_t t = 0; !analytic && t < d->n_triangles; ++t) {
        bool smooth = d->tri_flags && (d->tri_flags[t] & DR_TRI_SMOOTH) && d->normals;
        bool hasUV = d->texcoords && !(d->tri_flags && (d->tri_flags[t] & DR_TRI_NO_TEXCOORDS));
        groups[Key(d->tri_material[t], d->tri_emitter[t], smooth, hasUV)].push_back(t);
    }
    TR("bsdfs");
    for (auto &g : groups) {
        const std::vector<uint32_t> &tris = g.second;
        const bool smooth = std::get<2>(g.first), hasUV = std::get<3>(g.first);
        std::map<uint32_t, uint32_t> remap;
        for (uint32_t t : tris) for (int k = 0; k < 3; ++k) { uint32_t v = d->indices[3 * t + k]; if (!remap.count(v)) { uint32_t n = (uint32_t) remap.size(); remap[v] = n; } }
        /* a mesh with texture coordinates always gets UV tangents in this reference (TriMesh::configure, trimesh.cpp:400-402), which
         * is what DR_TRI_UV_TANGENTS says: the caller must have flagged every triangle */
        if (hasUV) for (uint32_t t : tris) if (!d->tri_flags || !(d->tri_flags[t] & DR_TRI_UV_TANGENTS)) { fprintf(stderr, "ref: texcoords without DR_TRI_UV_TANGENTS\n"); return NULL; }
        ref<TriMesh> mesh = new TriMesh("mesh", tris.size(), remap.size(), smooth, hasUV, false, false, !smooth);
        for (auto &kv : remap) {
            if (hasUV) mesh->getVertexTexcoords()[kv.second] = Point2(d->texcoords[2 * kv.first], d->texcoords[2 * kv.first + 1]);
            mesh->getVertexPositions()[kv.second] = Point(d->positions[3 * kv.first], d->positions[3 * kv.first + 1], d->positions[3 * kv.first + 2]);
            if (smooth) mesh->getVertexNormals()[kv.second] = Normal(d->normals[3 * kv.first], d->normals[3 * kv.first + 1], d->normals[3 * kv.first + 2]);
        }
        for (size_t i = 0; i < tris.size(); ++i) for (int k = 0; k < 3; ++k) mesh->getTriangles()[i].idx[k] = remap[d->indices[3 * tris[i] + k]];
        mesh->addChild(rs->bsdfs[std::get<0>(g.first)]);
        rs->bsdfs[std::get<0>(g.first)]->setParent(mesh);
        int32_t e = std::get<1>(g.first);
        if (e >= 0) {
            Properties ep("area");
            ep.setSpectrum("radiance", rgbSpectrum(d->emitters[e].radiance));
            ep.setFloat("samplingWeight", (Float) d->emitters[e].sampling_weight);
            ref<ConfigurableObject> em = pm->createObject(MTS_CLASS(Emitter), ep);
            mesh->addChild(em);
            em->setParent(mesh);       // what the scene loader does after every addChild (scenehandler.cpp)
        }
        mesh->configure();
        rs->scene->addChild(mesh);
    }
    TR("meshes");
    /* Scene::configure would otherwise instantiate the "direct" plugin (scene.cpp:273-277); PathSampler never uses it */
    ref<ConfigurableObject> integ = pm->createObject(MTS_CLASS(Integrator), integrator ? *integrator : Properties("path"));
    integ->configure();
    rs->scene->addChild(integ);
    rs->scene->configure();
    TR("scene configured");
    rs->scene->initialize();           // builds the SAH kd-tree (skdtree.cpp), emitter / area distributions
    TR("scene initialized");
    return rs;
}

void ref_scene_destroy(void *h) { delete (RefScene *) h; }

/* Texture2D::eval(its, filter) of a dr_texture at n intersection uv pairs (its.hasUVPartials = false, as on the bidirectional path) */
int ref_texture_eval(const dr_texture *t, const double *uv, int n, double *rgb, double *average) {
    initOnce();
    ref<PinTexture> tex = new PinTexture(*t);
    Intersection its;
    its.hasUVPartials = false;
    for (int i = 0; i < n; ++i) {
        its.uv = Point2(uv[2 * i], uv[2 * i + 1]);
        Float r, g, b;
        tex->eval(its, true).toLinearRGB(r, g, b);
        rgb[3 * i] = r; rgb[3 * i + 1] = g; rgb[3 * i + 2] = b;
    }
    if (average) { Float r, g, b; tex->getAverage().toLinearRGB(r, g, b); average[0] = r; average[1] = g; average[2] = b; }
    return 0;
}
void *ref_scene_ptr(void *h) { return ((RefScene *) h)->scene.get(); }     // the mitsuba::Scene, for the drivers in other translation units

// findMaxDimensions (pssmlt_utils.h:27-77) on the scene built from the caller's dr_scene_desc: the primary-sample space sizes of
// the three samplers, which depend on the scene (a RoughDielectric BSDF anywhere adds a dimension per vertex).
int ref_max_dimensions(void *h, int max_depth, int rr_depth, int depth, int technique, int direct_sampling, int *out3) {
    try {
        const PathSampler::ETechnique tech = technique == DR_TECH_MMLT ? PathSampler::EMMLT
            : technique == DR_TECH_PATH ? PathSampler::EUnidirectional : PathSampler::EBidirectional;
        MaxDim md = findMaxDimensions(((RefScene *) h)->scene.get(), max_depth, rr_depth, depth, tech, direct_sampling != 0);
        out3[0] = md.sensor; out3[1] = md.emitter; out3[2] = md.direct;
        return 0;
    } catch (const std::exception &e) { fprintf(stderr, "oracle/_ref: %s\n", e.what()); return 1; }
}

/* PathSampler::sampleSplats on replayed primary-sample vectors (pathsampler.cpp:79-571). */
int ref_eval_paths(void *h, int technique, int max_depth, int rr_depth, int sample_direct, int light_image,
                   const float *u_sensor, int dim_sensor, const float *u_emitter, int dim_emitter,
                   const float *u_direct, int dim_direct, const int32_t *depth, int64_t n,
                   dr_path_result *out, double *lum_out, int32_t *consumed /* 3 per path, may be NULL */) {
    RefScene *rs = (RefScene *) h;
    ref<ReplaySampler> se = new ReplaySampler(), em = new ReplaySampler(), di = new ReplaySampler();
    PathSampler::ETechnique tech = technique == DR_TECH_MMLT ? PathSampler::EMMLT
                                 : technique == DR_TECH_BDPT ? PathSampler::EBidirectional : PathSampler::EUnidirectional;
    ref<PathSampler> ps = new PathSampler(tech, rs->scene, em, se, di, max_depth, rr_depth, false, sample_direct != 0, light_image != 0);
    SplatList list;
    for (int64_t j = 0; j < n; ++j) {
        se->set(u_sensor + j * dim_sensor, dim_sensor);
        em->set(u_emitter + j * dim_emitter, dim_emitter);
        di->set(u_direct + j * dim_direct, dim_direct);
        ps->sampleSplats(Point2i(-1), list, depth ? depth[j] : -1);
        dr_path_result &r = out[j];
        memset(&r, 0, sizeof(r));
        r.luminance = (float) list.luminance;
        r.s = list.s; r.t = list.t;
        r.n_splats = (int32_t) std::min(list.splats.size(), (size_t) DR_MAX_SPLATS);
        for (int k = 0; k < r.n_splats; ++k) {
            r.pos[k][0] = (float) list.splats[k].first.x; r.pos[k][1] = (float) list.splats[k].first.y;
            Float R, G, B; list.splats[k].second.toLinearRGB(R, G, B);
            r.value[k][0] = (float) R; r.value[k][1] = (float) G; r.value[k][2] = (float) B;
        }
        if (lum_out) lum_out[j] = list.luminance;
        if (consumed) { consumed[3 * j] = se->consumed(); consumed[3 * j + 1] = em->consumed(); consumed[3 * j + 2] = di->consumed(); }
    }
    return (se->overflow || em->overflow || di->overflow) ? 1 : 0;
}

/* The reference's OWN integrator, end to end: DRMLT::render / PSSMLT::render (drmlt.cpp:393-611) through a RenderJob on
 * `threads` local workers, as src/mitsuba/mitsuba.cpp runs it.  Returns the developed image, the wall time of the job
 * and the integrator's statistics counters (acceptance rates) as text.  The reference seeds its generators from
 * /dev/urandom (random.cpp:473-489): results are comparable statistically, not sample by sample. */
int ref_render(const dr_scene_desc *d, const dr_config *c, int sample_count, int threads, float *image_rgb,
               double *seconds, double *scene_seconds, char *stats, int stats_len) {
    try {
        initOnce();
        Properties ip(c->integrator == DR_INTEGRATOR_DRMLT ? "drmlt" : "pssmlt");
        ip.setString("technique", c->technique == DR_TECH_MMLT ? "mmlt" : c->technique == DR_TECH_BDPT ? "bdpt" : "path");
        if (c->integrator == DR_INTEGRATOR_DRMLT) {
            const char *types[] = { "green", "mira", "orbital", "mirasym" };
            ip.setString("type", types[c->type]);
            ip.setBoolean("acceptanceMap", c->acceptance_map != 0);
            ip.setBoolean("timidAfterLarge", c->timid_after_large != 0);
            ip.setBoolean("fixEmitterPath", c->fix_emitter_path != 0);
            ip.setBoolean("useMixture", c->use_mixture != 0);
            ip.setFloat("scaleSecond", (Float) c->scale_second);
        } else {
            ip.setBoolean("kelemenStyleMutation", c->kelemen_style_mutation != 0);
            ip.setFloat("mutationSizeLow", (Float) c->mutation_size_low);
            ip.setFloat("mutationSizeHigh", (Float) c->mutation_size_high);
        }
        ip.setInteger("maxDepth", c->max_depth);
        ip.setInteger("rrDepth", c->rr_depth);
        ip.setBoolean("directSampling", c->direct_sampling != 0);
        ip.setInteger("directSamples", c->direct_samples);
        ip.setInteger("luminanceSamples", c->luminance_samples);
        ip.setFloat("pLarge", (Float) c->p_large);
        ip.setInteger("workUnits", c->work_units);
        ip.setBoolean("kelemenStyleWeights", c->kelemen_style_weights != 0);
        ip.setBoolean("lightImage", c->light_image != 0);
        ip.setFloat("sigma", (Float) c->sigma);
        ip.setInteger("timeout", c->timeout);
        ip.setBoolean("twoStage", c->two_stage != 0);                      // mltLuminancePass: the nested film is the "hdrfilm" alias above
        ip.setInteger("firstStageSizeReduction", c->first_stage_size_reduction);

        Scheduler *sched = Scheduler::getInstance();
        for (int i = 0; i < threads; ++i) sched->registerWorker(new LocalWorker(i, formatString("wrk%i", i)));
        sched->start();
        ref<Timer> timer = new Timer();
        /* REF_ANALYTIC_SCENE=1: the same job on a room of analytic shapes (scene_create) -- exercises the plugin's tessellation */
        RefScene *rs = (RefScene *) scene_create(d, c->rfilter, &ip, sample_count, getenv("REF_ANALYTIC_SCENE") != NULL);
        *scene_seconds = timer->getMilliseconds() / 1000.0;
        Statistics::getInstance()->resetAll();
        ref<RenderQueue> queue = new RenderQueue();
        int sceneResID = sched->registerResource(rs->scene);
        ref<RenderJob> job = new RenderJob("rend", rs->scene, queue, sceneResID, -1, -1, false, false);
        timer->reset();
        TR("job start");
        job->start();
        TR("job started");
        queue->waitLeft(0);
        TR("queue waited");
        queue->join();
        TR("queue joined");
        *seconds = timer->getMilliseconds() / 1000.0;
        sched->unregisterResource(sceneResID);
        sched->stop();                       // as src/mitsuba/mitsuba.cpp does after its jobs; workers are re-registered per call
        for (size_t i = sched->getWorkerCount(); i-- > 0; ) sched->unregisterWorker(sched->getWorker((int) i));
        const std::string st = Statistics::getInstance()->getStats();
        snprintf(stats, stats_len, "%s", st.c_str());
        if (rs->film->image.empty()) { delete rs; return 2; }
        memcpy(image_rgb, rs->film->image.data(), rs->film->image.size() * sizeof(float));
        delete rs;
        return 0;
    } catch (const std::exception &e) {
        fprintf(stderr, "oracle/_ref: %s\n", e.what());
        try {                                /* leave no worker thread behind: the calling process must be able to exit */
            Scheduler *sched = Scheduler::getInstance();
            if (sched->isRunning()) sched->stop();
            for (size_t i = sched->getWorkerCount(); i-- > 0; ) sched->unregisterWorker(sched->getWorker((int) i));
        } catch (...) {}
        return 1;
    }
}

/* The separate direct-illumination image: BidirectionalUtils::renderDirectComponent (src/libbidir/util.cpp:30-94) itself -- the
 * `direct` integrator (src/integrators/direct/direct.cpp) with the pixelSamples x shadingSamples split, an `ldsampler` per worker,
 * SamplingIntegrator::render on `threads` local workers, film->develop.  The ldsampler scrambles from /dev/urandom: comparable
 * statistically. */
int ref_direct_image(const dr_scene_desc *d, int rfilter, int direct_samples, int threads, float *image_rgb) {
    try {
        initOnce();
        Scheduler *sched = Scheduler::getInstance();
        for (int i = 0; i < threads; ++i) sched->registerWorker(new LocalWorker(i, formatString("wrk%i", i)));
        sched->start();
        RefScene *rs = (RefScene *) scene_create(d, rfilter);
        ref<RenderQueue> queue = new RenderQueue();
        int sceneResID = sched->registerResource(rs->scene);
        int sensorResID = sched->registerResource(rs->scene->getSensor());
        ref<RenderJob> job = new RenderJob("dire", rs->scene, queue, sceneResID, sensorResID, -1, false, false);
        rs->scene->getFilm()->clear();
        ref<Bitmap> img = BidirectionalUtils::renderDirectComponent(rs->scene, sceneResID, sensorResID, queue, job, (size_t) direct_samples);
        sched->unregisterResource(sceneResID);
        sched->unregisterResource(sensorResID);
        sched->stop();
        for (size_t i = sched->getWorkerCount(); i-- > 0; ) sched->unregisterWorker(sched->getWorker((int) i));
        if (!img) { delete rs; return 2; }
        const Float *src = img->getFloatData();
        const size_t n = (size_t) img->getWidth() * img->getHeight();
        for (size_t i = 0; i < n; ++i) {
            Spectrum sp; for (int q = 0; q < SPECTRUM_SAMPLES; ++q) sp[q] = src[i * SPECTRUM_SAMPLES + q];
            Float R, G, B; sp.toLinearRGB(R, G, B);
            image_rgb[3 * i] = (float) R; image_rgb[3 * i + 1] = (float) G; image_rgb[3 * i + 2] = (float) B;
        }
        delete rs;
        return 0;
    } catch (const std::exception &e) {
        fprintf(stderr, "oracle/_ref: %s\n", e.what());
        try {
            Scheduler *sched = Scheduler::getInstance();
            if (sched->isRunning()) sched->stop();
            for (size_t i = sched->getWorkerCount(); i-- > 0; ) sched->unregisterWorker(sched->getWorker((int) i));
        } catch (...) {}
        return 1;
    }
}

/* The up-sampling of the first-stage luminance map, as mltLuminancePass does it (util.cpp:175-196): Bitmap::resample
 * (src/libcore/bitmap.cpp:2230-2400) of an ELuminance float bitmap with the gaussian filter plugin, EClamp borders, clamped to
 * [0, inf). */
int ref_resample_luminance(const double *lum, int w, int h, int W, int H, double *out) {
    try {
        initOnce();
        ref<ReconstructionFilter> rfilter = static_cast<ReconstructionFilter *>(
            PluginManager::getInstance()->createObject(MTS_CLASS(ReconstructionFilter), Properties("gaussian")));
        rfilter->configure();
        ref<Bitmap> src = new Bitmap(Bitmap::ELuminance, Bitmap::EFloat, Vector2i(w, h));
        Float *sd = src->getFloatData();
        for (size_t i = 0; i < (size_t) w * h; ++i) sd[i] = (Float) lum[i];
        ref<Bitmap> dst = src->resample(rfilter, ReconstructionFilter::EClamp, ReconstructionFilter::EClamp, Vector2i(W, H),
                                        0.0f, std::numeric_limits<Float>::infinity());
        const Float *dd = dst->getFloatData();
        for (size_t i = 0; i < (size_t) W * H; ++i) out[i] = dd[i];
        return 0;
    } catch (const std::exception &e) { fprintf(stderr, "oracle/_ref: %s\n", e.what()); return 1; }
}

/* BSDF::sample / eval / pdf of the reference's plugins in the local frame (what orc_bsdf_sample / orc_bsdf_eval
 * expose of the oracle); `extra` is the number an EUsesSampler BSDF draws from bRec.sampler (roughdielectric.cpp:555). */
static void makeIts(Intersection &its) {
    its.p = Point(0, 0, 0); its.t = 1; its.time = 0;
    its.geoFrame = its.shFrame = Frame(Vector(1, 0, 0), Vector(0, 1, 0), Normal(0, 0, 1));
    its.uv = Point2(0.5, 0.5); its.dpdu = Vector(1, 0, 0); its.dpdv = Vector(0, 1, 0); its.hasUVPartials = false;
}
void ref_bsdf_sample(const dr_material *m, const double *wi, int mode, double u1, double u2, double extra,
                     double *wo, double *weight, double *pdf, int *sampledType, double *eta) {
    initOnce();
    ref<BSDF> bsdf = makeBSDF(*m);
    Intersection its; makeIts(its);
    its.wi = Vector(wi[0], wi[1], wi[2]);
    float ex = (float) extra;
    ref<ReplaySampler> smp = new ReplaySampler();
    smp->set(&ex, 1);
    BSDFSamplingRecord bRec(its, smp, mode == 0 ? ERadiance : EImportance);
    Float p = 0;
    Spectrum w = bsdf->sample(bRec, p, Point2(u1, u2));
    wo[0] = bRec.wo.x; wo[1] = bRec.wo.y; wo[2] = bRec.wo.z;
    Float R, G, B; w.toLinearRGB(R, G, B);
    weight[0] = R; weight[1] = G; weight[2] = B;
    *pdf = p; *sampledType = (int) bRec.sampledType; *eta = bRec.eta;
}
void ref_bsdf_eval(const dr_material *m, const double *wi, const double *wo, int mode, int measure, double *value, double *pdf) {
    initOnce();
    ref<BSDF> bsdf = makeBSDF(*m);
    Intersection its; makeIts(its);
    BSDFSamplingRecord bRec(its, Vector(wi[0], wi[1], wi[2]), Vector(wo[0], wo[1], wo[2]), mode == 0 ? ERadiance : EImportance);
    Spectrum v = bsdf->eval(bRec, (EMeasure) measure);
    Float R, G, B; v.toLinearRGB(R, G, B);
    value[0] = R; value[1] = G; value[2] = B;
    *pdf = bsdf->pdf(bRec, (EMeasure) measure);
}

}  // extern "C"
