// mts_plugin.cpp -- Mitsuba 0.6 integrator plugins `drmlt` and `pssmlt` backed by libdrmlt_b200.so.
//
// This is the reference-side half of the drop-in boundary.  Built once per plugin
// name inside a drmlt-mitsuba source tree (it needs Mitsuba's headers and links libmitsuba-core /
// libmitsuba-render, which are absent from the build image of this repository -- see INTEGRATION.md):
//
//   g++ -std=c++11 -shared -fPIC -DDR_PLUGIN_DRMLT  mts_plugin.cpp -I<mitsuba>/include -I<repo>/include \
//       -L<repo>/drmlt-mitsuba_b200/csrc -ldrmlt_b200 -lmitsuba-core -lmitsuba-render -o plugins/drmlt.so
//   g++ ... -DDR_PLUGIN_PSSMLT ... -o plugins/pssmlt.so
//
// It replaces src/integrators/drmlt/drmlt.cpp:176-621 and src/integrators/pssmlt/pssmlt.cpp:164-556 as the
// object PluginManager::createObject instantiates for <integrator type="drmlt|pssmlt"> (plugin.cpp:180-196):
// same class names, same parameters (they are forwarded verbatim to dr_config_set, which knows the
// reference's names and defaults), same output convention (film->setBitmap + queue->signalRefresh,
// drmlt_proc.cpp:850-853).  Scene loading (XML, meshes), the film and its file output stay Mitsuba's.
#include <mitsuba/render/scene.h>
#include <mitsuba/render/renderjob.h>
#include <mitsuba/render/renderqueue.h>
#include <mitsuba/render/trimesh.h>
#include <mitsuba/render/bsdf.h>
#include <mitsuba/render/emitter.h>
#include <mitsuba/render/sensor.h>
#include <mitsuba/render/film.h>
#include <mitsuba/core/bitmap.h>
#include <mitsuba/core/plugin.h>
#include <mitsuba/core/sched.h>
#include <mitsuba/core/rfilter.h>
#include <mitsuba/core/statistics.h>
#include <mitsuba/core/lock.h>
#include <mitsuba/core/mstream.h>
#include <mitsuba/core/serialization.h>
#include <mitsuba/core/half.h>
#include <mitsuba/render/texture.h>
#include <drmlt_b200.h>
#include "src/bsdfs/rtrans.h"          // RoughTransmittance, as src/bsdfs/roughplastic.cpp includes it (compile with -I<mitsuba root>)
#include <cstdlib>
#include <cctype>
#include <cstring>
#include <sstream>
#include <vector>
#include <map>
#include <algorithm>

MTS_NAMESPACE_BEGIN

// InstanceManager keeps the objects it has serialized in a private map (include/mitsuba/core/serialization.h:84-88).  A BSDF exposes
// no getter for its textures, but it hands every one of them to the manager when it is serialized (e.g. diffuse.cpp:161-165) -- which is
// the only enumeration of a BSDF's children Mitsuba offers.  Access to the private member through an explicit template instantiation
// (the one place the standard lets a private member be named).
template <typename Tag, typename Tag::type M> struct DrPrivateAccess { friend typename Tag::type get(Tag) { return M; } };
struct DrObjToId { typedef std::map<const SerializableObject *, unsigned int> InstanceManager::*type; friend type get(DrObjToId); };
template struct DrPrivateAccess<DrObjToId, &InstanceManager::m_objToId>;

namespace {

// ---- BSDF flattening ------------------------------------------------------------------------------
// Parameters come from the Properties the BSDF was constructed with (ConfigurableObject::getProperties,
// include/mitsuba/core/cobject.h:77) -- exact values.  The `twosided` adapter owns its nested BSDF privately
// (src/bsdfs/twosided.cpp:85-100) and exposes no getter, so the nested model is recovered from toString()
// (twosided.cpp: "nestedBRDF[0] = <nested toString>"), which prints spectra with 6 significant digits.
struct ToStringParser {
    const std::string &s;
    explicit ToStringParser(const std::string &str) : s(str) {}
    std::string className(size_t from = 0) const {
        size_t b = s.find_first_not_of(" \n\t", from), e = s.find('[', b);
        return s.substr(b, e - b);
    }
    // "key = [a, b, c]" or "key = ConstantSpectrumTexture[value=[a, b, c]]" / "ConstantFloatTexture[value=a]"
    bool spectrum(const char *key, size_t from, float out[3]) const {
        size_t p = s.find(std::string(key) + " = ", from);
        if (p == std::string::npos) return false;
        size_t lb = s.find('[', p);
        if (s.compare(lb + 1, 6, "value=") == 0 || s.find("Texture[", p) < lb + 1) lb = s.find('[', lb + 1);
        float a, b, c;
        if (sscanf(s.c_str() + lb, "[%f, %f, %f]", &a, &b, &c) != 3) return false;
        out[0] = a; out[1] = b; out[2] = c;
        return true;
    }
    bool scalar(const char *key, size_t from, float &out) const {
        size_t p = s.find(std::string(key) + " = ", from);
        if (p == std::string::npos) return false;
        p += strlen(key) + 3;
        size_t v = s.find("value=", p), nl = s.find('\n', p);
        if (v != std::string::npos && v < nl) p = v + 6;
        return sscanf(s.c_str() + p, "%f", &out) == 1;
    }
};

void toRGB(const Spectrum &sp, float out[3]) {
    Float r, g, b;
    sp.toLinearRGB(r, g, b);
    out[0] = (float) r; out[1] = (float) g; out[2] = (float) b;
}

// The rough-transmittance table of one roughplastic material (include/drmlt_b200.h, DR_ROUGH_TABLE_*), from the host's own
// RoughTransmittance and data/microfacet/*.dat, reduced exactly as RoughPlastic::configure reduces it (roughplastic.cpp:283-301)
void roughTable(bool ggx, Float eta, Float alpha, std::vector<double> &tables) {
    ref<RoughTransmittance> ext = new RoughTransmittance(ggx ? MicrofacetDistribution::EGGX : MicrofacetDistribution::EBeckmann);
    ext->checkEta(eta); ext->checkAlpha(alpha);
    ref<RoughTransmittance> in = ext->clone();
    ext->setEta(eta);
    in->setEta(1 / eta);
    struct Peek : public RoughTransmittance {
        static const Float *trans(const RoughTransmittance *r) { return r->*(&Peek::m_trans); }
        static size_t thetaSamples(const RoughTransmittance *r) { return r->*(&Peek::m_thetaSamples); }
    };
    const size_t base = tables.size();
    tables.resize(base + DR_ROUGH_TABLE_DOUBLES, 0.0);
    tables[base + 100] = in->evalDiffuse(alpha);
    ext->setAlpha(alpha);
    tables[base + 101] = ext->evalDiffuse(alpha);
    if (Peek::thetaSamples(ext.get()) != DR_ROUGH_TABLE_THETA) SLog(EError, "Unexpected rough transmittance table size");
    for (int k = 0; k < DR_ROUGH_TABLE_THETA; ++k) tables[base + k] = (double) Peek::trans(ext.get())[k];
}

// ---- bitmap textures (include/drmlt_b200.h: dr_texture) --------------------------------------------------------------
struct FlatTextures {
    std::vector<dr_texture> textures;
    std::vector<std::vector<float> > texels;              // one array per texture (stable addresses: filled, then pointers taken)
    std::map<const Texture2D *, uint32_t> index;
    std::map<const BSDF *, std::vector<const Texture2D *> > ofBsdf;   // a BSDF shared by many meshes is serialized once (BitmapTexture::serialize reads its file)
};

// The non-constant 2D textures of a BSDF in the order it serializes them (see DrPrivateAccess above); false: a texture this path cannot take.
bool texturesOf(const BSDF *bsdf, std::vector<const Texture2D *> &out, std::string &why) {
    ref<MemoryStream> ms = new MemoryStream();
    ref<InstanceManager> im = new InstanceManager();
    im->serialize(ms, bsdf);
    const std::map<const SerializableObject *, unsigned int> &objs = (*im).*get(DrObjToId());
    std::vector<std::pair<unsigned int, const Texture2D *> > found;
    for (std::map<const SerializableObject *, unsigned int>::const_iterator it = objs.begin(); it != objs.end(); ++it) {
        const Class *cls = it->first->getClass();
        if (!cls->derivesFrom(MTS_CLASS(Texture))) continue;
        const Texture *t = static_cast<const Texture *>(it->first);
        if (t->isConstant()) continue;
        if (!cls->derivesFrom(MTS_CLASS(Texture2D)) || t->toString().find("TMIPMap[") == std::string::npos) {
            why = "texture " + cls->getName() + " (only bitmap textures are flattened)"; return false;
        }
        found.push_back(std::make_pair(it->second, static_cast<const Texture2D *>(t)));
    }
    std::sort(found.begin(), found.end());
    for (size_t i = 0; i < found.size(); ++i) out.push_back(found[i].second);
    return true;
}

// One bitmap texture -> dr_texture: MIP level 0 as the texture holds it (BitmapTexture::getBitmap = TMIPMap::toBitmap, bitmap.cpp:483-485:
// half precision), filter and boundary conditions from the mipmap's description (mipmap.h:718-740), Texture2D's uv scale / offset.
bool flattenTexture(const Texture2D *t, FlatTextures &ft, uint32_t &idx, std::string &why) {
    std::map<const Texture2D *, uint32_t>::const_iterator known = ft.index.find(t);
    if (known != ft.index.end()) { idx = known->second; return true; }
    struct Peek : public Texture2D {
        static Point2 offset(const Texture2D *x) { return x->*(&Peek::m_uvOffset); }
        static Vector2 scale(const Texture2D *x) { return x->*(&Peek::m_uvScale); }
    };
    const std::string str = t->toString();
    ref<Bitmap> bmp = t->getBitmap();
    const int ch = bmp->getChannelCount();
    if (!(bmp->getPixelFormat() == Bitmap::ERGB || bmp->getPixelFormat() == Bitmap::ELuminance) || (ch != 1 && ch != 3)) { why = "bitmap texture with an unsupported pixel format"; return false; }
    const size_t n = (size_t) bmp->getWidth() * (size_t) bmp->getHeight();
    std::vector<float> texels(3 * n);
    for (size_t i = 0; i < n; ++i)
        for (int c = 0; c < 3; ++c) {
            const size_t k = i * ch + (ch == 3 ? c : 0);
            float v;
            switch (bmp->getComponentFormat()) {
                case Bitmap::EFloat16: v = (float) bmp->getFloat16Data()[k]; break;
                case Bitmap::EFloat32: v = bmp->getFloat32Data()[k]; break;
                case Bitmap::EFloat64: v = (float) bmp->getFloat64Data()[k]; break;
                default: why = "bitmap texture with an unsupported component format"; return false;
            }
            texels[3 * i + c] = v;
        }
    dr_texture d;
    memset(&d, 0, sizeof(d));
    d.width = (uint32_t) bmp->getWidth(); d.height = (uint32_t) bmp->getHeight();
    // "filterType = bilinear," and "bc = [repeat, mirror]," (ReconstructionFilter::EBoundaryCondition by name, rfilter.cpp operator<<)
    size_t f = str.find("filterType = "), b = str.find("bc = [");
    if (f == std::string::npos || b == std::string::npos) { why = "cannot parse the texture's MIP map description"; return false; }
    d.nearest = str.compare(f + 13, 7, "nearest") == 0;      // ewa / trilinear / bilinear: bilinear in level 0 without ray differentials (bitmap.cpp:432-455)
    const size_t comma = str.find(", ", b), close = str.find(']', b);
    if (comma == std::string::npos || close == std::string::npos || comma > close) { why = "cannot parse the texture's wrap modes"; return false; }
    const std::string names[2] = { str.substr(b + 6, comma - (b + 6)), str.substr(comma + 2, close - (comma + 2)) };
    uint32_t wrap[2];
    for (int a = 0; a < 2; ++a) {
        if (names[a] == "repeat") wrap[a] = DR_WRAP_REPEAT;
        else if (names[a] == "clamp") wrap[a] = DR_WRAP_CLAMP;
        else if (names[a] == "mirror") wrap[a] = DR_WRAP_MIRROR;
        else if (names[a] == "zero") wrap[a] = DR_WRAP_ZERO;
        else if (names[a] == "one") wrap[a] = DR_WRAP_ONE;
        else { why = "unknown texture wrap mode " + names[a]; return false; }
    }
    d.wrap_u = wrap[0]; d.wrap_v = wrap[1];
    d.uv_scale[0] = (double) Peek::scale(t).x; d.uv_scale[1] = (double) Peek::scale(t).y;
    d.uv_offset[0] = (double) Peek::offset(t).x; d.uv_offset[1] = (double) Peek::offset(t).y;
    idx = (uint32_t) ft.textures.size();
    if (idx >= DR_MAX_TEXTURES) { why = "too many textures"; return false; }
    ft.textures.push_back(d);
    ft.texels.push_back(std::vector<float>());
    ft.texels.back().swap(texels);
    ft.index[t] = idx;
    return true;
}

bool flattenBSDF(const BSDF *bsdf, dr_material &m, std::string &why, std::vector<double> &roughTables, FlatTextures &ft) {
    memset(&m, 0, sizeof(m));
    m.reflectance[0] = m.reflectance[1] = m.reflectance[2] = 1.f;
    m.transmittance[0] = m.transmittance[1] = m.transmittance[2] = 1.f;
    const std::string cls = bsdf->getClass()->getName();
    const Properties &props = bsdf->getProperties();
    const std::string str = bsdf->toString();
    ToStringParser ts(str);
    size_t from = 0;
    std::string model = cls;
    bool nested = false;
    if (cls == "TwoSidedBRDF") {
        m.flags |= DR_MAT_TWOSIDED;
        from = str.find("nestedBRDF[0] = ");
        if (from == std::string::npos) { why = "cannot parse twosided"; return false; }
        from += strlen("nestedBRDF[0] = ");
        {   // a different BSDF on the back side (nestedBRDF[1], twosided.cpp:85-100) is not representable: fail instead of using the front's
            const size_t second = str.find("nestedBRDF[1] = ", from);
            if (second == std::string::npos) { why = "cannot parse twosided"; return false; }
            std::string a, b;
            for (size_t i = from; i < second; ++i) if (!isspace((unsigned char) str[i])) a += str[i];
            for (size_t i = second + strlen("nestedBRDF[1] = "); i < str.size(); ++i) if (!isspace((unsigned char) str[i])) b += str[i];
            if (!a.empty() && a[a.size() - 1] == ',') a.erase(a.size() - 1);          // "...]," before nestedBRDF[1]
            if (!b.empty() && b[b.size() - 1] == ']') b.erase(b.size() - 1);          // the closing bracket of TwoSided[...]
            if (a != b) { why = "twosided with different front and back BSDFs"; return false; }
        }
        model = ts.className(from);
        if (model.compare(0, 4, "ref<") == 0) {          // ref<T>::toString: "ref<SmoothDiffuse>[ref=2, ptr=SmoothDiffuse[..."
            size_t ptr = str.find("ptr=", from);
            if (ptr == std::string::npos) { why = "cannot parse twosided"; return false; }
            from = ptr + 4;
            model = ts.className(from);
        }
        nested = true;
    }
    // Textured colour parameters: the BSDF's non-constant textures, in the order it serializes them (diffuse.cpp:161-165, dielectric.cpp:176-182,
    // conductor.cpp:201-205, roughconductor.cpp:215-224, roughdielectric.cpp:226-237, plastic.cpp:176-184, roughplastic.cpp:242-253), matched with
    // the parameters toString() prints as something other than a Constant*Texture.  slot 0 = dr_material.reflectance, 1 = .transmittance.
    struct Param { const char *name; int slot; };
    static const Param pDiffuse[] = { { "reflectance", 0 }, { NULL, 0 } };
    static const Param pDielectric[] = { { "specularReflectance", 0 }, { "specularTransmittance", 1 }, { NULL, 0 } };
    static const Param pConductor[] = { { "specularReflectance", 0 }, { NULL, 0 } };
    static const Param pRoughConductor[] = { { "alphaU", -1 }, { "alphaV", -1 }, { "specularReflectance", 0 }, { NULL, 0 } };
    static const Param pRoughDielectric[] = { { "alphaU", -1 }, { "alphaV", -1 }, { "specularReflectance", 0 }, { "specularTransmittance", 1 }, { NULL, 0 } };
    static const Param pPlastic[] = { { "specularReflectance", 1 }, { "diffuseReflectance", 0 }, { NULL, 0 } };
    static const Param pRoughPlastic[] = { { "specularReflectance", 1 }, { "diffuseReflectance", 0 }, { "alpha", -1 }, { NULL, 0 } };
    const Param *params = model == "SmoothDiffuse" ? pDiffuse : model == "SmoothDielectric" ? pDielectric : model == "SmoothConductor" ? pConductor :
        model == "RoughConductor" ? pRoughConductor : model == "RoughDielectric" ? pRoughDielectric : model == "SmoothPlastic" ? pPlastic :
        model == "RoughPlastic" ? pRoughPlastic : NULL;
    std::vector<const Texture2D *> texs;
    {
        std::map<const BSDF *, std::vector<const Texture2D *> >::const_iterator cached = ft.ofBsdf.find(bsdf);
        if (cached != ft.ofBsdf.end()) texs = cached->second;
        else {
            if (!texturesOf(bsdf, texs, why)) return false;
            ft.ofBsdf[bsdf] = texs;
        }
    }
    if (!texs.empty()) {
        size_t k = 0;
        bool texRset = false, texTset = false;
        for (const Param *p = params; p && p->name; ++p) {
            size_t pos = str.find(std::string(p->name) + " = ", from);
            if (pos == std::string::npos) continue;
            // constants print as "Constant*Texture[...]", a bare spectrum "[r, g, b]" or a bare number (ConstantFloatTexture, basicshader.h:125-129)
            const size_t vpos = str.find_first_not_of(" \n\t", pos + strlen(p->name) + 3);
            if (vpos == std::string::npos || str.compare(vpos, 8, "Constant") == 0 || strchr("[-+.0123456789", str[vpos])) continue;
            if (p->slot < 0) { why = "textured roughness"; return false; }
            if (k >= texs.size()) { why = "textures shared between parameters"; return false; }
            const Texture2D *t = texs[k++];
            uint32_t idx = 0;
            if ((p->slot == 0 ? texRset : texTset)) { why = "two textures on one parameter slot"; return false; }
            (p->slot == 0 ? texRset : texTset) = true;
            if (!flattenTexture(t, ft, idx, why)) return false;
            m.flags |= p->slot == 0 ? DR_MAT_TEX_REFLECTANCE(idx) : DR_MAT_TEX_TRANSMITTANCE(idx);
            toRGB(t->getAverage(), p->slot == 0 ? m.reflectance : m.transmittance);   // what the sampling weights of plastic / roughplastic read
        }
        if (k != texs.size()) { why = "texture on a parameter of " + model + " this path does not take (scaled or nested textures): " + str.substr(0, 300); return false; }
    }
    const bool texR = ((m.flags >> 8) & 0xfffu) != 0, texT = (m.flags >> 20) != 0;
    if (model == "SmoothDiffuse") {                       // src/bsdfs/diffuse.cpp
        m.type = DR_BSDF_DIFFUSE;
        // the exact constant when it was given as a property; a <texture> child (no property) must not fall back to the default silently
        if (!nested && (props.hasProperty("reflectance") || props.hasProperty("diffuseReflectance")))
            toRGB(props.getSpectrum(props.hasProperty("reflectance") ? "reflectance" : "diffuseReflectance", Spectrum(.5f)), m.reflectance);
        else if (!texR && !ts.spectrum("reflectance", from, m.reflectance)) { why = "cannot parse diffuse reflectance"; return false; }
    } else if (model == "SmoothDielectric") {             // src/bsdfs/dielectric.cpp
        m.type = DR_BSDF_DIELECTRIC;
        float eta = 0.f;
        if (!ts.scalar("eta", from, eta)) { why = "cannot parse dielectric eta"; return false; }
        m.eta[0] = eta;                                   // intIOR / extIOR, as printed by SmoothDielectric::toString
        ts.spectrum("specularReflectance", from, m.reflectance);
        ts.spectrum("specularTransmittance", from, m.transmittance);
    } else if (model == "SmoothConductor" || model == "RoughConductor") {   // conductor.cpp / roughconductor.cpp
        m.type = model == "SmoothConductor" ? DR_BSDF_CONDUCTOR : DR_BSDF_ROUGHCONDUCTOR;
        if (!ts.spectrum("eta", from, m.eta) || !ts.spectrum("k", from, m.k)) { why = "cannot parse conductor eta/k"; return false; }
        ts.spectrum("specularReflectance", from, m.reflectance);
        if (m.type == DR_BSDF_ROUGHCONDUCTOR) {
            float au = 0.f, av = 0.f;
            if (!ts.scalar("alphaU", from, au) || !ts.scalar("alphaV", from, av) || au != av) { why = "anisotropic or textured roughness"; return false; }
            m.alpha = au;
            size_t d = str.find("distribution = ", from);
            if (d == std::string::npos) { why = "no distribution"; return false; }
            if (str.compare(d + 15, 3, "ggx") == 0) m.flags |= DR_MAT_GGX;
            else if (str.compare(d + 15, 8, "beckmann") != 0) { why = "phong distribution"; return false; }
            size_t v = str.find("sampleVisible = ", from);
            if (v != std::string::npos && str[v + 16] == '1') m.flags |= DR_MAT_SAMPLE_VISIBLE;
        }
    } else if (model == "SmoothPlastic") {                // src/bsdfs/plastic.cpp (toString :479-493)
        m.type = DR_BSDF_PLASTIC;
        float eta = 0.f;
        if (!ts.scalar("eta", from, eta)) { why = "cannot parse plastic eta"; return false; }
        m.eta[0] = eta;                                   // intIOR / extIOR
        if (!texR) m.reflectance[0] = m.reflectance[1] = m.reflectance[2] = 0.5f;       // diffuseReflectance default (plastic.cpp:160)
        if (!texR && !ts.spectrum("diffuseReflectance", from, m.reflectance)) { why = "cannot parse plastic diffuseReflectance"; return false; }
        ts.spectrum("specularReflectance", from, m.transmittance);                      // dr_material: transmittance = specularReflectance
        size_t nl = str.find("nonlinear = ", from);
        if (nl != std::string::npos && str[nl + 12] == '1') m.flags |= DR_MAT_NONLINEAR;
    } else if (model == "RoughPlastic") {                 // src/bsdfs/roughplastic.cpp (toString :600-615)
        m.type = DR_BSDF_ROUGHPLASTIC;
        float eta = 0.f, alpha = 0.f;
        if (!ts.scalar("eta", from, eta)) { why = "cannot parse rough plastic eta"; return false; }
        m.eta[0] = eta;                                   // intIOR / extIOR
        // a constant alpha prints as a bare number (ConstantFloatTexture::toString, basicshader.h:125-129); a textured one was rejected above
        if (!ts.scalar("alpha", from, alpha)) { why = "textured roughness"; return false; }
        m.alpha = alpha;
        if (!texR) m.reflectance[0] = m.reflectance[1] = m.reflectance[2] = 0.5f;       // diffuseReflectance default (roughplastic.cpp:200)
        if (!texR && !ts.spectrum("diffuseReflectance", from, m.reflectance)) { why = "cannot parse rough plastic diffuseReflectance"; return false; }
        if (!texT && !ts.spectrum("specularReflectance", from, m.transmittance)) { why = "cannot parse rough plastic specularReflectance"; return false; }
        size_t d = str.find("distribution = ", from);
        if (d == std::string::npos) { why = "no distribution"; return false; }
        if (str.compare(d + 15, 3, "ggx") == 0) m.flags |= DR_MAT_GGX;
        else if (str.compare(d + 15, 8, "beckmann") != 0) { why = "phong distribution"; return false; }
        size_t v = str.find("sampleVisible = ", from);
        if (v != std::string::npos && str[v + 16] == '1') m.flags |= DR_MAT_SAMPLE_VISIBLE;
        size_t nl = str.find("nonlinear = ", from);
        if (nl != std::string::npos && str[nl + 12] == '1') m.flags |= DR_MAT_NONLINEAR;
        // the reference averages the constant alpha texture's Spectrum with a float third before it reaches the tables (spectrum.h:481-486)
        Float av = 0; av += (Float) m.alpha; av += (Float) m.alpha; av += (Float) m.alpha; av *= (1.0f / 3);
        m.table = (uint32_t) (roughTables.size() / DR_ROUGH_TABLE_DOUBLES);
        roughTable((m.flags & DR_MAT_GGX) != 0, (Float) m.eta[0], av, roughTables);
    } else if (model == "RoughDielectric") {              // src/bsdfs/roughdielectric.cpp (toString :659-672)
        if (nested) { why = "twosided rough dielectric"; return false; }
        m.type = DR_BSDF_ROUGHDIELECTRIC;
        float eta = 0.f, au = 0.f, av = 0.f;
        if (!ts.scalar("eta", from, eta)) { why = "cannot parse rough dielectric eta"; return false; }
        m.eta[0] = eta;                                   // intIOR / extIOR
        if (!ts.scalar("alphaU", from, au) || !ts.scalar("alphaV", from, av) || au != av) { why = "anisotropic or textured roughness"; return false; }
        m.alpha = au;
        ts.spectrum("specularReflectance", from, m.reflectance);
        ts.spectrum("specularTransmittance", from, m.transmittance);
        size_t d = str.find("distribution = ", from);
        if (d == std::string::npos) { why = "no distribution"; return false; }
        if (str.compare(d + 15, 3, "ggx") == 0) m.flags |= DR_MAT_GGX;
        else if (str.compare(d + 15, 8, "beckmann") != 0) { why = "phong distribution"; return false; }
        size_t v = str.find("sampleVisible = ", from);
        if (v != std::string::npos && str[v + 16] == '1') m.flags |= DR_MAT_SAMPLE_VISIBLE;
    } else { why = "unsupported BSDF " + model; return false; }
    return true;
}

} // namespace

#if defined(DR_PLUGIN_PSSMLT)
#define DR_CLASS PSSMLT
#define DR_NAME "pssmlt"
// the reference's statistics counters, under the reference's own names (pssmlt_proc.cpp:33-40): `mitsuba` prints the same table
static StatsCounter largeStepRatio("Primary sample space MLT", "Accepted large steps", EPercentage);
static StatsCounter smallStepRatio("Primary sample space MLT", "Accepted small steps", EPercentage);
static StatsCounter acceptanceRate("Primary sample space MLT", "Overall acceptance rate", EPercentage);
static void publishStats(const dr_stats &st) {
    largeStepRatio += st.large_accept; largeStepRatio.incrementBase(st.large_base);
    smallStepRatio += st.bold_accept; smallStepRatio.incrementBase(st.bold_base);
    acceptanceRate += st.accept; acceptanceRate.incrementBase(st.accept_base);
}
#else
#define DR_CLASS DRMLT
#define DR_NAME "drmlt"
// the reference's statistics counters, under the reference's own names (drmlt_proc.cpp:34-49)
static StatsCounter firstLevelRatio("Delayed Rejection MLT", "Accepted 1st-stage mutations", EPercentage);
static StatsCounter largeStepRatio("Delayed Rejection MLT", "Accepted large mutations in the 1st-stage mutations", EPercentage);
static StatsCounter boldStepRatio("Delayed Rejection MLT", "Accepted bold mutation in the 1st-stage mutations", EPercentage);
static StatsCounter secondLevelRatio("Delayed Rejection MLT", "Accepted 2nd-stage mutations", EPercentage);
static StatsCounter secondLevelLargeRatio("Delayed Rejection MLT", "Accepted 2nd-stage mutations after large mutation", EPercentage);
static StatsCounter secondLevelBoldRatio("Delayed Rejection MLT", "Accepted 2nd-stage mutations after bold mutation", EPercentage);
static StatsCounter acceptanceRate("Delayed Rejection MLT", "Overall acceptance rate", EPercentage);
static void publishStats(const dr_stats &st) {
    firstLevelRatio += st.first_accept; firstLevelRatio.incrementBase(st.first_base);
    largeStepRatio += st.large_accept; largeStepRatio.incrementBase(st.large_base);
    boldStepRatio += st.bold_accept; boldStepRatio.incrementBase(st.bold_base);
    secondLevelRatio += st.second_accept; secondLevelRatio.incrementBase(st.second_base);
    secondLevelLargeRatio += st.second_large_accept; secondLevelLargeRatio.incrementBase(st.second_large_base);
    secondLevelBoldRatio += st.second_bold_accept; secondLevelBoldRatio.incrementBase(st.second_bold_base);
    acceptanceRate += st.accept; acceptanceRate.incrementBase(st.accept_base);
}
#endif

class DR_CLASS : public Integrator {
public:
    DR_CLASS(const Properties &props) : Integrator(props), m_scene(NULL) {
        m_mutex = new Mutex();
        dr_config_default(&m_config);
        check(dr_config_set(&m_config, "integrator", DR_NAME));
        // forward every parameter under the reference's own name (drmlt.cpp:193-349, pssmlt.cpp:181-307)
        std::vector<std::string> names;
        props.putPropertyNames(names);
        for (size_t i = 0; i < names.size(); ++i) {
            const std::string &k = names[i];
            std::ostringstream v;
            switch (props.getType(k)) {
                case Properties::EBoolean: v << (props.getBoolean(k) ? "true" : "false"); break;
                case Properties::EInteger: v << props.getInteger(k); break;
                case Properties::EFloat: v.precision(17); v << props.getFloat(k); break;
                case Properties::EString: v << props.getString(k); break;
                default: continue;
            }
            check(dr_config_set(&m_config, k.c_str(), v.str().c_str()));
        }
    }
    DR_CLASS(Stream *stream, InstanceManager *manager) : Integrator(stream, manager), m_scene(NULL) {
        Log(EError, "Network rendering is not supported by the B200 plugin");
    }
    void serialize(Stream *stream, InstanceManager *manager) const {
        Integrator::serialize(stream, manager);
        Log(EError, "Network rendering is not supported by the B200 plugin");
    }

    bool preprocess(const Scene *scene, RenderQueue *, const RenderJob *, int, int, int) {
        if (scene->getSubsurfaceIntegrators().size() > 0)       // drmlt.cpp:377-378
            Log(EError, "Subsurface integrators are not supported by MLT!");
        if (scene->getSampler()->getClass()->getName() != "IndependentSampler")   // drmlt.cpp:380-381
            Log(EError, "Metropolis light transport requires the independent sampler");
        // participating media are not built on the GPU path: fail, do not render the scene without them
        if (!scene->getMedia().empty() || scene->getSensor()->getMedium() != NULL)
            Log(EError, "Participating media are not supported by the B200 plugin");
        return true;
    }

    bool render(Scene *scene, RenderQueue *queue, const RenderJob *job, int, int, int) {
        ref<Sensor> sensor = scene->getSensor();
        ref<Film> film = sensor->getFilm();
        const Vector2i size = film->getCropSize();
        // film + crop window (src/librender/film.cpp:30-48): the camera carries the full film, the job renders the crop
        m_config.film_width = film->getSize().x; m_config.film_height = film->getSize().y;
        m_config.crop_offset_x = film->getCropOffset().x; m_config.crop_offset_y = film->getCropOffset().y;
        m_config.crop_width = size.x; m_config.crop_height = size.y;
        m_config.sample_count = (int32_t) sensor->getSampler()->getSampleCount();   // drmlt.cpp:400
        // the film's reconstruction filter, whatever plugin and parameters it was built from: after configure() it is a radius and
        // a 32-entry table (src/libcore/rfilter.cpp:37-55), read back through evalDiscretized (include/mitsuba/core/rfilter.h:76-77)
        const ReconstructionFilter *rf = film->getReconstructionFilter();
        m_config.rfilter = DR_FILTER_TABLE;
        m_config.filter_radius = (double) rf->getRadius();
        for (int i = 0; i < 32; ++i)
            m_config.filter_table[i] = (double) rf->evalDiscretized((i + (Float) 0.5) * rf->getRadius() / MTS_FILTER_RESOLUTION);
        check(dr_config_validate(&m_config));

        // ---- flatten Scene -> dr_scene_desc
        std::vector<float> P, N, UV;
        bool anyTexcoords = false;
        std::vector<uint32_t> I, triMat, triFlags;
        std::vector<int32_t> triEm;
        std::vector<dr_material> mats;
        std::vector<double> roughTables;           // DR_ROUGH_TABLE_DOUBLES per roughplastic material
        FlatTextures flatTextures;                 // bitmap textures bound to colour parameters (dr_texture)
        std::map<const BSDF *, uint32_t> materialOf;
        std::vector<dr_emitter> ems;
        bool anyNormals = false;
        // triangle meshes as they are; analytic shapes (rectangle, sphere, disk, cylinder, heightfield: SURVEY 8f rank 4) through the
        // shape's own tessellation (Shape::createTriMesh, include/mitsuba/render/shape.h:243 -- what the reference's preview uses)
        std::vector<ref<TriMesh> > tessellated;
        std::vector<std::pair<const TriMesh *, const Shape *> > meshes;          // geometry, owner of BSDF / emitter
        const ref_vector<Shape> &shapes = scene->getShapes();
        for (size_t si = 0; si < shapes.size(); ++si) {
            const Shape *shape = shapes[si].get();
            if (shape->getClass()->derivesFrom(MTS_CLASS(TriMesh))) { meshes.push_back(std::make_pair(static_cast<const TriMesh *>(shape), shape)); continue; }
            ref<TriMesh> tm = const_cast<Shape *>(shape)->createTriMesh();          // NotImplemented for shapes without a tessellation
            Log(EWarn, "Shape \"%s\" (%s) is tessellated into %u triangles for the GPU path", shape->getName().c_str(),
                shape->getClass()->getName().c_str(), (unsigned) tm->getTriangleCount());
            tessellated.push_back(tm);
            meshes.push_back(std::make_pair(tm.get(), shape));
        }
        for (size_t mi = 0; mi < meshes.size(); ++mi) {
            const TriMesh *mesh = meshes[mi].first;
            const Shape *owner = meshes[mi].second;
            // texture coordinates and UV tangents (skdtree.h:373-405; every mesh with texture coordinates has tangents, trimesh.cpp:400-402):
            // its.uv and the shading frame of the GPU path then agree with the host's
            const Point2 *tex = mesh->getVertexTexcoords();
            const bool tangents = tex && mesh->getUVTangents() != NULL;
            anyTexcoords |= tex != NULL;
            // one material per distinct BSDF object, however many meshes share it (a roughplastic's table costs a data-file reduction)
            uint32_t matIndex;
            std::map<const BSDF *, uint32_t>::const_iterator seen = materialOf.find(owner->getBSDF());
            if (owner->getBSDF() && seen != materialOf.end()) matIndex = seen->second;
            else {
                dr_material mat; std::string why;
                if (!owner->getBSDF() || !flattenBSDF(owner->getBSDF(), mat, why, roughTables, flatTextures)) Log(EError, "Mesh \"%s\": %s", owner->getName().c_str(), why.c_str());
                matIndex = (uint32_t) mats.size();
                mats.push_back(mat);
                materialOf[owner->getBSDF()] = matIndex;
            }
            const uint32_t base = (uint32_t) (P.size() / 3), firstTri = (uint32_t) triMat.size();
            const Point *pos = mesh->getVertexPositions();
            const Normal *nrm = mesh->getVertexNormals();
            for (size_t v = 0; v < mesh->getVertexCount(); ++v) {
                P.push_back((float) pos[v].x); P.push_back((float) pos[v].y); P.push_back((float) pos[v].z);
                N.push_back(nrm ? (float) nrm[v].x : 0.f); N.push_back(nrm ? (float) nrm[v].y : 0.f); N.push_back(nrm ? (float) nrm[v].z : 0.f);
                UV.push_back(tex ? (float) tex[v].x : 0.f); UV.push_back(tex ? (float) tex[v].y : 0.f);
            }
            anyNormals |= nrm != NULL;
            int32_t em = -1;
            if (owner->isEmitter()) {
                const Emitter *e = owner->getEmitter();
                if (e->getClass()->getName() != "AreaLight") Log(EError, "Only area emitters are supported (pathsampler.cpp:65-71)");
                dr_emitter de;
                de.first_tri = firstTri; de.n_tris = (uint32_t) mesh->getTriangleCount();
                toRGB(e->getProperties().getSpectrum("radiance", Spectrum::getD65()), de.radiance);   // area.cpp:77
                de.sampling_weight = (float) e->getSamplingWeight();
                em = (int32_t) ems.size();
                ems.push_back(de);
            }
            const Triangle *tri = mesh->getTriangles();
            for (size_t t = 0; t < mesh->getTriangleCount(); ++t) {
                for (int k = 0; k < 3; ++k) I.push_back(base + tri[t].idx[k]);
                triMat.push_back(matIndex); triEm.push_back(em); triFlags.push_back((nrm ? DR_TRI_SMOOTH : 0u) | (tangents ? DR_TRI_UV_TANGENTS : 0u) | (tex ? 0u : DR_TRI_NO_TEXCOORDS));
            }
        }
        if (scene->getEmitters().size() != ems.size()) Log(EError, "Only area emitters attached to triangle meshes are supported");
        const PerspectiveCamera *cam = dynamic_cast<const PerspectiveCamera *>(sensor.get());
        if (!cam || sensor->getClass()->getName() != "PerspectiveCameraImpl") Log(EError, "Only the perspective (pinhole) sensor is supported");
        dr_scene_desc desc;
        memset(&desc, 0, sizeof(desc));
        desc.n_vertices = (uint32_t) (P.size() / 3); desc.n_triangles = (uint32_t) triMat.size();
        desc.n_materials = (uint32_t) mats.size(); desc.n_emitters = (uint32_t) ems.size();
        desc.positions = P.data(); desc.normals = anyNormals ? N.data() : NULL; desc.indices = I.data();
        desc.tri_material = triMat.data(); desc.tri_emitter = triEm.data(); desc.tri_flags = triFlags.data();
        desc.materials = mats.data(); desc.emitters = ems.empty() ? NULL : ems.data();
        desc.texcoords = anyTexcoords ? UV.data() : NULL;
        for (size_t t = 0; t < flatTextures.textures.size(); ++t) flatTextures.textures[t].texels = flatTextures.texels[t].data();
        desc.textures = flatTextures.textures.empty() ? NULL : flatTextures.textures.data(); desc.n_textures = (uint32_t) flatTextures.textures.size();
        desc.rough_tables = roughTables.empty() ? NULL : roughTables.data(); desc.n_rough_tables = (uint32_t) (roughTables.size() / DR_ROUGH_TABLE_DOUBLES);
        if (!cam->getWorldTransform()->isStatic()) Log(EError, "An animated sensor transform (motion blur) is not supported by the B200 plugin");
        const Matrix4x4 &tw = cam->getWorldTransform()->eval(0).getMatrix();
        for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) desc.camera.to_world[4 * r + c] = (float) tw(r, c);
        desc.camera.xfov_deg = (float) cam->getXFov();
        desc.camera.near_clip = (float) cam->getNearClip(); desc.camera.far_clip = (float) cam->getFarClip();
        desc.camera.film_width = film->getSize().x; desc.camera.film_height = film->getSize().y;

        // DRMLT_DUMP_SCENE=<file>: the flattened description as raw arrays (tests compare it with what the scene was built from)
        if (const char *dump = getenv("DRMLT_DUMP_SCENE")) {
            FILE *f = fopen(dump, "wb");
            if (f) {
                const uint32_t head[6] = { desc.n_vertices, desc.n_triangles, desc.n_materials, desc.n_emitters, desc.n_textures, desc.texcoords ? 1u : 0u };
                fwrite(head, sizeof(head), 1, f);
                fwrite(desc.positions, sizeof(float), 3 * (size_t) desc.n_vertices, f);
                if (desc.texcoords) fwrite(desc.texcoords, sizeof(float), 2 * (size_t) desc.n_vertices, f);
                fwrite(desc.indices, sizeof(uint32_t), 3 * (size_t) desc.n_triangles, f);
                fwrite(desc.tri_material, sizeof(uint32_t), desc.n_triangles, f);
                fwrite(desc.tri_flags, sizeof(uint32_t), desc.n_triangles, f);
                fwrite(desc.materials, sizeof(dr_material), desc.n_materials, f);
                for (uint32_t t = 0; t < desc.n_textures; ++t) {
                    const dr_texture &tx = desc.textures[t];
                    const uint32_t th[5] = { tx.width, tx.height, tx.wrap_u, tx.wrap_v, tx.nearest };
                    fwrite(th, sizeof(th), 1, f);
                    fwrite(tx.uv_scale, sizeof(double), 2, f); fwrite(tx.uv_offset, sizeof(double), 2, f);
                    fwrite(tx.texels, sizeof(float), 3 * (size_t) tx.width * tx.height, f);
                }
                const uint32_t nrt = desc.n_rough_tables;
                fwrite(&nrt, sizeof(nrt), 1, f);
                if (nrt) fwrite(desc.rough_tables, sizeof(double), (size_t) nrt * DR_ROUGH_TABLE_DOUBLES, f);
                const uint32_t hasN = desc.normals ? 1u : 0u;
                fwrite(&hasN, sizeof(hasN), 1, f);
                if (hasN) fwrite(desc.normals, sizeof(float), 3 * (size_t) desc.n_vertices, f);
                fwrite(desc.tri_emitter, sizeof(int32_t), desc.n_triangles, f);
                if (desc.n_emitters) fwrite(desc.emitters, sizeof(dr_emitter), desc.n_emitters, f);
                fwrite(&desc.camera, sizeof(dr_camera), 1, f);
                const uint32_t cfgBytes = (uint32_t) sizeof(dr_config);
                fwrite(&cfgBytes, sizeof(cfgBytes), 1, f);
                fwrite(&m_config, sizeof(dr_config), 1, f);
                fclose(f);
            }
        }
        // DRMLT_DEVICE=<g>: the GPU of the job; DRMLT_DEVICES=<g0>,<g1>,...: several GPUs of this node -- chains sharded across them,
        // b all-reduced and the films reduced with NCCL inside the library (dr_render_multi)
        std::vector<int> devices;
        if (const char *list = getenv("DRMLT_DEVICES")) {
            std::istringstream is(list);
            std::string tok;
            while (std::getline(is, tok, ',')) if (!tok.empty()) devices.push_back(atoi(tok.c_str()));
        }
        if (devices.empty()) { const char *dev = getenv("DRMLT_DEVICE"); devices.push_back(dev ? atoi(dev) : 0); }
        dr_scene created = NULL;
        // The plugin renders a scene once, so the BVH build is on the critical path (as the kd-tree build of Scene::initialize is
        // for the reference, scene.cpp:289-356): build it on the device (~10 ms per million triangles instead of ~0.3-0.6 s on the
        // host, for a ~6 % slower traversal).  DRMLT_BVH=host keeps the host's binned-SAH build.
        const char *builder = getenv("DRMLT_BVH");
        check(dr_scene_create_ex(&desc, devices[0], builder && !strcmp(builder, "host") ? DR_SCENE_BVH_HOST : DR_SCENE_BVH_GPU, &created));
        { LockGuard lock(m_mutex); m_scene = created; }

        // ---- render on the GPU, hand the developed image to the film (drmlt_proc.cpp:850-853)
        std::vector<float> image((size_t) size.x * size.y * 3);
        dr_stats st;
        // interactive jobs and `mitsuba -r` see partial results: develop + signalRefresh every <= 2 s (drmlt_proc.cpp:856-867);
        // Scene::flush (scene.cpp:468-511) then dumps whatever bitmap the film holds, plus _time.csv / _stats.txt
        RefreshCtx ctx = { film.get(), queue, job, size };
        dr_status status;
        std::vector<dr_scene> replicas(1, created);
        if (devices.size() > 1) {
            for (size_t g = 1; g < devices.size(); ++g) {
                dr_scene r = NULL;
                const dr_status cs = dr_scene_clone(created, devices[g], &r);
                if (cs != DR_OK) { for (size_t k = 1; k < replicas.size(); ++k) dr_scene_destroy(replicas[k]); check(cs); }
                replicas.push_back(r);
            }
            status = dr_render_multi(replicas.data(), (int32_t) replicas.size(), &m_config, image.data(), &st);
        } else {
            // (`mitsuba -r` marks the job interactive: mitsuba.cpp:395; non-interactive jobs only develop at the end, drmlt.cpp:608)
            status = dr_render_progressive(m_scene, &m_config, image.data(), &st, job->isInteractive() ? 2.0 : 0.0, &DR_CLASS::refresh, &ctx);
        }
        { LockGuard lock(m_mutex); m_scene = NULL; }            // cancel() from another thread never sees a scene being destroyed
        for (size_t k = 0; k < replicas.size(); ++k) dr_scene_destroy(replicas[k]);
        if (status == DR_ERR_CANCELLED) return false;
        check(status);
        refresh(image.data(), size.x, size.y, 0.0, &st, &ctx);
        publishStats(st);
        // same figures as the reference's StatsCounters (drmlt_proc.cpp:34-49)
        Log(EInfo, "Normalization factor b = %f; %llu mutations, first stage accepted %.2f %%, second stage %.2f %%, %.1f ms on the GPU",
            st.luminance, (unsigned long long) st.mutations, st.first_base ? 100.0 * st.first_accept / st.first_base : 0.0,
            st.second_base ? 100.0 * st.second_accept / st.second_base : 0.0, st.total_ms);
        return true;
    }

    struct RefreshCtx { Film *film; RenderQueue *queue; const RenderJob *job; Vector2i size; };
    // DRMLTProcess::develop, last lines (drmlt_proc.cpp:850-853): hand the developed image to the film and signal a refresh
    static int refresh(const float *image, int32_t w, int32_t h, double, const dr_stats *, void *user) {
        RefreshCtx *ctx = static_cast<RefreshCtx *>(user);
        ref<Bitmap> bitmap = new Bitmap(Bitmap::ESpectrum, Bitmap::EFloat, Vector2i(w, h));
        Spectrum *target = (Spectrum *) bitmap->getData();
        for (size_t i = 0; i < (size_t) w * h; ++i)
            target[i].fromLinearRGB(image[3 * i], image[3 * i + 1], image[3 * i + 2]);
        ctx->film->setBitmap(bitmap);
        ctx->queue->signalRefresh(ctx->job);
        return 0;
    }

    void cancel() { LockGuard lock(m_mutex); if (m_scene) dr_cancel(m_scene); }   // Integrator::cancel (drmlt.cpp:386-391), any thread

    MTS_DECLARE_CLASS()
private:
    void check(dr_status st) const { if (st != DR_OK) Log(EError, "%s", dr_last_error()); }   // EError throws (renderjob.cpp:110-114)
    dr_config m_config;
    dr_scene m_scene;
    ref<Mutex> m_mutex;
};

// (one more macro level, so that the class registers under its expanded name -- "DRMLT" / "PSSMLT", the reference's own)
#define DR_IMPLEMENT(cls) MTS_IMPLEMENT_CLASS_S(cls, false, Integrator)
#define DR_EXPORT(cls, descr) MTS_EXPORT_PLUGIN(cls, descr)
DR_IMPLEMENT(DR_CLASS)
DR_EXPORT(DR_CLASS, "B200 " DR_NAME " integrator");
MTS_NAMESPACE_END
