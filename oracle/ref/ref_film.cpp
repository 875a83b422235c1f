/* oracle/_ref film driver -- TEST INFRASTRUCTURE ONLY.
 * The reference's own ImageBlock::put (include/mitsuba/render/imageblock.h:149-196) with its own reconstruction filter plugin
 * (rfilters/gaussian.cpp, rfilters/box.cpp; the 32-entry table of ReconstructionFilter::configure, librender/rfilter.cpp), on a
 * block created the way DRMLTProcess::createWorkResult creates it (drmlt_proc.cpp:80-81).  Returns the block's interior (the
 * border pixels are what DRMLTProcess::processResult clips when it merges a block into the accumulation buffer). */
#include <mitsuba/mitsuba.h>
#include <mitsuba/core/plugin.h>
#include <mitsuba/core/bitmap.h>
#include <mitsuba/render/imageblock.h>
#include <mitsuba/core/rfilter.h>

using namespace mitsuba;

extern "C" void ref_init();     // ref_path.cpp

/* filter: 0 box, 1 gaussian, 3 tent, 4 mitchell, 5 catmullrom, 6 lanczos (3.. = dr_filter); radius_out / table_out (optional):
 * what the plugin boils down to after configure() -- getRadius() and the 32 discretised values read back through evalDiscretized */
extern "C" int ref_splat(int w, int h, int filter, const float *pos, const float *rgb, int64_t n, double *film_out /* [h][w][3] */,
                         int *accepted /* [n] */, double *radius_out, double *table_out /* [32] */) {
    try {
        ref_init();
        const char *names[] = { "box", "gaussian", "gaussian", "tent", "mitchell", "catmullrom", "lanczos" };
        ref<ReconstructionFilter> rf = static_cast<ReconstructionFilter *>(
            PluginManager::getInstance()->createObject(MTS_CLASS(ReconstructionFilter), Properties(names[filter])));
        rf->configure();
        if (radius_out) *radius_out = rf->getRadius();
        if (table_out) for (int i = 0; i < 32; ++i) table_out[i] = rf->evalDiscretized((i + (Float) 0.5) * rf->getRadius() / MTS_FILTER_RESOLUTION);
        ref<ImageBlock> block = new ImageBlock(Bitmap::ESpectrum, Vector2i(w, h), rf.get());
        block->clear();
        for (int64_t i = 0; i < n; ++i) {
            Spectrum value; value.fromLinearRGB(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]);
            accepted[i] = block->put(Point2(pos[2 * i], pos[2 * i + 1]), &value[0]) ? 1 : 0;
        }
        const Bitmap *bmp = block->getBitmap();
        const int border = block->getBorderSize(), bw = bmp->getSize().x, ch = bmp->getChannelCount();
        if (ch != 3) return 2;
        const Float *data = bmp->getFloatData();
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x)
                for (int k = 0; k < 3; ++k)
                    film_out[((size_t) y * w + x) * 3 + k] = data[((size_t) (y + border) * bw + (x + border)) * ch + k];
        return 0;
    } catch (const std::exception &e) { fprintf(stderr, "oracle/_ref: %s\n", e.what()); return 1; }
}
