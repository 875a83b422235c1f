/* render_box.c -- the C ABI from a plain C99 host (no Python, no C++, no Mitsuba): a closed box lit by a ceiling panel, with a textured
 * floor, rendered with `drmlt type=orbital technique=mmlt` -- the calls a host makes are the ones the Mitsuba-side plugin makes
 * (drmlt-mitsuba_b200/shim/mts_plugin.cpp): dr_config_default / dr_config_set (the reference's own `-D key=value` names,
 * drmlt.cpp:178-351), dr_scene_create, dr_render, dr_scene_destroy.
 *
 *   gcc -std=c99 -O2 -Iinclude examples/render_box.c -Ldrmlt-mitsuba_b200/csrc -ldrmlt_b200 -Wl,-rpath,$PWD/drmlt-mitsuba_b200/csrc -lm -o /tmp/render_box
 *   /tmp/render_box out.ppm [mutations per pixel]
 *
 * Exit code: 0 = rendered, 2 = no CUDA device (the library has no CPU fallback and says so), 1 = any other error.
 */
#include <drmlt_b200.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MAX_V 64
#define MAX_T 64
static float P[3 * MAX_V], UV[2 * MAX_V];
static uint32_t I[3 * MAX_T], tri_material[MAX_T], tri_flags[MAX_T];
static int32_t tri_emitter[MAX_T];
static uint32_t nv = 0, nt = 0;

/* quad a b c d, counter-clockwise seen from its front side; uv (0,0) (1,0) (1,1) (0,1) */
static void quad(const float *a, const float *b, const float *c, const float *d, uint32_t material, int32_t emitter, int textured) {
    const float *q[4] = { a, b, c, d };
    const float uv[8] = { 0, 0, 1, 0, 1, 1, 0, 1 };
    for (int k = 0; k < 4; ++k) {
        memcpy(P + 3 * (nv + k), q[k], 3 * sizeof(float));
        UV[2 * (nv + k)] = uv[2 * k]; UV[2 * (nv + k) + 1] = uv[2 * k + 1];
    }
    const uint32_t idx[6] = { nv, nv + 1, nv + 2, nv, nv + 2, nv + 3 };
    for (int k = 0; k < 2; ++k) {
        memcpy(I + 3 * nt, idx + 3 * k, 3 * sizeof(uint32_t));
        tri_material[nt] = material; tri_emitter[nt] = emitter;
        /* a mesh with texture coordinates carries UV tangents in the reference (trimesh.cpp:400-402); the others have none */
        tri_flags[nt] = textured ? DR_TRI_UV_TANGENTS : DR_TRI_NO_TEXCOORDS;
        ++nt;
    }
    nv += 4;
}

static int fail(const char *what, dr_status st) {
    fprintf(stderr, "%s: status %d: %s\n", what, (int) st, dr_last_error());
    return st == DR_ERR_NO_DEVICE ? 2 : 1;
}

int main(int argc, char **argv) {
    const char *out = argc > 1 ? argv[1] : "render_box.ppm";
    const char *spp = argc > 2 ? argv[2] : "64";
    enum { W = 128, H = 128, TW = 8, TH = 8 };

    /* materials: white, red, green diffuse walls; the floor's reflectance is a bitmap texture (an 8 x 8 checkerboard, bilinear, repeated 3 x) */
    static float texels[3 * TW * TH];
    for (int y = 0; y < TH; ++y)
        for (int x = 0; x < TW; ++x) {
            const float v = ((x / 2 + y / 2) & 1) ? 0.75f : 0.15f;
            texels[3 * (y * TW + x)] = v; texels[3 * (y * TW + x) + 1] = v * 0.9f; texels[3 * (y * TW + x) + 2] = v * 0.6f;
        }
    dr_texture tex;
    memset(&tex, 0, sizeof(tex));
    tex.width = TW; tex.height = TH; tex.texels = texels;
    tex.wrap_u = tex.wrap_v = DR_WRAP_REPEAT;
    tex.uv_scale[0] = tex.uv_scale[1] = 3.0; tex.uv_offset[0] = tex.uv_offset[1] = 0.0;

    dr_material mats[4];
    memset(mats, 0, sizeof(mats));
    const float refl[4][3] = { { 0.73f, 0.73f, 0.73f }, { 0.63f, 0.065f, 0.05f }, { 0.14f, 0.45f, 0.091f }, { 0.45f, 0.40f, 0.27f } /* the texture's average */ };
    for (int m = 0; m < 4; ++m) {
        mats[m].type = DR_BSDF_DIFFUSE;
        memcpy(mats[m].reflectance, refl[m], sizeof(refl[m]));
        mats[m].transmittance[0] = mats[m].transmittance[1] = mats[m].transmittance[2] = 1.f;
    }
    mats[3].flags |= DR_MAT_TEX_REFLECTANCE(0);

    const float a[3] = { -1, -1, 1 }, b[3] = { 1, -1, 1 }, c[3] = { 1, -1, -1 }, d[3] = { -1, -1, -1 };       /* floor corners */
    const float e[3] = { -1, 1, 1 }, f[3] = { 1, 1, 1 }, g[3] = { 1, 1, -1 }, h[3] = { -1, 1, -1 };           /* ceiling corners */
    quad(a, b, c, d, 3, -1, 1);          /* floor (+y), textured */
    quad(h, g, f, e, 0, -1, 0);          /* ceiling (-y) */
    quad(d, c, g, h, 0, -1, 0);          /* back wall (+z) */
    quad(a, d, h, e, 1, -1, 0);          /* left wall (+x), red */
    quad(c, b, f, g, 2, -1, 0);          /* right wall (-x), green */
    const float l0[3] = { -0.25f, 0.995f, -0.25f }, l1[3] = { 0.25f, 0.995f, -0.25f }, l2[3] = { 0.25f, 0.995f, 0.25f }, l3[3] = { -0.25f, 0.995f, 0.25f };
    dr_emitter light;
    memset(&light, 0, sizeof(light));
    light.first_tri = nt; light.n_tris = 2;
    light.radiance[0] = light.radiance[1] = light.radiance[2] = 15.f;
    light.sampling_weight = 1.f;
    quad(l0, l1, l2, l3, 0, 0, 0);       /* area light, facing down */

    dr_scene_desc desc;
    memset(&desc, 0, sizeof(desc));
    desc.n_vertices = nv; desc.n_triangles = nt; desc.n_materials = 4; desc.n_emitters = 1; desc.n_textures = 1;
    desc.positions = P; desc.texcoords = UV; desc.indices = I;
    desc.tri_material = tri_material; desc.tri_emitter = tri_emitter; desc.tri_flags = tri_flags;
    desc.materials = mats; desc.emitters = &light; desc.textures = &tex;
    /* pinhole camera at (0, 0, 3.9) looking down -z (Mitsuba's lookAt: left-handed frame; columns = left, up, direction, origin) */
    const float to_world[16] = { -1, 0, 0, 0,   0, 1, 0, 0,   0, 0, -1, 3.9f,   0, 0, 0, 1 };
    memcpy(desc.camera.to_world, to_world, sizeof(to_world));
    desc.camera.xfov_deg = 39.f; desc.camera.near_clip = 1e-2f; desc.camera.far_clip = 1e4f;
    desc.camera.film_width = W; desc.camera.film_height = H;

    /* the integrator's parameters, under the reference's names */
    dr_config cfg;
    dr_config_default(&cfg);
    const char *params[][2] = { { "integrator", "drmlt" }, { "technique", "mmlt" }, { "type", "orbital" }, { "maxDepth", "8" },
                                { "directSamples", "-1" }, { "sampleCount", spp }, { "seed", "1" } };
    for (size_t i = 0; i < sizeof(params) / sizeof(params[0]); ++i) {
        dr_status st = dr_config_set(&cfg, params[i][0], params[i][1]);
        if (st) return fail(params[i][0], st);
    }
    dr_status st = dr_config_validate(&cfg);
    if (st) return fail("dr_config_validate", st);

    dr_scene scene = NULL;
    st = dr_scene_create(&desc, 0, &scene);
    if (st) return fail("dr_scene_create", st);
    static float image[3 * W * H];
    dr_stats stats;
    st = dr_render(scene, &cfg, image, &stats);
    if (st) { int rc = fail("dr_render", st); dr_scene_destroy(scene); return rc; }
    dr_scene_destroy(scene);

    double mean = 0.0;                   /* mean luminance of the developed image = the normalisation b (drmlt_proc.cpp:813-854) */
    for (int i = 0; i < W * H; ++i) mean += 0.212671 * image[3 * i] + 0.715160 * image[3 * i + 1] + 0.072169 * image[3 * i + 2];
    mean /= (double) W * H;
    printf("RENDER_BOX mutations=%llu b=%.6f mean=%.6f accept=%.4f seconds=%.3f\n", (unsigned long long) stats.mutations, stats.luminance, mean,
           stats.accept_base ? (double) stats.accept / (double) stats.accept_base : 0.0, stats.total_ms * 1e-3);
    FILE *fp = fopen(out, "wb");
    if (!fp) { perror(out); return 1; }
    fprintf(fp, "P6\n%d %d\n255\n", W, H);
    for (int i = 0; i < 3 * W * H; ++i) {
        const float v = image[i] <= 0.0031308f ? 12.92f * image[i] : 1.055f * powf(image[i], 1.f / 2.4f) - 0.055f;      /* sRGB */
        fputc((int) (255.f * (v < 0.f ? 0.f : v > 1.f ? 1.f : v) + 0.5f), fp);
    }
    fclose(fp);
    return 0;
}
