"""One bounded end-to-end run of the REFERENCE'S OWN integrator (oracle/_ref/libref_path.so: DRMLT::render through a
RenderJob on the host cores) on a bench scene; prints one JSON line.  bench.py runs this in a subprocess for its
`cpu_baseline` leg and its `--impl reference` arm, so that the reference's scheduler threads and statics stay out of
the CUDA process.  Test / measurement infrastructure only.

mutations/s = crop area x mutations per pixel / the job's render time (bootstrap included), which is how the reference's
own log states it (renderjob.cpp:108)."""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C5", help="BASELINE.json configuration (bench.py CONFIGS)")
    ap.add_argument("--spp", type=int, default=8)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--film", default="", help="WxH (door scene only; default: the scene's own film)")
    ap.add_argument("--out-image", default="", dest="out_image", help="write the developed image to this .npy file")
    args = ap.parse_args()
    import ref_path_cases as RP
    import bench
    from drmlt_mitsuba_b200 import scenes
    scene_name, kw, params = bench.CONFIGS[args.config]
    kw = dict(kw)
    if args.film:
        kw["film"] = tuple(int(x) for x in args.film.split("x"))
    data = scenes.SCENES[scene_name](**kw)
    W, H = data.film
    # the reference hands out one seed per work unit of 1e5 mutations (2e5 for technique=path) and floor(workUnits / cores) seeds
    # per init thread (drmlt.cpp:430-450, 498-546): fewer work units than cores would render nothing
    unit = 200000 if params["technique"] == "path" else 100000
    spp = max(args.spp, -(-(args.threads + 1) * unit // (W * H)))
    params = dict(params)
    lib = C.CDLL(RP.REF_PATH)
    img, sec, scene_sec, stats = RP.run_render_ref(lib, params, spp, threads=args.threads, data=data)
    out = {"mutations_per_s": W * H * spp / sec, "render_s": sec, "scene_build_s": scene_sec, "spp": spp, "threads": args.threads,
           "mutations": W * H * spp, "mean_luminance": float(RP.luminance(img).mean()), "stats_percent": stats}
    if args.out_image:
        import numpy as np
        np.save(args.out_image, img)
    sys.stdout.write("\nREF_SAMPLE " + json.dumps(out) + "\n")
    sys.stdout.flush()
    os._exit(0)          # the reference's worker threads are not torn down


if __name__ == "__main__":
    main()
