"""Helpers for the tests that run the REFERENCE'S OWN renderer (oracle/_ref/pbrt_ref_cpu / pbrt_ref_b200: pbrt-v3 +
StatPathIntegrator compiled unmodified by oracle/Makefile) on a small scene of ours and read back its PFM dumps."""
import os
import subprocess

import numpy as np

from statmc_b200 import pfm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PBRT_CPU = os.path.join(ROOT, "oracle", "_ref", "pbrt_ref_cpu")
PBRT_B200 = os.path.join(ROOT, "oracle", "_ref", "pbrt_ref_b200")

# Our own scene (pbrt-v3 syntax): matte floor and wall, a rough-metal, a plastic and a glass sphere, one large and one
# small bright light -- smooth regions, edges, glossy highlights and fireflies at 16 spp.  Integrator block = the
# reference's scenes/render-denoise.pbrt parameters (statpath.cpp:902-1020 reads them).
SCENE = """
Integrator "statpath"
  "integer maxdepth" [16] "bool expiterations" ["true"] "integer iterations" [{iterations}]
  "integer trackedbounces" [{trackedbounces}] "bool multichannelstats" ["{multichannelstats}"]
  "bool denoiseimage" ["{denoiseimage}"] "bool acrr" ["{acrr}"] "bool smis" ["{smis}"]
  "bool calcstats" ["false"] "bool calcprodenstats" ["{calcprodenstats}"] "bool calcmoonstats" ["false"] "bool calcgbuffers" ["false"]
  "bool calcitstats" ["false"]
  "float filtersd" [{sd}] "integer filterradius" [{radius}]
  "string filterbuffers" ["albedo" "normal"] "float filterbuffersds" [0.02 0.1]
  "string outputregex" ["{outputregex}"]
Sampler "random" "integer pixelsamples" [4]
LookAt 0 3.2 8.5  0 0.7 0  0 1 0
Camera "perspective" "float fov" [36]
Film "image" "integer xresolution" [{width}] "integer yresolution" [{height}] "string filename" ["{stem}.pfm"]
WorldBegin
  AttributeBegin
    AreaLightSource "diffuse" "rgb L" [16 15 13]
    Translate -2.5 5 2
    Shape "sphere" "float radius" [0.7]
  AttributeEnd
  AttributeBegin
    AreaLightSource "diffuse" "rgb L" [900 850 800]
    Translate 3 3.5 -1
    Shape "sphere" "float radius" [0.06]
  AttributeEnd
  Material "matte" "rgb Kd" [0.6 0.55 0.5]
  Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [-12 0 -12  12 0 -12  12 0 12  -12 0 12]
  Material "matte" "rgb Kd" [0.7 0.25 0.2]
  Shape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point P" [-12 0 -4  12 0 -4  12 9 -4  -12 9 -4]
  AttributeBegin
    Material "metal" "float roughness" [0.08]
    Translate -1.5 1 0
    Shape "sphere" "float radius" [1]
  AttributeEnd
  AttributeBegin
    Material "plastic" "rgb Kd" [0.2 0.3 0.8] "rgb Ks" [0.4 0.4 0.4] "float roughness" [0.1]
    Translate 1.4 0.8 0.4
    Shape "sphere" "float radius" [0.8]
  AttributeEnd
  AttributeBegin
    Material "glass"
    Translate 0 0.5 2.2
    Shape "sphere" "float radius" [0.5]
  AttributeEnd
WorldEnd
"""


def write_scene(directory, width=96, height=64, radius=8, sd=4.0, iterations=3, trackedbounces=0, multichannelstats=True,
                denoiseimage=True, acrr=False, smis=False, calcprodenstats=False, outputregex=".*"):
    """scenes/render-denoise.pbrt by default; acrr.pbrt = trackedbounces 5, multichannelstats / denoiseimage false, acrr true;
    smis.pbrt = trackedbounces 6, multichannelstats / denoiseimage false, smis true; render-for-proden.pbrt = denoiseimage
    false, calcprodenstats true."""
    stem = os.path.join(str(directory), "smc")
    path = os.path.join(str(directory), "scene.pbrt")
    b = lambda v: "true" if v else "false"
    with open(path, "w") as f:
        f.write(SCENE.format(width=width, height=height, radius=radius, sd=sd, iterations=iterations, stem=stem,
                             trackedbounces=trackedbounces, multichannelstats=b(multichannelstats),
                             denoiseimage=b(denoiseimage), acrr=b(acrr), smis=b(smis), calcprodenstats=b(calcprodenstats),
                             outputregex=outputregex))
    return path, stem


def read_planes(stem, spp, type_index, bounce, channels):
    """Statistic planes of type `type_index`, bounce `bounce` (scalar planes come back H x W)."""
    pre = "%s-%d-t%d-b%d-" % (stem, spp, type_index, bounce)
    rd = lambda k: pfm.read(pre + k + ".pfm")
    sq = (lambda a: a[..., 0] if a.ndim == 3 and channels == 1 else a)
    out = {k.replace("-", "_"): sq(rd(k)) for k in ("mean", "m2", "m3", "film-mean", "film-mean-f")}
    out["n"] = pfm.read(pre + "n.pfm", np.int32)
    return out


def run_pbrt(exe, scene, *flags, env=None, nthreads=8):
    e = dict(os.environ)
    if env:
        e.update(env)
    p = subprocess.run([exe, "--nthreads", str(nthreads), *flags, scene], capture_output=True, text=True, timeout=900, env=e,
                       cwd=os.path.dirname(scene))
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    return p


def read_dump(stem, spp):
    """The planes of iteration `spp` as the dict the oracle / denoise_host take (+ film_f, film_mean, film_m2)."""
    pre = "%s-%d-" % (stem, spp)
    rd = lambda k: pfm.read(pre + k + ".pfm")
    return {"n": pfm.read(pre + "t0-b0-n.pfm", np.int32), "mean": rd("t0-b0-mean"), "m2": rd("t0-b0-m2"),
            "m3": rd("t0-b0-m3"), "film_mean": rd("t0-b0-film-mean"), "film_m2": rd("t0-b0-film-m2"), "film": rd("film"),
            "normal": rd("t1-b0-film-mean"), "albedo": rd("t2-b0-film-mean"), "film_f": rd("film-f")}


def reference_factors(sd, normal_sd=0.1, albedo_sd=0.02):
    """-.5f / (sd * sd) evaluated in float32 like the reference (estimator.h:259, estimator.cpp:16): the double-precision
    value rounds differently for sd = 0.1 (-49.999996 instead of -50)."""
    f = np.float32
    g = lambda s: float(f(-0.5) / (f(s) * f(s)))
    return g(sd), [g(normal_sd), g(albedo_sd)]
