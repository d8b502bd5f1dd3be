"""Generates tests/golden/kdtree_golden.npz by running the REFERENCE's own kdtree.py
(/root/reference/kdtree.py, imported -- never copied) on seeded inputs.

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_kdtree_golden.py
Two shims are applied before import because the file is Python 2 / NumPy 1 code:
builtins.xrange (kdtree.py:47,136) and numpy.Inf (kdtree.py:22-23).
"""
import builtins
import importlib.util
import os
import sys

import numpy as np

REF = "/root/reference/kdtree.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kdtree_golden.npz")


def load_reference():
    builtins.xrange = range
    if not hasattr(np, "Inf"):
        np.Inf = np.inf
    spec = importlib.util.spec_from_file_location("ref_kdtree", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def cases():
    rng = np.random.RandomState(77)
    out = {}
    out["uniform_2000_b50"] = (rng.uniform([25, -124], [49, -67], size=(2000, 2)), 50)
    out["uniform_5685_b50"] = (rng.uniform([25, -124], [49, -67], size=(5685, 2)), 50)
    # heavy duplicates: users sharing a city centroid, as in real geotagged data
    cent = rng.uniform([25, -124], [49, -67], size=(40, 2))
    out["duplicates_3000_b64"] = (cent[rng.randint(0, 40, 3000)], 64)
    # mixture of exact duplicates and jittered points
    pts = cent[rng.randint(0, 40, 4000)].copy()
    jit = rng.rand(4000) < 0.5
    pts[jit] += rng.standard_normal((int(jit.sum()), 2)) * 0.05
    out["mixed_4000_b100"] = (pts, 100)
    # ties between lat and lon widths (lat must win), grid points -> medians hit existing values
    g = np.stack(np.meshgrid(np.arange(32.0), np.arange(32.0)), -1).reshape(-1, 2)
    out["grid_1024_b10"] = (g[rng.permutation(len(g))], 10)
    # zero width in one dimension
    z = np.stack([np.full(500, 37.5), rng.uniform(-120, -70, 500)], 1)
    out["zero_lat_width_500_b20"] = (z, 20)
    # everything identical: root stays an over-full leaf (kdtree.py:90-91)
    out["identical_300_b10"] = (np.tile([[10.0, 20.0]], (300, 1)), 10)
    # median equals max (most points at the max value) -> split value falls back to min (:94-95)
    m = np.stack([np.concatenate([np.full(70, 5.0), rng.uniform(0, 5, 30)]), np.zeros(100)], 1)
    out["median_eq_max_100_b10"] = (m[rng.permutation(100)], 10)
    # n <= bucket: single leaf; bucket 1; "two Melbournes" toy (lang2loc.py:530-569 style)
    out["small_30_b50"] = (rng.uniform(0, 1, size=(30, 2)), 50)
    out["bucket1_257"] = (rng.uniform(0, 1, size=(257, 2)), 1)
    mel = np.concatenate([rng.standard_normal((400, 2)) * 0.3 + [-37.8, 144.9],
                          rng.standard_normal((400, 2)) * 0.3 + [28.1, -80.6]])
    out["two_melbournes_800_b25"] = (mel[rng.permutation(800)], 25)
    # three dimensions
    out["dims3_1000_b30"] = (rng.standard_normal((1000, 3)), 30)
    return out


def main():
    ref = load_reference()
    blob = {}
    for name, (pts, bucket) in cases().items():
        c = ref.KDTreeClustering(bucket_size=bucket)
        c.fit(pts)
        labels = np.asarray(c.get_clusters(), dtype=np.int64)
        blob[name + "/points"] = pts.astype(np.float64)
        blob[name + "/bucket"] = np.int64(bucket)
        blob[name + "/labels"] = labels
        blob[name + "/n_clusters"] = np.int64(c.num_clusters)
        print("%-28s n=%5d bucket=%4d -> %4d clusters" % (name, len(pts), bucket, c.num_clusters))
    np.savez_compressed(OUT, **blob)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    sys.setrecursionlimit(100000)
    main()
