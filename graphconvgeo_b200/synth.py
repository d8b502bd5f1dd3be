"""Seeded synthetic workloads shaped like the reference's datasets.

The reference trains on GEOTEXT / Twitter-US / Twitter-World dumps we do not have
(data.py:273-299 reads three gz TSVs).  These generators produce inputs with the
same *contract* as DataLoader's outputs:
  * a binary symmetric user-user mention graph without self loops
    (data.py:226-250, 302-375), power-law degrees;
  * X: CSR float32 TF-IDF rows, binary tf * smooth idf, L2-normalised
    (data.py:254-256, 378-397: btf=True, idf=True, norm='l2', dtype float32);
  * labels: k-d tree regions over training coordinates, dev/test assigned to the
    nearest region median under haversine (data.py:399-421).
Everything here is host-side NumPy; A_hat and the k-d tree come from libgcg.so's
host entry points (gcg_ahat_build_host, gcg_kdtree_fit_host).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp

# name -> shape parameters (BASELINE.json configs; SURVEY.md section 8 "G/U/W")
WORKLOADS = {
    # GEOTEXT-shaped: 9,475 users (5,685/1,895/1,895), ~9k vocab, bucket 50 -> 128 regions
    "geotext": dict(n_train=5685, n_dev=1895, n_test=1895, vocab=9000, terms=60, avg_deg=8,
                    bucket=50, hidden=300, n_cities=60),
    # Twitter-US-shaped: 450k users, bucket 2400 -> 256 regions
    "twitter-us": dict(n_train=430000, n_dev=10000, n_test=10000, vocab=250000, terms=150, avg_deg=20,
                       bucket=2400, hidden=600, n_cities=200),
    # Twitter-World-shaped: 1.4M users, bucket 2400 -> 1024 regions (distinct coordinates)
    "twitter-world": dict(n_train=1380000, n_dev=10000, n_test=10000, vocab=500000, terms=150, avg_deg=20,
                          bucket=2400, hidden=600, n_cities=400),
    # tiny case for unit tests / smoke
    "tiny": dict(n_train=600, n_dev=200, n_test=200, vocab=500, terms=20, avg_deg=6,
                 bucket=40, hidden=32, n_cities=8),
}


@dataclass
class Workload:
    name: str
    X: sp.csr_matrix              # [N, V] float32
    A_hat: sp.csr_matrix          # [N, N] float32, sym-normalised with self loops
    Y: np.ndarray                 # int64 region per node
    train_indices: np.ndarray     # int32
    dev_indices: np.ndarray
    test_indices: np.ndarray
    locs: np.ndarray              # float64 [N, 2] (lat, lon)
    medians: np.ndarray           # float64 [C, 2]
    hidden: int
    n_classes: int
    meta: dict = field(default_factory=dict)


# --------------------------------------------------------------------- graph
def powerlaw_graph(n, avg_deg, seed=77, alpha=1.5, communities=None, intra=0.8, max_deg=None):
    """Binary symmetric adjacency (CSR, no self loops) with Chung-Lu style power-law
    degrees: endpoint i ~ w (w = Pareto(alpha)+1), endpoint j uniform -- or, with
    ``communities`` (int array, one id per node), j is drawn inside i's community with
    probability ``intra`` (users mostly mention users of their own region)."""
    rng = np.random.RandomState(seed)
    m = int(n * avg_deg / 2)
    w = rng.pareto(alpha, size=n) + 1.0
    cdf = np.cumsum(w)
    cdf /= cdf[-1]
    src = np.searchsorted(cdf, rng.random_sample(m)).astype(np.int64)
    np.minimum(src, n - 1, out=src)
    dst = rng.randint(0, n, size=m).astype(np.int64)
    if communities is not None:
        communities = np.asarray(communities)
        order = np.argsort(communities, kind="stable")
        sorted_c = communities[order]
        starts = np.searchsorted(sorted_c, sorted_c, side="left")        # per sorted position
        ends = np.searchsorted(sorted_c, sorted_c, side="right")
        pos_of = np.empty(n, np.int64)
        pos_of[order] = np.arange(n)
        sp_ = pos_of[src]
        lo, hi = starts[sp_], ends[sp_]
        local = order[lo + (rng.random_sample(m) * (hi - lo)).astype(np.int64)]
        use_local = rng.random_sample(m) < intra
        dst = np.where(use_local, local, dst)
    keep = src != dst
    src, dst = src[keep], dst[keep]
    a = np.minimum(src, dst)
    b = np.maximum(src, dst)
    key = np.unique(a * n + b)
    a, b = key // n, key % n
    if max_deg is not None:      # celebrity-style cap (data.py:364-370 removes over-mentioned nodes)
        deg = np.bincount(np.concatenate([a, b]), minlength=n)
        ok = (deg[a] <= max_deg) & (deg[b] <= max_deg)
        a, b = a[ok], b[ok]
    rows = np.concatenate([a, b])
    cols = np.concatenate([b, a])
    adj = sp.csr_matrix((np.ones(len(rows), np.float64), (rows, cols)), shape=(n, n))
    adj.sort_indices()
    return adj


# --------------------------------------------------------------------- TF-IDF
LOCAL_TERM_FRACTION = 0.15      # share of a user's term draws that come from the vocabulary of the user's city
LOCAL_TERMS = 2048              # size of a city's own vocabulary slice


def _local_columns(ranks, city, vocab, xp):
    """term id of a 'local' draw: the city's own slice of the vocabulary (place names, dialect words --
    what makes text geolocation possible at all), Zipf-ranked inside the slice"""
    base = (city * 7919 * 64) % vocab
    return (base + (ranks % LOCAL_TERMS) * 17) % vocab


def tfidf_matrix(n, vocab, mean_terms, seed=77, zipf_s=1.1, city=None):
    """CSR float32 [n, vocab]: binary tf, smooth idf (ln((1+n)/(1+df))+1), rows L2-normalised
    -- sklearn TfidfVectorizer(binary=True, use_idf=True, norm='l2') as configured at
    data.py:254-256,386-389.  Row lengths ~ max(1, LogNormal) with the given mean; term ids
    ~ truncated Zipf(zipf_s), de-duplicated per row.  With ``city`` (one id per row) a fixed share of the
    draws is redirected to the city's own vocabulary slice, so the text carries location signal."""
    rng = np.random.RandomState(seed + 1)
    sigma = 0.6
    mu = np.log(mean_terms) - 0.5 * sigma * sigma
    lens = np.maximum(1, rng.lognormal(mu, sigma, size=n)).astype(np.int64)
    lens = np.minimum(lens, vocab // 2)
    total = int(lens.sum())
    ranks = np.arange(1, vocab + 1, dtype=np.float64)
    p = ranks ** (-zipf_s)
    cdf = np.cumsum(p)
    cdf /= cdf[-1]
    cols = np.searchsorted(cdf, rng.random_sample(total)).astype(np.int64)
    np.minimum(cols, vocab - 1, out=cols)
    rows = np.repeat(np.arange(n, dtype=np.int64), lens)
    if city is not None:
        local = rng.random_sample(total) < LOCAL_TERM_FRACTION
        cols = np.where(local, _local_columns(cols, np.asarray(city, dtype=np.int64)[rows], vocab, np), cols)
    key = np.unique(rows * vocab + cols)
    rows, cols = key // vocab, key % vocab
    df = np.bincount(cols, minlength=vocab).astype(np.float64)
    idf = np.log((1.0 + n) / (1.0 + df)) + 1.0
    vals = idf[cols]
    sq = np.bincount(rows, weights=vals * vals, minlength=n)
    vals = vals / np.sqrt(sq[rows])
    X = sp.csr_matrix((vals.astype(np.float32), (rows, cols)), shape=(n, vocab))
    X.sort_indices()
    return X


# --------------------------------------------------------------------- labels
def city_locations(n, n_cities, seed=77):
    """lat/lon float64 from a seeded mixture of Gaussians ("cities"); returns (locs, city id)."""
    rng = np.random.RandomState(seed + 2)
    centers = np.stack([rng.uniform(25.0, 49.0, n_cities), rng.uniform(-124.0, -67.0, n_cities)], axis=1)
    weights = rng.pareto(1.2, n_cities) + 1.0
    weights /= weights.sum()
    city = rng.choice(n_cities, size=n, p=weights)
    spread = rng.uniform(0.05, 0.6, n_cities)
    locs = centers[city] + rng.standard_normal((n, 2)) * spread[city][:, None]
    return locs.astype(np.float64), city


def _haversine_km(lat1, lon1, lat2, lon2):
    lat1, lon1, lat2, lon2 = (np.radians(a) for a in (lat1, lon1, lat2, lon2))
    d = np.sin((lat2 - lat1) * 0.5) ** 2 + np.cos(lat1) * np.cos(lat2) * np.sin((lon2 - lon1) * 0.5) ** 2
    return 2 * 6371.0088 * np.arcsin(np.sqrt(d))


def assign_classes(train_locs, other_locs, bucket):
    """data.py:399-421: k-d tree regions on the training points, per-region median
    coordinates, every other point goes to the nearest median (haversine, first minimum)."""
    from . import ops
    labels, n_leaves = ops.kdtree_fit(train_locs, bucket)
    med = np.zeros((n_leaves, 2))
    order = np.argsort(labels, kind="stable")
    bounds = np.searchsorted(labels[order], np.arange(n_leaves + 1))
    for c in range(n_leaves):
        pts = train_locs[order[bounds[c]:bounds[c + 1]]]
        med[c] = (np.median(pts[:, 0]), np.median(pts[:, 1]))
    out = np.zeros(len(other_locs), dtype=np.int64)
    for s in range(0, len(other_locs), 4096):
        blk = other_locs[s:s + 4096]
        d = _haversine_km(blk[:, 0:1], blk[:, 1:2], med[None, :, 0], med[None, :, 1])
        out[s:s + 4096] = d.argmin(axis=1)
    return labels, out, med


# ------------------------------------------------------------------- workload
def make_workload(name="geotext", seed=77, community=True, scale=1.0, **overrides) -> Workload:
    """Build one of WORKLOADS.  ``scale`` < 1 shrinks node / vocabulary counts (bounded CPU
    samples of the same shape); ``community`` wires 80% of the edges inside the node's city,
    as geography does for real mention graphs (the premise of the reference model)."""
    from .sparse import build_ahat_host
    cfg = dict(WORKLOADS[name])
    cfg.update(overrides)
    if scale != 1.0:
        for k in ("n_train", "n_dev", "n_test", "vocab"):
            cfg[k] = max(8, int(round(cfg[k] * scale)))
        cfg["n_cities"] = max(4, int(round(cfg["n_cities"] * max(scale, 0.05))))
    n_train, n_dev, n_test = cfg["n_train"], cfg["n_dev"], cfg["n_test"]
    n = n_train + n_dev + n_test
    locs, city = city_locations(n, cfg["n_cities"], seed)
    adj = powerlaw_graph(n, cfg["avg_deg"], seed, communities=city if community else None)
    a_hat = build_ahat_host(adj)
    X = tfidf_matrix(n, cfg["vocab"], cfg["terms"], seed, city=city)
    y_train, y_other, med = assign_classes(locs[:n_train], locs[n_train:], cfg["bucket"])
    Y = np.concatenate([y_train, y_other]).astype(np.int64)
    return Workload(
        name=name, X=X, A_hat=a_hat, Y=Y,
        train_indices=np.arange(0, n_train, dtype=np.int32),                       # tensormain.py:225
        dev_indices=np.arange(n_train, n_train + n_dev, dtype=np.int32),           # :229
        test_indices=np.arange(n_train + n_dev, n, dtype=np.int32),                # :230
        locs=locs, medians=med, hidden=cfg["hidden"], n_classes=int(Y.max()) + 1,
        meta=dict(cfg, n=n, nnz_A=int(a_hat.nnz), nnz_X=int(X.nnz), seed=seed, community=community,
                  max_degree=int(np.diff(a_hat.indptr).max())))


# --------------------------------------------------------------------------- #
# Large workloads: the same generators written with torch ops so that the 1.4M-node
# Twitter-World shape is produced on the GPU in seconds (NumPy on 8 host cores needs
# ~8 minutes).  Data generation is plumbing, not the product; the arrays are then
# handed to the product exactly like host data would be.
# --------------------------------------------------------------------------- #
def _torch_graph(n, avg_deg, gen, device, city=None, intra=0.8, alpha=1.5):
    import torch
    m = int(n * avg_deg / 2)
    u = torch.rand(n, generator=gen, device=device, dtype=torch.float64)
    w = (1.0 - u).pow(-1.0 / alpha)                       # Pareto(alpha) + 1
    cdf = torch.cumsum(w, 0)
    cdf = cdf / cdf[-1]
    src = torch.searchsorted(cdf, torch.rand(m, generator=gen, device=device, dtype=torch.float64)).clamp_(max=n - 1)
    dst = torch.randint(0, n, (m,), generator=gen, device=device)
    if city is not None:
        order = torch.argsort(city, stable=True)
        sorted_c = city[order]
        starts = torch.searchsorted(sorted_c, sorted_c, right=False)
        ends = torch.searchsorted(sorted_c, sorted_c, right=True)
        pos_of = torch.empty(n, dtype=torch.int64, device=device)
        pos_of[order] = torch.arange(n, device=device)
        sp_ = pos_of[src]
        lo, hi = starts[sp_], ends[sp_]
        r = torch.rand(m, generator=gen, device=device, dtype=torch.float64)
        local = order[lo + (r * (hi - lo).double()).long().clamp_(max=n)]
        use_local = torch.rand(m, generator=gen, device=device) < intra
        dst = torch.where(use_local, local, dst)
    keep = src != dst
    src, dst = src[keep], dst[keep]
    a, b = torch.minimum(src, dst), torch.maximum(src, dst)
    key = torch.unique(a * n + b)
    a, b = key // n, key % n
    rows = torch.cat([a, b])
    cols = torch.cat([b, a])
    key = torch.sort(rows * n + cols).values
    rows, cols = key // n, key % n
    indptr = torch.zeros(n + 1, dtype=torch.int64, device=device)
    indptr[1:] = torch.cumsum(torch.bincount(rows, minlength=n), 0)
    return indptr.to(torch.int32), cols.to(torch.int32)


def _torch_tfidf(n, vocab, mean_terms, gen, device, zipf_s=1.1, city=None):
    import torch
    sigma = 0.6
    mu = float(np.log(mean_terms) - 0.5 * sigma * sigma)
    lens = torch.exp(mu + sigma * torch.randn(n, generator=gen, device=device, dtype=torch.float64))
    lens = lens.clamp_(min=1, max=vocab // 2).long()
    total = int(lens.sum().item())
    ranks = torch.arange(1, vocab + 1, dtype=torch.float64, device=device)
    cdf = torch.cumsum(ranks.pow(-zipf_s), 0)
    cdf = cdf / cdf[-1]
    rows = torch.repeat_interleave(torch.arange(n, device=device), lens)
    keys = []
    step = 1 << 26                                          # bound peak memory
    for s in range(0, total, step):
        e = min(total, s + step)
        c = torch.searchsorted(cdf, torch.rand(e - s, generator=gen, device=device, dtype=torch.float64)).clamp_(max=vocab - 1)
        if city is not None:
            local = torch.rand(e - s, generator=gen, device=device) < LOCAL_TERM_FRACTION
            c = torch.where(local, _local_columns(c, city[rows[s:e]], vocab, torch), c)
        keys.append(rows[s:e] * vocab + c)
    del rows
    key = torch.unique(torch.cat(keys))
    del keys
    rows, cols = key // vocab, key % vocab
    del key
    df = torch.bincount(cols, minlength=vocab).double()
    idf = torch.log((1.0 + n) / (1.0 + df)) + 1.0
    vals = idf[cols]
    sq = torch.zeros(n, dtype=torch.float64, device=device).index_add_(0, rows, vals * vals)
    vals = (vals / torch.sqrt(sq[rows])).float()
    indptr = torch.zeros(n + 1, dtype=torch.int64, device=device)
    indptr[1:] = torch.cumsum(torch.bincount(rows, minlength=n), 0)
    return indptr.to(torch.int32), cols.to(torch.int32), vals


def make_raw_device(name="twitter-world", device="cuda", seed=77, community=True, scale=1.0, **overrides):
    """The seeded raw inputs of a workload, generated with torch on ``device`` and touching nothing of
    libgcg.so: binary adjacency pattern (indptr, indices int32), TF-IDF X (indptr, indices int32, values
    float32) and coordinates (host float64).  ``make_workload_device`` builds A_hat / labels from these with
    the product's entry points; ``bench.py --impl reference`` builds them with the oracle's (tensormain.py:
    170-180, kdtree.py) so that the reference arm runs the same workload without loading the product."""
    import torch
    cfg = dict(WORKLOADS[name])
    cfg.update(overrides)
    if scale != 1.0:
        for k in ("n_train", "n_dev", "n_test", "vocab"):
            cfg[k] = max(8, int(round(cfg[k] * scale)))
        cfg["n_cities"] = max(4, int(round(cfg["n_cities"] * max(scale, 0.05))))
    n = cfg["n_train"] + cfg["n_dev"] + cfg["n_test"]
    dev = torch.device(device)
    gen = torch.Generator(device=dev).manual_seed(seed)
    locs, city = city_locations(n, cfg["n_cities"], seed)
    city_t = torch.from_numpy(city).to(dev) if community else None
    ip, ix = _torch_graph(n, cfg["avg_deg"], gen, dev, city=city_t)
    xip, xix, xv = _torch_tfidf(n, cfg["vocab"], cfg["terms"], gen, dev,
                                city=torch.from_numpy(city).to(dev))
    return cfg, n, locs, (ip, ix), (xip, xix, xv)


def make_workload_device(name="twitter-world", device="cuda", seed=77, community=True, scale=1.0, **overrides):
    """Same shapes as make_workload, generated with torch on ``device``.  Returns a Workload whose
    X / A_hat are ``CSRMatrix`` objects already resident on the device (their host copies are
    materialised lazily, only where the product needs them: transposes and row gathers)."""
    import torch
    from .sparse import CSRMatrix, _np_ptr
    from . import _lib
    cfg, n, locs, (ip, ix), (xip, xix, xv) = make_raw_device(name, device, seed, community, scale, **overrides)
    n_train, n_dev, n_test = cfg["n_train"], cfg["n_dev"], cfg["n_test"]
    dev = torch.device(device)
    if dev.type == "cuda":
        # A_hat through the product's device builder (gcg_ahat_*_device): float64 normalise, cast
        from .sparse import build_ahat_device
        a_hat = build_ahat_device(ip, ix, n)
        oip = a_hat.indptr.cpu().numpy()
    else:
        # A_hat through the product's host builder (gcg_ahat_build_host)
        hip, hix = ip.cpu().numpy(), ix.cpu().numpy()
        L = _lib.lib()
        nnz = L.gcg_ahat_nnz_host(n, _np_ptr(hip), _np_ptr(hix))
        oip, oix, ov = np.empty(n + 1, np.int32), np.empty(nnz, np.int32), np.empty(nnz, np.float32)
        _lib.check(L.gcg_ahat_build_host(n, _np_ptr(hip), _np_ptr(hix), None, _np_ptr(oip), _np_ptr(oix), _np_ptr(ov)),
                   "gcg_ahat_build_host")
        a_hat = CSRMatrix.from_host((oip, oix, ov), (n, n), dev)
    del ip, ix
    X = CSRMatrix(xip, xix, xv, (n, cfg["vocab"]), long_row_threshold=1024)
    y_train, y_other, med = assign_classes(locs[:n_train], locs[n_train:], cfg["bucket"])
    Y = np.concatenate([y_train, y_other]).astype(np.int64)
    return Workload(
        name=name, X=X, A_hat=a_hat, Y=Y,
        train_indices=np.arange(0, n_train, dtype=np.int32),
        dev_indices=np.arange(n_train, n_train + n_dev, dtype=np.int32),
        test_indices=np.arange(n_train + n_dev, n, dtype=np.int32),
        locs=locs, medians=med, hidden=cfg["hidden"], n_classes=int(Y.max()) + 1,
        meta=dict(cfg, n=n, nnz_A=int(a_hat.nnz), nnz_X=int(X.nnz), seed=seed, community=community,
                  max_degree=int(np.diff(oip).max())))
