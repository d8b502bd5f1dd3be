"""CPU restatement of the reference k-d tree region labelling (TEST INFRASTRUCTURE).

Follows /root/reference/kdtree.py (KDTree.nodeSplit :84-118,
KDTreeClustering.fit :126-147) and its caller data.py:399-421.  PINNED: checked
bit-for-bit against the reference file itself (tests/golden/make_kdtree_golden.py
imports it from /root/reference; outputs committed under tests/golden/).
"""
from __future__ import annotations

import numpy as np


def kdtree_labels(X, bucket_size):
    """Leaf ordinal per point, int64 -- KDTreeClustering.fit (kdtree.py:126-147).

    Semantics restated from kdtree.py:
      * split iff count > bucket_size                                  (:85)
      * split dim = argmax(max-min), first index wins ties             (:51-54, :86)
      * split value = np.median of that coordinate                     (:88)
      * zero width -> stay a leaf                                      (:90-91)
      * split value == max -> split value = min                        (:94-95)
      * value > split -> right, else left, original order kept         (:106-114)
      * recurse left then right; leaves numbered depth-first L-before-R (:72-79,117-118)
    """
    X = np.asarray(X, dtype=np.float64)
    n, dims = X.shape
    labels = np.zeros(n, dtype=np.int64)           # kdtree.py:139 (dtype=int)
    next_leaf = 0
    stack = [np.arange(n)]
    while stack:
        ids = stack.pop()
        if len(ids) == 0:
            # an empty child is still a leaf (getLeaves counts it); cannot happen
            # after the max->min rule, kept for faithfulness
            next_leaf += 1
            continue
        pts = X[ids]
        leaf = True
        if len(ids) > bucket_size:
            mins, maxs = pts.min(axis=0), pts.max(axis=0)
            d = int(np.argmax([maxs[i] - mins[i] for i in range(dims)]))
            sv = np.median(pts[:, d])
            if mins[d] != maxs[d]:
                if sv == maxs[d]:
                    sv = mins[d]
                right = pts[:, d] > sv
                # depth-first, left before right: push right first
                stack.append(ids[right])
                stack.append(ids[~right])
                leaf = False
        if leaf:
            labels[ids] = next_leaf
            next_leaf += 1
    return labels, next_leaf


def cluster_medians(train_locs, labels):
    """data.py:404-413: per cluster (median lat, median lon), sorted cluster order."""
    train_locs = np.asarray(train_locs, dtype=np.float64)
    k = int(labels.max()) + 1
    med = np.zeros((k, 2), dtype=np.float64)
    order = np.argsort(labels, kind="stable")
    bounds = np.searchsorted(labels[order], np.arange(k + 1))
    for c in range(k):
        pts = train_locs[order[bounds[c]:bounds[c + 1]]]
        med[c, 0] = np.median(pts[:, 0])
        med[c, 1] = np.median(pts[:, 1])
    return med


AVG_EARTH_RADIUS_KM = 6371.0088   # haversine package constant [3P]


def haversine_km(lat1, lon1, lat2, lon2):
    """``haversine`` package formula [3P] (data.py:416, tensormain.py:47)."""
    lat1, lon1, lat2, lon2 = (np.radians(np.asarray(a, dtype=np.float64)) for a in (lat1, lon1, lat2, lon2))
    d = np.sin((lat2 - lat1) * 0.5) ** 2 + np.cos(lat1) * np.cos(lat2) * np.sin((lon2 - lon1) * 0.5) ** 2
    return 2 * AVG_EARTH_RADIUS_KM * np.arcsin(np.sqrt(d))


def nearest_median_labels(locs, medians, chunk=4096):
    """data.py:416-419: brute-force 1-NN under haversine; ties -> lowest index."""
    locs = np.asarray(locs, dtype=np.float64)
    out = np.zeros(len(locs), dtype=np.int64)
    for s in range(0, len(locs), chunk):
        blk = locs[s:s + chunk]
        d = haversine_km(blk[:, 0:1], blk[:, 1:2], medians[None, :, 0], medians[None, :, 1])
        out[s:s + chunk] = d.argmin(axis=1)
    return out


def geo_eval(true_locs, y_pred, medians):
    """tensormain.py:38-54: mean / median haversine km and Acc@161 (strict <)."""
    true_locs = np.asarray(true_locs, dtype=np.float64)
    assert len(y_pred) == len(true_locs)                                  # :39
    pred = medians[np.asarray(y_pred)]
    dist = haversine_km(true_locs[:, 0], true_locs[:, 1], pred[:, 0], pred[:, 1])
    acc161 = 100.0 * np.count_nonzero(dist < 161) / float(len(dist))      # :50
    return float(np.mean(dist)), float(np.median(dist)), acc161
