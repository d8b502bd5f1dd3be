"""Label pipeline and evaluation around the GCN hot path.

``assign_classes`` mirrors DataLoader.assignClasses (data.py:399-421): k-d tree regions over the
training coordinates (bit-exact, gcg_kdtree_fit_host), per-region median coordinates, dev/test users
to the nearest median under haversine.  ``geo_eval`` mirrors tensormain.py:38-54 (mean / median km,
Acc@161).  The distance work runs on the GPU in float64 (gcg_haversine_*_f64).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib, ops


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def cluster_medians(train_locs, labels):
    """data.py:404-413: (median lat, median lon) per cluster, sorted cluster order (host, float64)."""
    train_locs = np.asarray(train_locs, dtype=np.float64)
    k = int(labels.max()) + 1 if len(labels) else 0
    med = np.zeros((k, 2), dtype=np.float64)
    order = np.argsort(labels, kind="stable")
    bounds = np.searchsorted(labels[order], np.arange(k + 1))
    for c in range(k):
        pts = train_locs[order[bounds[c]:bounds[c + 1]]]
        med[c] = (np.median(pts[:, 0]), np.median(pts[:, 1]))
    return med


def nearest_median(locs, medians, device="cuda", return_km=False):
    """argmin_c haversine(loc, median_c), first minimum (data.py:416-419) -- gcg_haversine_nearest_f64."""
    pts = torch.from_numpy(np.ascontiguousarray(np.asarray(locs, dtype=np.float64))).to(device)
    med = torch.from_numpy(np.ascontiguousarray(np.asarray(medians, dtype=np.float64))).to(device)
    n = pts.shape[0]
    out = torch.empty(n, dtype=torch.int64, device=device)
    km = torch.empty(n, dtype=torch.float64, device=device) if return_km else None
    _lib.check(_lib.lib().gcg_haversine_nearest_f64(pts.data_ptr(), n, med.data_ptr(), med.shape[0], out.data_ptr(),
                                                    km.data_ptr() if km is not None else None, _stream()),
               "gcg_haversine_nearest_f64")
    return (out.cpu().numpy(), km.cpu().numpy()) if return_km else out.cpu().numpy()


def assign_classes(train_locs, dev_locs, test_locs, bucket_size, device="cuda"):
    """-> (train_classes, dev_classes, test_classes, cluster_median) as DataLoader.assignClasses leaves them."""
    train_classes, _ = ops.kdtree_fit(train_locs, bucket_size)                 # data.py:400-403
    med = cluster_medians(train_locs, train_classes)                           # :404-413
    dev_classes = nearest_median(dev_locs, med, device)                        # :416-419
    test_classes = nearest_median(test_locs, med, device)
    return train_classes, dev_classes, test_classes, med


def geo_eval(true_locs, y_pred, medians, device="cuda"):
    """tensormain.py:38-54: (mean km, median km, Acc@161 in percent) of predicted-region medians."""
    true_locs = np.asarray(true_locs, dtype=np.float64)
    y_pred = np.asarray(y_pred)
    assert len(y_pred) == len(true_locs), "#preds: %d, #users: %d" % (len(y_pred), len(true_locs))   # :39
    a = torch.from_numpy(np.ascontiguousarray(true_locs)).to(device)
    b = torch.from_numpy(np.ascontiguousarray(np.asarray(medians, dtype=np.float64)[y_pred])).to(device)
    km = torch.empty(len(y_pred), dtype=torch.float64, device=device)
    _lib.check(_lib.lib().gcg_haversine_pairs_f64(a.data_ptr(), b.data_ptr(), len(y_pred), km.data_ptr(), _stream()),
               "gcg_haversine_pairs_f64")
    d = km.cpu().numpy()
    acc161 = 100.0 * np.count_nonzero(d < 161) / float(len(d))                # :50
    return float(np.mean(d)), float(np.median(d)), acc161
