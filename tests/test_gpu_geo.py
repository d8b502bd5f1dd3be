"""Label pipeline / geo_eval (data.py:399-421, tensormain.py:38-54) on the GPU vs the oracle."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import kdtree_oracle as ko  # noqa: E402


def test_assign_classes_and_geo_eval_match_the_oracle():
    from graphconvgeo_b200 import geo, synth
    locs, _ = synth.city_locations(30000, 60, seed=5)
    tr, dv, te = locs[:24000], locs[24000:27000], locs[27000:]
    ytr, ydv, yte, med = geo.assign_classes(tr, dv, te, bucket_size=300)
    ref_tr, k = ko.kdtree_labels(tr, 300)
    assert np.array_equal(ytr, ref_tr)                                  # bit-exact regions
    ref_med = ko.cluster_medians(tr, ref_tr)
    assert np.array_equal(med, ref_med)
    ref_dv = ko.nearest_median_labels(dv, ref_med)
    d_gpu = ko.haversine_km(dv[:, 0], dv[:, 1], med[ydv, 0], med[ydv, 1])
    d_ref = ko.haversine_km(dv[:, 0], dv[:, 1], med[ref_dv, 0], med[ref_dv, 1])
    agree = ydv == ref_dv
    assert agree.mean() > 0.9995                                         # only float64 near-ties may differ
    assert np.all(np.abs(d_gpu - d_ref)[~agree] < 1e-6)
    idx, km = geo.nearest_median(te, med, return_km=True)
    assert np.array_equal(idx, yte)
    assert np.allclose(km, ko.haversine_km(te[:, 0], te[:, 1], med[idx, 0], med[idx, 1]), rtol=1e-10, atol=1e-9)
    mean, median, acc = geo.geo_eval(te, yte, med)
    rmean, rmedian, racc = ko.geo_eval(te, yte, med)
    assert abs(mean - rmean) < 1e-8 and abs(median - rmedian) < 1e-8 and acc == racc
    with pytest.raises(AssertionError, match="#preds"):
        geo.geo_eval(te, yte[:-1], med)
