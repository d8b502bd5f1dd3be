"""scripts/l2_lru.c (the LRU cache model behind profiles/r02_spmm_spans.md) against a plain-Python LRU, and the
round-robin interleaving of spans that models 2,368 warps advancing together."""
import collections
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))


def py_lru_misses(acc, cap):
    cache = collections.OrderedDict()
    miss = 0
    for r in acc:
        if r in cache:
            cache.move_to_end(r)
        else:
            miss += 1
            if len(cache) == cap:
                cache.popitem(last=False)
            cache[r] = True
    return miss


def test_c_lru_equals_python_lru():
    import l2_model
    lib = l2_model.lru_lib()
    rng = np.random.RandomState(0)
    for n_rows, n_acc, cap in ((50, 2000, 1), (50, 2000, 7), (300, 20000, 64), (300, 20000, 300), (10, 500, 50)):
        # a skewed stream (hot rows) with some locality, like gathers of a community-ordered graph
        base = (rng.zipf(1.3, size=n_acc) % n_rows).astype(np.int32)
        local = ((np.arange(n_acc) // 40 + rng.randint(0, 5, n_acc)) % n_rows).astype(np.int32)
        acc = np.ascontiguousarray(np.where(rng.random_sample(n_acc) < 0.5, base, local).astype(np.int32))
        got = int(lib.lru_misses(acc.ctypes.data, len(acc), n_rows, cap))
        assert got == py_lru_misses(acc.tolist(), cap), (n_rows, n_acc, cap)
    one = np.zeros(100, np.int32)
    assert int(lib.lru_misses(one.ctypes.data, 100, 1, 1)) == 1


def test_interleave_is_round_robin_over_the_spans_in_flight():
    import l2_model
    s = np.arange(2 * 3 * 4 + 5, dtype=np.int32)          # two full groups of 3 spans x 4 non-zeros, then a tail
    out = l2_model.interleave(s, span=4, in_flight=3)
    assert out[:12].tolist() == [0, 4, 8, 1, 5, 9, 2, 6, 10, 3, 7, 11]
    assert out[12:24].tolist() == [12, 16, 20, 13, 17, 21, 14, 18, 22, 15, 19, 23]
    assert out[24:].tolist() == [24, 25, 26, 27, 28] and sorted(out.tolist()) == s.tolist()
