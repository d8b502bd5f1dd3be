"""cProfile of one minibatch-MLP epoch (host side) -- finds host/sync overheads of the minibatch path."""
import cProfile
import os
import pstats
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from graphconvgeo_b200 import synth  # noqa: E402
from graphconvgeo_b200.mlp import MLP  # noqa: E402
from graphconvgeo_b200.sparse import spgemm  # noqa: E402

wl = synth.make_workload_device(sys.argv[1] if len(sys.argv) > 1 else "geotext", device="cuda:0", seed=77)
Xc = spgemm(wl.A_hat, wl.X, a_values=wl.A_hat.data.double())
ntr = len(wl.train_indices)
y = wl.Y.astype(np.int32)
uniq = np.unique(y[:ntr])
remap = -np.ones(int(y.max()) + 1, np.int64)
remap[uniq] = np.arange(len(uniq))
clf = MLP(n_epochs=1, batch_size=500, regul_coefs=[1e-6, 1e-6], hidden_layer_size=500, seed=0)
clf.prepare(Xc.to_scipy()[:ntr], remap[y[:ntr]].astype(np.int32))
clf.train_epoch()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
clf.train_epoch()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
