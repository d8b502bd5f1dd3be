"""Experiment only (not product code): run bench.py with ops.put_rows replaced by torch's index_copy_ (what the
hot path used before gcg_put_rows_f32) to check that the two placements give the same training trajectory."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graphconvgeo_b200 import ops


def put_rows_torch(src, idx, dst):
    dst.index_copy_(0, idx.long(), src)
    return dst


ops.put_rows = put_rows_torch
import bench
bench.main()
