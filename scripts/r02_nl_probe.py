"""Neighbour rebuild of the bench's Cu configuration (1 M atoms, rc 5.5) three times -- run under
`ncu --metrics gpu__time_duration.sum` to see the kernels of a build.  Measurement helper."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from atomistica_b200 import native  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 63
rc = float(sys.argv[2]) if len(sys.argv) > 2 else 5.50679
pos, cell = bench.fcc_positions(3.615, n)
pos += np.random.RandomState(12345).normal(scale=0.35, size=pos.shape)
p = native.from_arrays(np.full(len(pos), 29, dtype=np.int32), pos, cell, True)
p._sync()
nl = native.Neighbors(200)
nl.request_interaction_range(rc)
for k in range(3):
    nl.rebuild(p)
print(len(pos), nl.info())
