"""One short invocation of the hot path for ncu (GPU box): build the index of a bench workload, then run
`batches` device-resident query batches (the first is the allocation warm-up).  Usage:
    python tools/ncu_step.py [workload=c2] [batches=2]
The launch list / --set full captures committed under profiles/ are taken from this command."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from cgx_b200.extractor import GrammarExtractor  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "c2"
batches = int(sys.argv[2]) if len(sys.argv) > 2 else 2
import torch  # noqa: E402  (device memory for the resident query batch only)

lay = bench.make_inputs(workload, 0)
ex = GrammarExtractor(0)
info = ex.build_index(lay)
print("index", info, flush=True)
dev = torch.device("cuda", 0)
Q, T = len(lay["qry_off"]) - 1, len(lay["qry_tok"])
qt = torch.from_numpy(np.ascontiguousarray(lay["qry_tok"], dtype=np.int32)).to(dev)
qo = torch.from_numpy(np.ascontiguousarray(lay["qry_off"], dtype=np.int32)).to(dev)
t2q = torch.from_numpy(np.repeat(np.arange(Q, dtype=np.int32), np.diff(lay["qry_off"]))).to(dev)
for b in range(batches):
    r = ex.extract_dev(qt.data_ptr(), qo.data_ptr(), t2q.data_ptr(), Q, T)
    print("batch", b, r, flush=True)
ex.close()
