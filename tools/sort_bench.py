"""GPU box: the onesweep radix sort alone (cgx_debug_sort_u64) on random keys -- correctness against torch.sort and
achieved HBM GB/s per pass.  Usage: python tools/sort_bench.py [log2_n=27] [bits=52] [vals=0|1] [reps=3] [check=1]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402  (device memory + the reference sort only)
from cgx_b200 import _lib  # noqa: E402

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 27
bits = int(sys.argv[2]) if len(sys.argv) > 2 else 52
with_vals = int(sys.argv[3]) if len(sys.argv) > 3 else 0
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
check = int(sys.argv[5]) if len(sys.argv) > 5 else 1
n = 1 << lg
L = _lib.load()
h = C.c_void_p()
assert L.cgx_create(0, C.byref(h)) == 0
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)
g.manual_seed(7)
src = torch.randint(0, 1 << 62, (n,), dtype=torch.int64, device=dev, generator=g) & ((1 << bits) - 1)
ms, passes = C.c_float(), C.c_int()
for r in range(reps):
    keys = src.clone()
    vals = torch.arange(n, dtype=torch.int32, device=dev) if with_vals else None
    torch.cuda.synchronize()
    rc = L.cgx_debug_sort_u64(h, keys.data_ptr(), vals.data_ptr() if with_vals else None, n, 0, bits, C.byref(ms), C.byref(passes))
    assert rc == 0, L.cgx_last_error(h)
    per_key = 8 + (4 if with_vals else 0)
    gb = n * (2.0 * per_key * passes.value + 8) / 1e9          # every pass reads+writes the data, the histogram reads the keys once
    print("n=2^%d bits=%d vals=%d: %.3f ms, %d passes, %.1f Gkeys/s, %.0f GB/s algorithmic" % (lg, bits, with_vals, ms.value, passes.value, n / ms.value / 1e6, gb / (ms.value / 1e3)), flush=True)
if check:
    ref, perm = torch.sort(src, stable=True)
    assert torch.equal(keys, ref), "keys not sorted like torch.sort"
    if with_vals:
        assert torch.equal(vals.to(torch.int64), perm), "payloads differ from the stable permutation"
    print("check ok")
L.cgx_destroy(h)
