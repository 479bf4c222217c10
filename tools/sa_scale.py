"""GPU box: suffix-array build at BASELINE.json configs[3] scale (~1 G source tokens) with a size-independent check.
    python tools/sa_scale.py [tokens=1040000000] [vocab=50000] [checks=2000000]
Tokens are drawn on the GPU (log-uniform ranks = Zipf s=1, EOS every ~26 tokens, the reference's trailer `1, V+2, 0 0 0`,
Start.cu:321-330); the SA comes from cgx_sa_build_dev (tokens and SA resident in HBM).  Check: the SA is a permutation and
`checks` random adjacent pairs sa[k], sa[k+1] are in lexicographic order (compared on the GPU over 48 tokens; ties beyond
that are counted, not failed)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from cgx_b200.extractor import GrammarExtractor  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_040_000_000
V = int(sys.argv[2]) if len(sys.argv) > 2 else 50_000
checks = int(sys.argv[3]) if len(sys.argv) > 3 else 2_000_000
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)
g.manual_seed(1234)
s = torch.empty(n + 3, dtype=torch.int32, device=dev)
step = 1 << 27
for a in range(0, n, step):
    b = min(n, a + step)
    u = torch.rand(b - a, device=dev, generator=g)
    s[a:b] = (torch.exp(u * float(torch.log(torch.tensor(float(V))))).to(torch.int32).clamp_(1, V) + 1)
    del u
pos = torch.arange(25, n, 26, device=dev)
s[pos] = 1
del pos
s[n - 2] = 1
s[n - 1] = V + 2
s[n:] = 0
sa = torch.empty(n, dtype=torch.int32, device=dev)
ex = GrammarExtractor(0)
for rep in range(2):
    rounds, ms = ex.sa_build_dev(s.data_ptr(), n, V + 2, sa.data_ptr())
    print("n=%d tokens: SA build %.1f ms, %d doubling rounds, %.2f G tokens/s, %.0f GB/s at 16 B/token/round" % (n, ms, rounds, n / ms / 1e6, 16.0 * n * rounds / ms / 1e6), flush=True)
# permutation: every position exactly once
seen = torch.zeros(n, dtype=torch.uint8, device=dev)
seen[sa.long()] = 1
assert int(seen.sum()) == n, "suffix array is not a permutation"
del seen
k = torch.randint(0, n - 1, (checks,), device=dev, generator=g)
a, b = sa[k].long(), sa[k + 1].long()
undecided = torch.ones(checks, dtype=torch.bool, device=dev)
ok = torch.ones(checks, dtype=torch.bool, device=dev)
for j in range(48):
    ta, tb = s[(a + j).clamp_(max=n + 2)], s[(b + j).clamp_(max=n + 2)]
    lt, gt = undecided & (ta < tb), undecided & (ta > tb)
    ok &= ~gt
    undecided &= ~(lt | gt)
print("checked %d adjacent pairs: %d out of order, %d equal over 48 tokens" % (checks, int((~ok).sum()), int(undecided.sum())))
assert bool(ok.all())
ex.close()
