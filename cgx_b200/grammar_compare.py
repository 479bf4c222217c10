"""Compare two directories of cdec-style per-query grammar files (``grammar.<qid>.s``).

The reference's rule order inside a source pattern is nondeterministic (atomicAdd cursors + sorts
keyed only on the pattern id), so files are compared as *multisets of lines*:
``(source, target, IsSingletonF, IsSingletonFE)`` must match exactly and the five float features
within a tolerance (1e-5 relative per BASELINE.json; an absolute floor covers the 6-decimal "%f"
print and values near zero).
"""
from __future__ import annotations

import os
import re
from collections import defaultdict

_FEAT = re.compile(r"(\w+)=(-?[0-9.]+(?:e[-+]?\d+)?|-?inf|-?nan)")
FLOAT_KEYS = ("EgivenFCoherent", "SampleCountF", "CountEF", "MaxLexFgivenE", "MaxLexEgivenF")
INT_KEYS = ("IsSingletonF", "IsSingletonFE")


def parse_line(line: str):
    parts = line.rstrip("\n").split(" ||| ")
    if len(parts) != 4 or parts[0] != "[X]":
        raise ValueError(f"bad grammar line: {line!r}")
    feats = dict(_FEAT.findall(parts[3]))
    key = (parts[1], parts[2]) + tuple(int(feats[k]) for k in INT_KEYS)
    vals = tuple(float(feats[k]) for k in FLOAT_KEYS)
    return key, vals


def load(path: str):
    d = defaultdict(list)
    with open(path) as fh:
        for line in fh:
            if not line.strip():
                continue
            k, v = parse_line(line)
            d[k].append(v)
    return d


def compare_files(path_a: str, path_b: str, rtol: float = 1e-5, atol: float = 2e-6):
    """Returns dict(n_a, n_b, only_a, only_b, float_mismatch, examples)."""
    a, b = load(path_a), load(path_b)
    only_a = only_b = bad = 0
    by_feature = dict.fromkeys(FLOAT_KEYS, 0)      # float mismatches by the first feature that differs
    ex = []
    for k in set(a) | set(b):
        va, vb = sorted(a.get(k, [])), sorted(b.get(k, []))
        if len(va) != len(vb):
            only_a += max(0, len(va) - len(vb))
            only_b += max(0, len(vb) - len(va))
            if len(ex) < 10:
                ex.append(("count", k, len(va), len(vb)))
        for x, y in zip(va, vb):
            for i, (p, q) in enumerate(zip(x, y)):
                if abs(p - q) > atol + rtol * max(abs(p), abs(q)):
                    bad += 1
                    by_feature[FLOAT_KEYS[i]] += 1
                    if len(ex) < 10:
                        ex.append(("float", k, FLOAT_KEYS[i], p, q))
                    break
    na = sum(len(v) for v in a.values())
    nb = sum(len(v) for v in b.values())
    return dict(n_a=na, n_b=nb, only_a=only_a, only_b=only_b, float_mismatch=bad, float_mismatch_by_feature=by_feature, examples=ex)


def compare_dirs(dir_a: str, dir_b: str, rtol: float = 1e-5, atol: float = 2e-6):
    names = sorted(set(f for f in os.listdir(dir_a) if f.startswith("grammar.")) |
                   set(f for f in os.listdir(dir_b) if f.startswith("grammar.")))
    tot = dict(files=0, n_a=0, n_b=0, only_a=0, only_b=0, float_mismatch=0, missing_files=0, examples=[],
               float_mismatch_by_feature=dict.fromkeys(FLOAT_KEYS, 0))
    for n in names:
        pa, pb = os.path.join(dir_a, n), os.path.join(dir_b, n)
        if not (os.path.exists(pa) and os.path.exists(pb)):
            tot["missing_files"] += 1
            continue
        r = compare_files(pa, pb, rtol, atol)
        tot["files"] += 1
        for k in ("n_a", "n_b", "only_a", "only_b", "float_mismatch"):
            tot[k] += r[k]
        for k, v in r["float_mismatch_by_feature"].items():
            tot["float_mismatch_by_feature"][k] += v
        if len(tot["examples"]) < 10:
            tot["examples"] += [(n,) + e for e in r["examples"]][: 10 - len(tot["examples"])]
    matched = tot["n_b"] - tot["only_b"]
    tot["frac_equal"] = (matched - tot["float_mismatch"]) / max(1, max(tot["n_a"], tot["n_b"]))
    return tot


if __name__ == "__main__":
    import json
    import sys
    r = compare_dirs(sys.argv[1], sys.argv[2])
    print(json.dumps(r, indent=1, default=str))
