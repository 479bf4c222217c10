"""GPU box (development): run the instrumented reference (oracle/_ref/strmatchcuda_dump) on a named synthetic configuration and keep
its intermediate arrays as gpurun_out/r2/dump_<name>.npz, for offline comparison with the oracle (tie orders of the hit lists)."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cgx_b200 import synth  # noqa: E402
from _oracle import REF_DUMP_BIN, load_dump  # noqa: E402

CONFIGS = {
    "small": dict(n_sent=3000, n_qry=24, v_src=600, v_tgt=600, n_phrases=1200, seed=99, qry_seed=77),
    "mid": dict(n_sent=100_000, n_qry=60, v_src=20_000, v_tgt=20_000, seed=1234, qry_seed=4321),
}


def main():
    name = sys.argv[1]
    c = synth.generate(**CONFIGS[name])
    out = os.path.join(ROOT, "gpurun_out", "r2")
    os.makedirs(out, exist_ok=True)
    with tempfile.TemporaryDirectory() as work:
        paths = synth.write_text(c, work, "corpus")
        os.makedirs(os.path.join(work, "out"))
        os.makedirs(os.path.join(work, "dump"))
        env = dict(os.environ, CGX_DUMP_DIR=os.path.join(work, "dump"))
        r = subprocess.run([REF_DUMP_BIN, paths["f"], paths["q"], paths["e"], paths["a"], paths["lex"], os.path.join(work, "out")], env=env,
                           cwd=work, capture_output=True, text=True)
        print(r.stderr[-1500:])
        d = load_dump(os.path.join(work, "dump"))
        keep = ("oneGapSA", "oneGapSearch", "twoGapSA", "twoGapSearch", "precomp_index", "precomp_onegap", "featureMissingCount", "frequentList",
                "out_res", "oneGapRule", "twoGapRule", "separators", "blocks")
        np.savez_compressed(os.path.join(out, "dump_%s.npz" % name), **{k: d[k] for k in keep if k in d})
    print(os.path.getsize(os.path.join(out, "dump_%s.npz" % name)))


if __name__ == "__main__":
    main()
