"""Generate tests/golden/ fixtures by running the REFERENCE ITSELF (oracle/_ref/strmatchcuda_dump, the
unmodified reference plus fwrite hooks, see oracle/build_ref_dump.sh) on a B200.

    gpurun -- 'python tools/make_golden.py && cp -r tests/golden gpurun_out/golden'

The corpus is regenerated deterministically from the parameters stored in the fixture, so only the
reference's outputs are committed: its intermediate arrays (micro_ref.npz) and its grammar files
(micro_ref_grammars.txt.xz, all queries concatenated with '### <qid>' separators).
"""
import lzma
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

PARAMS = dict(n_sent=700, n_qry=8, v_src=260, v_tgt=260, n_phrases=500, mean_len=14.0, sd_len=5.0, max_len=40, qry_mean_len=9.0)


def main():
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    c = synth.generate(**PARAMS)
    with tempfile.TemporaryDirectory() as work:
        paths = synth.write_text(c, work, "corpus")
        os.makedirs(os.path.join(work, "out"))
        os.makedirs(os.path.join(work, "dump"))
        env = dict(os.environ, CGX_DUMP_DIR=os.path.join(work, "dump"))
        r = subprocess.run([REF_DUMP_BIN, paths["f"], paths["q"], paths["e"], paths["a"], paths["lex"], os.path.join(work, "out")], env=env,
                           cwd=work, capture_output=True, text=True)
        if r.returncode != 0 or "Start Printing Gappy Phrases" not in r.stderr:
            print(r.stderr[-3000:])
            raise SystemExit("reference run failed")
        d = load_dump(os.path.join(work, "dump"))
        keep = ("sa", "result_two", "connectoffset", "result_connect", "oneGapSA", "oneGapSearch", "onegapPattern", "twoGapSA", "twoGapSearch",
                "twogapPattern", "precomp_index", "precomp_onegap", "featureMissingCount", "frequentList", "out_res", "oneGapRule", "twoGapRule",
                "separators", "blocks")
        np.savez_compressed(os.path.join(out, "micro_ref.npz"), params=np.array(repr(PARAMS)), stderr=np.array(r.stderr), **{k: d[k] for k in keep})
        with lzma.open(os.path.join(out, "micro_ref_grammars.txt.xz"), "wt") as fh:
            for q in range(c.n_qry):
                fh.write("### %d\n" % q)
                with open(os.path.join(work, "out", "grammar.%d.s" % q)) as g:
                    fh.write(g.read())
    for f in os.listdir(out):
        print(f, os.path.getsize(os.path.join(out, f)))


if __name__ == "__main__":
    main()
