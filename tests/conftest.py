import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)

MICRO = dict(n_sent=700, n_qry=8, v_src=260, v_tgt=260, n_phrases=500, mean_len=14.0, sd_len=5.0, max_len=40, qry_mean_len=9.0)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Build (idempotent) the CUDA library, the host library and the oracle."""
    import __graft_entry__ as g
    g.build()
    return True


@pytest.fixture(scope="session")
def micro(built):
    from cgx_b200 import synth
    c = synth.generate(**MICRO)
    return c, synth.text_layout(c)


@pytest.fixture(scope="session")
def micro_files(micro, tmp_path_factory):
    from cgx_b200 import synth
    d = tmp_path_factory.mktemp("micro")
    return synth.write_text(micro[0], str(d), "corpus")


@pytest.fixture(scope="session")
def golden():
    import lzma

    import numpy as np
    g = os.path.join(ROOT, "tests", "golden")
    arrays = dict(np.load(os.path.join(g, "micro_ref.npz"), allow_pickle=False))
    grammars, cur = {}, None
    with lzma.open(os.path.join(g, "micro_ref_grammars.txt.xz"), "rt") as fh:
        for line in fh:
            if line.startswith("### "):
                cur = int(line[4:])
                grammars[cur] = []
            else:
                grammars[cur].append(line)
    return arrays, grammars


@pytest.fixture(scope="session")
def micro_oracle(micro):
    from _oracle import Oracle
    _, lay = micro
    o = Oracle.from_layout(lay)
    o.build_sa()
    o.run(lay["qry_tok"], lay["qry_off"])
    return o
