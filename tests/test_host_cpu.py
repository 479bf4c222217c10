"""CPU: the C-ABI library loads and exports every symbol include/cgx_b200.h declares (no compute calls without
a GPU), the product fails loudly without a device (no fallback), and the C host loaders reproduce the reference
layouts (checked against the independent numpy construction in cgx_b200/synth.py and the oracle's own loader)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_exports_every_declared_symbol(built):
    from cgx_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "cgx_b200.h")).read()
    declared = set(re.findall(r"\b(cgx_[a-z0-9_]+)\s*\(", hdr))
    lib = _lib.load()
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert lib.cgx_version() >= 100
    assert set(_lib.EXPORTED_SYMBOLS) <= declared


def test_no_cpu_fallback(built):
    """Without a CUDA device every entry point must fail loudly."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from cgx_b200.extractor import GrammarExtractor
    with pytest.raises(RuntimeError):
        GrammarExtractor(0)


def test_product_never_touches_the_oracle():
    """cgx_b200/ (kernels, ABI, host) must not include, link or import anything under oracle/."""
    pkg = os.path.join(ROOT, "cgx_b200")
    for dirpath, _, files in os.walk(pkg):
        if "lib" in dirpath.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".c", ".h", ".cu", ".cuh")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                for line in text.splitlines():
                    if re.search(r"#\s*include.*oracle|import\s+_oracle|from\s+_oracle|cgx_oracle|-loracle|oracle/_build|oracle/_ref", line):
                        if "oracle/cgx_oracle.c" in line and line.strip().startswith(("//", "*", "#", "/*")):
                            continue   # a comment pointing at the restatement
                        pytest.fail("%s references the oracle: %s" % (os.path.join(dirpath, f), line.strip()))
    out = subprocess.run(["ldd", os.path.join(pkg, "lib", "libcgx_host.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_host_loaders_match_reference_layout(micro, micro_files, built):
    from cgx_b200.host import HostCorpus
    _, lay = micro
    hc = HostCorpus(micro_files["f"], micro_files["q"], micro_files["e"], micro_files["a"], micro_files["lex"])
    h = hc.layout()
    assert h["n"] == lay["n"] and h["m"] == lay["m"]
    for k in ("str", "tgt", "P", "L_tar", "R_tar", "qry_tok", "qry_off"):
        assert np.array_equal(h[k], lay[k]), k
    # RLP[n-1] is never written by the reference (ExtractPair.cu:2721 stops at toklen-1)
    assert np.array_equal(h["RLP"][:-1], lay["RLP"][:-1])
    # lexical table: same multiset of (f, e, v1, v2)
    a = sorted(zip(h["lex_f"].tolist(), h["lex_e"].tolist(), h["lex_v1"].tolist(), h["lex_v2"].tolist()))
    b = sorted(zip(lay["lex_f"].tolist(), lay["lex_e"].tolist(), lay["lex_v1"].tolist(), lay["lex_v2"].tolist()))
    assert a == b
    assert hc.src_name(2) == "s%d" % lay["src_names"][0]
    assert h["src_last"] == lay["src_last"]


def test_loader_edge_cases(tmp_path, built):
    """Empty lines, OOV query words, ragged alignment lines, a token that starts with a tab (Start.cu:279)."""
    from cgx_b200.host import HostCorpus
    (tmp_path / "f").write_text("a b c\n\nd a\n")
    (tmp_path / "e").write_text("x y\n\nz\n")
    (tmp_path / "a").write_text("0-0 2-1\n\n0-0 1-0\n")
    (tmp_path / "q").write_text("a zzz b\n\nc \tq d\n")
    (tmp_path / "lex").write_text("a x 0.5 0.25\nNULL x 0.1 0.2\nb NULL 0.3 0.4\nunk x 1 1\nNULL NULL 1 1\n")
    hc = HostCorpus(*(str(tmp_path / k) for k in ("f", "q", "e", "a", "lex")))
    h = hc.layout()
    assert h["str"].tolist() == [2, 3, 4, 1, 1, 5, 2, 1, 1, 6, 0, 0, 0]
    assert h["tgt"].tolist() == [2, 3, 1, 1, 4, 1, 1, 5, 0, 0, 0]
    assert h["qry_tok"].tolist() == [2, -1, 3, 4] and h["qry_off"].tolist() == [0, 3, 3, 4]
    assert h["P"].tolist() == [0, 1, 2, 0, 0, 0, 1, 0, 0, 0]
    L = (h["RLP"] >> 24) & 255
    assert L[:3].tolist() == [0, 255, 1] and h["RLP"][3] == 3 and h["RLP"][4] == 4
    assert h["L_tar"].tolist()[:2] == [0, 2] and h["L_tar"][4] == 0 and h["R_tar"][4] == 1
    assert sorted(zip(h["lex_f"].tolist(), h["lex_e"].tolist())) == [(-1, -1), (-1, 2), (2, 2), (3, -1)]


def test_cli_usage_contract(built):
    """Exactly six positionals, otherwise help and exit 0 (Main.c:46-48)."""
    exe = os.path.join(ROOT, "bin", "strmatchcuda")
    r = subprocess.run([exe, "a", "b"], capture_output=True, text=True)
    assert r.returncode == 0 and "Please check your input arguments" in r.stdout
    r = subprocess.run([exe, "-t", "11", "a", "b", "c", "d", "e", "f"], capture_output=True, text=True)
    assert r.returncode == 0 and "finger length must be between 1 and 10" in r.stderr


def test_synth_is_deterministic_and_within_reference_limits(micro):
    from cgx_b200 import synth
    from conftest import MICRO
    c, lay = micro
    c2 = synth.generate(**MICRO)
    assert np.array_equal(c.src_words, c2.src_words) and np.array_equal(c.link_t, c2.link_t) and np.array_equal(c.qry_words, c2.qry_words)
    assert np.diff(c.src_off).max() < 255 and np.diff(c.tgt_off).max() < 255          # ExtractPair.cu:2683
    assert len(np.unique(c.src_words)) >= 100                                          # SuffixArray.cu:1175
    assert lay["str"][lay["n"] - 1] == lay["str"].max() and lay["str"][lay["n"]:].tolist() == [0, 0, 0]
