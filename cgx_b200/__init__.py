"""cgx_b200 -- B200-native hierarchical (Hiero) grammar extractor: the `strmatchcuda` hot path of
hohoCode/cgx rebuilt from scratch for sm_100a.  See DESIGN.md.

    csrc/   hand-written CUDA kernels + the C ABI (include/cgx_b200.h) -> lib/libcgx_b200.so
    host/   plain-C loaders, grammar writer, driver, strmatchcuda main -> lib/libcgx_host.so, bin/strmatchcuda
    extractor.py / host.py   ctypes mirrors of the reference's host interface (tests, bench)
    synth.py                 deterministic synthetic corpora (the reference ships no data)
"""
__version__ = "0.1.0"
