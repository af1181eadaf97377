"""The C-ABI shared library: it loads, exports every symbol include/strainscan_b200.h declares (and
nothing is bound in Python that the header lacks), fails loudly without a GPU, and the product
package never touches the oracle.  CPU only -- no compute calls."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    txt = open(os.path.join(ROOT, "include", "strainscan_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return set(re.findall(r"\b(ss_[a-z0-9_]+)\s*\(", txt))


def test_library_exports_every_declared_symbol():
    from strainscan_b200 import _lib
    names = _header_functions()
    assert len(names) >= 30
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "header declares %s but the .so does not export it" % n


def test_python_binding_matches_header():
    from strainscan_b200 import _lib
    assert set(_lib.SIGNATURES) == _header_functions()
    _lib.load()


def test_struct_layouts_match_header():
    """ss_stats / ss_synth_params sizes as the C compiler lays them out."""
    import subprocess
    import tempfile
    from strainscan_b200 import _lib
    src = '#include <stdio.h>\n#include "strainscan_b200.h"\nint main(){printf("%zu %zu\\n", sizeof(ss_stats), sizeof(ss_synth_params));return 0;}\n'
    with tempfile.TemporaryDirectory() as td:
        open(td + "/t.c", "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), td + "/t.c", "-o", td + "/t"])
        a, b = subprocess.check_output([td + "/t"]).split()
    assert int(a) == ctypes.sizeof(_lib.Stats) and int(b) == ctypes.sizeof(_lib.SynthParams)


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from strainscan_b200 import Engine, StrainScanB200Error
    with pytest.raises(StrainScanB200Error) as ei:
        Engine(0)
    assert ei.value.code == 1 and "no CPU fallback" in str(ei.value)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "strainscan_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")) or f == "Makefile":
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "oracle" not in txt.lower().replace("oracle/ is", ""), os.path.join(dp, f)


def test_shard_ranges_partition_the_records():
    """Host-only helper: ranges are record aligned, disjoint, and cover the text, even when quality
    lines open with '@' or '+'."""
    import numpy as np
    from strainscan_b200 import dist
    from tests import util
    rng = np.random.default_rng(3)
    G = util.rand_genome(rng, 5000)
    fq = util.make_reads(rng, G, 400, 80, var_len=True)
    recs = fq.count(b"\n") // 4
    for n in (1, 2, 3, 7, 64, 1000):
        pos, total = 0, 0
        for s in range(n):
            lo, hi = dist.shard_range(fq, s, n)
            assert lo == pos and hi >= lo
            chunk = fq[lo:hi]
            if chunk:
                assert chunk[:1] == b"@" and chunk.count(b"\n") % 4 == 0
                lines = chunk.split(b"\n")
                assert all(lines[i][:1] == b"+" for i in range(2, len(lines) - 1, 4))
            total += chunk.count(b"\n") // 4
            pos = hi
        assert pos == len(fq) and total == recs
