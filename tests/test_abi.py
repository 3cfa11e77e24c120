"""The C-ABI library loads on a CPU-only box and exports every symbol include/vrag_b200.h declares;
compute entry points fail loudly (no fallback) when there is no GPU."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT, gpu_available
from visual_rag_b200 import _native


def header_functions():
    src = open(os.path.join(ROOT, "include", "vrag_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vrag_[a-z0-9_]+)\s*\(", src)))


def test_library_is_built_in_tree():
    assert os.path.exists(_native.lib_path()), "run `python -c 'import __graft_entry__ as g; g.build()'` first"
    assert os.path.dirname(_native.lib_path()).startswith(ROOT)


def test_every_declared_symbol_is_exported_and_bound():
    lib = _native.load()
    names = header_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/vrag_b200.h but not exported"
    assert sorted(_native.SIGNATURES.keys()) == names, "ctypes binding and header disagree"
    assert lib.vrag_abi_version() >= 1


@pytest.mark.skipif(gpu_available(), reason="CPU-only behaviour")
def test_no_gpu_means_error_not_fallback():
    lib = _native.load()
    h = C.c_void_p()
    rc = lib.vrag_corpus_create(0, 0, C.byref(h))
    assert rc != 0
    assert b"no CUDA device" in lib.vrag_last_error() or b"CUDA" in lib.vrag_last_error()
    from visual_rag_b200.corpus import GpuCorpus

    with pytest.raises(_native.VragError):
        GpuCorpus(0)


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "visual-rag-toolkit_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("the oracle", "").replace("oracle will see", ""), f"{f} mentions oracle/"


# ------------------------------------------------------------------ a host with no Python in it (tests/c_host/vrag_host.c)
def _c_host():
    import subprocess

    d = os.path.join(ROOT, "tests", "c_host")
    subprocess.run(["make", "-C", d], check=True, capture_output=True)   # -std=c99 -pedantic -Werror: the header is plain C
    return os.path.join(d, "vrag_host")


def test_plain_c_host_builds_against_the_header_and_loads_the_library():
    import subprocess

    out = subprocess.run([_c_host(), "abi"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and out.stdout.startswith("abi "), out.stderr
    if not gpu_available():   # no fallback from C either
        bad = subprocess.run([_c_host(), "single", "100", "64"], capture_output=True, text=True, timeout=60)
        assert bad.returncode != 0 and "no CUDA device" in bad.stderr


@pytest.mark.gpu
def test_plain_c_host_two_stage_search():
    """The C program builds a corpus on the device, pools it, runs two-stage and exhaustive searches through the C ABI and
    checks the lists against vrag_score + a stable host sort (exit code 0 = all equal)."""
    import subprocess

    out = subprocess.run([_c_host(), "single", "20000", "256"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "equal score + stable sort" in out.stdout
