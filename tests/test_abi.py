"""The C-ABI library: it loads, exports every symbol include/elba_fe.h declares, and fails LOUDLY without a GPU
(no CPU fallback anywhere in the product path).  No compute calls here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "elba_fe.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(elba_fe_[a-zA-Z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from elba_b200 import frontend
    L = frontend.load_library()
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), f"libelba_fe.so does not export {n}"
    assert set(names) == set(frontend.ABI_SYMBOLS)
    assert L.elba_fe_version() == 1


def test_no_cpu_fallback():
    import torch
    from elba_b200 import frontend
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(frontend.FrontEndError) as e:
        frontend.Context(frontend.Params(k=17, lower=2, upper=8))
    assert "no CUDA device" in str(e.value) or "-2" in str(e.value)


def test_product_never_touches_the_oracle():
    """Nothing under elba_b200/ may import, link or execute oracle/."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "elba_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("the oracle", "").replace("CPU oracle", "").lower() or f == "frontend.py" and "import oracle" not in text, f
                assert "import oracle" not in text and "from oracle" not in text and "oracle/" not in text, f


def test_parameter_validation_is_in_the_library():
    """elba_fe_create rejects out-of-contract parameters before touching the device."""
    from elba_b200 import frontend
    L = frontend.load_library()
    for bad in (dict(k=33), dict(k=2), dict(lower=1), dict(lower=9, upper=8), dict(seed_count=3), dict(stride=0)):
        kw = dict(k=17, stride=1, seed_count=2, lower=2, upper=8)
        kw.update(bad)
        cfg = frontend._Config(kw["k"], kw["stride"], kw["seed_count"], kw["lower"], kw["upper"], 0, 0, 0)
        h = ctypes.c_void_p()
        assert L.elba_fe_create(ctypes.byref(cfg), ctypes.byref(h)) == -1, bad
        assert L.elba_fe_last_error(None)


@pytest.mark.reference
def test_cpp_shim_compiles_against_the_reference_headers():
    """elba_b200/host/elba_fe_shim.cpp re-implements the reference's five driver functions on the C ABI.  It must
    compile against the reference's OWN headers, unmodified (KmerOps.hpp, SharedSeeds.hpp, DnaBuffer.hpp, Kmer.hpp),
    with the MPI / CombBLAS stand-ins of the test tree on the include path (neither library exists in this image).
    Build container only: /root/reference is not on the GPU box."""
    import subprocess
    ref = "/root/reference"
    if not os.path.exists(os.path.join(ref, "include", "KmerOps.hpp")):
        pytest.skip("reference tree not present")
    for k, lo, up, extra in ((31, 15, 35, []), (17, 2, 8, []), (17, 2, 8, ["-DELBA_FE_SHIM_ALIGN"])):      # the last one also replaces PairwiseAlignment
        cmd = ["g++", "-fsyntax-only", "-std=c++17", "-Wno-deprecated-declarations", f"-DKMER_SIZE={k}", f"-DLOWER_KMER_FREQ={lo}", f"-DUPPER_KMER_FREQ={up}", "-DLOG_LEVEL=0", *extra,
               "-I", os.path.join(ROOT, "oracle", "stubs"), "-I", os.path.join(ref, "include"), "-I", os.path.join(ref, "src"),
               "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "elba_b200", "host", "elba_fe_shim.cpp")]
        p = subprocess.run(cmd, capture_output=True, text=True)
        assert p.returncode == 0, p.stderr[-3000:]
