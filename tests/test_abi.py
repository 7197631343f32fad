"""The C-ABI library loads, exports every symbol include/f5_b200.h declares, its struct layouts
match the ctypes mirrors, and — on a box without a GPU — every compute entry point fails loudly
instead of falling back to a CPU path.  No compute calls are made here."""
import ctypes as C
import os
import re

import pytest
import torch

from f5_tts_mlx_b200 import _lib
from f5_tts_mlx_b200.dit import DitBuffersC
from f5_tts_mlx_b200.duration import DurationBuffersC, DurationWeightsC
from f5_tts_mlx_b200.vocos import VocosBlockWeightsC, VocosBuffersC, VocosWeightsC
from f5_tts_mlx_b200.weights import ConvNextWeightsC, DitBlockWeightsC, DitWeightsC

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "f5_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(f5_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/f5_b200.h but not exported by libf5b200.so"
        assert n in _lib.SYMBOLS, f"{n} has no ctypes prototype in _lib.SYMBOLS"
    for n in _lib.SYMBOLS:
        assert n in names, f"{n} bound in _lib.py but not declared in the header"


def test_struct_layouts_match_ctypes():
    lib = _lib.load()
    out = (C.c_int32 * 10)()
    assert lib.f5_struct_sizes(out, 10) == 10
    mirrors = [_lib.GemmArgs, ConvNextWeightsC, DitBlockWeightsC, DitWeightsC, DitBuffersC, VocosBlockWeightsC,
               VocosWeightsC, VocosBuffersC, DurationWeightsC, DurationBuffersC]
    for got, m in zip(list(out), mirrors):
        assert got == C.sizeof(m), f"{m.__name__}: C sizeof {got} != ctypes {C.sizeof(m)}"


def test_abi_version():
    assert _lib.load().f5_abi_version() >= 1000


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback_entry_points_fail_loudly():
    lib = _lib.load()
    assert lib.f5_device_check() == -3                     # F5_ERR_NO_DEVICE
    g = _lib.GemmArgs()
    assert lib.f5_gemm_bf16(C.byref(g), None) == -3
    assert b"no CPU fallback" in lib.f5_last_error() or b"sm_100a" in lib.f5_last_error()
    assert lib.f5_attention_fwd(None, 0, None, 0, 1, 1, 1, 64, None, None) == -3
    assert lib.f5_dit_forward(None, None, 0, None) == -3
    assert lib.f5_mel_forward(None, 1, 1, None, None, 100, 256, None, 1, None) == -3
    assert lib.f5_vocos_decode(None, None, None, None, None) == -3
    with pytest.raises(_lib.F5Error):
        _lib.check(lib.f5_device_check())


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a box without a GPU")
def test_python_surface_refuses_cpu_tensors():
    from f5_tts_mlx_b200 import MelSpec
    with pytest.raises(_lib.F5Error):
        MelSpec()(torch.zeros(2048))


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: the product path must not route through it."""
    pkg = os.path.join(ROOT, "f5_tts_mlx_b200")
    pat = re.compile(r"^\s*(from\s+oracle|import\s+oracle)", re.M)
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, fn)).read()
                assert not pat.search(src), f"{fn} imports the oracle"


def test_header_is_plain_c99_and_links_from_c(tmp_path):
    """include/f5_b200.h compiles as strict C99 and a C program links against libf5b200.so: the boundary has no
    C++ or torch types.  The consumer also checks that the C compiler's struct layout equals the library's."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    _lib.load()                                            # builds the library if needed
    pkg = os.path.join(ROOT, "f5_tts_mlx_b200")
    exe = str(tmp_path / "consumer")
    cmd = [gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c_abi", "consumer.c"), "-o", exe, "-L", pkg, "-l:libf5b200.so",
           f"-Wl,-rpath,{pkg}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "abi=" in r.stdout and "structs=10" in r.stdout
    if not torch.cuda.is_available():
        assert "device_check=-3" in r.stdout and "gemm_rc=-3" in r.stdout     # F5_ERR_NO_DEVICE, no CPU fallback
