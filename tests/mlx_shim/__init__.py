"""Torch-backed stand-ins for the reference's unavailable imports (`mlx.core`, `mlx.nn`, `einx`,
`vocos_mlx`, `jieba`, `pypinyin`, `sounddevice`, `soundfile`) so that the UNMODIFIED files under
/root/reference/f5_tts_mlx/ can be imported and executed in this image.  TEST INFRASTRUCTURE: used
only by tests/golden/make_ref_golden.py (fixture generator, build container) and
tests/test_ref_pins.py (live comparison when /root/reference is present).  Nothing in the product
package or in bench.py imports it.

    import mlx_shim as shim                # with tests/ on sys.path
    ref = shim.import_reference()          # -> the real f5_tts_mlx package, running on the shim
    ref.dit.DiT(...), ref.cfm.F5TTS(...), ref.audio.log_mel_spectrogram(...)
"""
from __future__ import annotations

import importlib
import sys
import types
import wave as _wave
from pathlib import Path

import numpy as np

from . import core, einx as _einx, nn

REFERENCE_ROOT = Path("/root/reference")


def _module(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    return m


def _jieba_cut(text: str):
    """jieba.cut stand-in for ASCII text: split into words / whitespace / punctuation runs the way
    jieba's default mode does for Latin text (words, single spaces and symbols are separate tokens)."""
    import re
    return [s for s in re.findall(r"[A-Za-z0-9]+|\s|[^A-Za-z0-9\s]", text)]


class _VocosStub:
    """`vocos_mlx.Vocos` is third-party and absent; cfm.py only needs the symbol to import."""

    @classmethod
    def from_pretrained(cls, *a, **k):
        raise RuntimeError("mlx_shim: vocos_mlx is not available (no network, not vendored)")


def _sf_read(path):
    with _wave.open(str(path), "rb") as f:
        assert f.getsampwidth() == 2 and f.getnchannels() == 1
        pcm = np.frombuffer(f.readframes(f.getnframes()), dtype=np.int16)
        return pcm.astype(np.float64) / 32768.0, f.getframerate()


def _sf_write(path, data, sr):
    pcm = np.clip(np.asarray(data, dtype=np.float64) * 32767.0, -32768, 32767).astype(np.int16)
    with _wave.open(str(path), "wb") as f:
        f.setnchannels(1); f.setsampwidth(2); f.setframerate(sr); f.writeframes(pcm.tobytes())


def install() -> None:
    """Register the stand-ins in sys.modules (idempotent; real packages, if present, win)."""
    def absent(name):
        if name in sys.modules:
            return False
        try:
            return importlib.util.find_spec(name) is None
        except (ImportError, ValueError):
            return True

    if absent("mlx"):
        pkg = _module("mlx", __path__=[])
        pkg.core, pkg.nn = core, nn
        sys.modules["mlx"] = pkg
        sys.modules["mlx.core"] = core
        sys.modules["mlx.nn"] = nn
    if absent("einx"):
        sys.modules["einx"] = _einx
    if absent("vocos_mlx"):
        sys.modules["vocos_mlx"] = _module("vocos_mlx", Vocos=_VocosStub)
    if absent("jieba"):
        sys.modules["jieba"] = _module("jieba", setLogLevel=lambda *_: None, cut=_jieba_cut)
    if absent("pypinyin"):
        class Style:
            TONE3 = 8

        def lazy_pinyin(*a, **k):
            raise RuntimeError("mlx_shim: pypinyin is not available (ASCII text only)")
        sys.modules["pypinyin"] = _module("pypinyin", lazy_pinyin=lazy_pinyin, Style=Style)
    if absent("sounddevice"):
        sys.modules["sounddevice"] = _module("sounddevice", OutputStream=None)
    if absent("soundfile"):
        sys.modules["soundfile"] = _module("soundfile", read=_sf_read, write=_sf_write)


def reference_available() -> bool:
    return (REFERENCE_ROOT / "f5_tts_mlx" / "dit.py").exists()


def import_reference():
    """Import the real, unmodified `f5_tts_mlx` package from /root/reference on top of the shim and
    return a namespace with its submodules."""
    if not reference_available():
        raise RuntimeError(f"{REFERENCE_ROOT} is not present on this machine")
    install()
    if str(REFERENCE_ROOT) not in sys.path:
        sys.path.insert(0, str(REFERENCE_ROOT))
    mods = {}
    for name in ("utils", "rope", "convnext_v2", "audio", "dit", "duration", "cfm", "generate"):
        mods[name] = importlib.import_module(f"f5_tts_mlx.{name}")
    return types.SimpleNamespace(mx=core, nn=nn, **mods)
