"""F5TTS.from_pretrained — host-side mirror of cfm.py:404-520 (load-time only).

There is no network in the build/bench environment, so `hf_model_name_or_path` is a local directory
(or, when `huggingface_hub` is importable and online, a hub repo id exactly like the reference's
`fetch_from_hub`, utils.py:179-192) holding `model_v1.safetensors` (or `model_v1_{4,8}b.safetensors`, MLX affine
quantised, dequantised at load), `vocab.txt` and optionally `duration_v2.safetensors`.  The vocoder is resolved
separately, like the reference's `Vocos.from_pretrained("lucasnewman/vocos-mel-24khz")` (cfm.py:446): a
`vocos.safetensors` / `vocos-mel-24khz/` next to the model, `$F5_VOCOS_PATH`, or that hub repo; if none can be
found the load FAILS (the reference always has a vocoder) unless `vocoder=False` is passed explicitly.
The special name "random" builds the base model with seeded random weights (what the tests and bench.py use); its
vocabulary is `vocab_path` / `$F5_VOCAB_PATH` when given, else a printable-ASCII table, stated in `vocab_source`.

Multi-GPU: only rank 0 reads / converts / packs; every other rank allocates the same packed layout
and receives it in ONE broadcast (parallel.load_weights_distributed).
"""
from __future__ import annotations

from pathlib import Path
from typing import Optional

import torch

from .dit import DiT
from .parallel import load_weights_distributed
from .vocos import Vocos
from .weights import (BASE_CONFIG, VocosConfig, Weights, convert_upstream_keys, dequantize_mlx_checkpoint,
                      random_dit_weights, random_vocos_weights)

VOCOS_REPO = "lucasnewman/vocos-mel-24khz"
VOCOS_FILES = ("vocos.safetensors", "vocos-mel-24khz/model.safetensors", "vocos-mel-24khz/vocos.safetensors")


def _resolve(path_or_repo: str, quantization_bits: Optional[int]) -> Optional[Path]:
    p = Path(path_or_repo)
    if p.is_dir():
        return p
    try:                                               # same behaviour as utils.py:179-192 when online
        from huggingface_hub import snapshot_download  # type: ignore
        fn = "model_v1.safetensors" if quantization_bits is None else f"model_v1_{quantization_bits}b.safetensors"
        return Path(snapshot_download(repo_id=path_or_repo, allow_patterns=[fn, "duration_v2.safetensors", "*.txt"]))
    except Exception:
        return None


def _resolve_vocos(model_dir: Optional[Path]) -> Optional[Path]:
    """The vocoder checkpoint: next to the model, $F5_VOCOS_PATH (file or directory), or the hub repo the reference
    uses (cfm.py:446)."""
    import os
    cands = []
    env = os.environ.get("F5_VOCOS_PATH")
    if env:
        e = Path(env)
        cands += [e] if e.is_file() else [e / "model.safetensors", e / "vocos.safetensors"]
    if model_dir is not None:
        cands += [model_dir / f for f in VOCOS_FILES]
    for c in cands:
        if c.is_file():
            return c
    try:
        from huggingface_hub import snapshot_download  # type: ignore
        d = Path(snapshot_download(repo_id=VOCOS_REPO, allow_patterns=["*.safetensors", "*.yaml", "*.json"]))
        for c in sorted(d.glob("*.safetensors")):
            return c
    except Exception:
        pass
    return None


def ascii_vocab() -> dict:
    """Printable ASCII, in the file layout read_vocab() expects (trailing '' entry)."""
    chars = [chr(i) for i in range(32, 127)] + [""]
    return {v: i for i, v in enumerate(chars)}


def read_vocab(vocab_path: Path) -> dict:
    """cfm.py:418-421 — note the trailing '' entry: text_num_embeds = len(vocab) - 1."""
    vocab = {v: i for i, v in enumerate(Path(vocab_path).read_text().split("\n"))}
    if len(vocab) == 0:
        raise ValueError(f"Could not load vocab from {vocab_path}")
    return vocab


def convert_vocos_upstream(w: Weights) -> Weights:
    """Upstream (PyTorch) Vocos checkpoint names/layouts -> the names used here; conv weights
    (O, I/g, K) -> (O, K, I/g)."""
    out: Weights = {}
    for k, v in w.items():
        if k.startswith("feature_extractor") or "istft.window" in k:
            continue
        if k.endswith("dwconv.weight") or k == "backbone.embed.weight":
            v = v.transpose(1, 2)
        out["vocos." + k] = v
    return out


def from_pretrained(cls, hf_model_name_or_path: str, convert_weights=None, quantization_bits: Optional[int] = None,
                    device: str | torch.device = "cuda", vocab_path: Optional[str] = None, vocoder=None):
    """`vocoder`: None = resolve and REQUIRE one (reference behaviour), False = none (sample() returns mels),
    or a callable mel -> waveform."""
    import os
    if quantization_bits is not None and quantization_bits not in (4, 8):
        raise ValueError(f"quantization_bits must be 4 or 8 (generate.py --q), got {quantization_bits}")
    if hf_model_name_or_path == "random":
        vp = vocab_path or os.environ.get("F5_VOCAB_PATH")
        if vp is not None:
            vocab, vocab_source = read_vocab(Path(vp)), str(vp)     # missing file -> FileNotFoundError, not a silent switch
        else:
            vocab, vocab_source = ascii_vocab(), "ascii"
        cfg = BASE_CONFIG
        dit = DiT(dim=cfg.dim, depth=cfg.depth, heads=cfg.heads, ff_mult=cfg.ff_mult, text_dim=cfg.text_dim,
                  conv_layers=cfg.conv_layers, text_num_embeds=cfg.text_num_embeds, device=device)
        load_weights_distributed(dit, lambda: random_dit_weights(cfg, seed=1234))
        if vocoder is None:
            vocoder = Vocos(VocosConfig(), device).load_weights(random_vocos_weights()).decode
        m = cls(transformer=dit, vocab_char_map=vocab, vocoder=vocoder or None)
        m.vocab_source = vocab_source
        return m

    path = _resolve(hf_model_name_or_path, quantization_bits)
    if path is None:
        raise ValueError(f"Could not find model {hf_model_name_or_path}")        # cfm.py:413-414
    from safetensors.torch import load_file
    vocab = read_vocab(path / "vocab.txt")
    convert = True if convert_weights is None else convert_weights               # cfm.py:455
    model_file = "model_v1.safetensors"
    if quantization_bits is not None:                                            # cfm.py:450-453
        model_file, convert = f"model_v1_{quantization_bits}b.safetensors", False
    dit = DiT(dim=1024, depth=22, heads=16, ff_mult=2, text_dim=512, conv_layers=4,
              text_num_embeds=len(vocab) - 1, text_mask_padding=True, device=device)     # cfm.py:459-469

    def weights_fn() -> Weights:
        w = load_file(str(path / model_file))
        if quantization_bits is not None:     # nn.quantize + load_weights (cfm.py:510-517): dense again at pack time
            return dequantize_mlx_checkpoint(w, quantization_bits)
        return convert_upstream_keys(w) if convert else w

    load_weights_distributed(dit, weights_fn)
    if vocoder is None:                                                          # cfm.py:446: always present
        vpath = _resolve_vocos(path)
        if vpath is None:
            raise FileNotFoundError(
                f"no Vocos checkpoint found (looked for {', '.join(VOCOS_FILES)} in {path}, $F5_VOCOS_PATH and the hub "
                f"repo {VOCOS_REPO}); pass vocoder=False to get mel spectrograms from sample() instead")
        vocoder = Vocos(VocosConfig(), device).load_weights(convert_vocos_upstream(load_file(str(vpath)))).decode
    elif vocoder is False:
        vocoder = None
    duration_predictor = None
    dpath = path / "duration_v2.safetensors"
    if dpath.exists():                                                           # cfm.py:425-442
        from .duration import DurationPredictor, DurationTransformer
        duration_predictor = DurationPredictor(
            transformer=DurationTransformer(dim=512, depth=8, heads=8, text_dim=512, ff_mult=2, conv_layers=2,
                                            text_num_embeds=len(vocab) - 1),
            vocab_char_map=vocab, device=device).load_weights(load_file(str(dpath)))
    return cls(transformer=dit, vocab_char_map=vocab, vocoder=vocoder, duration_predictor=duration_predictor)
