"""F5TTS.from_pretrained — host-side mirror of cfm.py:404-520 (load-time only).

There is no network in the build/bench environment, so `hf_model_name_or_path` is a local directory
(or, when `huggingface_hub` is importable and online, a hub repo id exactly like the reference's
`fetch_from_hub`, utils.py:179-192) holding `model_v1.safetensors`, `vocab.txt` and optionally
`duration_v2.safetensors` / `vocos.safetensors`.  The special name "random" builds the base model
with seeded random weights (what the tests and bench.py use).

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
from .weights import BASE_CONFIG, VocosConfig, Weights, convert_upstream_keys, random_dit_weights, random_vocos_weights

LOCAL_VOCAB = Path("/root/reference/data/Emilia_ZH_EN_pinyin/vocab.txt")


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
                    device: str | torch.device = "cuda"):
    if quantization_bits is not None:
        raise NotImplementedError("MLX affine 4/8-bit checkpoints (cfm.py:510-515) are out of scope on this path")
    if hf_model_name_or_path == "random":
        vocab = read_vocab(LOCAL_VOCAB) if LOCAL_VOCAB.exists() else {chr(i): i for i in range(32, 127)}
        cfg = BASE_CONFIG
        dit = DiT(dim=cfg.dim, depth=cfg.depth, heads=cfg.heads, ff_mult=cfg.ff_mult, text_dim=cfg.text_dim,
                  conv_layers=cfg.conv_layers, text_num_embeds=cfg.text_num_embeds, device=device)
        load_weights_distributed(dit, lambda: random_dit_weights(cfg, seed=1234))
        vocos = Vocos(VocosConfig(), device).load_weights(random_vocos_weights())
        return cls(transformer=dit, vocab_char_map=vocab, vocoder=vocos.decode)

    path = _resolve(hf_model_name_or_path, quantization_bits)
    if path is None:
        raise ValueError(f"Could not find model {hf_model_name_or_path}")        # cfm.py:413-414
    from safetensors.torch import load_file
    vocab = read_vocab(path / "vocab.txt")
    convert = True if convert_weights is None else convert_weights               # cfm.py:455
    dit = DiT(dim=1024, depth=22, heads=16, ff_mult=2, text_dim=512, conv_layers=4,
              text_num_embeds=len(vocab) - 1, text_mask_padding=True, device=device)     # cfm.py:459-469

    def weights_fn() -> Weights:
        w = load_file(str(path / "model_v1.safetensors"))
        return convert_upstream_keys(w) if convert else w

    load_weights_distributed(dit, weights_fn)
    vocoder = None
    vpath = path / "vocos.safetensors"
    if vpath.exists():
        vocoder = Vocos(VocosConfig(), device).load_weights(convert_vocos_upstream(load_file(str(vpath)))).decode
    duration_predictor = None
    dpath = path / "duration_v2.safetensors"
    if dpath.exists():                                                           # cfm.py:425-442
        from .duration import DurationPredictor, DurationTransformer
        duration_predictor = DurationPredictor(
            transformer=DurationTransformer(dim=512, depth=8, heads=8, text_dim=512, ff_mult=2, conv_layers=2,
                                            text_num_embeds=len(vocab) - 1),
            vocab_char_map=vocab, device=device).load_weights(load_file(str(dpath)))
    return cls(transformer=dit, vocab_char_map=vocab, vocoder=vocoder, duration_predictor=duration_predictor)
