"""Shared helpers for the tests (oracle <-> package config mapping, synthetic inputs)."""
import numpy as np
import torch

from oracle import f5_oracle as O


def ocfg_of(cfg) -> O.DiTConfig:
    return O.DiTConfig(dim=cfg.dim, depth=cfg.depth, heads=cfg.heads, ff_mult=cfg.ff_mult,
                       text_num_embeds=cfg.text_num_embeds, text_dim=cfg.text_dim, conv_layers=cfg.conv_layers)


def rel(a: torch.Tensor, b: torch.Tensor) -> float:
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def make_dit(cfg, W, device="cuda", fused_adaln=True):
    from f5_tts_mlx_b200 import DiT
    return DiT(dim=cfg.dim, depth=cfg.depth, heads=cfg.heads, ff_mult=cfg.ff_mult, mel_dim=cfg.mel_dim,
               text_num_embeds=cfg.text_num_embeds, text_dim=cfg.text_dim, conv_layers=cfg.conv_layers,
               device=device, fused_adaln=fused_adaln).load_weights(W)


def synth_audio(length: int, seed: int = 0) -> torch.Tensor:
    rng = np.random.default_rng(seed)
    t = np.arange(length) / 24000.0
    f0 = 110 + 110 * rng.random()
    x = sum(np.sin(2 * np.pi * f0 * (h + 1) * t + rng.random() * 6.28) / (h + 1) for h in range(8))
    x = x * (0.6 + 0.4 * np.sin(2 * np.pi * 1.3 * t)) + 0.01 * rng.standard_normal(length)
    return torch.from_numpy((x * 0.1 / np.sqrt(np.mean(x ** 2))).astype(np.float32))
