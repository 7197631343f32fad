"""GPU bring-up: mel front-end and Vocos back-end vs the CPU oracle."""
import sys, wave, struct
import numpy as np, torch
from f5_tts_mlx_b200 import MelSpec
from f5_tts_mlx_b200.vocos import Vocos
from f5_tts_mlx_b200.weights import VocosConfig, random_vocos_weights
from oracle import f5_oracle as O
dev = "cuda"; fails = []
def rel(a, b): return ((a - b).norm() / (b.norm() + 1e-30)).item()

rng = np.random.default_rng(0)
def synth(L):
    t = np.arange(L) / 24000.0
    f0 = 110 + 110 * rng.random()
    x = sum(np.sin(2 * np.pi * f0 * (h + 1) * t + rng.random() * 6.28) / (h + 1) for h in range(8))
    x = x * (0.6 + 0.4 * np.sin(2 * np.pi * 1.3 * t)) + 0.01 * rng.standard_normal(L)
    return torch.from_numpy((x * 0.1 / np.sqrt(np.mean(x ** 2))).astype(np.float32))

ms = MelSpec()
for L in [240000, 127987, 1024, 700, 84000]:
    x = synth(L)
    ref = O.log_mel_spectrogram(x)
    got = ms(x.to(dev)).cpu()
    ok = got.shape == ref.shape and (got - ref).abs().max().item() < 2e-3
    print(f"{'OK  ' if ok else 'FAIL'} mel L={L}: shape {tuple(got.shape)} vs {tuple(ref.shape)} max_abs={(got-ref).abs().max().item() if got.shape==ref.shape else -1:.3e}", flush=True)
    if not ok: fails.append(("mel", L))
xb = torch.stack([synth(24000), synth(24000)])
ok = (ms(xb.to(dev)).cpu() - O.log_mel_spectrogram(xb)).abs().max().item() < 2e-3
print(f"{'OK  ' if ok else 'FAIL'} mel batched"); 
if not ok: fails.append("mel batched")

# Vocos
for norm, trim in [("window", False), ("window_sq", True)]:
    vc = VocosConfig(istft_norm=norm, istft_trim=trim)
    ovc = O.VocosConfig(istft_norm=norm, istft_trim=trim)
    W = random_vocos_weights(vc)
    voc = Vocos(vc, dev).load_weights(W)
    for n in [1, 5, 499, 937]:
        mel = (torch.randn(1, n, 100) * 2.24 - 1.27).clamp(-11.5, 5)
        ref = O.vocos_decode(mel, W, ovc)
        ref16 = O.vocos_decode(mel, W, ovc, O.Precision(True))
        got = voc.decode(mel.to(dev)).cpu()
        drift = rel(ref16, ref)
        ok = got.shape == ref.shape and rel(got, ref) < max(3 * drift, 2e-2)
        snr = 10 * np.log10((ref.norm() ** 2 / ((got - ref).norm() ** 2 + 1e-30)).item()) if got.shape == ref.shape else -1
        print(f"{'OK  ' if ok else 'FAIL'} vocos {norm} trim={trim} n={n}: shape {tuple(got.shape)} vs {tuple(ref.shape)} rel={rel(got, ref) if got.shape==ref.shape else -1:.3e} drift={drift:.3e} SNR={snr:.1f} dB", flush=True)
        if not ok: fails.append(("vocos", norm, n))
    # pure ISTFT exactness: feed the oracle's own head output
    # batch > 1
    melb = (torch.randn(3, 200, 100) * 2.24 - 1.27).clamp(-11.5, 5)
    gotb = voc.decode(melb.to(dev)).cpu()
    refb = torch.stack([O.vocos_decode(melb[i:i+1], W, ovc) for i in range(3)])
    ok = rel(gotb, refb) < 2e-2
    print(f"{'OK  ' if ok else 'FAIL'} vocos batched rel={rel(gotb, refb):.3e}")
    if not ok: fails.append(("vocos batched", norm))
print("FAILS:", fails); sys.exit(1 if fails else 0)
