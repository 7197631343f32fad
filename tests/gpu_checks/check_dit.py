"""GPU bring-up: DiT forward and sample() of the CUDA path vs the CPU oracle (gate config)."""
import sys, time
import torch
from f5_tts_mlx_b200 import DiT, F5TTS, GATE_CONFIG, BASE_CONFIG
from f5_tts_mlx_b200.weights import random_dit_weights
from oracle import f5_oracle as O

torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"
fails = []

def rel(a, b):
    return ((a - b).norm() / (b.norm() + 1e-30)).item()

def mk(cfg, dev=dev):
    W = random_dit_weights(cfg, seed=1234)
    m = DiT(dim=cfg.dim, depth=cfg.depth, heads=cfg.heads, ff_mult=cfg.ff_mult, mel_dim=cfg.mel_dim,
            text_num_embeds=cfg.text_num_embeds, text_dim=cfg.text_dim, conv_layers=cfg.conv_layers, device=dev)
    m.load_weights(W)
    return W, m

cfg = GATE_CONFIG
ocfg = O.DiTConfig(dim=cfg.dim, depth=cfg.depth, heads=cfg.heads, ff_mult=cfg.ff_mult, text_num_embeds=cfg.text_num_embeds,
                   text_dim=cfg.text_dim, conv_layers=cfg.conv_layers)
W, model = mk(cfg)
g = torch.Generator().manual_seed(1)

# ---- single forward, B=1 and B=2 (masked), drop flags ----
for (B, N, nt, drops, lens) in [(1, 200, 40, (False, False), None), (1, 200, 40, (True, True), None),
                                (2, 300, 60, (False, False), [300, 211])]:
    x = torch.randn(B, N, 100, generator=g); cond = torch.randn(B, N, 100, generator=g) * 2 - 1
    text = torch.randint(0, 2545, (B, nt), generator=g, dtype=torch.int32)
    if B > 1: text[1, nt - 17:] = -1
    t = torch.tensor(0.37)
    mask = None
    if lens is not None:
        mask = torch.arange(N)[None, :] < torch.tensor(lens)[:, None]
    ref = O.dit_forward(x, cond, text, t, drops[0], drops[1], mask, W, ocfg)
    ref16 = O.dit_forward(x, cond, text, t, drops[0], drops[1], mask, W, ocfg, O.Precision(True))
    got = model(x.to(dev), cond.to(dev), text.to(dev), t, drops[0], drops[1], mask.to(dev) if mask is not None else None).cpu()
    if mask is not None:   # padded query rows: reference computes them too; compare valid rows only as well
        vm = mask[..., None].float()
    r_fp32, r_emul, drift = rel(got, ref), rel(got, ref16), rel(ref16, ref)
    ok = r_fp32 < max(3 * drift, 2e-2) and got.isfinite().all().item()
    print(f"{'OK  ' if ok else 'FAIL'} dit fwd B{B} N{N} drops={drops} lens={lens}: rel(cuda,fp32)={r_fp32:.3e} rel(cuda,bf16emu)={r_emul:.3e} drift(bf16emu,fp32)={drift:.3e}", flush=True)
    if not ok: fails.append(("fwd", B, N, drops))

# ---- sample(): config 1 shape (N=937 is slow-ish on CPU; use the real gate config, Euler, 8 grid pts) ----
for (method, steps, N, nref) in [("euler", 8, 937, 328), ("midpoint", 4, 300, 100), ("rk4", 3, 200, 80)]:
    cond = (torch.randn(1, nref, 100, generator=g) * 2.24 - 1.27).clamp(-11.51, 5)
    text = torch.randint(0, 2545, (1, 152 if N == 937 else 40), generator=g, dtype=torch.int32)
    y0 = torch.randn(1, 100, N, generator=g).permute(0, 2, 1).contiguous()
    t0 = time.time()
    ref_out, ref_traj = O.sample(cond, text, N, W, ocfg, steps=steps, method=method, cfg_strength=2.0, sway_sampling_coef=-1.0, y0=y0)
    t_cpu = time.time() - t0
    ref16_out, _ = O.sample(cond, text, N, W, ocfg, steps=steps, method=method, cfg_strength=2.0, sway_sampling_coef=-1.0, y0=y0, prec=O.Precision(True))
    f5 = F5TTS(model)
    for graph in (False, True):
        f5.use_cuda_graph = graph
        out, traj = f5.sample(cond.to(dev), text, N, steps=steps, method=method, cfg_strength=2.0, sway_sampling_coef=-1.0, y0=y0)
        torch.cuda.synchronize()
        out, traj = out.cpu(), traj.cpu()
        r, r16, drift = rel(out, ref_out), rel(out, ref16_out), rel(ref16_out, ref_out)
        mx = (out - ref_out).abs().max().item()
        ok = r < max(3 * drift, 2e-2) and traj.shape == ref_traj.shape
        print(f"{'OK  ' if ok else 'FAIL'} sample {method} steps={steps} N={N} graph={graph}: rel(cuda,fp32)={r:.3e} rel(cuda,bf16emu)={r16:.3e} drift={drift:.3e} max_abs={mx:.3e} traj_rel={rel(traj, ref_traj):.3e} (oracle {t_cpu:.1f}s)", flush=True)
        if not ok: fails.append(("sample", method, graph))

# ---- base config timing (B=1, N=937, euler 32) ----
try:
    cfgb = BASE_CONFIG
    Wb, mb = mk(cfgb)
    f5 = F5TTS(mb)
    cond = (torch.randn(1, 328, 100, generator=g) * 2.24 - 1.27).clamp(-11.51, 5)
    text = torch.randint(0, 2545, (1, 152), generator=g, dtype=torch.int32)
    y0 = torch.randn(1, 937, 100, generator=g)
    for it in range(4):
        torch.cuda.synchronize(); t0 = time.time()
        out, _ = f5.sample(cond.to(dev), text, 937, steps=32, method="euler", cfg_strength=2.0, y0=y0, return_trajectory=False)
        torch.cuda.synchronize(); dt = time.time() - t0
        print(f"base B=1 N=937 euler32 iter{it}: {dt*1e3:.1f} ms -> {937/dt:.0f} mel-frames/s finite={out.isfinite().all().item()}", flush=True)
except Exception as e:
    import traceback; traceback.print_exc(); fails.append(("base", str(e)))
print("FAILS:", fails)
sys.exit(1 if fails else 0)
