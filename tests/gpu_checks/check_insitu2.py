"""In-situ GEMM phase times inside the captured graph of one base-model B=1 step, averaged per GEMM of the DiT block
(QKV, out-proj, FF1, FF2) over all layers of the second forward.  F5_FUSED=0 -> separate LayerNorm kernels."""
import os, ctypes as C, numpy as np, torch
from f5_tts_mlx_b200 import DiT, F5TTS, BASE_CONFIG, _lib
from f5_tts_mlx_b200.weights import random_dit_weights
dev = "cuda"; lib = _lib.load()
cfg = BASE_CONFIG
fused = os.environ.get("F5_FUSED", "1") != "0"
B = int(os.environ.get("F5_B", "1"))
fp8 = os.environ.get("F5_FP8", "0") == "1"
model = DiT(dim=cfg.dim, depth=cfg.depth, heads=cfg.heads, ff_mult=cfg.ff_mult, text_num_embeds=cfg.text_num_embeds,
            text_dim=cfg.text_dim, conv_layers=cfg.conv_layers, device=dev, fused_adaln=fused, fp8=fp8).load_weights(random_dit_weights(cfg))
g = torch.Generator().manual_seed(3)
N, nref = 937, 328
cond = (torch.randn(B, nref, 100, generator=g) * 2.24 - 1.27).to(dev)
text = torch.randint(0, 2545, (B, 152), generator=g, dtype=torch.int32)
y0 = torch.randn(B, N, 100, generator=g).to(dev)
f5 = F5TTS(model)
NCALL, NCTA = 400, 1024
ts = torch.zeros(NCALL, NCTA * 10, dtype=torch.int64, device=dev)
kw = dict(steps=4, method="euler", cfg_strength=2.0, return_trajectory=False, y0=y0)
f5.use_cuda_graph = False
f5.sample(cond, text, N, **kw)
plan = f5.last_plan
lib.f5_debug_gemm_ts(ts.data_ptr(), ts.stride(0) * 8, NCALL)
plan.capture(f5)
lib.f5_debug_gemm_ts(None, 0, 0)
for _ in range(3):
    plan.y.copy_(y0); plan.graph.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
plan.y.copy_(y0); e0.record(); plan.graph.replay(); e1.record(); torch.cuda.synchronize()
print(f"fp8={fp8} fused={fused} B={B} graph replay (3 intervals): {e0.elapsed_time(e1):.3f} ms")
t = ts.cpu().numpy().reshape(NCALL, NCTA, 10).astype(np.float64)
used = [i for i in range(NCALL) if t[i, :, 0].max() > 0]
pre = 2 * 4 + 1 + 1 + (2 * cfg.depth + 1 if fused else 0)      # precompute GEMMs (text blocks, hoist, mod table, LN tables)
per_fwd = 3 + 4 * cfg.depth + 1
names = {0: "QKV", 1: "out", 2: "FF1", 3: "FF2"}
acc = {k: [] for k in names.values()}
fwd0 = pre + per_fwd      # second forward
for l in range(cfg.depth):
    for j in range(4):
        i = used[fwd0 + 3 + 4 * l + j]
        m = t[i][:, 0] > 0
        a = t[i][m]
        beg = a[:, 2].min()       # first CTA past the PDL wait
        row = dict(ctas=int(m.sum()), prologue=(np.nanmean(np.where(a[:, 5] > 0, a[:, 5], np.nan)) - beg) / 1e3,
                   main=(np.nanmean(np.where(a[:, 6] > 0, a[:, 6], np.nan)) - np.nanmean(np.where(a[:, 5] > 0, a[:, 5], np.nan))) / 1e3,
                   epi=(a[:, 8].max() - np.nanmean(np.where(a[:, 6] > 0, a[:, 6], np.nan))) / 1e3,
                   total=(a[:, 9].max() - beg) / 1e3, start_skew=(a[:, 2].max() - beg) / 1e3)
        acc[names[j]].append(row)
tot = 0.0
for k, rows in acc.items():
    mean = {f: float(np.mean([r[f] for r in rows])) for f in rows[0]}
    tot += mean["total"]
    print(f"{k:4s} ctas {mean['ctas']:.0f}: wait->mma0 {mean['prologue']:.2f}  main {mean['main']:.2f}  mma_last->epi_end {mean['epi']:.2f}  "
          f"total(wait->exit) {mean['total']:.2f} us  (start skew {mean['start_skew']:.2f})")
print(f"sum of the block's 4 GEMMs: {tot:.2f} us")
# whole second forward: first GEMM start to last GEMM exit
i0, i1 = used[fwd0], used[fwd0 + per_fwd - 1]
f_beg = t[i0][t[i0][:, 0] > 0][:, 0].min(); f_end = t[i1][t[i1][:, 0] > 0][:, 9].max()
print(f"one forward: {(f_end - f_beg) / 1e3:.1f} us")
