"""Debug: in-situ GEMM timelines inside the captured CUDA graph of one base-model B=1 step."""
import ctypes as C, numpy as np, torch
from f5_tts_mlx_b200 import DiT, F5TTS, BASE_CONFIG, _lib
from f5_tts_mlx_b200.weights import random_dit_weights
dev = "cuda"; lib = _lib.load()
cfg = BASE_CONFIG
model = DiT(dim=cfg.dim, depth=cfg.depth, heads=cfg.heads, ff_mult=cfg.ff_mult, text_num_embeds=cfg.text_num_embeds,
            text_dim=cfg.text_dim, conv_layers=cfg.conv_layers, device=dev).load_weights(random_dit_weights(cfg))
g = torch.Generator().manual_seed(3)
N, nref = 937, 328
cond = (torch.randn(1, nref, 100, generator=g) * 2.24 - 1.27).to(dev)
text = torch.randint(0, 2545, (1, 152), generator=g, dtype=torch.int32)
y0 = torch.randn(1, N, 100, generator=g).to(dev)
f5 = F5TTS(model)
NCALL, NCTA = 700, 512
ts = torch.zeros(NCALL, NCTA * 10, dtype=torch.int64, device=dev)
kw = dict(steps=4, method="euler", cfg_strength=2.0, return_trajectory=False, y0=y0)
f5.use_cuda_graph = False
f5.sample(cond, text, N, **kw)           # eager warm-up (sets attributes), no timestamps
f5.use_cuda_graph = True
f5._plans.clear()
# the graph path runs one eager pass, then captures: give only the CAPTURE pass the timestamp slices
orig = f5.last_plan.__class__.run_eager
state = {"n": 0}
def patched(self, m):
    state["n"] += 1
    if state["n"] == 2:
        lib.f5_debug_gemm_ts(ts.data_ptr(), ts.stride(0) * 8, NCALL)
    orig(self, m)
    if state["n"] == 2:
        lib.f5_debug_gemm_ts(None, 0, 0)
f5.last_plan.__class__.run_eager = patched
f5.sample(cond, text, N, **kw)           # eager + capture + replay
for _ in range(3): f5.sample(cond, text, N, **kw)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); f5.last_plan.graph.replay(); e1.record(); torch.cuda.synchronize()
print("graph replay (3 intervals):", e0.elapsed_time(e1), "ms")
t = ts.cpu().numpy().reshape(NCALL, NCTA, 10).astype(np.float64)
used = t[:, :, 0].max(axis=1) > 0
idx = np.nonzero(used)[0]
rows = []
for i in idx:
    m = t[i][:, 0] > 0
    st = t[i][m][:, 0].min(); en = t[i][m][:, 9].max()
    rows.append((i, int(m.sum()), st, en, t[i][m][:, 5].mean() - st, t[i][m][:, 6].mean() - st, t[i][m][:, 7].mean() - st))
base = rows[0][2]
# one DiT forward = 3 (input embed + 2 conv) + 22*4 + 1 GEMMs = 92 ; precompute has 2*4+... GEMMs before
print("calls logged:", len(rows))
prev_end = None
out = []
for (i, n, st, en, mma0, mmal, epi0) in rows:
    gap = (st - prev_end) / 1e3 if prev_end else 0.0
    out.append((i, n, (st - base) / 1e3, (en - st) / 1e3, gap, mma0 / 1e3, mmal / 1e3))
    prev_end = en
# skip precompute GEMMs: find the first call with 240 ctas? print a window of the second forward
k0 = 11 + 92          # precompute (10 GEMMs + mod table = 10) ... approximate: print from there
for o in out[k0:k0 + 12]:
    print(f"call {o[0]:4d} ctas {o[1]:4d} t={o[2]:9.1f}us dur={o[3]:6.2f}us gap_before={o[4]:6.2f}us mma0={o[5]:5.2f} mma_last={o[6]:5.2f}")
NAMES = ["start", "setup", "pdl_wait", "tma0", "tma_last", "mma0", "mma_last", "epi0", "epi_end", "exit"]
print("     " + " ".join(f"{n:>9s}" for n in NAMES))
for (i, n, *_rest) in rows[k0 + 30:k0 + 38]:
    m = t[i][:, 0] > 0
    st0 = t[i][m][:, 0].min()
    rel = (t[i][m] - st0) / 1e3
    rel[t[i][m] == 0] = np.nan
    print(f"call {i} ctas {n}")
    print("mean " + " ".join(f"{np.nanmean(rel[:, k]):9.2f}" for k in range(10)))
    print("max  " + " ".join(f"{np.nanmax(rel[:, k]):9.2f}" for k in range(10)))
durs = np.array([o[3] for o in out[k0:k0 + 92]]); gaps = np.array([o[4] for o in out[k0:k0 + 92]])
print("one forward (92 GEMMs): sum dur", durs.sum(), "us; sum gaps (incl. LN/attn/launch)", gaps.sum(), "us")
