"""Debug: batched-CFG sample step vs two unbatched DiT passes (base model), per GEMM variant."""
import os, sys, torch
from f5_tts_mlx_b200 import DiT, F5TTS, BASE_CONFIG
from f5_tts_mlx_b200.weights import random_dit_weights
dev = "cuda"
def rel(a, b): return ((a - b).norm() / (b.norm() + 1e-30)).item()
cfg = BASE_CONFIG
W = random_dit_weights(cfg, seed=1234)
model = DiT(dim=cfg.dim, depth=cfg.depth, heads=cfg.heads, ff_mult=cfg.ff_mult, text_num_embeds=cfg.text_num_embeds,
            text_dim=cfg.text_dim, conv_layers=cfg.conv_layers, device=dev).load_weights(W)
g = torch.Generator().manual_seed(3)
N, nref = 937, 328
cond = (torch.randn(1, nref, 100, generator=g) * 2.24 - 1.27).clamp(-11.51, 5).to(dev)
text = torch.randint(0, 2545, (1, 152), generator=g, dtype=torch.int32)
y0 = torch.randn(1, N, 100, generator=g).to(dev)
f5 = F5TTS(model); f5.use_cuda_graph = False
out, traj = f5.sample(cond, text, N, steps=2, method="euler", cfg_strength=2.0, sway_sampling_coef=None, y0=y0)
s = f5.last_plan.session
v = s.v.clone(); x_b = s.x.clone(); hoist_b = s.hoist.clone(); qkv_b = s.qkv_bf16.float().clone(); mod_b = s.mod_table.clone(); h_b = s.h.clone()
step_cond = torch.zeros(1, N, 100, device=dev); step_cond[:, :nref] = cond
t0 = torch.tensor(0.0)
pred = model(y0, step_cond, text.to(dev), t0, False, False)
s1 = model._sessions[(1, N, 1, False, 152, False)]
x_p = s1.x.clone(); hoist_p = s1.hoist.clone(); qkv_p = s1.qkv_bf16.float().clone(); mod_p = s1.mod_table.clone(); h_p = s1.h.clone()
null = model(y0, step_cond, text.to(dev), t0, True, True)
x_n = s1.x.clone(); hoist_n = s1.hoist.clone()
print("variant env", os.environ.get("F5_GEMM_VARIANT"))
print("mod table", rel(mod_b, mod_p))
print("hoist cond rows", rel(hoist_b[:N], hoist_p), "uncond rows", rel(hoist_b[N:], hoist_n))
print("h (after in-proj) cond", rel(h_b[:N], h_p))
print("x final cond", rel(x_b[:N], x_p), "uncond", rel(x_b[N:], x_n))
print("qkv last block cond", rel(qkv_b[:N], qkv_p))
print("v cond", rel(v[:N], pred.view(N, 100)), "v uncond", rel(v[N:], null.view(N, 100)))
y1 = y0 + 1.0 * (pred + (pred - null) * 2.0)
print("y1", rel(traj[-1], y1))
