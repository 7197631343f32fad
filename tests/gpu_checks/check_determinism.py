"""Debug: is f5_dit_precompute deterministic run to run, and does the CFG-batched session agree
with single sessions bit for bit?"""
import torch
from f5_tts_mlx_b200 import DiT, BASE_CONFIG
from f5_tts_mlx_b200.weights import random_dit_weights
dev = "cuda"
cfg = BASE_CONFIG
W = random_dit_weights(cfg, seed=1234)
model = DiT(dim=cfg.dim, depth=cfg.depth, heads=cfg.heads, ff_mult=cfg.ff_mult, text_num_embeds=cfg.text_num_embeds,
            text_dim=cfg.text_dim, conv_layers=cfg.conv_layers, device=dev).load_weights(W)
g = torch.Generator().manual_seed(3)
N, nref = 937, 328
cond = torch.zeros(1, N, 100); cond[:, :nref] = (torch.randn(1, nref, 100, generator=g) * 2.24 - 1.27)
text = torch.randint(0, 2545, (1, 152), generator=g, dtype=torch.int32)
tv = torch.tensor([0.0])
def mx(a, b): return (a.float() - b.float()).abs().max().item()
for use_cfg in (True, False):
    s = model.session(1, N, 1, use_cfg, 152, False)
    s.set_inputs(text, cond.to(dev), tv.to(dev), None)
    runs = []
    for i in range(4):
        model.precompute(s); torch.cuda.synchronize()
        runs.append((s.text_x.clone(), s.ct_bf16.clone(), s.hoist.clone(), s.text_h.clone(), s.text_g.clone(), s.grn_nx.clone()))
    for i in range(1, 4):
        print(f"cfg={use_cfg} run{i} vs run0: text_x {mx(runs[i][0], runs[0][0]):.3e} ct {mx(runs[i][1], runs[0][1]):.3e} hoist {mx(runs[i][2], runs[0][2]):.3e} "
              f"text_h {mx(runs[i][3], runs[0][3]):.3e} text_g {mx(runs[i][4], runs[0][4]):.3e} nx {mx(runs[i][5], runs[0][5]):.3e}")
    if use_cfg:
        bat = runs[0]
    else:
        print("batched cond half vs single: text_x", mx(bat[0][:N], runs[0][0]), "hoist", mx(bat[2][:N], runs[0][2]), "hoist scale", runs[0][2].abs().max().item())
