"""Fixtures produced by the REFERENCE'S OWN CODE (build container only: needs /root/reference).

MLX cannot be installed in this image, so the unmodified reference sources
(/root/reference/f5_tts_mlx/{utils,rope,convnext_v2,audio,dit,duration,cfm,generate}.py) are imported
on top of tests/mlx_shim (a torch-CPU implementation of the ~60 MLX / einx calls they make) and
executed with the repo's seeded synthetic weights.  The outputs written to ref_*.npz are therefore
what the reference's Python computes — module wiring, argument order, masks, quirks and all — with
only the leaf array primitives supplied by the shim.  tests/test_ref_pins.py requires the CPU oracle
to match them to 1e-5 relative (fp32 both sides) and the CUDA path to match them within its derived
bf16 tolerance.  Nothing here is copied from the reference; the fixtures hold numbers only.

    python tests/golden/make_ref_golden.py          # rewrites tests/golden/ref_*.npz, prints oracle deviations
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import mlx_shim as shim                                                  # noqa: E402
from oracle import f5_oracle as O                                        # noqa: E402
from f5_tts_mlx_b200.weights import (GATE_CONFIG, DiTConfig as PkgDiTConfig, random_dit_weights,   # noqa: E402
                                     random_duration_weights)

torch.set_num_threads(8)
ref = shim.import_reference()
mx = ref.mx
A = mx.array


def t2a(t):
    return A(t)


def a2n(a):
    return np.asarray(a)


def rel(a, b):
    a, b = torch.as_tensor(np.asarray(a)).double(), torch.as_tensor(np.asarray(b)).double()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def build_ref_dit(cfg, W):
    m = ref.dit.DiT(dim=cfg.dim, depth=cfg.depth, heads=cfg.heads, ff_mult=cfg.ff_mult, mel_dim=cfg.mel_dim,
                    text_num_embeds=cfg.text_num_embeds, text_dim=cfg.text_dim, conv_layers=cfg.conv_layers)
    m.load_weights([(k[len("transformer."):], A(v)) for k, v in W.items() if k.startswith("transformer.")])
    return m


def ocfg_of(cfg):
    return O.DiTConfig(dim=cfg.dim, depth=cfg.depth, heads=cfg.heads, ff_mult=cfg.ff_mult, text_num_embeds=cfg.text_num_embeds,
                       text_dim=cfg.text_dim, conv_layers=cfg.conv_layers)


report = {}

# ------------------------------------------------------------------------------------------------
# (1) DiT.__call__ (dit.py:374-401), gate config (4 layers / 512), batch 1 (mask=None, as sample() calls it)
# ------------------------------------------------------------------------------------------------
cfg = GATE_CONFIG
W = random_dit_weights(cfg, seed=1234)
dit = build_ref_dit(cfg, W)
ocfg = ocfg_of(cfg)
g = torch.Generator().manual_seed(101)
N, nt = 80, 24
x = torch.randn(1, N, 100, generator=g)
cond = torch.randn(1, N, 100, generator=g)
text = torch.randint(0, 2545, (1, nt), generator=g, dtype=torch.int32)
text[0, 19:] = -1                                     # trailing pad inside the row
tval = torch.tensor(0.4321)
out = dit(x=A(x), cond=A(cond), text=A(text), time=A(tval), drop_audio_cond=False, drop_text=False, mask=None)
out_drop = dit(x=A(x), cond=A(cond), text=A(text), time=A(tval), drop_audio_cond=True, drop_text=True, mask=None)
# (a text longer than the mel cannot reach DiT.__call__ through sample() — duration >= text_len + 1, cfm.py:301-317 —
# and fails upstream in mx.pad with a negative width at dit.py:204; the shim raises there too)
try:
    dit(x=A(x), cond=A(cond), text=A(torch.zeros(1, N + 9, dtype=torch.int32)), time=A(tval), drop_audio_cond=False,
        drop_text=False, mask=None)
    raise SystemExit("expected the reference to reject text longer than the mel")
except ValueError:
    pass
report["dit_forward"] = rel(O.dit_forward(x, cond, text, tval, False, False, None, W, ocfg), a2n(out))
report["dit_forward_drop"] = rel(O.dit_forward(x, cond, text, tval, True, True, None, W, ocfg), a2n(out_drop))

# batch 2 with a key-padding mask: the reference's branch calls `.expand` (dit.py:162), which mlx.core.array does
# not have, so it cannot run upstream; with the shim's opt-in `expand` the INTENDED semantics execute.
shim.core.ALLOW_EXPAND = True
B = 2
x2 = torch.randn(B, N, 100, generator=g); cond2 = torch.randn(B, N, 100, generator=g)
text2 = torch.randint(0, 2545, (B, nt), generator=g, dtype=torch.int32); text2[1, 13:] = -1
lens2 = torch.tensor([N, 57]); mask2 = torch.arange(N)[None, :] < lens2[:, None]
out_b2 = dit(x=A(x2), cond=A(cond2), text=A(text2), time=A(tval), drop_audio_cond=False, drop_text=False, mask=A(mask2))
shim.core.ALLOW_EXPAND = False
report["dit_forward_b2_masked(intended .expand)"] = rel(O.dit_forward(x2, cond2, text2, tval, False, False, mask2, W, ocfg),
                                                        a2n(out_b2))
np.savez_compressed(os.path.join(HERE, "ref_dit_gate.npz"), x=x.numpy(), cond=cond.numpy(), text=text.numpy(),
                    t=tval.numpy(), out=a2n(out), out_drop=a2n(out_drop),
                    x2=x2.numpy(), cond2=cond2.numpy(), text2=text2.numpy(), lens2=lens2.numpy(), out_b2=a2n(out_b2),
                    weight_seed=1234)

# ------------------------------------------------------------------------------------------------
# (2) F5TTS.sample (cfm.py:264-402), batch 1, mel conditioning and raw-wave conditioning
# ------------------------------------------------------------------------------------------------
f5 = ref.cfm.F5TTS(transformer=dit)
nref, N = 40, 96
cond = (torch.randn(1, nref, 100, generator=g) * 2.24 - 1.27).clamp(-11.51, 5)
text = torch.randint(0, 2545, (1, 24), generator=g, dtype=torch.int32)
samples = {}
for name, kw in {
    "euler_cfg": dict(steps=4, method="euler", cfg_strength=2.0, sway_sampling_coef=-1.0, seed=7),
    "midpoint_nocfg": dict(steps=3, method="midpoint", cfg_strength=0.0, sway_sampling_coef=None, seed=7),
    "rk4_cfg": dict(steps=3, method="rk4", cfg_strength=2.0, sway_sampling_coef=-1.0, seed=11),
}.items():
    o, tr = f5.sample(A(cond), A(text), N, **kw)
    oo, otr = O.sample(cond, text, N, W, ocfg, **kw)
    report[f"sample_{name}"] = rel(oo, a2n(o))
    report[f"sample_{name}_traj"] = rel(otr, a2n(tr))
    samples[name + "_out"] = a2n(o)
    samples[name + "_traj"] = a2n(tr)
# duration shorter than the text/cond (clamped to lens+1, cfm.py:317) and a text longer than the cond
text_l = torch.randint(0, 2545, (1, 50), generator=g, dtype=torch.int32)
o, tr = f5.sample(A(cond), A(text_l), 10, steps=3, method="euler", cfg_strength=2.0, seed=3)
oo, otr = O.sample(cond, text_l, 10, W, ocfg, steps=3, method="euler", cfg_strength=2.0, seed=3)
report["sample_short_duration"] = rel(oo, a2n(o))
samples["short_out"], samples["short_traj"], samples["text_l"] = a2n(o), a2n(tr), text_l.numpy()
# raw-wave conditioning (cfm.py:283-286): 1 s of the reference's fixture clip
pcm = np.load(os.path.join(HERE, "mel_fixture.npz"))["pcm"]
wavef = torch.from_numpy(pcm.astype(np.float32) / 32768.0)[None]
o, tr = f5.sample(A(wavef), A(text), 150, steps=3, method="euler", cfg_strength=2.0, seed=5)
oo, otr = O.sample(wavef, text, 150, W, ocfg, steps=3, method="euler", cfg_strength=2.0, seed=5)
report["sample_raw_wave"] = rel(oo, a2n(o))
samples["wave_out"], samples["wave_traj_last"] = a2n(o), a2n(tr)[-1]
np.savez_compressed(os.path.join(HERE, "ref_sample_gate.npz"), cond=cond.numpy(), text=text.numpy(), duration=N,
                    weight_seed=1234, **samples)

# ------------------------------------------------------------------------------------------------
# (3) log_mel_spectrogram / MelSpec (audio.py:12-230) on the fixture clip, batch of two lengths
# ------------------------------------------------------------------------------------------------
mel = ref.audio.log_mel_spectrogram(A(wavef[0]))
report["mel_fixture"] = float(np.abs(a2n(mel) - O.log_mel_spectrogram(wavef[0]).numpy()).max())
filt = ref.audio.mel_filters(sample_rate=24000, n_fft=1024, n_mels=100, norm=None, mel_scale="htk")
report["mel_filters"] = float(np.abs(a2n(filt) - O.mel_filters(24000, 1024, 100).numpy()).max())
report["hanning"] = float(np.abs(a2n(ref.audio.hanning(1024)) - O.hanning(1024).numpy()).max())
odd = wavef[0, :5000 + 131]
mel_odd = ref.audio.MelSpec()(A(odd))
report["mel_odd_length"] = float(np.abs(a2n(mel_odd) - O.log_mel_spectrogram(odd).numpy()).max())
np.savez_compressed(os.path.join(HERE, "ref_mel.npz"), mel=a2n(mel), mel_odd=a2n(mel_odd), odd_len=odd.numel(),
                    filters=a2n(filt).astype(np.float32))

# ------------------------------------------------------------------------------------------------
# (4) DurationPredictor (duration.py:198-253)
# ------------------------------------------------------------------------------------------------
dW = random_duration_weights(seed=777)
dcfg = O.DurationConfig()
dp = ref.duration.DurationPredictor(
    transformer=ref.duration.DurationTransformer(dim=dcfg.dim, depth=dcfg.depth, heads=dcfg.heads, text_dim=dcfg.text_dim,
                                                 ff_mult=dcfg.ff_mult, conv_layers=dcfg.conv_layers,
                                                 text_num_embeds=dcfg.text_num_embeds))
dp.load_weights([(k, A(v)) for k, v in dW.items()], strict=False)   # (rotary inv_freq: the module's own constant)
gd = torch.Generator().manual_seed(21)
dmel = (torch.randn(2, 90, 100, generator=gd) * 2.24 - 1.27)
dtext = torch.randint(0, 2545, (2, 30), generator=gd, dtype=torch.int32); dtext[1, 18:] = -1
dlens = torch.tensor([90, 61])
sec = dp(A(dmel), A(dtext), lens=A(dlens))
sec_nolens = dp(A(dmel[:1]), A(dtext[:1]))
dWo = {"duration." + k: v for k, v in dW.items()}
report["duration"] = rel(O.duration_predictor(dmel, dtext, dWo, dcfg, lens=dlens), a2n(sec))
report["duration_nolens"] = rel(O.duration_predictor(dmel[:1], dtext[:1], dWo, dcfg), a2n(sec_nolens))
np.savez_compressed(os.path.join(HERE, "ref_duration.npz"), mel=dmel.numpy(), text=dtext.numpy(), lens=dlens.numpy(),
                    seconds=a2n(sec), seconds_nolens=a2n(sec_nolens), weight_seed=777)

# ------------------------------------------------------------------------------------------------
# (5) operators and host utilities: rope.py, convnext_v2.py, dit.py pieces, utils.py, solvers
# ------------------------------------------------------------------------------------------------
tiny = PkgDiTConfig(dim=128, depth=1, heads=2, ff_mult=2, text_dim=64, conv_layers=1, text_num_embeds=50)
Wt = random_dit_weights(tiny, seed=77)
tdit = build_ref_dit(tiny, Wt)
tcfg = ocfg_of(tiny)
g = torch.Generator().manual_seed(5)
B, N = 2, 40
xx = torch.randn(B, N, 128, generator=g)
xt = torch.randn(B, N, 64, generator=g)
tt = torch.tensor([0.25, 0.9])
ttext = torch.randint(0, 50, (B, 12), generator=g, dtype=torch.int32); ttext[1, 7:] = -1
ops = {}
ops["time_embed"] = a2n(tdit.time_embed(A(tt)))
report["op_time_embed"] = rel(O.timestep_embedding(tt, Wt), ops["time_embed"])
blk = tdit.text_embed.text_blocks.layers[0]
ops["grn"] = a2n(blk.grn(A(xx)))
report["op_grn"] = rel(O.grn(xx, Wt["transformer.text_embed.text_blocks.layers.0.grn.gamma"],
                             Wt["transformer.text_embed.text_blocks.layers.0.grn.beta"]), ops["grn"])
ops["convnext"] = a2n(blk(A(xt)))
report["op_convnext"] = rel(O.convnext_v2_block(xt, Wt, "transformer.text_embed.text_blocks.layers.0."), ops["convnext"])
ops["text_embed"] = a2n(tdit.text_embed(A(ttext), N, drop_text=False))
ops["text_embed_drop"] = a2n(tdit.text_embed(A(ttext), N, drop_text=True))
report["op_text_embed"] = rel(O.text_embedding(ttext, N, False, Wt, tcfg), ops["text_embed"])
report["op_text_embed_drop"] = rel(O.text_embedding(ttext, N, True, Wt, tcfg), ops["text_embed_drop"])
ops["conv_pos"] = a2n(tdit.input_embed.conv_pos_embed(A(xx)))
report["op_conv_pos"] = rel(O.conv_position_embedding(xx, Wt), ops["conv_pos"])
rope_ref = tdit.rotary_embed.forward_from_seq_len(N)
ops["rope_freqs"] = a2n(rope_ref[0])
report["op_rope_freqs"] = rel(O.rotary_freqs(N), ops["rope_freqs"])
ops["rope_q"] = a2n(ref.rope.apply_rotary_pos_emb(A(xx[:, None, :, :64]), rope_ref[0]))
report["op_rope_apply"] = rel(O.apply_rotary_pos_emb(xx[:, None, :, :64], O.rotary_freqs(N)), ops["rope_q"])
ops["attention_nomask"] = a2n(tdit.transformer_blocks[0].attn(A(xx), mask=None, rope=rope_ref))
report["op_attention"] = rel(O.attention(xx, None, O.rotary_freqs(N), Wt, "transformer.transformer_blocks.0.attn.", 2),
                             ops["attention_nomask"])
temb = tdit.time_embed(A(tt))
ops["dit_block"] = a2n(tdit.transformer_blocks[0](A(xx), temb, mask=None, rope=rope_ref))
report["op_dit_block"] = rel(O.dit_block(xx, O.timestep_embedding(tt, Wt), None, O.rotary_freqs(N), Wt, 0, tcfg), ops["dit_block"])
ops["freqs_cis"] = a2n(ref.rope.precompute_freqs_cis(64, 96))
report["op_freqs_cis"] = rel(O.precompute_freqs_cis(64, 96), ops["freqs_cis"])
ops["pos_idx"] = a2n(ref.rope.get_pos_embed_indices(A(torch.zeros(2, dtype=torch.int32)), 20, max_pos=16))
report["op_pos_idx_equal"] = float((O.get_pos_embed_indices(torch.zeros(2, dtype=torch.int32), 20, 16).numpy() != ops["pos_idx"]).sum())
# utils.py
lens = torch.tensor([3.0, 7.0, 5.0])
ops["lens_to_mask"] = a2n(ref.utils.lens_to_mask(A(lens)))
report["util_lens_to_mask_equal"] = float((O.lens_to_mask(lens).numpy() != ops["lens_to_mask"]).sum())
vocab = {c: i for i, c in enumerate(" abcdefghijklmnopqrstuvwxyz")}
ops["list_str_to_idx"] = a2n(ref.utils.list_str_to_idx([list("hello w?rld"), list("abc")], vocab))
report["util_list_str_to_idx_equal"] = float((O.list_str_to_idx([list("hello w?rld"), list("abc")], vocab).numpy()
                                              != ops["list_str_to_idx"]).sum())
ops["list_str_to_tensor"] = a2n(ref.utils.list_str_to_tensor(["héllo", "ab"]))
report["util_list_str_to_tensor_equal"] = float((O.list_str_to_tensor(["héllo", "ab"]).numpy() != ops["list_str_to_tensor"]).sum())
# solvers on dy/dt = -y + sin(3t)
tg = O.time_grid(9, -1.0)
f_t = lambda t, y: -y + torch.sin(3 * t)                                   # noqa: E731
f_m = lambda t, y: -y + mx.sin(3 * t)                                      # noqa: E731
y0 = torch.linspace(-1, 1, 7)
for nm in ("euler", "midpoint", "rk4"):
    r = a2n(getattr(ref.cfm, f"odeint_{nm}")(f_m, A(y0), A(tg)))
    ops[f"odeint_{nm}"] = r
    report[f"odeint_{nm}"] = rel(getattr(O, f"odeint_{nm}")(f_t, y0, tg), r)
# the sway grid the way sample() builds it (cfm.py:377-381)
for steps in (2, 8, 32):
    t_ref = mx.linspace(0, 1, steps)
    t_ref = t_ref + (-1.0) * (mx.cos(mx.pi / 2 * t_ref) - 1 + t_ref)
    ops[f"tgrid_{steps}"] = a2n(t_ref)
    report[f"tgrid_{steps}_maxabs"] = float(np.abs(O.time_grid(steps, -1.0).numpy() - a2n(t_ref)).max())
np.savez_compressed(os.path.join(HERE, "ref_ops.npz"), x=xx.numpy(), xt=xt.numpy(), t=tt.numpy(), text=ttext.numpy(),
                    weight_seed=77, **ops)

# ------------------------------------------------------------------------------------------------
# (6) generate.py host logic (generate.py:30-36, 104-111, 158-232): what reaches F5TTS.sample
# ------------------------------------------------------------------------------------------------
calls = []


class _Rec:
    def sample(self, audio, text, duration, **kw):
        calls.append(dict(text="".join(text[0]), duration=-1 if duration is None else int(duration), audio_len=audio.shape[1],
                          audio_rms=float(np.sqrt(np.mean(np.asarray(audio) ** 2))), **{k: v for k, v in kw.items()
                                                                                        if k in ("steps", "method", "speed", "seed")}))
        return A(torch.zeros(audio.shape[1] + 256 * 10)), None


orig = ref.cfm.F5TTS.from_pretrained
ref.generate.F5TTS.from_pretrained = classmethod(lambda cls, *a, **k: _Rec())
import tempfile, wave                                                      # noqa: E402
tmp = tempfile.mkdtemp()
refwav = os.path.join(tmp, "ref.wav")
with wave.open(refwav, "wb") as f:
    f.setnchannels(1); f.setsampwidth(2); f.setframerate(24000); f.writeframes((pcm // 8).astype(np.int16).tobytes())
gen_text = "Hello there. This is a longer sentence; it has parts: two of them! Is it done?"
ref.generate.generate(gen_text, ref_audio_path=refwav, ref_audio_text="A reference.", steps=4, method="euler",
                      estimate_duration=True, speed=1.25, seed=3, output_path=os.path.join(tmp, "o.wav"))
ref.generate.generate("Only one sentence here", ref_audio_path=refwav, ref_audio_text="A reference.", duration=2.5,
                      output_path=os.path.join(tmp, "o2.wav"))
ref.generate.F5TTS.from_pretrained = orig
import json                                                                # noqa: E402
json.dump({"gen_text": gen_text, "calls": calls, "split": ref.generate.split_sentences(gen_text),
           "pcm_divisor": 8}, open(os.path.join(HERE, "ref_generate_calls.json"), "w"), indent=1)

print(json.dumps(report, indent=1))
bad = {k: v for k, v in report.items() if v > 1e-5}
print("oracle deviations > 1e-5:", bad)
for f in sorted(os.listdir(HERE)):
    if f.startswith("ref_"):
        print(f, os.path.getsize(os.path.join(HERE, f)))
