"""Long-text RTF through the public generate() (generate.py:113-244): 12 sentences of different length with the same
reference clip, the reference CLI's defaults (rk4, steps=8, CFG 2, sway -1), estimated durations.  Compares the serial
loop with one CUDA graph per length bucket (frame_bucket=128, default), with exact shapes (frame_bucket=0: a new plan —
eager pass + capture — for every new length) and the one-ragged-batch extension.  Prints one JSON line."""
import json, time, wave, os, tempfile
import numpy as np, torch
from f5_tts_mlx_b200 import F5TTS
from f5_tts_mlx_b200 import generate as G

torch.manual_seed(0)
sr = 24000
t = np.arange(5 * sr) / sr
ref = (0.1 * np.sin(2 * np.pi * 180 * t) * (0.6 + 0.4 * np.sin(2 * np.pi * 1.3 * t))).astype(np.float32)
tmp = tempfile.mkdtemp()
G.write_wav(os.path.join(tmp, "ref.wav"), torch.from_numpy(ref))
ref_text = "this is the reference transcript of about five seconds."
sents = ["Short one.", "A somewhat longer sentence follows here, with a clause.", "Tiny.", "The quick brown fox jumps over the lazy dog again and again.",
         "Numbers like forty two appear.", "Another medium length sentence for the test!", "Is this a question?", "Yes; it is: really.",
         "A long sentence that keeps going for quite a while so that its estimated duration is clearly larger than the others in the list.",
         "Back to short.", "Penultimate sentence of the long text.", "The end."]
text = " ".join(sents)
f5 = F5TTS.from_pretrained("random")
res = {}
for name, kw in (("bucket128", dict(frame_bucket=128)), ("exact_shapes", dict(frame_bucket=0)), ("one_ragged_batch", dict(batch_sentences=True))):
    f5._plans.clear(); f5.transformer._sessions.clear(); torch.cuda.empty_cache()
    walls = []
    for rep in range(2):           # rep 0 pays the captures, rep 1 is the steady state of a serving process
        torch.cuda.synchronize(); t0 = time.perf_counter()
        w = G.generate(text, estimate_duration=True, ref_audio_path=os.path.join(tmp, "ref.wav"), ref_audio_text=ref_text,
                       seed=1, f5tts=f5, **kw)
        torch.cuda.synchronize(); walls.append(time.perf_counter() - t0)
    secs = w.shape[0] / sr
    res[name] = {"audio_s": round(secs, 2), "first_call_s": round(walls[0], 3), "steady_s": round(walls[1], 3),
                 "rtf_first": round(walls[0] / secs, 4), "rtf_steady": round(walls[1] / secs, 4), "plans": len(f5._plans)}
print("GENERATE_RTF " + json.dumps(res))
