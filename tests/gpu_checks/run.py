"""Bring-up runner for the GPU box: runs each check in its own process under a timeout so that a
trapped kernel (sticky CUDA error) or a hang cannot take the other checks down.

    python tests/gpu_checks/run.py [name ...]      # default: all check_*.py in this directory
Writes gpurun_out/checks.log and gpurun_out/checks.json.
"""
import json, os, subprocess, sys, time
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
OUT = ROOT / "gpurun_out"
OUT.mkdir(exist_ok=True)

def main():
    names = sys.argv[1:] or sorted(p.stem for p in HERE.glob("check_*.py"))
    results = {}
    log = open(OUT / "checks.log", "a")
    for name in names:
        t0 = time.time()
        env = dict(os.environ, PYTHONPATH=str(ROOT))
        try:
            r = subprocess.run([sys.executable, str(HERE / f"{name}.py")], capture_output=True, text=True,
                               timeout=int(os.environ.get("F5_CHECK_TIMEOUT", "240")), env=env, cwd=str(ROOT))
            rc, out = r.returncode, r.stdout + r.stderr
        except subprocess.TimeoutExpired as e:
            rc, out = -999, (e.stdout or b"").decode() if isinstance(e.stdout, bytes) else str(e.stdout) + "\nTIMEOUT"
        dt = time.time() - t0
        results[name] = {"rc": rc, "seconds": round(dt, 1)}
        msg = f"===== {name}: rc={rc} ({dt:.1f}s)\n{out}\n"
        log.write(msg); log.flush()
        print(msg[-6000:])
    (OUT / "checks.json").write_text(json.dumps(results, indent=1))
    bad = [k for k, v in results.items() if v["rc"] != 0]
    print("FAILED:" if bad else "ALL OK", bad)
    return 1 if bad else 0

if __name__ == "__main__":
    sys.exit(main())
