"""Run-to-run and window-to-window spread of bench.py's driver-style headline (20 steps + join): `ms_windows` of three runs."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for i in range(3):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--gpus", "1", "--steps", "20", "--warmup", "5", "--no-cpu-baseline", "--no-extras",
                        "--windows", "9"], capture_output=True, text=True)
    try:
        b = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
        print("reported", round(b["ms_per_step"] * 20, 4), "windows", b.get("ms_windows"), "clocks", b.get("clocks"), flush=True)
    except Exception as ex:
        print("ERR", ex, r.stderr[-500:], flush=True)
