"""A/B of library environment switches on the bench workload: every argument is one run, "K=V,K2=V2" (or "-" for the defaults);
each run is tools/quad_ab.py's child (steady step, 20-step window, tick / regeneration launch, host-buffer step, 1-env reset).
    python tools/env_ab.py - FFMP_TICK_WARPS=22 FFMP_TICK_WARPS=33"""
import json, os, subprocess, sys
HERE = os.path.dirname(os.path.abspath(__file__))
for arg in sys.argv[1:]:
    e = {} if arg == "-" else dict(kv.split("=", 1) for kv in arg.split(","))
    r = subprocess.run([sys.executable, os.path.join(HERE, "quad_ab.py"), "child", "16", "3"], env=dict(os.environ, **e), capture_output=True, text=True)
    print(json.dumps(e), r.stdout.strip() or r.stderr[-600:], flush=True)
