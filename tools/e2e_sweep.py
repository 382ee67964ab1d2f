"""Sweep the slab size of the streamed e2e path: python tools/e2e_sweep.py 128 256 512 ..."""
import json, subprocess, sys, os
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for sl in sys.argv[1:]:
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--parse-only", "--no-cpu", "--steps", "1", "--e2e-steps", "3",
                          "--slab-bytes", str(int(sl) << 20)], capture_output=True, text=True).stdout
    try:
        d = json.loads(out.strip().splitlines()[-1])
        print(sl, "MiB:", round(d["e2e"]["value"] / 1e9, 2), "G calls/s; step", round(d["ms_per_step"], 2), "ms", flush=True)
    except Exception as e:
        print(sl, "failed", e, out[-300:])
