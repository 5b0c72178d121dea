"""Dev helper: isolated correlate-stage time (ms) and frac for the current environment: python tools/iso.py [hours]"""
import json
import subprocess
import sys

hours = sys.argv[1] if len(sys.argv) > 1 else "24"
out = subprocess.run(["python", "bench.py", "--hours", hours, "--steps", "2", "--warmup", "2", "--no-cpu-baseline"],
                     capture_output=True, text=True).stdout
d = json.loads([l for l in out.splitlines() if l.startswith("{")][-1])
r = d["roofline"]
print(f"value={d['value']:.2f} ms={d['ms_per_step']:.1f} e2e={d['e2e']['value']:.2f} alone_ms={r['stage_ms_alone']:.1f} "
      f"frac={r['frac']:.3f} in_step_frac={r['in_step']['frac']:.3f} clk={d['clocks']['sm_mhz']}")
