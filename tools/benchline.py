"""Prints a short summary of the JSON line bench.py wrote to stdin (dev helper)."""
import json
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else ""
line = [l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1]
d = json.loads(line)
st = d.get("stage_ms", {})
print(tag, f"value={d['value']:.2f} h/s ms={d['ms_per_step']:.1f} e2e={d['e2e']['value']:.2f}",
      "stages:", {k: round(v, 1) for k, v in st.items()}, f"roofline={d['roofline']['frac']:.3f}",
      f"launches={d.get('gpu_launches')}", f"det={d['config'].get('detections_per_step')}",
      f"clk={d.get('clocks', {}).get('sm_mhz')}", "p2:", d.get("phase2_work_per_step"))
