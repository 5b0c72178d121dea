"""Dev helper: static opcode mix of kernels matching a substring in an object/.so (cuobjdump -sass)."""
import collections
import re
import subprocess
import sys

obj, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
name = None
mix = collections.defaultdict(collections.Counter)
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and name and pat in name:
        mix[name][m.group(2).split('.')[0]] += 1
for n, c in mix.items():
    tot = sum(c.values())
    fp = sum(v for k, v in c.items() if k in ("FADD", "FMUL", "FFMA", "FADD2", "FMUL2", "FFMA2"))
    print(n[:110])
    print("  total", tot, "fp", fp, " ".join(f"{k}:{v}" for k, v in c.most_common(18)))
