"""Dev helper: opcode mix and top stalled instructions of one captured launch.
usage: python tools/ncu_source.py report.ncu-rep <launch-skip>"""
import collections
import csv
import subprocess
import sys

rep, skip = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
print(rows[0][1][:100])
hdr = rows[1]
data = [r for r in rows[2:] if r and r[0].startswith("0x") and len(r) > 40]
i_src, i_s, i_ex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
tot = sum(int(r[i_s]) for r in data)
ops, samp = collections.Counter(), collections.Counter()
for r in data:
    toks = r[i_src].split()
    op = toks[1] if toks[0].startswith('@') else toks[0]
    ops[op] += int(r[i_ex])
    samp[op] += int(r[i_s])
te = sum(ops.values())
print('samples', tot, 'static instr', len(data), 'warp-instr executed', te)
for op, c in ops.most_common(24):
    print(f"{op:24s} exec {100*c/te:5.1f}%  samples {100*samp[op]/tot:5.1f}%")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
for r in sorted(data, key=lambda r: -int(r[i_s]))[:22]:
    st = sorted(((int(r[i]), hdr[i]) for i in stall_cols), reverse=True)[:2]
    print(f"{int(r[i_s]):6d} {100*int(r[i_s])/tot:4.1f}% {r[i_src].strip()[:64]:64s} {st}")
