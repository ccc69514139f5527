"""Developer tool: opcode mix / top stall instructions of one kernel from an .ncu-rep source page.
usage: python tools/ncu_opmix.py <rep> <kernel regex> [topN]"""
import collections, csv, re, subprocess, sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name-base", "demangled", "-k", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# several launches may follow each other: take the first block
hdr = rows[1]
body = []
for r in rows[2:]:
    if len(r) != len(hdr) or r[0] == 'Address':
        if body:
            break
        continue
    body.append(r)
si, ns, ie = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
tot = sum(int(r[ns]) for r in body); toti = sum(int(r[ie]) for r in body)
print("kernel", rows[0][1][:100]); print("instructions", len(body), "samples", tot, "warp-instr executed", toti)
ops, samp = collections.Counter(), collections.Counter()
for r in body:
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[si]); op = m.group(2).split('.')[0] if m else '?'
    ops[op] += int(r[ie]); samp[op] += int(r[ns])
for op, c in ops.most_common(22):
    print(f"{op:10s} exec {c:9d} {100*c/toti:5.1f}%   samples {samp[op]:7d} {100*samp[op]/tot:5.1f}%")
print("--- top sampled instructions")
for r in sorted(body, key=lambda r: -int(r[ns]))[:top]:
    print(f"{int(r[ns]):6d} {100*int(r[ns])/tot:4.1f}% exec {int(r[ie]):7d}  {r[si].strip()[:100]}")
