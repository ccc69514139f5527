"""Developer tool: executed-instruction mix by opcode of one kernel in an .ncu-rep."""
import csv, subprocess, sys, collections
rep, kernel = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name-base", "demangled", "-k", f"regex:{kernel}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
body = [r for r in rows[2:] if len(r) == len(hdr) and r[0] != 'Address']
si, ie = hdr.index('Source'), hdr.index('Instructions Executed')
mix = collections.Counter()
seen_first = {}
tot = 0
for r in body:
    try: n = int(r[ie])
    except: n = 0
    op = r[si].split()[0] if r[si].split() else '?'
    if op.startswith('@'): op = r[si].split()[1]
    op = op.split('.')[0]
    mix[op] += n; tot += n
print(kernel, "warp instructions executed (first launch in file):", tot)
for op, n in mix.most_common(25):
    print(f"  {op:12s} {n:12d} {100*n/tot:5.1f}%")
