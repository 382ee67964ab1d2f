"""Attribute ncu per-SASS-instruction counts to CUDA source lines.
usage: ncu_lines.py <report.ncu-rep> <kernel regex> <cubin> <function substring> [launch count divisor]
(the SASS order of `ncu --page source` and of `nvdisasm -g` is the same; line info comes from -lineinfo)"""
import csv, re, subprocess, sys
rep, kre, cubin, fn = sys.argv[1:5]
div = float(sys.argv[5]) if len(sys.argv) > 5 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
h = rows[hi]
ci, si = h.index("Instructions Executed"), h.index("Source")
sti = h.index("# Samples") if "# Samples" in h else None
sass = []
for r in rows[hi + 1:]:
    if len(r) <= ci or r[0].startswith("Kernel Name") :
        if r and r[0].startswith("Kernel Name"): break
        continue
    try: sass.append((r[si].strip(), int(r[ci]), int(r[sti]) if sti is not None else 0))
    except ValueError: pass
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# find function section
lines = []
cur = None
infn = False
for ln in dis:
    if ln.startswith(".text.") or ln.strip().startswith(".section"):
        infn = fn in ln
    if not infn: continue
    m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', ln)
    if m: cur = (m.group(1), int(m.group(2))); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", ln)
    if m: lines.append((cur, m.group(1).strip()))
print("sass rows", len(sass), "disasm instrs", len(lines))
agg = {}
n = min(len(sass), len(lines))
for k in range(n):
    key = lines[k][0]
    a = agg.setdefault(key, [0, 0])
    a[0] += sass[k][1]; a[1] += sass[k][2]
tot = sum(v[0] for v in agg.values())
src = {}
for key in agg:
    if key and key[0] not in src:
        try: src[key[0]] = open([p for p in subprocess.run(["find", "/root/repo", "-name", key[0]], capture_output=True, text=True).stdout.split() if p][0]).read().splitlines()
        except Exception: src[key[0]] = []
print("total instr", tot, "per unit", tot / div)
for key, (v, smp) in sorted(agg.items(), key=lambda kv: (kv[0] or ("", 0))):
    if v < tot * float(__import__("os").environ.get("NCU_LINES_MIN", "0.003")): continue
    text = ""
    if key and src.get(key[0]) and key[1] - 1 < len(src[key[0]]): text = src[key[0]][key[1] - 1].strip()[:110]
    print(f"{v / div:9.1f} {100.0 * v / tot:5.1f}% smp {smp:6d}  {key[0] if key else '?'}:{key[1] if key else 0:<5d} {text}")
