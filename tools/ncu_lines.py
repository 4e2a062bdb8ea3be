"""Join an `ncu --page source --csv` (SASS view) dump with nvdisasm line info: stall samples per
source line.  usage: ncu_lines.py <src.csv> <lib.so> <kernel-mangled-substring> [top]"""
import csv, collections, os, re, subprocess, sys, tempfile
src_csv, lib, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
lines_of = {}
for f in os.listdir(tmp):
    if not f.endswith(".cubin"): continue
    out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    cur, line, fn = None, None, None
    for l in out.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", l)
        if m: fn = m.group(1); continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m: line = (os.path.basename(m.group(1)), int(m.group(2))); continue
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
        if m and fn and kern in fn:
            lines_of.setdefault(fn, []).append((int(m.group(1), 16), line, m.group(2)))
if not lines_of: sys.exit("kernel not found")
fn = max(lines_of, key=lambda k: len(lines_of[k]))
dis = lines_of[fn]
rows = list(csv.reader(open(src_csv)))
h = rows[1]; idx = {n: i for i, n in enumerate(h)}
recs = [r for r in rows[2:] if len(r) >= len(h) and r[idx['# Samples']].isdigit()]
base = int(recs[0][idx['Address']], 16)
by_line = collections.Counter(); ex_line = collections.Counter(); stall_line = collections.defaultdict(collections.Counter)
stalls = [n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
off2line = {o: ln for o, ln, _ in dis}
for r in recs:
    off = int(r[idx['Address']], 16) - base
    ln = off2line.get(off)
    s = int(r[idx['# Samples']])
    by_line[ln] += s
    ex_line[ln] += int(r[idx['Instructions Executed']] or 0)
    for st in stalls: stall_line[ln][st] += int(r[idx[st]] or 0)
tot = sum(by_line.values())
print(fn, "samples", tot)
src_cache = {}
for ln, s in by_line.most_common(top):
    text = ""
    if ln:
        p = os.path.join(os.path.dirname(os.path.abspath(lib)), "csrc", ln[0])
        if p not in src_cache and os.path.exists(p): src_cache[p] = open(p).read().splitlines()
        if p in src_cache and ln[1] <= len(src_cache[p]): text = src_cache[p][ln[1] - 1].strip()[:80]
    st = ' '.join(f"{k[6:]}:{v}" for k, v in stall_line[ln].most_common(3) if v)
    print(f"{s:5d} {100*s/tot:5.1f}% {str(ln):28s} ex={ex_line[ln]:8d} [{st}] {text}")
