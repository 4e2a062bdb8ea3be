"""Static SASS instruction count per source line of one kernel (code-size budget: at ~10^4 targets
the per-target kernels are instruction-fetch bound after an L2 flush).
usage: sass_lines.py <lib.so> <kernel-mangled-substring> [top]"""
import collections, os, re, subprocess, sys, tempfile
lib, kern = sys.argv[1:3]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cnt = collections.defaultdict(collections.Counter)
for f in os.listdir(tmp):
    if not f.endswith(".cubin"): continue
    out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    line, fn = None, None
    for l in out.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", l) or re.match(r"(\$[_A-Za-z]\S*):\s*$", l)   # kernels and their out-of-line callees
        if m: fn = m.group(1); continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m: line = (os.path.basename(m.group(1)), int(m.group(2))); continue
        if fn and kern in fn and re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(.*?);", l):
            cnt[fn][line] += 1
for fn, c in cnt.items():
    tot = sum(c.values())
    print(fn, tot, "instructions,", tot * 16 // 1024, "KB")
    src = {}
    for (f, ln), n in c.most_common(top):
        p = os.path.join(os.path.dirname(os.path.abspath(lib)), "csrc", f)
        if p not in src and os.path.exists(p): src[p] = open(p).read().splitlines()
        text = src[p][ln - 1].strip()[:90] if p in src and ln <= len(src[p]) else ""
        print(f"{n:6d} {f}:{ln:<5d} {text}")
