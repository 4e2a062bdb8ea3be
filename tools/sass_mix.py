"""Static SASS opcode mix per kernel of an object / shared library: python tools/sass_mix.py <file> [regex]"""
import collections, re, subprocess, sys
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
fn, mix = None, collections.defaultdict(collections.Counter)
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(.*?);", line)
    if m and fn:
        toks = m.group(1).split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        mix[fn][op.split(".")[0] + (".SPILL" if "SPILL" in op else "")] += 1
for fn, c in mix.items():
    if pat and not pat.search(fn): continue
    tot = sum(c.values())
    fp = c["DFMA"] + c["DMUL"] + c["DADD"]
    print(f"{fn}: {tot} instrs, FP64 {fp} ({100*fp/tot:.0f}%)")
    print("   " + "  ".join(f"{k}:{v}" for k, v in c.most_common(14)))
