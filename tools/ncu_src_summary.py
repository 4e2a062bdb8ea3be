"""Summarise an `ncu --page source --csv` dump: stall reasons and opcode mix of one kernel."""
import csv, collections, sys
rows=list(csv.reader(open(sys.argv[1])))
h=rows[1]
idx={n:i for i,n in enumerate(h)}
stalls=[n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
tot=collections.Counter(); ops=collections.Counter(); opsamp=collections.Counter(); n=0
for r in rows[2:]:
    if len(r)<len(h) or not r[idx['# Samples']].isdigit(): continue
    n+=1
    for s in stalls:
        tot[s]+=int(r[idx[s]] or 0)
    src=r[idx['Source']].split()
    op=src[1] if src[0].startswith('@') else src[0]
    op=op.split('.')[0]
    ops[op]+=int(r[idx['Instructions Executed']] or 0)
    opsamp[op]+=int(r[idx['# Samples']] or 0)
T=sum(tot.values())
print(rows[0][1][:90])
for s,v in tot.most_common(9): print(f"  {s:26s}{v:8d} {100*v/T:5.1f}%")
I=sum(ops.values())
for o,v in ops.most_common(16): print(f"  {o:12s}{v:10d} {100*v/I:5.1f}%  samples {opsamp[o]}")
print(' ',n,'SASS instrs,',I,'warp-instrs executed')
