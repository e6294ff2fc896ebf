"""Summarise an `ncu --page source --csv` dump (SASS view): stall-reason totals, instruction
mix by opcode and the hottest instructions.  usage: ncu_src_summary.py file.csv [top]"""
import csv, sys, collections
allrows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0  # index of the kernel block in the dump
starts = [i for i, r in enumerate(allrows) if r and r[0] == 'Kernel Name'] + [len(allrows)]
rows = allrows[starts[which]:starts[which + 1]]
print('kernel:', rows[0][1][:100])
h = rows[1]
ix = {k: i for i, k in enumerate(h)}
stalls = [k for k in h if k.startswith('stall_') and 'Not Issued' not in k]
tot = collections.Counter(); ops = collections.Counter(); opsamp = collections.Counter()
data = []
for r in rows[2:]:
    if len(r) < len(h): continue
    try:
        samp = int(r[ix['# Samples']] or 0); ex = int(r[ix['Instructions Executed']] or 0)
    except ValueError:
        continue
    op = r[ix['Source']].split()[0] if r[ix['Source']].strip() else '?'
    if op.startswith('@'): op = r[ix['Source']].split()[1]
    op = op.split('.')[0]
    ops[op] += ex; opsamp[op] += samp
    for s in stalls:
        try: tot[s] += int(r[ix[s]] or 0)
        except ValueError: pass
    data.append((samp, ex, r[ix['Source']][:90]))
S = sum(tot.values()) or 1
print('stall totals:', ', '.join(f'{k[6:]} {100*v/S:.1f}%' for k, v in tot.most_common(10)))
E = sum(ops.values()) or 1
print('inst mix (executed %, sample %):', ', '.join(f'{k} {100*v/E:.1f}/{100*opsamp[k]/max(1,sum(opsamp.values())):.1f}' for k, v in ops.most_common(18)))
print('total warp instructions', E)
for samp, ex, src in sorted(set(data), reverse=True)[:top]:
    print(f'{samp:7d} {ex:10d}  {src}')
