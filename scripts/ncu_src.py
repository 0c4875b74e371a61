"""Developer tool: summarise `ncu --page source --csv --print-source sass` output: stall totals and top instructions per stall."""
import csv, sys
path = sys.argv[1]; which = int(sys.argv[2]) if len(sys.argv) > 2 else 0; topn = int(sys.argv[3]) if len(sys.argv) > 3 else 6
rows = list(csv.reader(open(path)))
kern = []; cur = None
for r in rows:
    if r and r[0] == 'Kernel Name': cur = {'name': r[1], 'rows': []}; kern.append(cur); continue
    if r and r[0] == 'Address': cur['hdr'] = r; continue
    if cur is not None and len(r) > 10: cur['rows'].append(r)
k = kern[which]; h = k['hdr']; idx = {n: i for i, n in enumerate(h)}
cols = [n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
tot = sum(int(r[idx['# Samples']] or 0) for r in k['rows'])
print(k['name'], 'n_instr', len(k['rows']), 'samples', tot)
for c in cols:
    s = sum(int(r[idx[c]] or 0) for r in k['rows'])
    if s < 0.01 * tot: continue
    print(f'== {c}: {s} ({100*s/tot:.1f}%)')
    top = sorted(range(len(k['rows'])), key=lambda i: -int(k['rows'][i][idx[c]] or 0))[:topn]
    for i in top:
        r = k['rows'][i]
        print(f"   #{i:5d} {int(r[idx[c]] or 0):6d}  x{r[idx['Instructions Executed']]:>8s}  {r[idx['Source']].strip()[:80]}")
