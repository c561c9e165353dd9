"""Reads `ncu --page source --csv` output and prints the instructions with the most stall samples."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
H = rows[hdr]
data = [r for r in rows[hdr + 1:] if len(r) == len(H) and r[0] != 'Address']
si, so, ie = H.index('# Samples'), H.index('Source'), H.index('Instructions Executed')
stalls = [i for i, h in enumerate(H) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[si]) for r in data)
print('total samples', tot, 'instructions', len(data))
agg = {}
for r in data:
    for i in stalls:
        agg[H[i]] = agg.get(H[i], 0) + int(r[i])
print('stall mix:', sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:8])
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for r in sorted(data, key=lambda r: -int(r[si]))[:n]:
    st = sorted([(int(r[i]), H[i]) for i in stalls if int(r[i]) > 0], reverse=True)[:3]
    print('%6d %5.1f%% exec=%8s  %-64s %s' % (int(r[si]), 100 * int(r[si]) / max(tot, 1), r[ie], r[so].strip()[:64], st))
