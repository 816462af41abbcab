#!/usr/bin/env python
"""Top stalled SASS instructions of one captured launch:  tools/ncu_hot.py <report.ncu-rep> <section-index> [N]"""
import csv, io, subprocess, sys
rep, idx = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
allrows = list(csv.reader(io.StringIO(out)))
starts = [i for i, r in enumerate(allrows) if r and r[0] == 'Kernel Name']
sec = int(idx)            # section index (ncu prints two sections per captured launch)
rows = allrows[starts[sec]:starts[sec + 1] if sec + 1 < len(starts) else None]
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
body = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(int(r[col['# Samples']] or 0) for r in body)
print(rows[0][1][:150])
print('total samples', tot, ' instructions', len(body))
for i, r in enumerate(body):
    r.append(i)
top = sorted(body, key=lambda r: -int(r[col['# Samples']] or 0))[:N]
for r in top:
    s = int(r[col['# Samples']] or 0)
    why = sorted(((int(r[col[k]] or 0), k[6:]) for k in stalls), reverse=True)[:3]
    print('%5.1f%% #%-5d %-78s x%-8s %s' % (100.0 * s / max(tot, 1), r[-1], r[col['Source']][:78], r[col['Instructions Executed']],
                                    ' '.join('%s=%d' % (k, v) for v, k in why if v)))
