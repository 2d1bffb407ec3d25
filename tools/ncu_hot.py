"""Print the hottest SASS lines (by stall samples) per kernel from `ncu --page source --csv --print-source sass`."""
import csv, sys
path, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25
only = sys.argv[3] if len(sys.argv) > 3 else None
rows = list(csv.reader(open(path)))
i = 0
seen = set()
while i < len(rows):
    if rows[i] and rows[i][0] == 'Kernel Name':
        name = rows[i][1]; hdr = rows[i + 1]; j = i + 2
        body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == 'Kernel Name'):
            body.append(rows[j]); j += 1
        i = j
        if name in seen or (only and only not in name): continue
        seen.add(name)
        ix = {h: k for k, h in enumerate(hdr)}
        s = ix['# Samples']; ex = ix['Instructions Executed']
        tot = sum(int(r[s] or 0) for r in body)
        print('=====', name[:120], 'samples', tot, 'sass lines', len(body))
        stall_cols = [k for k, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
        order = sorted(range(len(body)), key=lambda k: -int(body[k][s] or 0))[:top]
        for k in sorted(order):
            r = body[k]
            st = sorted(((int(r[c] or 0), hdr[c][6:]) for c in stall_cols), reverse=True)[:2]
            print(f'{k:5d} {int(r[s] or 0):6d} {100.0 * int(r[s] or 0) / max(tot, 1):5.1f}% ex={r[ex]:>8} {r[1][:90]:90s} {st}')
    else:
        i += 1
