"""Debug tool: stall samples per device function (regions delimited by CALL targets) from
`ncu -i X.ncu-rep --page source --csv --print-source sass,cuda`."""
import csv, collections, re, sys, bisect
rows = list(csv.reader(open(sys.argv[1])))
secs = []; cur = None
for r in rows:
    if r and r[0] == 'File Path': cur = {'rows': [], 'hdr': None}; secs.append(cur)
    elif r and r[0] == 'Line No': cur['hdr'] = r
    elif cur is not None and cur['hdr'] is not None and r: cur['rows'].append(r)
sass = {}
for s in secs:
    h = s['hdr']; ia = h.index('Address')
    for r in s['rows']:
        if r[ia].startswith('0x'): sass[int(r[ia], 16)] = (r, h)
addrs = sorted(sass); base = addrs[0]; h = sass[base][1]
iS = h.index('# Samples'); iI = h.index('Instructions Executed'); isrc = h.index('Address') + 1
stall_cols = [i for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
targets = {0}
for a in addrs:
    t = sass[a][0][isrc]
    m = re.search(r'CALL\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)', t)
    if m: targets.add(int(m.group(1), 16) - base)
targets = sorted(targets)
reg = collections.defaultdict(lambda: [0, 0, collections.Counter(), 0, collections.Counter()])
tot = 0
for a in addrs:
    off = a - base
    i = bisect.bisect_right(targets, off) - 1
    r = sass[a][0]
    n = int(r[iS] or 0); tot += n
    e = reg[targets[i]]
    e[0] += n; e[1] += 1; e[3] += int(r[iI] or 0)
    for c in stall_cols: e[2][h[c]] += int(r[c] or 0)
    t = r[isrc].strip(); op = (t.split()[1] if t.startswith('@') else t.split()[0]).split('.')[0]
    if op in ('DFMA', 'DMUL', 'DADD'): e[4]['fp64'] += int(r[iI] or 0)
print('total samples', tot)
for k, e in sorted(reg.items(), key=lambda x: -x[1][0]):
    if e[0] < 0.005 * tot: continue
    print(hex(k), 'ninstr', e[1], 'samples %.1f%%' % (100 * e[0] / tot), 'warp-inst %.2e' % e[3], 'fp64 %.2e' % e[4]['fp64'],
          [(n[6:], '%.0f%%' % (100 * v / e[0])) for n, v in e[2].most_common(6)])
