"""Debug tool: per-loop stall breakdown from `ncu -i X.ncu-rep --page source --csv --print-source sass,cuda`."""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
secs = []; cur = None
for r in rows:
    if r and r[0] == 'File Path': cur = {'file': r[1], 'rows': [], 'hdr': None}; secs.append(cur)
    elif r and r[0] == 'Line No': cur['hdr'] = r
    elif cur is not None and cur['hdr'] is not None and r: cur['rows'].append(r)
sass = {}
for s in secs:
    h = s['hdr']; ia = h.index('Address')
    for r in s['rows']:
        if r[ia].startswith('0x'): sass[int(r[ia], 16)] = (r, h)
addrs = sorted(sass)
base = addrs[0]
h = sass[base][1]
iS = h.index('# Samples'); iI = h.index('Instructions Executed'); isrc = h.index('Address') + 1
stall_cols = [i for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
tot = sum(int(sass[a][0][iS] or 0) for a in addrs)
print('total samples', tot, 'n instr', len(addrs))
ins = [(a - base, sass[a][0][isrc].strip()) for a in addrs]
loops = []
for off, t in ins:
    m = re.search(r'BRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)', t)
    if m:
        tgt = int(m.group(1), 16) - base
        if 0 <= tgt < off: loops.append((tgt, off))
def region(lo, hi):
    c = collections.Counter(); n = 0; smp = 0; ex = 0; ops = collections.Counter()
    for a in addrs:
        off = a - base
        if lo <= off <= hi:
            r = sass[a][0]; smp += int(r[iS] or 0); n += 1
            ex = max(ex, int(r[iI] or 0))
            t = r[isrc].strip(); op = (t.split()[1] if t.startswith('@') else t.split()[0]).split('.')[0]
            ops[op] += 1
            for i in stall_cols: c[h[i]] += int(r[i] or 0)
    return n, smp, ex, c, ops
for lo, hi in sorted(loops):
    n, smp, ex, c, ops = region(lo, hi)
    if smp > thr * tot:
        f = ops['DFMA'] + ops['DMUL'] + ops['DADD']
        print(hex(lo), hex(hi), 'n', n, 'fp64', f, 'LDS', ops['LDS'], 'LDL', ops['LDL'], 'STL', ops['STL'], 'samples %.1f%%' % (100 * smp / tot), 'maxexec', ex,
              [(k[6:], '%.0f%%' % (100 * v / smp)) for k, v in c.most_common(7)])
