"""Debug tool: stall samples per CUDA source line from
`ncu -i X.ncu-rep --page source --csv --print-source sass,cuda` (top N lines)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None; hdr = None
lines = collections.Counter(); text = {}; inst = collections.Counter()
for r in rows:
    if not r: continue
    if r[0] == 'File Path': cur_file = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No': hdr = r; iS = hdr.index('# Samples'); iI = hdr.index('Instructions Executed'); continue
    if hdr is None or r[0] == '': continue
    key = (cur_file, int(r[0]))
    lines[key] += int(r[iS]) if r[iS].isdigit() else 0; inst[key] += int(r[iI]) if r[iI].isdigit() else 0; text[key] = r[1].strip()[:110]
tot = sum(lines.values())
print('total samples', tot)
byfile = collections.Counter()
for (f, l), n in lines.items(): byfile[f] += n
print({f: '%.1f%%' % (100 * n / tot) for f, n in byfile.most_common()})
for key, n in lines.most_common(topn):
    print('%-12s %5d %5.1f%% inst %.2e  %s' % (key[0], key[1], 100 * n / tot, inst[key], text[key]))
