"""Summarise an `ncu --page source --csv` SASS export: executed warp-instructions and stall samples per opcode."""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
si, ii, wi = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
mix, samp = defaultdict(int), defaultdict(int)
tot = tots = 0
lines = []
for r in rows[2:]:
    try:
        n, s = int(r[ii]), int(r[wi])
    except Exception:
        continue
    ops = r[si].split()
    op = ops[1] if ops and ops[0].startswith('@') else (ops[0] if ops else '?')
    op = op.split('.')[0] + ('.' + op.split('.')[1] if op.startswith(('MUFU', 'LDG', 'STG', 'LDS', 'STS', 'ATOM', 'RED')) and '.' in op else '')
    mix[op] += n
    samp[op] += s
    tot += n
    tots += s
    lines.append((s, n, r[si].strip()[:90]))
print('total warp-instructions', tot, 'samples', tots, 'static SASS lines', len(lines))
for op, n in sorted(mix.items(), key=lambda kv: -kv[1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print('%-12s %6.2f%% instr  %6.2f%% samples' % (op, 100.0 * n / tot, 100.0 * samp[op] / max(1, tots)))
print('--- top stall lines')
for s, n, t in sorted(lines, reverse=True)[:12]:
    print('%5.2f%% samples %5.2f%% instr  %s' % (100.0 * s / max(1, tots), 100.0 * n / tot, t))
