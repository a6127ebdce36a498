"""Single-warp in-order issue model of a SASS region: how many cycles one warp needs for it when it runs alone.

    cuobjdump -sass -fun <mangled> obj.o | python tools/sass_sim.py <start_hex> <end_hex>

Latency / issue numbers are the ones measured on B200 (tools/micro/fp64_latency.cu: dependent DFMA every 8 cycles,
one FP64 warp instruction per 2 cycles per scheduler); the others are round figures (ALU 4, LDS 30, SHFL 25, LDG 500).
Registers of 64-bit operands are tracked as pairs.  The output is (issue-bound cycles, latency-bound cycles): when the
second is much larger than the first, the compiler's schedule leaves the FP64 pipe waiting on dependencies.
"""
import re
import sys

lo, hi = int(sys.argv[1], 16), int(sys.argv[2], 16)
LAT = {'DFMA': 8, 'DMUL': 8, 'DADD': 8, 'DSETP': 10, 'LDS': 30, 'SHFL': 25, 'LDG': 500, 'LD': 500, 'MUFU': 20}
ISSUE = {'DFMA': 2, 'DMUL': 2, 'DADD': 2, 'DSETP': 2}
WIDE = ('DFMA', 'DMUL', 'DADD', 'DSETP')
ready = {}
t = 0
issue_total = 0
n = 0
fp64 = 0
for line in sys.stdin:
    m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);', line)
    if not m:
        continue
    a = int(m.group(1), 16)
    if a < lo or a > hi:
        continue
    txt = m.group(2).strip()
    if txt.startswith('@'):
        txt = txt.split(None, 1)[1]
    op = txt.split()[0]
    base = op.split('.')[0]
    args = txt[len(op):]
    regs = re.findall(r'\bR(\d+)\b', args)
    if not regs and base not in ('BRA', 'BSSY', 'BSYNC'):
        pass
    wide = base in WIDE
    w128 = '.128' in op
    w64 = '.64' in op
    dst = []
    src = []
    if regs and base not in ('ST', 'STG', 'STS', 'BRA', 'ISETP', 'DSETP', 'FSETP'):
        d = int(regs[0])
        width = 2 if (wide or w64) else (4 if w128 else 1)
        dst = list(range(d, d + width))
        srcs = regs[1:]
    else:
        srcs = regs
    for r in srcs:
        r = int(r)
        src += [r, r + 1] if wide else [r]
    start = max([t] + [ready.get(r, 0) for r in src])
    lat = LAT.get(base, 4)
    for r in dst:
        ready[r] = start + lat
    ic = ISSUE.get(base, 1)
    t = start + ic
    issue_total += ic
    n += 1
    fp64 += wide
print('instructions %d (FP64 %d): issue-bound %d cycles, single-warp latency-bound %d cycles' % (n, fp64, issue_total, t))
