"""Opcode histogram of the large straight-line regions of a kernel's SASS (cuobjdump -sass output).

    cuobjdump -sass -fun <mangled> obj.o | python tools/sass_blocks.py [min_instructions]

Regions are delimited by backward-branch targets and unconditional control flow; forward conditional branches (rare
tail paths) stay inside a region.  Used to count FP64 / LDS / integer instructions per frame of the lane kernels.
"""
import collections
import re
import sys

minlen = int(sys.argv[1]) if len(sys.argv) > 1 else 150
ins = []
for line in sys.stdin:
    m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);', line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr_idx = {a: i for i, (a, _) in enumerate(ins)}
cuts = {0, len(ins)}
for i, (a, t) in enumerate(ins):
    op = t.split()[1] if t.startswith('@') else t.split()[0]
    if op.startswith(('BRA', 'EXIT', 'RET', 'CALL')):
        m = re.search(r'0x([0-9a-f]+)', t)
        tgt = int(m.group(1), 16) if m else None
        uncond = not t.startswith('@') and not op.startswith('CALL') and 'U' not in op.split('.')[1:]
        if uncond:
            cuts.add(i + 1)
        if tgt is not None and tgt in addr_idx:
            if tgt <= a:                     # loop head
                cuts.add(addr_idx[tgt])
            elif addr_idx[tgt] - i > 400:    # long forward jump: a dispatch between code paths
                cuts.add(addr_idx[tgt]); cuts.add(i + 1)
cuts = sorted(cuts)
for lo, hi in zip(cuts, cuts[1:]):
    if hi - lo < minlen:
        continue
    h = collections.Counter()
    for a, t in ins[lo:hi]:
        op = t.split()[1] if t.startswith('@') else t.split()[0]
        h[op.split('.')[0]] += 1
    fp64 = sum(v for k, v in h.items() if k in ('DFMA', 'DMUL', 'DADD', 'DSETP', 'DMNMX'))
    print('region %05x-%05x: %d instr, FP64 %d' % (ins[lo][0], ins[hi - 1][0], hi - lo, fp64))
    print('   ' + ', '.join('%s %d' % kv for kv in h.most_common(24)))
