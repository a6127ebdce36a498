"""Turns a .cu source of bhmm_b200/csrc into host C++ for the warp emulator (tests/emu/warp_emu.h + cuda_fake.h):

    kernel<targs><<<grid, block, smem, stream>>>(args)   ->   emu::launch(emu::cfg(grid, block, smem, stream), [&] { kernel<targs>(args); })
    extern __shared__ double sm[];                        ->   double* sm = emu::dyn_smem();
    asm volatile("prefetch....                            ->   (dropped: a hint)

Test infrastructure only (tests/test_engine_emulated_cpu.py): it lets the library's HOST logic -- chain planning, workspace
layout, certification loops, the C ABI -- run on the CPU together with the kernels' emulated source."""
import re
import sys


def _match_back(s, i):
    """s[i-1] ends a kernel expression `name` or `name<...>`; return its start index."""
    j = i
    while j > 0 and s[j - 1].isspace():
        j -= 1
    if s[j - 1] == '>':
        depth = 0
        while j > 0:
            j -= 1
            if s[j] == '>':
                depth += 1
            elif s[j] == '<':
                depth -= 1
                if depth == 0:
                    break
    while j > 0 and (s[j - 1].isalnum() or s[j - 1] in '_:'):
        j -= 1
    return j


def _match_paren(s, i):
    """s[i] == '('; return index after the matching ')'."""
    depth = 0
    while True:
        if s[i] == '(':
            depth += 1
        elif s[i] == ')':
            depth -= 1
            if depth == 0:
                return i + 1
        i += 1


def hostify(src):
    src = re.sub(r'extern\s+__shared__\s+((?:unsigned\s+)?(?:long\s+long|\w+))\s+(\w+)\[\];', r'\1* \2 = reinterpret_cast<\1*>(emu::dyn_smem());', src)
    src = re.sub(r'^[^\n]*asm volatile\("prefetch[^\n]*\n', '\n', src, flags=re.M)
    out = []
    pos = 0
    while True:
        k = src.find('<<<', pos)
        if k < 0:
            out.append(src[pos:])
            break
        start = _match_back(src, k)
        end_cfg = src.find('>>>', k)
        cfg = src[k + 3:end_cfg]
        p = end_cfg + 3
        while src[p].isspace():
            p += 1
        assert src[p] == '(', src[k - 40:k + 80]
        end_args = _match_paren(src, p)
        kernel = src[start:k].strip()
        args = src[p:end_args]
        out.append(src[pos:start])
        out.append('emu::launch(emu::cfg(%s), [&] { %s%s; })' % (cfg, kernel, args))
        pos = end_args
    return ''.join(out)


if __name__ == '__main__':
    sys.stdout.write(hostify(open(sys.argv[1]).read()))
