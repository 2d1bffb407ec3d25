"""Instruction mix of the innermost loops of one kernel in an object file (static SASS count).

    python tools/sass_loop_mix.py <obj-or-so> <substring of the mangled kernel name> [min loop length]

Used to budget issue slots / FMA-pipe cycles of rank_pairs before spending GPU time."""
import re
import subprocess
import sys
from collections import Counter

obj, pat = sys.argv[1], sys.argv[2]
minlen = int(sys.argv[3]) if len(sys.argv) > 3 else 150
txt = subprocess.run(['cuobjdump', '-sass', obj], capture_output=True, text=True).stdout
for f in re.split(r'\n\s*Function : ', txt):
    name = f.split('\n')[0]
    if pat not in name:
        continue
    lines = [l for l in f.split('\n') if re.match(r'\s+/\*[0-9a-f]{4,5}\*/', l)]
    addr = lambda l: int(re.match(r'\s+/\*([0-9a-f]{4,5})\*/', l).group(1), 16)
    idx = {addr(l): i for i, l in enumerate(lines)}
    print(name, len(lines), 'instructions')
    for i, l in enumerate(lines):
        m = re.search(r'BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)', l)
        if not m:
            continue
        tgt = int(m.group(1), 16)
        if tgt >= addr(l) or tgt not in idx or i - idx[tgt] < minlen:
            continue
        seg = lines[idx[tgt]:i + 1]
        c = Counter()
        for s in seg:
            mm = re.match(r'\s+/\*[0-9a-f]{4,5}\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)', s)
            if mm:
                c[mm.group(2).split('.')[0]] += 1
        packed = sum(c[k] for k in ('FFMA2', 'FMUL2', 'FADD2'))
        scalar = sum(c[k] for k in ('FFMA', 'FMUL', 'FADD', 'HFMA2', 'IMAD', 'HADD2'))
        print(f'  loop {tgt:#x}..{addr(l):#x}: {len(seg)} instr; packed fp32 {packed}, scalar fma-pipe {scalar}, '
              f'MUFU {c["MUFU"]}, SHFL {c["SHFL"]}, LDS/STS {c["LDS"] + c["STS"]}, LDL/STL {c["LDL"] + c["STL"]}; '
              f'fma-pipe cycles {2 * packed + scalar}')
        print('   ', c.most_common(30))
