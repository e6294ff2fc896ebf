"""Offline SASS of the NVRTC-specialised component kernels (needs no GPU).

    python tools/jit_sass.py [cfg2|cfg1|cfg3|cfg4] [--dump DIR]

Builds the model description of a BASELINE configuration, lets the library generate and compile its
translation unit (elisa_b200_specialise_check), dumps the cubin (ELISA_B200_DUMP_CUBIN) and prints,
per kernel, the static instruction mix of the hottest loop (the back edge with the largest body) --
what `ncu --page source` shows dynamically, available here before any GPU time is spent.
"""
import collections
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))


def model_of(name):
    from elisa_b200 import models as M
    return {
        'cfg1': lambda: M.PowerLaw(),
        'cfg2': lambda: M.CutoffPL() + M.Blackbody(),
        'cfg3': lambda: M.Band(),
        'cfg4': lambda: M.Gauss() + M.PowerLaw(),
    }[name]()


def main():
    name = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith('-') else 'cfg2'
    dump = sys.argv[sys.argv.index('--dump') + 1] if '--dump' in sys.argv else tempfile.mkdtemp()
    os.makedirs(dump, exist_ok=True)
    cubin = os.path.join(dump, f'{name}.cubin')
    os.environ['ELISA_B200_DUMP_CUBIN'] = cubin
    from elisa_b200 import _lib
    lib = _lib.load()
    cm = model_of(name).compile()
    md, _keep = cm.to_desc()
    buf = C.create_string_buffer(1 << 18)
    n = lib.elisa_b200_specialise_check(C.byref(md), max(cm.nparam, 1), buf, len(buf))
    if n <= 0:
        raise SystemExit(lib.elisa_b200_last_error().decode())
    open(os.path.join(dump, f'{name}.cu'), 'w').write(buf.value.decode())
    sass = subprocess.run(['cuobjdump', '-sass', cubin], capture_output=True, text=True, check=True).stdout
    open(os.path.join(dump, f'{name}.sass'), 'w').write(sass)
    res = subprocess.run(['cuobjdump', '-res-usage', cubin], capture_output=True, text=True).stdout
    for fn in re.split(r'\n\s*Function : ', sass)[1:]:
        fname = fn.split('\n', 1)[0].strip()
        ins = []  # (address, mnemonic, text)
        for line in fn.split('\n'):
            m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', line)
            if m:
                text = m.group(2)
                text = re.sub(r'^@!?U?P\d+\s+', '', text)
                ins.append((int(m.group(1), 16), text.split()[0].split('.')[0], m.group(2)))
        # back edges: BRA to a lower address; take the largest loop body
        best = None
        for addr, mn, text in ins:
            if mn == 'BRA':
                t = re.search(r'0x([0-9a-f]+)', text)
                if t and int(t.group(1), 16) < addr:
                    span = (int(t.group(1), 16), addr)
                    if best is None or span[1] - span[0] > best[1] - best[0]:
                        best = span
        regs = re.search(re.escape(fname) + r".*?REG:(\d+)", res, re.S)
        print(f'== {fname}: {len(ins)} instructions, REG {regs.group(1) if regs else "?"}')
        if best is None:
            continue
        body = [(mn, text) for addr, mn, text in ins if best[0] <= addr <= best[1]]
        cnt = collections.Counter(mn for mn, _ in body)
        fp64 = sum(v for k, v in cnt.items() if k in ('DFMA', 'DMUL', 'DADD', 'DSETP', 'MUFU'))
        print(f'   hottest loop: {len(body)} instructions, FP64-pipe {fp64}')
        print('   ' + ', '.join(f'{k} {v}' for k, v in cnt.most_common(24)))
    print('dumped to', dump)


if __name__ == '__main__':
    main()
