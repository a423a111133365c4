"""List the loops (backward branches) of a kernel's SASS with instruction mix and spill counts.
usage: python tools/sass_loops.py <lib.so> <mangled-kernel-name-substring>"""
import re
import subprocess
import sys

lib, pat = sys.argv[1], sys.argv[2]
names = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
blocks = names.split("Function : ")
for blk in blocks[1:]:
    name = blk.split("\n", 1)[0]
    if pat not in name:
        continue
    ins = []
    for l in blk.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2)))
    print(name, "instructions:", len(ins), "LDL", sum("LDL" in t for _, t in ins), "STL", sum("STL" in t for _, t in ins))
    for a, t in ins:
        m = re.search(r"BRA\S*\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a:
            lo = int(m.group(1), 16)
            body = [x for x in ins if lo <= x[0] <= a]
            cnt = lambda rx: sum(bool(re.search(rx, x[1])) for x in body)
            print(f"  loop {lo:#x}-{a:#x}: n={len(body)} f64={cnt(r'D(FMA|MUL|ADD|SETP)')} mufu={cnt('MUFU')} lds={cnt('LDS')} "
                  f"ldl={cnt('LDL')} stl={cnt('STL')} ldc={cnt('LDC')} ldgsts={cnt('LDGSTS')} imad={cnt('IMAD')}")
