#!/usr/bin/env bash
# Opcode mix of the main loop (largest backward branch) of the production march kernel in an object file.
# usage: tools/loop_mix.sh [kernels_fast.o] [mangled-kernel-substring]
obj="${1:-$(dirname "$0")/../par2d_b200/csrc/build/kernels_fast.o}"
pat="${2:-march_kernelILb0ELb0ELb1ELb0}"
fn=$(cuobjdump -elf "$obj" 2>/dev/null | grep -o "_ZN5par2d12${pat}[A-Za-z0-9_]*" | head -1)
tmp=$(mktemp)
cuobjdump -sass -fun "$fn" "$obj" > "$tmp"
python - "$tmp" <<'PY'
import re, sys, collections
ins = []
for l in open(sys.argv[1]):
    m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);', l)
    if m:
        ins.append((int(m.group(1), 16), re.sub(r'^@!?U?P\d+\s+', '', m.group(2).strip())))
best = (0, 0, 0)
for a, t in ins:
    if t.startswith('BRA'):
        m = re.search(r'0x([0-9a-f]+)', t)
        if m and int(m.group(1), 16) < a and a - int(m.group(1), 16) > best[0]:
            best = (a - int(m.group(1), 16), int(m.group(1), 16), a)
_, lo, hi = best
ops = collections.Counter(t.split()[0].split('.')[0] for a, t in ins if lo <= a <= hi)
tot = sum(ops.values())
fp64 = sum(v for k, v in ops.items() if k in ('DMUL', 'DFMA', 'DADD', 'DSETP'))
print(f"kernel total {len(ins)} instrs; main loop {lo:#x}..{hi:#x}: {tot} instrs (2 rows), fp64-pipe {fp64}, other {tot - fp64}")
print("  ".join(f"{v} {k}" for k, v in ops.most_common(40)))
PY
rm -f "$tmp"
