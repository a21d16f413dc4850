"""Opcode mix of an address range of a cuobjdump -sass listing: sass_mix.py file lo hi (hex)."""
import re, collections, sys
lo, hi = int(sys.argv[2], 16), int(sys.argv[3], 16)
ops = collections.Counter()
for l in open(sys.argv[1]):
    m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);', l)
    if not m: continue
    addr = int(m.group(1), 16)
    if lo <= addr <= hi:
        ins = re.sub(r'^@!?U?P\d+\s+', '', m.group(2).strip())
        ops[ins.split()[0]] += 1
tot = sum(ops.values())
fp64 = sum(v for k, v in ops.items() if k.startswith(('DMUL', 'DFMA', 'DADD', 'DSETP')))
mov = sum(v for k, v in ops.items() if k.startswith(('MOV', 'IMAD.MOV')))
print(f"static instrs {tot}  fp64-pipe {fp64}  moves {mov}")
print("  ".join(f"{v} {k}" for k, v in ops.most_common(32)))
