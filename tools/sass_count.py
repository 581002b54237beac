"""Dev tool: instruction mix of a kernel's hottest loop from `cuobjdump -sass parcompfin_b200/libpcf.so`.
usage: python tools/sass_count.py <mangled-name-fragment>   (prints the largest backward-branch loop's mix)"""
import collections, re, subprocess, sys
frag = sys.argv[1]
txt = subprocess.check_output(["cuobjdump", "-sass", "parcompfin_b200/libpcf.so"]).decode()
funcs = re.split(r"\n\s*Function : ", txt)
for f in funcs:
    name = f.split("\n", 1)[0]
    if frag not in name:
        continue
    ins = []
    for line in f.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(3), m.group(4)))
    loops = []
    for addr, op, rest in ins:
        if op.startswith("BRA"):
            t = re.search(r"0x([0-9a-f]+)", rest)
            if t and int(t.group(1), 16) < addr:
                loops.append((addr - int(t.group(1), 16), int(t.group(1), 16), addr))
    # hottest loop = the innermost loop with the most FP64 work: take the largest body that contains no other loop head
    loops.sort(reverse=True)
    for size, lo, hi in loops:
        body = [(a, o) for a, o, _ in ins if lo <= a <= hi]
        inner = [l for l in loops if l[1] > lo and l[2] < hi]
        if inner:
            continue
        c = collections.Counter(o.split(".")[0] for _, o in body)
        wide = sum(1 for _, o in body if o.startswith("IMAD.WIDE") or o.startswith("IMAD.HI"))
        fp64 = sum(c[k] for k in ("DFMA", "DMUL", "DADD", "DSETP"))
        print(f"{name[:100]}\n  loop 0x{lo:x}-0x{hi:x}: {len(body)} instr, FP64 {fp64}, IMAD.WIDE/HI {wide}, "
              f"FP64-pipe cycles {2 * fp64 + 4 * wide}\n  {dict(c.most_common(12))}")
        break
