"""Dev tool: every loop (backward branch) of one kernel in parcompfin_b200/libpcf.so with its instruction mix --
`python tools/sass_loops.py <mangled-name-fragment> [min-instructions] [lib]`. Used to see whether spills (STL/LDL)
sit inside a hot loop."""
import collections, re, subprocess, sys
frag = sys.argv[1]
least = int(sys.argv[2]) if len(sys.argv) > 2 else 100
lib = sys.argv[3] if len(sys.argv) > 3 else "parcompfin_b200/libpcf.so"
txt = subprocess.check_output(["cuobjdump", "-sass", lib]).decode()
for f in re.split(r"\n\s*Function : ", txt):
    name = f.split("\n", 1)[0]
    if frag not in name:
        continue
    ins = []
    for line in f.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(3), m.group(4)))
    print(name[:110], len(ins), "instructions")
    for a, o, r in ins:
        if o.startswith("BRA"):
            t = re.search(r"0x([0-9a-f]+)", r)
            if t and int(t.group(1), 16) < a:
                lo = int(t.group(1), 16)
                body = [y for x, y, _ in ins if lo <= x <= a]
                if len(body) < least:
                    continue
                c = collections.Counter(y.split(".")[0] for y in body)
                print(f"  0x{lo:x}-0x{a:x}: {len(body)} instr, FP64 {sum(c[k] for k in ('DFMA', 'DMUL', 'DADD', 'DSETP'))}, "
                      f"STL {c['STL']}, LDL {c['LDL']}, LDG {c['LDG']}, LDS {c['LDS']}, STS {c['STS']}, LDGSTS {c['LDGSTS']}, "
                      f"VOTE {c['VOTE']}, BAR {c['BAR']}")
