"""Print head/tail of the first hot loop (>= 90 packed instr) of k_quad<float,1024,frame in smem>: sass_loop.py LIB [nhead] [ntail] [function-substring]"""
import re, subprocess, sys
lib = sys.argv[1]; nh = int(sys.argv[2]) if len(sys.argv) > 2 else 36; nt = int(sys.argv[3]) if len(sys.argv) > 3 else 24
want = sys.argv[4] if len(sys.argv) > 4 else "k_quadIfLi1024ELb1"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur = None; ins = []
for l in txt.splitlines():
    m = re.search(r"Function : (\S+)", l)
    if m: cur = m.group(1); continue
    if cur and want in cur:
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", l)
        if m: ins.append((int(m.group(1), 16), m.group(2)))
for a, t in ins:
    if "BRA" in t:
        m = re.search(r"0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a:
            tgt = int(m.group(1), 16)
            body = [x for x in ins if tgt <= x[0] <= a]
            n2 = sum(1 for x in body if re.search(r"F(FMA|MUL|ADD)2", x[1]))
            if n2 >= 90 and len(body) < 700:
                for x in body[:nh]: print(hex(x[0]), x[1])
                print("...")
                for x in body[-nt:]: print(hex(x[0]), x[1])
                break
