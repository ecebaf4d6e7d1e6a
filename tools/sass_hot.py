"""Static size of the k_quad hot loops of a library build: usage sass_hot.py LIB [function-substring]"""
import re, subprocess, sys, collections
lib = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else "k_quadIfLi1024ELb1"   # the default f32 configuration
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur = None; fns = {}
for l in txt.splitlines():
    m = re.search(r"Function : (\S+)", l)
    if m: cur = m.group(1); fns[cur] = []; continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", l)
    if m and cur: fns[cur].append((int(m.group(1), 16), m.group(2)))
for name, ins in fns.items():
    if want not in name: continue
    print(name, len(ins), "instructions")
    for a, t in ins:
        if "BRA" in t:
            m = re.search(r"0x([0-9a-f]+)", t)
            if m and int(m.group(1), 16) < a:
                tgt = int(m.group(1), 16)
                body = [x[1] for x in ins if tgt <= x[0] <= a]
                n2 = sum(1 for x in body if re.search(r"F(FMA|MUL|ADD)2", x))
                if n2 >= 90 and len(body) < 700:
                    c = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", x).split()[0].split(".")[0] for x in body)
                    print(hex(tgt), hex(a), len(body), "packed", n2, dict(c.most_common(12)))
