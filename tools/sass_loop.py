#!/usr/bin/env python
"""Static instruction mix of every backward-branch loop of one kernel in libqo100net.so:
   python tools/sass_loop.py '<mangled-name substring>' [min_instructions]"""
import collections, re, subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "qo-100-tools_b200", "lib", "libqo100net.so")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
minn = int(sys.argv[2]) if len(sys.argv) > 2 else 100
for blk in re.split(r"\n\s*Function : ", sass)[1:]:
    name = blk.split("\n", 1)[0].strip()
    if sys.argv[1] not in name:
        continue
    ins = []
    for ln in blk.split("\n"):
        m = re.search(r"/\*([0-9a-f]{4,6})\*/\s+(.*?);", ln)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    print(name, len(ins), "instructions")
    for a, t in ins:
        m = re.search(r"\bBRA\b.*0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a:
            lo = int(m.group(1), 16)
            c = collections.Counter()
            for b, u in ins:
                if lo <= b <= a:
                    p = u.split()
                    c[(p[1] if p[0].startswith("@") else p[0]).split(".")[0]] += 1
            if sum(c.values()) >= minn:
                print("  loop %05x-%05x: %d" % (lo, a, sum(c.values())), dict(c.most_common(14)))
