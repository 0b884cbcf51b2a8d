#!/usr/bin/env python
"""Per-kernel hash of the machine code in libqo100net.so (cuobjdump -sass, instruction text only: addresses, encodings and
comments dropped).  `python tools/sass_hash.py [lib] > out.json`.  Used to show that a source change elsewhere left the profiled
kernels' SASS untouched (profiles/executed_fp64.json: "_sass"), no GPU needed.  `--write`: the hashes of the object files the
library was linked from -> qo-100-tools_b200/lib/sass_hashes.json (what build() does)."""
import hashlib
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "qo-100-tools_b200", "lib", "libqo100net.so")


def sass_hashes(lib=LIB):
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    out = {}
    for blk in re.split(r"\n\s*Function : ", sass)[1:]:
        name = blk.split("\n", 1)[0].strip()
        h = hashlib.sha256()
        n = 0
        for ln in blk.split("\n"):
            m = re.search(r"/\*[0-9a-f]{4,6}\*/\s+(.*?);", ln)
            if m:
                h.update(m.group(1).strip().encode() + b"\n")
                n += 1
        out[name] = {"sha": h.hexdigest()[:16], "instructions": n}
    return out


def by_kernel(hashes):
    """family name (text before the template arguments of the demangled name) -> one hash over its instantiations"""
    names = sorted(hashes)
    dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.split("\n")
    fam = {}
    for n, d in zip(names, dem):
        m = re.search(r"(qo_\w+)", d)
        fam.setdefault(m.group(1) if m else d, []).append(hashes[n]["sha"])
    return {k: {"sha": hashlib.sha256("".join(sorted(v)).encode()).hexdigest()[:16], "instantiations": len(v)} for k, v in sorted(fam.items())}


def object_hashes(libdir=os.path.dirname(LIB)):
    """the same per-kernel hashes from the object files the library is linked from, dumped in parallel (~15 s instead of ~55 s)"""
    from concurrent.futures import ThreadPoolExecutor
    objs = sorted(os.path.join(libdir, n) for n in os.listdir(libdir) if n.endswith(".o"))
    merged = {}
    with ThreadPoolExecutor(max_workers=min(16, len(objs) or 1)) as ex:
        for h in ex.map(lambda o: sass_hashes(o) if b"nv_fatbin" in open(o, "rb").read() else {}, objs):
            merged.update(h)
    return merged


def write_lib_hashes():
    """qo-100-tools_b200/lib/sass_hashes.json next to the library it describes (build() calls this; bench.py reads it)"""
    h = object_hashes()
    out = {"kernels": by_kernel(h), "lib_sha256": hashlib.sha256(open(LIB, "rb").read()).hexdigest()}
    json.dump(out, open(os.path.join(os.path.dirname(LIB), "sass_hashes.json"), "w"), indent=1, sort_keys=True)
    return out


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--write":
        print(json.dumps(write_lib_hashes()["kernels"], indent=1, sort_keys=True))
        sys.exit(0)
    h = sass_hashes(sys.argv[1] if len(sys.argv) > 1 else LIB)
    json.dump({"functions": h, "kernels": by_kernel(h)}, sys.stdout, indent=1, sort_keys=True)
