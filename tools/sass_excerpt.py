"""SASS evidence for profiles/: mnemonic histogram and the first occurrences of the instructions that prove TMA
(UTMALDG), mbarrier (SYNCS), integer dot products (IDP.4A / IDP.2A) etc. in the shipped kernels.
Usage: python tools/sass_excerpt.py [lib.so] > profiles/rN_sass_excerpt.txt   (needs cuobjdump, no GPU)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "rs-vio_b200", "lib", "librsvio_b200.so")
text = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs, cur = {}, None
for line in text.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
    elif cur:
        c = re.sub(r"/\* 0x[0-9a-f]+ \*/", "", re.sub(r"/\*[0-9a-f]{4}\*/", "", line)).strip()
        if c and not c.startswith("/*") and not c.startswith(".") and ";" in c:
            funcs[cur].append(c)
print(f"# SASS excerpts of {os.path.relpath(lib, ROOT)} (cuobjdump -sass, sm_100a); regenerate: tools/sass_excerpt.py")
WANT = (("pyr_fast_kernelILi2", ["UTMALDG", "SYNCS", "IDP.4A", "IDP.2A", "STG.E.128", "LDS.128"], 2),
        ("lk_track_fb_kernel", ["LDG.E.U8", "SHFL", "MUFU", "FFMA", "I2F"], 0),
        ("box_pass_kernel", ["LDS", "STG"], 0), ("trackb_kernel", ["LDG", "SHFL"], 0))
for key, pats, ctx in WANT:
    name = [k for k in funcs if key in k][0]
    ins = funcs[name]
    hist = collections.Counter((i.split()[1] if i.startswith("@") else i.split()[0]) for i in ins)
    print(f"\n## {name}\n{len(ins)} instructions; mnemonic histogram (top 24): " + ", ".join(f"{k} x{v}" for k, v in hist.most_common(24)))
    for p in pats:
        hits = [i for i, x in enumerate(ins) if p in x]
        print(f"-- {p}: {len(hits)} occurrences; first ones:")
        for h in hits[:3]:
            for j in range(max(0, h - ctx), min(len(ins), h + ctx + 1)):
                print("    " + ins[j])
            if ctx:
                print("    ...")
