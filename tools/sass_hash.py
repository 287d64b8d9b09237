#!/usr/bin/env python
"""Prints one md5 per kernel of the built objects (SASS text without addresses/encodings): a cheap way to see whether a source change
-- a new knob, a refactor -- altered the code of a kernel that was already verified on the GPU.  usage: tools/sass_hash.py [objects...]"""
import glob, hashlib, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
objs = sys.argv[1:] or sorted(glob.glob(os.path.join(ROOT, "build", "obj", "*.o")))
for o in objs:
    out = subprocess.run(["cuobjdump", "-sass", o], capture_output=True, text=True).stdout
    name, body = None, {}
    for l in out.splitlines():
        m = re.search(r"Function : (\S+)", l)
        if m:
            name = re.sub(r"_GLOBAL__N__[0-9a-f_]+?_cu_[0-9a-f]+", "", m.group(1)); body[name] = []; continue
        if name and re.search(r"/\*[0-9a-f]{4}\*/", l):
            body[name].append(re.sub(r"/\*[0-9a-f]+\*/", "", l).strip())
    for k, v in body.items():
        dem = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()
        dem = re.sub(r"\(anonymous namespace\)::", "", dem).split("(")[0].replace("void ", "")
        print(f"{os.path.basename(o):16s} {hashlib.md5(chr(10).join(v).encode()).hexdigest()[:12]} {len(v):5d} instr  {dem}")
