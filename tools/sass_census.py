#!/usr/bin/env python
"""Per-kernel SASS mnemonic census of the built objects (cuobjdump -sass; no GPU needed):
    python tools/sass_census.py > profiles/r02c_sass_census.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neuromancer_b200 import build, configs  # noqa: E402

KEYS = ["UTCHMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "SYNCS", "LDGMC", "REDG", "MUFU", "F2FP", "HADD2", "FFMA2", "FFMA", "STG", "LDG", "LDS"]


def census(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], stdout=subprocess.PIPE, text=True).stdout
    per, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
            base = re.sub(r"<.*", "", name).replace("void ", "")
            tmpl = re.search(r"<(.*)>\(", name)
            extra = ""
            if tmpl:                                     # tell the instantiations of one template apart by their trailing arguments
                args = [a.strip() for a in tmpl.group(1).split(",")]
                extra = "<" + ",".join(a for a in args[1:] if re.fullmatch(r"[\w() ]+", a)) + ">" if len(args) > 1 else ""
            cur = base + extra
            k = 2
            while cur in per:
                cur = f"{base}{extra}#{k}"
                k += 1
            per[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            op = m.group(1)
            for k in KEYS:
                if op.startswith(k):
                    per[cur][k] += 1
                    break
    return per


print("# SASS opcode census of the hand-written kernels (cuobjdump -sass on neuromancer_b200/_C/obj/*.o, sm_100a; tools/sass_census.py).")
print("# UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (1-D TMA), SYNCS = mbarrier ops,")
print("# LDGMC = multimem.ld_reduce (NVLS), REDG = red.global.add, MUFU = ex2/rcp, F2FP = cvt.rn.f16x2.f32, HADD2 = fp16 -> fp32 unpack, FFMA2 = packed fp32 pair")
objs = [("C ABI object (scoring / gradient / C_* / AdamW / exchange kernels)", os.path.join(build.OBJ_DIR, "nm_api.o"))]
for name in ("c5", "c2", "c3", "c4", "c1", "t_ude_brusselator"):
    hdr = build.write_spec(configs.build_case(name).plan().c)
    objs.append((name, os.path.join(build.OBJ_DIR, os.path.splitext(os.path.basename(hdr))[0] + ".o")))
for title, obj in objs:
    print(f"\n## {title}  ({os.path.basename(obj)})")
    for fn, c in census(obj).items():
        if sum(c.values()):
            print(f"   {fn:42s} " + "  ".join(f"{k} {c[k]}" for k in KEYS if c[k]))
