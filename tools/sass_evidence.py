"""Mnemonic counts per kernel of the shipped librlvi_b200.so + the instructions that prove the Blackwell-native path
(cuobjdump -sass; run in the build container):  python tools/sass_evidence.py > profiles/r02_sass_evidence.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "rlvi_b200", "librlvi_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
KEYS = ["UTCHMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UTMALDG", "UBLKCP", "SYNCS", "USETMAXREG", "UCGABAR", "DMMA",
        "FMUL2", "FFMA2", "MUFU", "DFMA", "FFMA", "LDS", "STS", "LDG", "STG", "BAR", "F2F"]
SHOW = ["UTCHMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "UTMALDG", "UBLKCP", "DMMA", "USETMAXREG", "UCGABAR"]
print("# SASS evidence for the shipped librlvi_b200.so (cuobjdump -sass; sm_100a), round 2 -- tools/sass_evidence.py")
print("# mnemonic counts per kernel + the first instructions that prove the Blackwell-native path (B200_PROFILING.md table):")
print("#   tcgen05.mma -> UTCHMMA (kind::tf32 shares the H opcode; .2CTA = cta_group::2), tcgen05.commit -> UTCBAR (.MULTICAST),")
print("#   tcgen05.ld -> LDTM, tcgen05.alloc/dealloc -> UTCATOMSWS, cp.async.bulk.tensor -> UTMALDG, cp.async.bulk -> UBLKCP,")
print("#   mbarrier -> SYNCS, barrier.cluster -> UCGABAR, mma.sync.m8n8k4.f64 -> DMMA (tcgen05 has no f64 kind),")
print("#   setmaxnreg -> USETMAXREG, mul/fma.f32x2 -> FMUL2 / FFMA2\n")
name, body = None, []


def flush():
    if name is None:
        return
    cnt = collections.Counter()
    shown = collections.defaultdict(list)
    for ln in body:
        m = re.search(r"^\s*/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if not m:
            continue
        op = m.group(2)
        for k in KEYS:
            if op.startswith(k):
                cnt[k] += 1
                if k in SHOW and len(shown[k]) < 2:
                    shown[k].append(ln.split("/*", 2)[1].split("*/", 1)[-1].strip().rstrip(";").strip()[:110])
                break
    if not cnt:
        return
    print("== " + demangle(name)[:200])
    print("   " + "  ".join(f"{k}={cnt[k]}" for k in KEYS if cnt[k]))
    for k in SHOW:
        for ln in shown[k]:
            print("      " + ln)
    print()


for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        flush()
        name, body = m.group(1), []
    elif name is not None:
        body.append(ln)
flush()
