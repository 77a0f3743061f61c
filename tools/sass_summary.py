#!/usr/bin/env python3
"""Per-object counts of the SASS mnemonics that prove what each kernel file runs on (tcgen05 = UTC*MMA / LDTM / UTCBAR,
bulk TMA = UBLKCP / UBLKRED / UBLKPF, packed fp32 = FFMA2 / FADD2 / FMUL2, mbarrier = SYNCS, setmaxnreg = USETMAXREG).
    python tools/sass_summary.py > profiles/sass_r02.txt      (needs cuobjdump; reads tec_mollm_b200/build/*.o)"""
import collections, glob, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAT = ["UTCHMMA", "UTCQMMA", "UTCOMMA", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "UBLKCP", "UBLKRED", "UBLKPF", "SYNCS", "USETMAXREG",
       "FFMA2", "FADD2", "FMUL2", "FFMA", "DFMA", "DMUL", "MUFU", "HMMA", "LDG", "LDS", "STS", "STG", "ATOM", "RED"]
print("# SASS mnemonic counts per object (cuobjdump -sass tec_mollm_b200/build/*.o), sm_100a\n")
print("| object | kernels | " + " | ".join(PAT) + " |")
print("|---|---|" + "---|" * len(PAT))
for obj in sorted(glob.glob(os.path.join(ROOT, "tec_mollm_b200", "build", "*.o"))):
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    kernels = len(re.findall(r"^\s*Function :", sass, flags=re.M))
    c = collections.Counter()
    for line in sass.splitlines():
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            for p in PAT:
                if op == p or (p in ("UBLKCP", "UBLKRED", "UBLKPF", "SYNCS", "USETMAXREG", "UTCBAR", "LDTM", "MUFU", "ATOM", "RED", "LDG", "LDS", "STS", "STG") and op.startswith(p)):
                    c[p] += 1
                    break
    print(f"| {os.path.basename(obj)} | {kernels} | " + " | ".join(str(c[p]) for p in PAT) + " |")
