"""Per-kernel SASS instruction-class histogram of libia_b200.so (cuobjdump -sass) -> profiles/sass_r2.txt.
Evidence for which hardware paths the kernels use: HMMA (legacy mma.sync tensor path), UTC*MMA / LDTM / STTM (tcgen05 +
TMEM), UBLKCP / UTMALDG (TMA engine), LDG.E.*.256 (256-bit global loads), RED / ATOM (gradient scatter)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "instantavatar_b200", "libia_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
CLASSES = [("tcgen05 MMA (UTC*MMA)", r"\bUTC\w*MMA"), ("TMEM load/store (LDTM/STTM)", r"\b(LDTM|STTM)"), ("TMA tensor (UTMALDG/UTMASTG)", r"\bUTMA(LDG|STG)"),
           ("TMA bulk copy (UBLKCP)", r"\bUBLKCP"), ("mbarrier (SYNCS)", r"\bSYNCS"), ("legacy tensor MMA (HMMA)", r"\bHMMA"), ("LDSM (ldmatrix)", r"\bLDSM"),
           ("LDG 256-bit", r"\bLDG\.E[\w.]*\.256"), ("LDG 128-bit", r"\bLDG\.E[\w.]*\.128"), ("LDG other", r"\bLDG\b"), ("STG", r"\bSTG\b"),
           ("LDS", r"\bLDS\b"), ("STS", r"\bSTS\b"), ("RED/ATOM global", r"\b(RED|ATOMG|ATOM)\b"), ("ATOMS", r"\bATOMS\b"), ("SHFL", r"\bSHFL\b"),
           ("FFMA", r"\bFFMA\b"), ("FMUL/FADD", r"\b(FMUL|FADD)\b"), ("MUFU", r"\bMUFU\b"), ("IMAD/IADD3/LEA", r"\b(IMAD|IADD3|LEA)\b"), ("BRA/branches", r"\b(BRA|BSSY|BSYNC)\b")]
kern, hist, total = None, {}, {}
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        hist[kern], total[kern] = collections.Counter(), 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,8}\*/\s+(@!?U?P\d+\s+)?([A-Z][\w.]*)", line)
    if kern and m:
        op = m.group(2)
        total[kern] += 1
        for name, pat in CLASSES:
            if re.search(pat, op):
                hist[kern][name] += 1
                break
out = ["# SASS instruction classes per kernel (cuobjdump -sass instantavatar_b200/libia_b200.so; static counts)", ""]
for k in sorted(hist, key=lambda k: -total[k]):
    out.append(f"{k}  [{total[k]} instructions]")
    out.append("   " + ", ".join(f"{n}: {c}" for n, c in hist[k].most_common() if c))
agg = collections.Counter()
for h in hist.values():
    agg.update(h)
out += ["", "## whole library", "   " + ", ".join(f"{n}: {c}" for n, c in agg.most_common())]
path = os.path.join(ROOT, "profiles", "sass_r2.txt")
open(path, "w").write("\n".join(out) + "\n")
print("\n".join(out[-3:]))
