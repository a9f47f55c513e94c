"""Per-kernel counts of the SASS mnemonics that prove the Blackwell-native paths (B200_PROFILING.md):
UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA tensor copies, UBLKCP = cp.async.bulk,
SYNCS = mbarrier ops, UCGABAR = cluster barrier, plus HMMA (legacy mma.sync -- must be 0).

    python tools/sass_summary.py [libb200mvs.so] > profiles/r2_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(REPO, "multi_view_stereonet_b200", "libb200mvs.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
PAT = [("UTCHMMA", r"\bUTCHMMA\b"), ("UTC*MMA(other)", r"\bUTC(?!HMMA)\w*MMA\b"), ("LDTM", r"\bLDTM\b"), ("STTM", r"\bSTTM\b"),
       ("UTMALDG", r"\bUTMALDG\b"), ("UTMASTG", r"\bUTMASTG\b"), ("UBLKCP", r"\bUBLKCP\b"),
       ("SYNCS", r"\bSYNCS\b"), ("UCGABAR", r"\bUCGABAR\w*"), ("HMMA", r"\bHMMA\b"), ("FFMA", r"\bFFMA\b")]
kernels = collections.OrderedDict()
name = None
arch = set()
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = name.replace("b200mvs::(anonymous namespace)::", "").replace("b200mvs::", "").replace("void ", "")
        name = re.sub(r"\((?:[^()]|\([^()]*\))*\)$", "", name)
        kernels[name] = collections.Counter()
        continue
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch.add(m.group(1))
    if name:
        for key, pat in PAT:
            if re.search(pat, line):
                kernels[name][key] += 1
print(f"# SASS summary of {os.path.relpath(lib, REPO)} ({os.path.getsize(lib)} bytes), arch {sorted(arch)}")
print(f"# {len(kernels)} kernels; columns: " + " ".join(k for k, _ in PAT))
tot = collections.Counter()
w = max(len(k) for k in kernels)
print(f"{'kernel':{w}s} " + " ".join(f"{k:>8s}" for k, _ in PAT))
for k, c in kernels.items():
    tot.update(c)
    print(f"{k:{w}s} " + " ".join(f"{c[key]:8d}" for key, _ in PAT))
print(f"{'TOTAL':{w}s} " + " ".join(f"{tot[key]:8d}" for key, _ in PAT))
