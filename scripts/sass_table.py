"""Per-kernel SASS evidence table of libhavatar_b200.so (read here, no GPU): counts of the Blackwell-native mnemonics
(UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTCCP = tcgen05.cp, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk,
UTMALDG / UTMASTG = tensor-map TMA, SYNCS = mbarrier, UCGABAR = cluster barrier) and of the legacy tensor path (HMMA).
usage: python scripts/sass_table.py > profiles/r02_sass_table.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "havatar_b200", "libhavatar_b200.so")
PATS = [("UTC*MMA", r"\bUTC[A-Z]*MMA"), ("..2CTA", r"\bUTC[A-Z]*MMA\.2CTA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTCCP", r"\bUTCCP"),
        ("UTCBAR", r"\bUTCBAR"), ("UBLKCP", r"\bUBLKCP"), ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"), ("SYNCS", r"\bSYNCS"),
        ("UCGABAR", r"\bUCGABAR"), ("HMMA", r"\bHMMA"), ("LDG", r"\bLDG"), ("STG", r"\bSTG"), ("RED", r"\bRED\b|\bREDG")]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
blocks = re.split(r"\n\s*Function : \S+\n", sass)[1:]
print("libhavatar_b200.so (sm_100a): %d kernels" % len(blocks))
print("%-72s %s  instr" % ("kernel", " ".join("%7s" % p[0] for p in PATS)))
for name, body in sorted(zip(names, blocks)):
    instr = [l for l in body.split("\n") if re.search(r"/\*[0-9a-f]{4,}\*/", l)]
    text = "\n".join(instr)
    counts = [len(re.findall(pat, text)) for _, pat in PATS]
    short = re.sub(r"\(hav::.*|\(float.*|\(unsigned.*|\(void.*", "", name).replace("void ", "").replace("hav::", "")
    print("%-72s %s  %5d" % (short[:72], " ".join("%7d" % c for c in counts), len(instr)))
