"""Opcode histogram per kernel of the shipped library (cuobjdump -sass unrealgo_b200/lib/libugeval.so): the committed
evidence that the tower runs on tcgen05 / TMEM / TMA (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG =
TMA tensor load / store, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops; HMMA / WGMMA would mean legacy tensor paths).
    python profiles/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "unrealgo_b200", "lib", "libugeval.so")
KEY = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "SYNCS", "UTCATOM", "HMMA", "WGMMA", "IMMA",
       "FFMA", "LDG", "STG", "LDS", "STS", "RED", "ATOM", "BAR", "UCGABAR", "ACQBULK", "ELECT"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
            if m.group(1) in ("UTCHMMA", "UTCBAR") and ".2CTA" in m.group(2):
                cur[m.group(1) + ".2CTA"] += 1
            cur["_total"] += 1
    print(f"# {os.path.relpath(LIB, ROOT)}: arch list = "
          + subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout.replace("\n", " ").strip())
    tot = collections.Counter()
    for name, c in kernels.items():
        short = re.sub(r"\(.*", "", demangle(name).replace("(anonymous namespace)::", ""))
        cols = " ".join(f"{k}={c[k]}" for k in KEY + ["UTCHMMA.2CTA"] if c[k])
        print(f"{short:<64s} instr={c['_total']:<6d} {cols}")
        tot.update(c)
    print("# totals: " + " ".join(f"{k}={tot[k]}" for k in KEY + ["UTCHMMA.2CTA", "UTCBAR.2CTA"] if tot[k]))


if __name__ == "__main__":
    sys.exit(main())
