#!/usr/bin/env python
"""SASS census of libpegainfer_kernels_b200.so: per object file, how many tcgen05 / TMEM / TMA / legacy-MMA
instructions the shipped sm_100a code contains (B200_PROFILING.md "What proves a Blackwell-native kernel").
No GPU needed:  python tools/sass_census.py > profiles/r2_sass_census.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "pegainfer_b200", "build")
PATTERNS = [("UTCHMMA", r"\bUTC[A-Z]*MMA"), ("UTCHMMA.2CTA", r"\bUTC[A-Z]*MMA\.2CTA"), ("UTCBAR (tcgen05.commit)", r"\bUTCBAR"),
            ("LDTM (tcgen05.ld)", r"\bLDTM"), ("UTMALDG (TMA tensor load)", r"\bUTMALDG"), ("UTMALDG.2CTA", r"\bUTMALDG\.[0-9]D\.2CTA"),
            ("UBLKCP (cp.async.bulk)", r"\bUBLKCP"), ("UTMAPF/UBLKPF (bulk prefetch)", r"\bU(TMA|BLK)PF"), ("SYNCS (mbarrier)", r"\bSYNCS"),
            ("HMMA (mma.sync)", r"\bHMMA"), ("LDSM (ldmatrix)", r"\bLDSM"), ("UCGABAR (cluster barrier)", r"\bUCGABAR"),
            ("LDGSTS (cp.async)", r"\bLDGSTS"), ("HGMMA/wgmma (must be 0)", r"\b[HQI]GMMA")]


def main():
    objs = sorted(f for f in os.listdir(OBJ) if f.endswith(".o"))
    print("# SASS census, sm_100a, cuobjdump -sass per object of pegainfer_b200/libpegainfer_kernels_b200.so")
    print("# (counts of static instructions; kernels listed with their dominant Blackwell instructions)\n")
    for o in objs:
        sass = subprocess.run(["cuobjdump", "-sass", os.path.join(OBJ, o)], capture_output=True, text=True).stdout
        if not sass.strip():
            continue
        counts = collections.OrderedDict((n, len(re.findall(p, sass))) for n, p in PATTERNS)
        kernels = re.findall(r"Function : (\S+)", sass)
        nz = ", ".join(f"{n} {c}" for n, c in counts.items() if c)
        print(f"{o:34s} kernels {len(kernels):2d} | {nz if nz else 'plain LDG/STG/FFMA'}")
    arch = subprocess.run(["cuobjdump", "-lelf", os.path.join(ROOT, "pegainfer_b200", "libpegainfer_kernels_b200.so")], capture_output=True, text=True).stdout
    print("\n# embedded cubins:", ", ".join(sorted(set(re.findall(r"sm_\d+a?", arch)))))
    need = subprocess.run(["readelf", "-d", os.path.join(ROOT, "pegainfer_b200", "libpegainfer_kernels_b200.so")], capture_output=True, text=True).stdout
    print("# NEEDED:", ", ".join(re.findall(r"NEEDED.*\[(.*?)\]", need)))


if __name__ == "__main__":
    main()
