"""cuobjdump -sass of libpvt.so -> profiles/sass_r2_summary.txt (+ gzipped full listings of the main kernels under profiles/sass_r2/).
Runs without a GPU.  python tools/sass_summary.py"""
import collections
import gzip
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "parallel-video-object-tracker_b200", "libpvt.so")
OUT = os.path.join(ROOT, "profiles", "sass_r2_summary.txt")
DIR = os.path.join(ROOT, "profiles", "sass_r2")
FULL = ("k_ingest", "k_ncc_finalize", "k_ncc_fringe", "k_ncc_local", "k_ncc_search", "k_ncc_tc", "k_step_fused", "k_update", "k_winstats")
KEY = ("ACQBULK", "ATOMG", "ATOMS", "BAR", "CCTL", "DADD", "DFMA", "DMUL", "ERRBAR", "FFMA", "IDP", "LDG", "LDGDEPBAR", "LDGSTS", "LDS", "LDTM", "MEMBAR",
       "NANOSLEEP", "PRMT", "REDG", "REDUX", "STG", "STS", "SYNCS", "UBLKCP", "UTCATOMSWS", "UTCBAR", "UTCIMMA", "UTCHMMA", "UTMALDG")

txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
kernels, cur = collections.OrderedDict(), None
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = []
    elif cur is not None:
        kernels[cur].append(line)


def short(mangled):
    m = re.match(r"_ZN3pvt(\d+)", mangled)
    return mangled[len(m.group(0)):len(m.group(0)) + int(m.group(1))] if m else mangled


os.makedirs(DIR, exist_ok=True)
with open(OUT, "w") as f:
    f.write("# cuobjdump -sass of libpvt.so (sm_100a), per kernel: instruction count, opcode histogram (top 12), and the mnemonics that prove\n"
            "# TMA (UTMALDG / UBLKCP), mbarrier (SYNCS), tcgen05 (UTCIMMA / UTCBAR / LDTM) and programmatic launch (ACQBULK / PREEXIT-style) use.\n"
            "# Full listings: profiles/sass_r2/<kernel>.sass.gz   (regenerate: python tools/sass_summary.py)\n")
    for name, lines in kernels.items():
        ops = collections.Counter()
        for ln in lines:
            m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", ln)
            if m:
                ops[m.group(1)] += 1
        n = sum(ops.values())
        f.write(f"\n## {short(name)}  ({name})\n  instructions: {n}\n")
        f.write("  top: " + ", ".join(f"{k} {v}" for k, v in ops.most_common(12)) + "\n")
        f.write("  key: " + ", ".join(f"{k} {ops[k]}" for k in KEY if ops[k]) + "\n")
        if short(name) in FULL:
            with gzip.open(os.path.join(DIR, short(name) + ".sass.gz"), "wt") as g:
                g.write("\n".join(lines) + "\n")
print("wrote", OUT)
