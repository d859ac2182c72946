"""Instruction-mix summary of the built library (cuobjdump -sass): per kernel the counts of the mnemonics that prove which
hardware paths the code uses -- TMA (UTMALDG / UBLKCP / UTMAPF), mbarriers (SYNCS), tensor cores + TMEM (UTCHMMA / LDTM /
UTCBAR / UTCATOMSWS), packed FP32 (FFMA2), MUFU -- and the total.  Writes profiles/sass_summary.txt.
    python tools/sass_summary.py [path/to/lib.so]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "huggingface_asr_b200", "libctcps_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
KEYS = ["UTMALDG", "UBLKCP", "UTMAPF", "SYNCS", "UTCHMMA", "LDTM", "UTCBAR", "UTCATOMSWS", "FFMA2", "FFMA", "MUFU", "LDS", "STG", "LDG", "ATOMS", "BAR"]
kern, counts = None, collections.OrderedDict()
arch = set()
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        kern = m.group(1)
        counts[kern] = collections.Counter()
        continue
    m = re.search(r"arch = (sm_\w+)", ln)
    if m:
        arch.add(m.group(1))
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
    if m and kern:
        op = m.group(1)
        counts[kern]["total"] += 1
        for k in KEYS:
            if op == k or (k in ("UTMALDG", "LDTM", "UTCHMMA", "MUFU", "SYNCS", "UTCATOMSWS") and op.startswith(k)):
                counts[kern][k] += 1


def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    except Exception:
        return n


lines = [f"{os.path.basename(lib)}: architectures {sorted(arch)}; {len(counts)} kernels", ""]
hdr = f"{'kernel':72s} {'total':>6s} " + " ".join(f"{k:>8s}" for k in KEYS)
lines.append(hdr)
tot = collections.Counter()
for k, c in counts.items():
    name = re.sub(r"\(anonymous namespace\)::", "", demangle(k))
    name = re.sub(r"\(.*", "", name).replace("void ", "")
    lines.append(f"{name[:72]:72s} {c['total']:6d} " + " ".join(f"{c[x]:8d}" for x in KEYS))
    tot.update(c)
lines.append(f"{'ALL':72s} {tot['total']:6d} " + " ".join(f"{tot[x]:8d}" for x in KEYS))
text = "\n".join(lines) + "\n"
path = os.path.join(ROOT, "profiles", "sass_summary.txt")
open(path, "w").write(text)
print(text)
