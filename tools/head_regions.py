"""Warp-state samples of k_head_gemm by warp role from the source page of an ncu capture (SASS order follows the roles):
    ncu -i X.ncu-rep --page source --csv > X.src.csv;  python tools/head_regions.py X.src.csv
Regions are found by their landmark instructions: UBLKCP (TMA producer), BAR.SYNC 0x1 + the tmem_full wait + LDTM (drain),
UTCHMMA (MMA issuer)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
S = [int(r[ix["# Samples"]]) for r in data]
src = [r[ix["Source"]].strip() for r in data]
ex = [int(r[ix["Instructions Executed"]]) for r in data]


def first(pat, start=0):
    for k in range(start, len(src)):
        if pat in src[k]:
            return k
    return len(src)


def last(pat):
    for k in range(len(src) - 1, -1, -1):
        if pat in src[k]:
            return k
    return 0


marks = sorted([(first("UBLKCP"), "TMA producer"), (first("BAR.SYNC.DEFER_BLOCKING 0x1") - 40, "drain: bias + wait for the tile"),
                (first("LDTM"), "drain: TMEM loads, statistics, stores"), (first("UTCHMMA") - 80, "MMA issuer")])
marks = [(max(k, 0), n) for k, n in marks]
bounds = [(0, "prologue")] + marks
tot = sum(S)
print(f"total samples {tot}")
for i, (k, name) in enumerate(bounds):
    end = bounds[i + 1][0] if i + 1 < len(bounds) else len(S)
    if name == "drain: TMEM loads, statistics, stores":
        end = min(end, last("UTMASTG") + 40) if "UTMASTG" in " ".join(src) else end
    print(f"{name:40s} [{k:5d},{end:5d})  {sum(S[k:end]):7d}  {sum(S[k:end]) / max(tot, 1):6.1%}")
print("hottest instructions:")
for k in sorted(sorted(range(len(S)), key=lambda k: -S[k])[:12]):
    print(f"  {k:5d} {S[k]:6d} {ex[k]:9d}  {src[k][:90]}")
