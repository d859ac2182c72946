"""Warp-state samples of k_head_gemm by code region (TMA / MMA / drain / band barrier / row normalisation) from the
source page of an ncu capture:  ncu -i X.ncu-rep --page source --csv > X.src.csv;  python tools/head_regions.py X.src.csv"""
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


k_tma = first("UBLKCP")
k_ldtm = first("LDTM")
k_bar2 = first("BAR.SYNC.DEFER_BLOCKING 0x2")
k_call = first("CALL.REL", k_bar2)
k_exit = first("EXIT", k_call)
k_ret = first("RET.REL", k_exit)
k_bar1 = first("BAR.SYNC.DEFER_BLOCKING 0x1")
k_mma = first("UTCHMMA")
regions = [("prologue", 0, k_tma - 30), ("TMA producer", k_tma - 30, min(k_mma, k_bar1) - 60)]
if k_mma < k_bar1:
    regions += [("MMA issuer", k_mma - 60, k_bar1 - 30), ("epilogue: bias + wait for the tile", k_bar1 - 30, k_ldtm), ("epilogue: drain", k_ldtm, k_bar2 - 10)]
else:
    regions += [("epilogue: bias + wait for the tile", k_bar1 - 30, k_ldtm), ("epilogue: drain", k_ldtm, k_bar2 - 10)]
regions += [("band report + wait", k_bar2 - 10, k_call), ("tail", k_call, k_exit + 1)]
if k_mma > k_exit:
    regions += [("MMA issuer", k_exit + 1, k_ret - 200 if k_ret - 200 > k_exit else k_exit + 1)]
regions += [("rest (incl. row normalisation)", regions[-1][2], len(S))]
tot = sum(S)
print(f"total samples {tot}")
for name, a, b in regions:
    print(f"{name:36s} [{a:5d},{b:5d})  {sum(S[a:b]):7d}  {sum(S[a:b]) / tot:6.1%}")
top = sorted(range(len(S)), key=lambda k: -S[k])[:14]
for k in sorted(top):
    print(k, S[k], ex[k], src[k][:90])
