"""Summarise `ncu --page source --csv` (SASS view) of one kernel: stall samples per code region and the hottest instructions.
Regions are split at the packed-FMA main loop (first / last FFMA2) -- before it: prologue, after it: epilogue."""
import csv
import sys


def main(path, top=25):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    body = [r for r in rows[2:] if len(r) == len(hdr)]
    col = {h: i for i, h in enumerate(hdr)}
    samp = [int(r[col["# Samples"]] or 0) for r in body]
    src = [r[col["Source"]] for r in body]
    execd = [int(r[col["Instructions Executed"]] or 0) for r in body]
    tot = sum(samp)
    ffma = [i for i, s in enumerate(src) if "FFMA2" in s]
    lo, hi = (ffma[0], ffma[-1]) if ffma else (0, 0)
    # the main loop also holds the LDS / MUFU of the frame: extend to the enclosing backward branch
    regions = {"before main loop": (0, lo), "main loop (first..last FFMA2)": (lo, hi + 1), "after main loop": (hi + 1, len(body))}
    print(f"{path}: {len(body)} SASS instructions, {tot} samples, {sum(execd)} warp instructions executed")
    for name, (a, b) in regions.items():
        s = sum(samp[a:b])
        e = sum(execd[a:b])
        print(f"  {name:34s} instrs {b - a:5d}  samples {s:7d} ({s / max(tot, 1):5.1%})  executed {e:10d} ({e / max(sum(execd), 1):5.1%})")
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    for name, (a, b) in regions.items():
        agg = {h: sum(int(body[i][col[h]] or 0) for i in range(a, b)) for h in stall_cols}
        s = sum(agg.values()) or 1
        print(f"  {name}: " + ", ".join(f"{h[6:]} {v / s:.0%}" for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:6]))
    print("  hottest instructions:")
    for i in sorted(range(len(body)), key=lambda i: -samp[i])[:top]:
        where = "pre" if i < lo else ("main" if i <= hi else "epi")
        print(f"    {i:5d} {where:4s} {samp[i]:6d} ({samp[i] / max(tot, 1):5.1%})  {src[i][:90]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
