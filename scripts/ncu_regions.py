"""Aggregate an `ncu --page source --csv` SASS listing into regions separated by BAR.SYNC: executed warp instructions,
stall samples, shared-memory wavefronts per region (to see which phase of a multi-phase kernel costs what)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {n: i for i, n in enumerate(hdr)}
regions = []
cur = dict(start=0, inst=0, samples=0, wf=0, wfx=0, n=0, ops=collections.Counter(), stalls=collections.Counter())
stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
for k, r in enumerate(rows[2:]):
    if len(r) < len(hdr): continue
    src = r[ix["Source"]].strip()
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    ie = int(r[ix["Instructions Executed"]] or 0)
    cur["inst"] += ie; cur["samples"] += int(r[ix["# Samples"]] or 0); cur["n"] += 1
    cur["wf"] += int(r[ix["L1 Wavefronts Shared"]] or 0); cur["wfx"] += int(r[ix["L1 Wavefronts Shared Excessive"]] or 0)
    cur["ops"][op.split(".")[0]] += ie
    for s in stall_cols:
        cur["stalls"][s] += int(r[ix[s]] or 0)
    if op.startswith("BAR") or op.startswith("EXIT"):
        cur["end"] = k; regions.append(cur)
        cur = dict(start=k + 1, inst=0, samples=0, wf=0, wfx=0, n=0, ops=collections.Counter(), stalls=collections.Counter())
regions.append(cur)
ti = sum(r["inst"] for r in regions); ts = sum(r["samples"] for r in regions)
for r in regions:
    if r["inst"] == 0: continue
    top = ", ".join(f"{o}:{c * 100 // max(r['inst'], 1)}%" for o, c in r["ops"].most_common(7))
    st = ", ".join(f"{s[6:]}:{c * 100 // max(r['samples'], 1)}%" for s, c in r["stalls"].most_common(5))
    print(f"sass[{r['start']:5d}..] n={r['n']:5d} inst={r['inst'] / ti * 100:5.1f}% samples={r['samples'] / max(ts, 1) * 100:5.1f}% smem_wf={r['wf'] / 1e6:7.1f}M (excess {r['wfx'] / 1e6:6.1f}M) | {top} | {st}")
print("total warp instructions", ti, "samples", ts)
