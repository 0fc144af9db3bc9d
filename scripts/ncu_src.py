"""Per-kernel summary of an `ncu --page source --csv` dump: executed warp instructions and shared-memory wavefronts by opcode,
and the top lines by stall samples."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 12
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        name = rows[i][1]; hdr = rows[i + 1]; ix = {n: k for k, n in enumerate(hdr)}
        j = i + 2
        body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            if len(rows[j]) >= len(hdr): body.append(rows[j])
            j += 1
        ops = collections.Counter(); wf = collections.Counter(); wfx = collections.Counter(); smp = collections.Counter()
        ti = ts = 0
        for r in body:
            src = r[ix["Source"]].strip()
            t = src.split()
            op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0] if t else "?"
            full = (t[1] if t[0].startswith("@") else t[0]) if t else "?"
            ie = int(r[ix["Instructions Executed"]] or 0); ti += ie
            s = int(r[ix["# Samples"]] or 0); ts += s
            key = full if op in ("LDS", "STS", "LDG", "STG", "SHFL") else op
            ops[key] += ie; smp[key] += s
            wf[key] += int(r[ix["L1 Wavefronts Shared"]] or 0); wfx[key] += int(r[ix["L1 Wavefronts Shared Excessive"]] or 0)
        print(name[:90]); print(f"  warp instructions {ti}, samples {ts}")
        for k, c in ops.most_common(top_n):
            print(f"    {k:22s} inst {c / ti * 100:5.1f}%  samples {smp[k] / max(ts, 1) * 100:5.1f}%  smem wavefronts {wf[k] / 1e6:8.2f}M (excess {wfx[k] / 1e6:6.2f}M)")
        print(f"    total smem wavefronts {sum(wf.values()) / 1e6:.1f}M")
        i = j
    else:
        i += 1
