"""Hot SASS instructions of an ncu report's source page:
    ncu -i rep.ncu-rep --page source --csv > src.csv; python profiles/ncu_hot.py src.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
# several kernels may be concatenated: a "Kernel Name" row then a header row
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        name, hdr = rows[i][1], rows[i + 1]
        ci = {h: k for k, h in enumerate(hdr)}
        j = i + 2
        body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            body.append(rows[j]); j += 1
        tot = sum(int(r[ci["# Samples"]] or 0) for r in body)
        print(f"== {name[:100]}  total samples {tot}")
        stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        for k, r in sorted(enumerate(body), key=lambda kr: -int(kr[1][ci["# Samples"]] or 0))[:top]:
            s = int(r[ci["# Samples"]] or 0)
            why = sorted(((int(r[ci[h]] or 0), h[6:]) for h in stalls), reverse=True)[:2]
            print(f"{s:6d} {100 * s / max(tot, 1):5.1f}%  #{k:4d} {r[ci['Source']].strip()[:70]:70s} {why}")
        i = j
    else:
        i += 1
