import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iS, iI, iT, iN = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
data = rows[2:]
tot = sum(int(r[iI]) for r in data); tots = sum(int(r[iN]) for r in data)
print("total", tot, "samples", tots)
# print contiguous regions with similar exec count: blocks split at branch targets; just print cumulative per 40-instr chunk and top-level listing
mode = sys.argv[2] if len(sys.argv) > 2 else "chunks"
if mode == "chunks":
    # segment into basic-block-like runs where count is equal
    runs = []
    cur = None
    for i, r in enumerate(data):
        c = int(r[iI])
        if cur and cur[2] == c:
            cur[1] = i; cur[3] += c; cur[4] += int(r[iN]); cur[5] += int(r[iT])
        else:
            cur = [i, i, c, c, int(r[iN]), int(r[iT])]; runs.append(cur)
    for a, b, c, s, n, t in runs:
        if s / tot > 0.004:
            print(f"{a:5d}-{b:5d} n={b-a+1:4d} count={c:12d} ins={s/tot*100:5.2f}% smp={n/tots*100:5.2f}% lanes={t/max(s,1):4.1f}  {data[a][iS].strip()[:50]}")
else:
    a, b = int(sys.argv[2]), int(sys.argv[3])
    for i in range(a, b + 1):
        r = data[i]
        print(f"{i:5d} {int(r[iI]):12d} {int(r[iT])/max(int(r[iI]),1):4.1f} {int(r[iN]):6d} {r[iS].strip()}")
