import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
iS, iI, iT, iN = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
data = rows[2:]
tot = sum(int(r[iI]) for r in data); tots = sum(int(r[iN]) for r in data)
step = int(sys.argv[2])
for a in range(0, len(data), step):
    seg = data[a:a+step]
    s = sum(int(r[iI]) for r in seg); n = sum(int(r[iN]) for r in seg); t = sum(int(r[iT]) for r in seg)
    fp = sum(int(r[iI]) for r in seg if r[iS].strip().split()[0 if not r[iS].strip().startswith('@') else 1][:2] in ("DF","DA","DM","DS"))
    print(f"{a:5d} ins={s/tot*100:5.2f}% smp={n/tots*100:5.2f}% lanes={t/max(s,1):4.1f} fp64={fp/max(s,1)*100:4.1f}%")
