# -*- coding: utf-8 -*-
""" Per-source-line share of executed warp instructions and stall samples of an
ncu source page dumped with
  ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > X.csv
usage: python tools/ncu_lines.py X.csv [top] """
import csv, sys, os

def main(path, top=60):
    rows = list(csv.reader(open(path)))
    hdr_i = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
    out = []
    for k, hi in enumerate(hdr_i):
        fname = os.path.basename(rows[hi - 2][1])
        h = rows[hi]
        iI, iS, iT = h.index("Instructions Executed"), h.index("# Samples"), h.index("Thread Instructions Executed")
        end = hdr_i[k + 1] - 2 if k + 1 < len(hdr_i) else len(rows)
        for r in rows[hi + 1:end]:
            if r[0] == "" or r[iI] in ("", "-"):
                continue
            try:
                out.append((fname, int(r[0]), r[1].strip(), int(r[iI]), int(r[iS]), int(r[iT])))
            except (ValueError, IndexError):
                pass    # a source line with quotes that broke the CSV row
    tot_i = sum(o[3] for o in out); tot_s = sum(o[4] for o in out)
    print(f"total warp instr {tot_i}  samples {tot_s}")
    for o in sorted(out, key=lambda o: -o[4])[:top]:
        print(f"{o[4] / tot_s * 100:5.2f}% smp {o[3] / tot_i * 100:5.2f}% ins  lanes {o[5] / max(o[3], 1):4.1f}  {o[0]}:{o[1]}  {o[2][:90]}")

if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 60)
