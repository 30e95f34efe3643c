# -*- coding: utf-8 -*-
""" Static look at the hot loops of k_perturb_m2_v2 in a built library: for every backward
branch whose body holds two warp votes and the FP64 iteration, print the instruction count
per trip (two iterations) and the classes that should not be there (moves, constant-bank
reloads).  usage: python tools/sass_hotloop.py lib.so [mangled-name-substring] """
import re, subprocess, sys

def kernels(lib, filt):
    out = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True).stdout
    cur, name = [], None
    for l in out.splitlines():
        m = re.search(r"Function : (\S+)", l)
        if m:
            if name and filt in name: yield name, cur
            name, cur = m.group(1), []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if m: cur.append((int(m.group(1), 16), m.group(2).strip()))
    if name and filt in name: yield name, cur

def main(lib, filt="k_perturb_m2_v2"):
    for name, ins in kernels(lib, filt):
        addr = {a: i for i, (a, _) in enumerate(ins)}
        print("==", name[:60], len(ins), "instructions")
        for i, (a, t) in enumerate(ins):
            m = re.search(r"BRA (0x[0-9a-f]+)", t)
            if not m or "BRA.DIV" in t: continue
            tgt = int(m.group(1), 16)
            if tgt >= a or tgt not in addr: continue
            body = [x for _, x in ins[addr[tgt]:i + 1]]
            votes = sum("VOTE.ANY" in x for x in body)
            fp = sum(bool(re.match(r"(@!?U?P\d+\s+)?D(FMA|ADD|MUL|SETP)", x)) for x in body)
            if votes not in (1, 2) or fp < 14: continue
            mov = sum(bool(re.search(r"\b(IMAD\.MOV|MOV|IMAD\.U32)\b", x)) for x in body)
            ldc = sum(bool(re.search(r"\b(LDC|LDCU|UMOV)", x)) for x in body)
            print(f"  loop @{tgt:#x}-{a:#x}: {len(body)} instr / {votes} iteration(s), FP64 {fp}, moves {mov}, const loads {ldc}")

if __name__ == "__main__":
    main(*sys.argv[1:])
