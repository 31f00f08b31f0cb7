#!/usr/bin/env python
"""Aggregate an `ncu --page source --print-source sass --csv` export by CUDA source line.

usage: ncu_lines.py <source.csv> <nvdisasm -g -c output> [top]
Joins the SASS listing of the profile with the line table of `nvdisasm -g -c <cubin>` (instruction text sequence), then
prints stall samples and executed instructions per source line and per function.
"""
import csv
import re
import sys
from collections import defaultdict


def parse_nvdisasm(path):
    out, func, line = [], None, None
    ins_re = re.compile(r"^\s*/\*([0-9a-f]{4,})\*/\s+(.*?);")
    for ln in open(path):
        if ln.startswith(".text."):
            func = ln.strip()[6:-1]
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            line = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = ins_re.match(ln)
        if m and func:
            out.append((func, int(m.group(1), 16), line, re.sub(r"\s+", " ", m.group(2).strip())))
    return out


def main():
    src, dis = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    rows = list(csv.reader(open(src)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    col = {n: i for i, n in enumerate(hdr)}
    prof = []
    for r in rows[hi + 1:]:
        if len(r) < len(hdr) or not r[0]:
            continue
        def num(name):
            try:
                return float(r[col[name]] or 0)
            except ValueError:
                return 0.0
        stalls = {n[6:]: num(n) for n in hdr if n.startswith("stall_") and "Not Issued" not in n}
        prof.append((r[0], re.sub(r"\s+", " ", r[col["Source"]].strip().rstrip(";").strip()), num("# Samples"),
                     num("Instructions Executed"), stalls))
    dis_l = parse_nvdisasm(dis)
    # align: per function in nvdisasm, find the run of the same length in prof whose first opcodes match
    by_func = defaultdict(list)
    for d in dis_l:
        by_func[d[0]].append(d)
    texts = [p[1].split(" ")[0] for p in prof]
    used = {}
    for f, ins in by_func.items():
        ops = [i[3].split(" ")[0].lstrip("@!P0123456789UT ") or i[3] for i in ins]
        ops = [i[3] for i in ins]
        n = len(ins)
        key = [re.sub(r"^@!?U?P\d+ ", "", o).split(" ")[0] for o in ops]
        for s in range(0, len(prof) - n + 1):
            if s in used:
                continue
            pk = [re.sub(r"^@!?U?P\d+ ", "", prof[s + k][1]).split(" ")[0] for k in (0, n // 2, n - 1)]
            if pk == [key[0], key[n // 2], key[n - 1]]:
                pk_all = [re.sub(r"^@!?U?P\d+ ", "", prof[s + k][1]).split(" ")[0] for k in range(n)]
                if pk_all == key:
                    used[s] = (f, ins)
                    break
    per_line, per_func = defaultdict(lambda: [0, 0, defaultdict(float)]), defaultdict(lambda: [0, 0])
    tot_s = sum(p[2] for p in prof) or 1
    tot_i = sum(p[3] for p in prof) or 1
    for s, (f, ins) in used.items():
        for k, d in enumerate(ins):
            p = prof[s + k]
            e = per_line[(f, d[2])]
            e[0] += p[2]; e[1] += p[3]
            for a, b in p[4].items():
                e[2][a] += b
            per_func[f][0] += p[2]; per_func[f][1] += p[3]
    print(f"total samples {tot_s:.0f}, instructions {tot_i:.0f}; matched functions {len(used)}/{len(by_func)}")
    print("\nper function: samples%  inst%")
    for f, (s_, i_) in sorted(per_func.items(), key=lambda x: -x[1][0]):
        print(f"  {100 * s_ / tot_s:6.2f}  {100 * i_ / tot_i:6.2f}  {f[:100]}")
    print(f"\ntop {top} lines: samples%  inst%  func  file:line  top stalls")
    for (f, ln), (s_, i_, st) in sorted(per_line.items(), key=lambda x: -x[1][0])[:top]:
        ts = sorted(st.items(), key=lambda x: -x[1])[:3]
        print(f"  {100 * s_ / tot_s:6.2f}  {100 * i_ / tot_i:6.2f}  {f[-40:]:40s} {ln}  " + " ".join(f"{a}:{b:.0f}" for a, b in ts if b))


if __name__ == "__main__":
    main()
