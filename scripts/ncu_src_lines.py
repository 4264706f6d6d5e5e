"""Attribute the per-SASS-instruction warp-stall samples of an `ncu --page source --csv` export to CUDA source lines.
The export carries SASS only; the line table comes from `nvdisasm -gi -c` on the cubin of the same build.
Usage: python scripts/ncu_src_lines.py <source.csv> <nvdisasm.sass> <kernel substring> [top N]"""
import collections
import csv
import re
import sys


def line_table(sass_path, kernel):
    table, cur, on = {}, None, False
    for ln in open(sass_path):
        if ln.startswith("//-----") and ".text." in ln:
            on = kernel in ln
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            # with `nvdisasm -gi` an inlined instruction carries 'inlined at "<file>", line N': report the call site
            m2 = re.search(r'inlined at "([^"]+)", line (\d+)', m.group(3))
            if m2:
                cur = (m2.group(1).split("/")[-1], int(m2.group(2)), " <- " + m.group(1).split("/")[-1] + ":" + m.group(2))
            else:
                cur = (m.group(1).split("/")[-1], int(m.group(2)), "")
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            table[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return table


def main():
    src, sass, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
    table = line_table(sass, kernel)
    rows = list(csv.reader(open(src)))
    hdr = rows[1]
    a_i, s_i = hdr.index("Address"), hdr.index("# Samples")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_")]
    base = int(rows[2][a_i], 16)
    per_line = collections.defaultdict(lambda: [0, collections.Counter(), collections.Counter()])
    total = 0
    for r in rows[2:]:
        off = int(r[a_i], 16) - base
        n = int(r[s_i] or 0)
        if not n:
            continue
        total += n
        loc, ins = table.get(off, (None, "?"))
        e = per_line[loc]
        e[0] += n
        for i in stall_cols:
            v = int(r[i] or 0)
            if v:
                e[1][hdr[i][6:]] += v
        e[2][ins.split()[0] if ins[0] != "@" else ins.split()[1]] += n
    print(f"{src}: {total} samples")
    for loc, (n, st, ops) in sorted(per_line.items(), key=lambda x: -x[1][0])[:top]:
        name = f"{loc[0]}:{loc[1]}{loc[2]}" if loc else "?"
        print(f"  {100 * n / total:5.1f}%  {name:58s} {dict(st.most_common(3))}  {dict(ops.most_common(3))}")


if __name__ == "__main__":
    main()
