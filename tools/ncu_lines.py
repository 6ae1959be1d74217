"""Per-source-line profile of one kernel from an ncu report: joins the SASS page of the report
(`ncu -i X.ncu-rep --page source --csv`, warp-state samples per instruction) with the line table of the same cubin
(`nvdisasm -gi -c`, which knows the inlining chain of every instruction), instruction by instruction in address
order.  A line's figure is INCLUSIVE: the samples of every instruction whose inlining chain passes through it, so a
call site carries what it calls.

    cuobjdump -xelf all metada_b200/_obj/nsp_10_10.o && nvdisasm -gi -c nsp_tu.sm_100a.cubin > nsp10.sass
    ncu -i gpurun_out/r02_nsp_d.ncu-rep --page source --csv > sass.csv
    python tools/ncu_lines.py sass.csv nsp10.sass 'letkf_nsp_kernelILi10ELi256ELi2ELb0ELb0' [file.cuh] [top]
"""
import collections
import csv
import re
import sys


def load_sass(path, fn_key):
    """[(opcode, [(file, line), ...innermost first])] of the function whose section name contains fn_key"""
    out, chain, pending, inside = [], [], [], False
    for ln in open(path, errors="replace"):
        if ln.startswith(".text."):
            inside = fn_key in ln
            chain, pending = [], []
            continue
        if not inside:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            pending.append((m.group(1), int(m.group(2))))
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(.*?);", ln)
        if m:
            if pending:
                chain, pending = pending, []
            txt = m.group(1).strip()
            mm = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", txt)
            out.append((mm.group(2) if mm else txt, chain))
    return out


def load_ncu(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = []
    for r in rows[2:]:
        mm = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]])
        def f(h):
            try:
                return float(r[ix[h]])
            except (ValueError, KeyError):
                return 0.0
        data.append({"op": mm.group(2) if mm else "?", "samples": f("# Samples"), "exec": f("Instructions Executed"),
                     "stalls": {h[6:]: f(h) for h in hdr if h.startswith("stall_") and "Not Issued" not in h}})
    return data


def main():
    ncu, sass = load_ncu(sys.argv[1]), load_sass(sys.argv[2], sys.argv[3])
    only = sys.argv[4] if len(sys.argv) > 4 else "letkf_nsp.cuh"
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 60
    if len(ncu) != len(sass):
        print(f"instruction counts differ: report {len(ncu)}, cubin {len(sass)}", file=sys.stderr)
    n = min(len(ncu), len(sass))
    bad = sum(1 for i in range(n) if ncu[i]["op"] != sass[i][0])
    if bad:
        print(f"{bad} of {n} opcodes differ: is this the cubin the report was taken from?", file=sys.stderr)
    total = sum(d["samples"] for d in ncu)
    incl, excl = collections.Counter(), collections.Counter()
    stall = collections.defaultdict(collections.Counter)
    for i in range(n):
        seen = set()
        for j, (f, l) in enumerate(sass[i][1]):
            if not f.endswith(only) or (f, l) in seen:
                continue
            seen.add((f, l))
            incl[l] += ncu[i]["samples"]
            if j == 0:
                excl[l] += ncu[i]["samples"]
            for s, v in ncu[i]["stalls"].items():
                stall[l][s] += v
    print(f"total samples {total:.0f}; inclusive share per line of {only}")
    for l, v in sorted(incl.items(), key=lambda kv: -kv[1])[:top]:
        ss = ", ".join(f"{s} {c / v:.2f}" for s, c in stall[l].most_common(3))
        print(f"line {l:5d}  incl {v / total:6.3f}  excl {excl[l] / total:6.3f}   {ss}")


if __name__ == "__main__":
    main()
