#!/usr/bin/env python3
"""Per-instruction view of an .ncu-rep captured with `ncu --set full --import-source on`: where the warps of a kernel spend
their time (warp-stall samples per SASS instruction), the executed instruction mix by opcode, and instructions / lanes per warp.

    python tools/ncu_sass_stalls.py gpurun_out/prof_r02g_steady.ncu-rep profiles/ncu_r02g_sass_stalls.md [top=20]

Reads the report here (no GPU needed): `ncu -i <rep> --page source --csv --print-source sass`."""
import collections
import csv
import io
import re
import subprocess
import sys


def I(x):
    try:
        return int(x)
    except ValueError:
        return 0


def main(rep, out, top=20):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    kernel = rows[0][1]
    hdr = rows[1]
    H = {n: i for i, n in enumerate(hdr)}
    recs = [r for r in rows[2:] if len(r) > 10]
    ns, ie, te = H["# Samples"], H["Instructions Executed"], H["Thread Instructions Executed"]
    stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
    samples = sum(I(r[ns]) for r in recs)
    first = max(I(r[ie]) for r in recs[:4])          # warps launched = executions of the first instructions
    inst = sum(I(r[ie]) for r in recs)
    thr = sum(I(r[te]) for r in recs)
    by_reason = collections.Counter()
    for r in recs:
        for k in stalls:
            by_reason[k[6:]] += I(r[H[k]])
    by_op, lanes_op = collections.Counter(), collections.Counter()
    for r in recs:
        t = re.sub(r"^@!?U?P\d+\s+", "", r[1].strip())
        op = t.split()[0].split(".")[0].rstrip(";")
        by_op[op] += I(r[ie]); lanes_op[op] += I(r[te])
    L = [f"# SASS-level summary of `{rep}`", "", f"Kernel: `{kernel}`", "",
         f"- static instructions: {len(recs)} ({len(recs) * 16 / 1024:.1f} KB)",
         f"- warps: {first}; executed warp instructions per warp: {inst / first:.1f}; active lanes per instruction: {thr / inst:.2f}",
         f"- warp-stall samples: {samples}", "",
         "## Where a warp's time goes (share of all stall samples by reason)", ""]
    L += [f"- {k}: {100 * v / samples:.1f} %" for k, v in by_reason.most_common(12)]
    L += ["", f"## The {top} instructions with the most samples (index in the kernel, share of samples, executions per warp, top reasons)", ""]
    for idx, r in sorted(enumerate(recs), key=lambda x: -I(x[1][ns]))[:top]:
        st = sorted(((k[6:], I(r[H[k]])) for k in stalls if I(r[H[k]]) > 0), key=lambda x: -x[1])[:2]
        L.append(f"- [{idx}] `{r[1].strip()[:70]}`: {100 * I(r[ns]) / samples:.2f} %, {I(r[ie]) / first:.2f} per warp, "
                 + ", ".join(f"{a} {b}" for a, b in st))
    L += ["", "## Executed instructions per warp by opcode (lanes active)", ""]
    L += [f"- {k}: {v / first:.1f} ({100 * v / inst:.1f} %), {lanes_op[k] / max(v, 1):.1f} lanes" for k, v in by_op.most_common(28)]
    open(out, "w").write("\n".join(L) + "\n")
    print("\n".join(L[:40]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 20)
