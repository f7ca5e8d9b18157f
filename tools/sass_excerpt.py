#!/usr/bin/env python
"""SASS of the hot inner loops, straight from the built library (evidence for DESIGN.md §4):

    python tools/sass_excerpt.py            -> profiles/r2_sass_<kernel>.txt

For every kernel: the innermost loop that contains the marker instruction (the backward branch with the most marker
instructions inside), its instruction histogram, instructions per DP cell pair where that applies, and the loop body."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "lambda_b200", "liblambda_b200.so")

KERNELS = [
    # (file tag, demangled-name regex, marker mnemonic, DP columns per thread (K) or 0)
    ("dp_score_8x19", r"swDpxKernel<8, 19, false, false>", "VIADDMNMX", 19),
    ("dp_trace_32x5", r"swDpxKernel<32, 5, true, true>", "VIADDMNMX", 5),
    ("dp_score_priv_8x10", r"swDpxKernel<8, 10, true, false>", "VIADDMNMX", 10),
    ("traceback_res", r"tracebackResKernel", "SHFL", 0),
    ("seed_spec", r"seedSpecKernel", "POPC", 0),
]


def functions():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    cur, body, res = None, [], {}
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if cur:
                res[cur] = body
            cur, body = m.group(1), []
        elif cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            body.append(line)
    if cur:
        res[cur] = body
    return res


def demangle(names):
    p = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True)
    return dict(zip(names, p.stdout.splitlines()))


def parse(line):
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", line)
    if not m:
        return None
    addr, text = int(m.group(1), 16), m.group(2).strip()
    toks = text.split()
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    return addr, text, op


def main():
    fns = functions()
    dm = demangle(list(fns))
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    for tag, rx, marker, K in KERNELS:
        cand = [n for n in fns if re.search(rx, dm.get(n, ""))]
        if not cand:
            print("not found:", rx)
            continue
        ins = [p for p in (parse(l) for l in fns[cand[0]]) if p]
        addr_idx = {a: i for i, (a, _, _) in enumerate(ins)}
        best = None
        for i, (a, text, op) in enumerate(ins):
            m = re.search(r"BRA(?:\.\w+)*\s+(?:\w+,\s*)?(0x[0-9a-f]+)", text)
            if not m:
                continue
            tgt = int(m.group(1), 16)
            if tgt >= a or tgt not in addr_idx:
                continue
            body = ins[addr_idx[tgt]:i + 1]
            n_mark = sum(1 for _, _, o in body if o.startswith(marker))
            if n_mark and (best is None or n_mark > best[0] or (n_mark == best[0] and len(body) < len(best[1]))):
                best = (n_mark, body)
        path = os.path.join(ROOT, "profiles", f"r2_sass_{tag}.txt")
        with open(path, "w") as f:
            f.write(f"kernel: {dm[cand[0]]}\nlibrary: lambda_b200/liblambda_b200.so (cuobjdump -sass), {len(ins)} instructions in total\n")
            if not best:
                f.write("no loop with the marker instruction found\n")
                continue
            body = best[1]
            hist = collections.Counter(o for _, _, o in body)
            f.write(f"hot loop: {len(body)} instructions, 0x{body[0][0]:x} .. 0x{body[-1][0]:x}\n\ninstruction mix of the loop:\n")
            for o, c in hist.most_common():
                f.write(f"  {c:5d}  {o}\n")
            if K:
                relu = sum(c for o, c in hist.items() if o.startswith("VIADDMNMX") and "RELU" in o)
                steps = max(relu // K, 1) if relu else 1  # unrolled wavefront steps inside the loop body
                pairs = steps * K
                dpx = sum(c for o, c in hist.items() if o.startswith(("VIADDMNMX", "VIMNMX3", "VIADD.16x2", "VIADD", "PRMT", "VIMNMX")))
                f.write(f"\n{steps} wavefront step(s) unrolled = {pairs} cell pairs per lane; half-rate ALU-pipe instructions "
                        f"(VIADDMNMX / VIMNMX3 / VIADD.16x2 / PRMT / VIMNMX): {dpx} = {dpx / pairs:.2f} per cell pair; "
                        f"all instructions: {len(body) / pairs:.2f} per cell pair\n")
            f.write("\nloop body:\n")
            for a, text, _ in body:
                f.write(f"  /*{a:04x}*/ {text}\n")
        print(path, len(best[1]) if best else 0)


if __name__ == "__main__":
    sys.exit(main())
