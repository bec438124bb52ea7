#!/usr/bin/env python3
"""Instruction mix of the loops of a kernel, from the SASS of the built objects (no GPU needed):

    python tools/sass_loop_mix.py force_paired.cu k_force_paired [min_loop_length]

For every backward branch whose body is at least `min_loop_length` instructions long, prints the address range, the
length and the opcode histogram.  Used for the FP64-instruction counts quoted in DESIGN.md / profiles/README.md."""
import os
import re
import subprocess
import sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def kernels(obj, pattern):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    cur, body = None, []
    for line in out.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            if cur and pattern in cur:
                yield cur, body
            cur, body = m.group(1), []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m and cur:
            body.append((int(m.group(1), 16), m.group(2).strip()))
    if cur and pattern in cur:
        yield cur, body


def main():
    src, pattern = sys.argv[1], sys.argv[2]
    min_len = int(sys.argv[3]) if len(sys.argv) > 3 else 100
    obj = os.path.join(ROOT, "essa_b200", "lib", "obj", src + ".o")
    for name, ins in kernels(obj, pattern):
        demangled = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
        print("%s  (%d instructions)" % (demangled, len(ins)))
        addr = {a: k for k, (a, _) in enumerate(ins)}
        for k, (a, t) in enumerate(ins):
            m = re.search(r"BRA\S*\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)", t)
            if not m:
                continue
            tgt = int(m.group(1), 16)
            if tgt < a and tgt in addr and k + 1 - addr[tgt] >= min_len:
                body = ins[addr[tgt]:k + 1]
                c = Counter(re.sub(r"^@!?U?P\w+\s+", "", x[1]).split()[0].split(".")[0] for x in body)
                fp64 = c["DFMA"] + c["DMUL"] + c["DADD"]
                print("  loop %#06x..%#06x  %4d instr, %3d FP64 (DFMA %d DMUL %d DADD %d)  |  %s" % (
                    tgt, a, len(body), fp64, c["DFMA"], c["DMUL"], c["DADD"],
                    " ".join("%s %d" % kv for kv in c.most_common() if kv[0] not in ("DFMA", "DMUL", "DADD"))))


if __name__ == "__main__":
    main()
