#!/usr/bin/env python
"""Derive the 32-entry table of gnn-mwvc_b200/csrc/gvc_expf.h from first principles.

glibc's expf (the function the reference's sigmoid calls, src/gnn_inference.cpp:51) looks up
T[i] = bits(2^(i/32)) - (i << 47) for i = 0..31, bits() being the IEEE-754 binary64 pattern of
the correctly rounded value.  Here 2^(i/32) is computed with 60 significant digits (decimal
arithmetic: exp(i/32 * ln 2)) and rounded to binary64 by Python's correctly rounding
Decimal -> float conversion.  Prints the table in the header's layout; with --check compares it
with the header and exits non-zero on a difference (tests/test_host_expf.py runs that).
"""
import re
import struct
import sys
from decimal import Decimal, getcontext
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def table():
    getcontext().prec = 60
    ln2 = Decimal(2).ln()
    out = []
    for i in range(32):
        v = (Decimal(i) / Decimal(32) * ln2).exp()           # 2^(i/32), 60 digits
        bits = struct.unpack("<Q", struct.pack("<d", float(v)))[0]
        out.append((bits - (i << 47)) & 0xFFFFFFFFFFFFFFFF)
    return out


def formatted(t):
    rows = []
    for r in range(0, 32, 4):
        rows.append("    " + " ".join(f"0x{v:016x}ULL," for v in t[r:r + 4]) + "\\")
    return "\n".join(rows)


def header_table():
    src = (ROOT / "gnn-mwvc_b200" / "csrc" / "gvc_expf.h").read_text()
    body = src[src.index("#define GVC_EXP2F_TAB"):src.index("static const uint64_t gvc_exp2f_tab_host")]
    return [int(x, 16) for x in re.findall(r"0x([0-9a-fA-F]{16})ULL", body)]


if __name__ == "__main__":
    t = table()
    if "--check" in sys.argv:
        h = header_table()
        bad = [i for i in range(32) if len(h) != 32 or h[i] != t[i]]
        print("expf table: header == derivation" if not bad else f"expf table differs at entries {bad}")
        sys.exit(1 if bad else 0)
    print(formatted(t))
