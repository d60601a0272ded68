"""CPU: the exact-mode expf (gnn-mwvc_b200/csrc/gvc_expf.h, host build of the very header the
kernels include) is bit-identical to glibc's expf, which reference sigmoid::forward calls
(src/gnn_inference.cpp:51)."""
import subprocess

from conftest import ROOT

SRC = r"""
#include <math.h>
#include <stdio.h>
#include "gvc_expf.h"
int main(void) {
    unsigned long bad = 0, tot = 0;
    for (uint32_t u = 0; u < 0xffffff00u; u += 211) {      /* stride over all bit patterns */
        float x = gvc_u2f(u);
        if (x != x) continue;
        tot++;
        if (gvc_f2u(expf(x)) != gvc_f2u(gvc_expf_glibc(x))) bad++;
    }
    for (int i = -2000000; i <= 2000000; i++) {           /* dense in the sigmoid's working range */
        float x = (float)i * 1e-5f;
        tot++;
        if (gvc_f2u(expf(x)) != gvc_f2u(gvc_expf_glibc(x))) bad++;
    }
    printf("%lu %lu\n", tot, bad);
    return 0;
}
"""


def test_expf_bit_identical_to_libm(tmp_path):
    c = tmp_path / "t.c"
    c.write_text(SRC)
    exe = tmp_path / "t"
    subprocess.check_call(["/usr/bin/gcc", "-O2", "-ffp-contract=off", "-I", str(ROOT / "gnn-mwvc_b200" / "csrc"),
                           str(c), "-o", str(exe), "-lm"])
    tot, bad = map(int, subprocess.check_output([str(exe)]).split())
    assert tot > 20_000_000 and bad == 0


def test_expf_table_is_reproducible():
    """The 32 constants exact mode depends on come out of tools/make_expf_table.py (60-digit 2^(i/32),
    correctly rounded to binary64), entry for entry."""
    import sys
    subprocess.check_call([sys.executable, str(ROOT / "tools" / "make_expf_table.py"), "--check"])
