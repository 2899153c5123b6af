"""CPU-tier guard for the programmatically launched kernels (griddepcontrol / PDL): a kernel that may be
scheduled while its predecessor drains must not touch global memory before `griddepcontrol.wait` (SASS:
ACQBULK).  Invariant loads (`const T* __restrict__` -> LDG.CONSTANT) are fair game for the compiler to hoist
above the wait -- that happened to cg_finish_x_kernel in round 2 (it read the iteration count of the previous
kernel's reduction tail before the tail had run).  This scans the SASS of every kernel in libspmv_b200.so."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cuda-spmv-benchmark_b200", "libspmv_b200.so")


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")
@pytest.mark.skipif(not os.path.exists(LIB), reason="libspmv_b200.so not built")
def test_no_global_access_above_griddepcontrol_wait():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], check=True, capture_output=True, text=True).stdout
    fn, seen_wait, early, with_wait, bad = None, False, [], set(), {}
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn, seen_wait, early = m.group(1), False, []
            continue
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if fn is None or not m:
            continue
        ins = m.group(1)
        if "ACQBULK" in ins:
            if not seen_wait and early:
                bad[fn] = list(early)
            seen_wait = True
            with_wait.add(fn)
        elif not seen_wait and re.search(r"\b(LDG|LD\.|ATOMG|ATOM\.|REDG|RED\.|STG|ST\.)", ins):
            early.append(ins)
    assert len(with_wait) >= 20, "expected the CG / STENCIL5 kernels to carry griddepcontrol.wait"
    assert not bad, "global memory touched before griddepcontrol.wait: %r" % {k: v[:2] for k, v in bad.items()}
