"""Checks the SASS of the cooperative Cholesky for local-memory traffic inside the pivot recurrence.

The 32-column register Cholesky of `factor_diag64` (potrf_coop.cu) is inlined into `panel_role`; its loop-carried state sits at
the edge of the 128-register budget and small changes elsewhere in that function have made ptxas spill one of its
loop-carried registers (one STL / LDL round trip per round ON the critical chain: 5.7 -> 11.2 us per 32 columns, see
r2_experiment_log.md).  Run after every change to potrf_coop.cu, before spending GPU time:

    make -C cornell-moe_b200 && python profiles/check_pivot_loop_spills.py

Prints, for each of the two pivot loops (located by their pair of MUFU.RSQ64H), the LDL / STL instructions in the window
around it.  The first loop must be clean; the second has carried two spilled words since the round-1 kernel without a
measurable cost."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
obj = os.path.join(ROOT, "cornell-moe_b200", "build", "potrf_coop.o")
sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout.split("\n")
mu = [i for i, line in enumerate(sass) if "MUFU.RSQ64H" in line]
if len(mu) < 4:
    sys.exit(f"expected two pivot loops (4 MUFU.RSQ64H), found {len(mu)}")
bad = False
for name, a in (("first half", mu[0]), ("second half", mu[2])):
    seg = sass[a - 40:a + 470]
    hits = [line.split("*/")[1].strip()[:40] for line in seg if ("LDL" in line or "STL" in line)]
    print(f"{name}: {len(hits)} local-memory instructions near the loop", hits[:6])
    bad = bad or (name == "first half" and hits)
sys.exit(1 if bad else 0)
