"""Build-time checks that need no GPU: properties of the generated SASS that measured performance depends on."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "cornell-moe_b200", "build")


def _sass(obj):
    path = os.path.join(BUILD, obj)
    if not os.path.exists(path) or shutil.which("cuobjdump") is None:
        pytest.skip("object file or cuobjdump not available (run __graft_entry__.build() first)")
    return subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout


def test_pivot_recurrence_of_the_cooperative_cholesky_is_not_spilled():
    """A spilled loop-carried register in the first pivot loop of factor_diag64 doubled the time of every panel step
    (profiles/r2_experiment_log.md); the check is the script the round used before spending GPU time."""
    if not os.path.exists(os.path.join(BUILD, "potrf_coop.o")) or shutil.which("cuobjdump") is None:
        pytest.skip("object file or cuobjdump not available")
    rc = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "check_pivot_loop_spills.py")],
                        capture_output=True, text=True)
    assert rc.returncode == 0, rc.stdout + rc.stderr


def test_hot_kernels_use_the_instructions_the_design_claims():
    """TMA tensor loads + DMMA + shared-memory operand loads in the Cholesky launch (no generic LD.E in
    the DMMA loop any more: LDS must dominate), TMA bulk copies in the fused q-KG kernel, LDGSTS in the d-KG kernel."""
    chol = _sass("potrf_coop.o")
    assert "UTMALDG" in chol and "DMMA.8x8x4" in chol and "UBLKCP" in chol
    assert chol.count(" LDS") > 4 * chol.count(" LD.E"), "operand fragments of the trailing update fell back to generic loads"
    kg = _sass("kg_mc_inst_8.o")
    assert "UBLKCP" in kg and "LDS.128" in kg
    gen = _sass("kg_mc_inst_4.o")
    assert "LDGSTS" in gen
