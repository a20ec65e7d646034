"""C++ host layer (include/mptg/*.hpp): the wave planners behind Planner<Scenario, Algorithm>.

CPU: the reference's planner integration tests (test/planner_integration_test.hpp:219-254) against a
TEST-ONLY mock of the C ABI backed by the oracle -- this checks the host logic (wave scheduling,
parent choice, rewiring, union-find, solution extraction, option packs), not the kernels.
GPU: the same program linked against libmptg.so.
"""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _run(path):
    r = subprocess.run([str(path)], capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:])
    print(r.stderr[-3000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "0 failures" in r.stdout
    return r.stdout


def test_wave_planners_host_logic_with_mock_backend():
    from tests import build_host

    out = _run(build_host.build_mock())
    for name in ("PRRT:", "PRRT wave 64:", "PRRT* k-nearest:", "PRRT* r-nearest:", "PPRM:", "PRRT* invariants:"):
        assert f"PASS {name}" in out


PARITY_CASES = ("PRRT range 20:", "PRRT unbounded:", "PRRT* k-nearest:", "PRRT* r-nearest:", "PPRM:")


def test_wave_planners_build_the_reference_planners_graphs():
    """Row a11: Planner<Scenario, PRRT / PRRTStar / PPRM <wave_size<1>>> over the CPU mock of the ABI against the
    reference's own planner classes in the same process (tests/cpp/reference_planner_parity.cpp): identical vertices,
    edges, solution paths and PRRT* solution costs.  Needs /root/reference to compile the reference side."""
    from tests import build_host

    prog = build_host.build_reference_parity(mock=True)
    if prog is None:
        pytest.skip("/root/reference not present and no prebuilt parity program")
    out = _run(prog)
    for name in PARITY_CASES:
        assert f"PASS {name}" in out


def test_demo_programs_compile():
    """The demo mains and the GPU test program build against include/mptg and libmptg.so."""
    from tests import build_host

    for p in build_host.build():
        assert p.exists()


@pytest.mark.gpu
def test_wave_planners_on_gpu():
    from tests import build_host

    prog = ROOT / "tests" / "cpp" / "_build" / "planner_test"
    if not prog.exists():
        build_host.build()
    out = _run(prog)
    for name in ("PRRT:", "PRRT device-resident:", "PRRT* device-resident:", "PRRT* device-resident r-nearest:", "PRRT* device-resident invariants:", "PPRM device-resident:", "PRRT* k-nearest:", "PRRT* r-nearest:", "PPRM:", "PRRT* invariants:"):
        assert f"PASS {name}" in out


@pytest.mark.gpu
def test_wave_planners_on_gpu_build_the_reference_planners_graphs():
    """The same comparison with the wave planners running on the device (libmptg.so): the reference's planner classes
    run on the host in the same process; graphs must be identical.  The program is built where /root/reference
    exists and travels with the snapshot."""
    from tests import build_host

    prog = build_host.build_reference_parity(mock=False)
    if prog is None:
        pytest.skip("reference parity program was not built (no /root/reference here and none shipped)")
    out = _run(prog)
    for name in PARITY_CASES:
        assert f"PASS {name}" in out


@pytest.mark.gpu
def test_demo_scenarios_on_gpu():
    """BASELINE.json configs[0..3]: the four demo scenarios solve on the device."""
    from tests import build_host

    prog = ROOT / "demos" / "_build" / "planning_demos"
    if not prog.exists():
        build_host.build()
    r = subprocess.run([str(prog), "--all", "--check", "--device-prrt"], capture_output=True, text=True, timeout=900)
    print(r.stdout[-4000:])
    print(r.stderr[-2000:])
    assert r.returncode == 0
    for name in ("holonomic_2d_point", "png_2d", "link_manipulator", "se3_rigid_body"):
        assert f"OK {name}" in r.stdout
    assert r.stdout.count("[PRRT, device-resident]") == 2 and "FAILED" not in r.stdout
