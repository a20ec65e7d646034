"""C++ host layer (include/mptg/*.hpp): the wave planners behind Planner<Scenario, Algorithm>.

CPU: the reference's planner integration tests (test/planner_integration_test.hpp:219-254) against a
TEST-ONLY mock of the C ABI backed by the oracle -- this checks the host logic (wave scheduling,
parent choice, rewiring, union-find, solution extraction, option packs), not the kernels.
GPU: the same program linked against libmptg.so.
"""
import subprocess
from pathlib import Path

import numpy as np
import pytest

from tests import build_host

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
    for name in ("PRRT:", "PRRT wave 64:", "PRRT* k-nearest:", "PRRT* r-nearest:", "PPRM:", "PRRT* invariants:", "PPRM-IRS:",
                 "PPRM-IRS keep_dense_edges, wave 64:", "PPRM-IRS spanner:", "PRRT on the Nao-cup scenario:", "PRRT* on the Nao-cup scenario:",
                 "scenario sampler option (sample(rng) + RNG):", "scenario sampler option (sampler()):"):
        assert f"PASS {name}" in out


PARITY_CASES = ("PRRT range 20:", "PRRT unbounded:", "PRRT* k-nearest:", "PRRT* r-nearest:", "PPRM:", "PPRM-IRS:", "PPRM-IRS keep_dense_edges:",
                "reference PRRT with mptg::GpuBatch:", "reference PRRT* with mptg::GpuBatch:", "reference PRRT* r-nearest with mptg::GpuBatch:",
                "reference PPRM with mptg::GpuBatch:", "binding on SE3Space<double, 50>:",
                "reference PRRT* (kNN + validity) with mptg::GpuBatch:", "reference PPRM (kNN + validity) with mptg::GpuBatch:")


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
    for name in ("PRRT on the Nao-cup scenario:", "PRRT* on the Nao-cup scenario:", "PPRM-IRS:", "PPRM-IRS keep_dense_edges, wave 64:", "PPRM-IRS spanner:",
                 "PRRT:", "PRRT device-resident:", "PPRM-IRS device-resident:", "PRRT* device-resident:", "PRRT* device-resident r-nearest:", "PRRT* device-resident invariants:", "PPRM device-resident:", "PRRT* k-nearest:", "PRRT* r-nearest:", "PPRM:", "PRRT* invariants:"):
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


def _write_dae(path, tris, up="Y_UP"):
    """A COLLADA file of one geometry (a <triangles> list) under one scene node."""
    pts = np.asarray(tris, dtype=np.float64).reshape(-1, 3)
    idx = " ".join(str(i) for i in range(pts.shape[0]))
    path.write_text(f"""<?xml version="1.0"?>
<COLLADA xmlns="http://www.collada.org/2005/11/COLLADASchema" version="1.4.1">
 <asset><up_axis>{up}</up_axis></asset>
 <library_geometries><geometry id="g"><mesh>
  <source id="p"><float_array id="pa" count="{pts.size}">{" ".join(repr(float(x)) for x in pts.ravel())}</float_array>
   <technique_common><accessor source="#pa" count="{pts.shape[0]}" stride="3"/></technique_common></source>
  <vertices id="v"><input semantic="POSITION" source="#p"/></vertices>
  <triangles count="{pts.shape[0] // 3}"><input semantic="VERTEX" source="#v" offset="0"/><p>{idx}</p></triangles>
 </mesh></geometry></library_geometries>
 <library_visual_scenes><visual_scene id="s"><node id="n"><instance_geometry url="#g"/></node></visual_scene></library_visual_scenes>
 <scene><instance_visual_scene url="#s"/></scene>
</COLLADA>
""")


@pytest.mark.gpu
def test_se3_demo_from_cfg_and_collada(tmp_path):
    """the SE(3) demo runs from the reference's input formats -- an OMPL-style .cfg naming COLLADA
    meshes (demo/se3_rigid_body_planning.cpp:240-262, demo/se3_rigid_body_scenario.hpp:164-204)."""
    from mpt_b200 import workloads as W

    robot, env, _, _ = W.alpha_puzzle_like(env_tris_target=600, robot_tris_target=200)
    _write_dae(tmp_path / "env.dae", env)
    _write_dae(tmp_path / "robot.dae", robot)
    (tmp_path / "problem.cfg").write_text("""[problem]
name = tubes
robot = robot.dae
world = env.dae
start.x = -52
start.y = -50
start.z = 0
start.theta = 0
start.axis.x = 1
start.axis.y = 0
start.axis.z = 0
goal.x = 52
goal.y = 50
goal.z = 5
goal.theta = 0
goal.axis.x = 1
goal.axis.y = 0
goal.axis.z = 0
volume.min.x = -60
volume.min.y = -60
volume.min.z = -40
volume.max.x = 60
volume.max.y = 60
volume.max.z = 40
[planner]
rrt.range = 40
""")
    prog = build_host.build()[1]
    r = subprocess.run([str(prog), "--demo", "se3_rigid_body", "--cfg", str(tmp_path / "problem.cfg"), "--time-ms", "20000", "--check"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    print(r.stdout[-1500:])
    assert r.returncode == 0, r.stdout[-3000:]
    assert "world env.dae" in r.stdout and "OK se3_rigid_body [PRRT*]" in r.stdout
