"""Input formats either side of the hot path: include/mptg/formats.hpp -- PNG decoding and
the reference's obstacle colour filter (demo/png_2d_scenario.hpp:50-69,192-265), OMPL .cfg files
(demo/scenario_config.hpp), OBJ triangle soups.  The decoder is checked against PIL on generated images of every
supported colour type / bit depth / scan-line filter mix, and -- where /root/reference exists -- on the reference's own
demo/png_planning_input.png, whose filtered occupancy must match the figures of SURVEY.md (3976 x 2603, 34.6 % obstacles,
start and goal free)."""
import math
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
BUILD = ROOT / "tests" / "cpp" / "_build"
REF_PNG = Path("/root/reference/demo/png_planning_input.png")
FILTERS = [(126, 106, 61, 15), (61, 53, 6, 15), (255, 255, 255, 5)]  # demo/png_2d_planning.cpp:69-72


@pytest.fixture(scope="module")
def tool():
    BUILD.mkdir(parents=True, exist_ok=True)
    out, src = BUILD / "formats_tool", ROOT / "tests" / "cpp" / "formats_tool.cpp"
    hdr = ROOT / "include" / "mptg" / "formats.hpp"
    if not out.exists() or max(src.stat().st_mtime, hdr.stat().st_mtime) > out.stat().st_mtime:
        subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-Wall", f"-I{ROOT / 'include'}", str(src), "-o", str(out), "-lz"], check=True)
    return out


def decode(tool, cmd, path, tmp_path):
    out = tmp_path / "out.raw"
    subprocess.run([str(tool), cmd, str(path), str(out)], check=True)
    raw = out.read_bytes()
    w, h = np.frombuffer(raw[:8], dtype=np.int32)
    return np.frombuffer(raw[8:], dtype=np.uint8).reshape((h, w, 3) if cmd == "png" else (h, w))


def filter_obstacles(rgb):
    r, g, b = (rgb[..., i].astype(np.int32) for i in range(3))
    occ = np.zeros(rgb.shape[:2], dtype=bool)
    for fr, fg, fb, tol in FILTERS:
        occ |= (abs(r - fr) <= tol) & (abs(g - fg) <= tol) & (abs(b - fb) <= tol)
    return occ.astype(np.uint8)


def test_png_decoder_matches_pil_on_every_supported_layout(tool, tmp_path):
    from PIL import Image

    rng = np.random.default_rng(3)
    h, w = 37, 53
    smooth = (np.add.outer(np.arange(h) * 3, np.arange(w) * 2) % 256).astype(np.uint8)  # makes the encoder pick Sub/Up/Avg/Paeth
    noise = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    rgba = np.stack([smooth, smooth[::-1], noise[..., 2], noise[..., 3]], axis=-1)
    cases = {
        "rgb": Image.fromarray(rgba[..., :3], "RGB"),
        "rgba": Image.fromarray(rgba, "RGBA"),
        "grey": Image.fromarray(smooth, "L"),
        "grey_alpha": Image.fromarray(np.stack([smooth, noise[..., 0]], axis=-1), "LA"),
        "palette": Image.fromarray(rgba[..., :3], "RGB").quantize(colors=200),
        "palette16": Image.fromarray(rgba[..., :3], "RGB").quantize(colors=13),   # 4-bit indices
        "bilevel": Image.fromarray((smooth > 127).astype(np.uint8) * 255, "L").convert("1"),
        "grey16": Image.fromarray((smooth.astype(np.uint16) * 257 + 3)),
    }
    for name, im in cases.items():
        p = tmp_path / f"{name}.png"
        im.save(p, optimize=(name != "rgb"))
        got = decode(tool, "png", p, tmp_path)
        ref = Image.open(p)
        if name == "grey16":
            want = np.repeat((np.asarray(ref).astype(np.uint16) >> 8).astype(np.uint8)[..., None], 3, axis=-1)  # strip_16: high byte
        else:
            want = np.asarray(ref.convert("RGB"))
        assert got.shape == want.shape and np.array_equal(got, want), name
        assert np.array_equal(decode(tool, "occ", p, tmp_path), filter_obstacles(want)), name


def test_png_errors(tool, tmp_path):
    bad = tmp_path / "bad.png"
    bad.write_bytes(b"not a png at all")
    assert subprocess.run([str(tool), "png", str(bad), str(tmp_path / "o")], capture_output=True).returncode == 1
    assert subprocess.run([str(tool), "png", str(tmp_path / "missing.png"), str(tmp_path / "o")], capture_output=True).returncode == 1
    from PIL import Image

    p = tmp_path / "interlaced.png"
    # PIL cannot write Adam7: flag the header of a plain image instead
    Image.fromarray(np.zeros((8, 8, 3), np.uint8), "RGB").save(p)
    raw = bytearray(p.read_bytes())
    raw[28] = 1  # IHDR interlace method (CRC is not checked by the reader)
    p.write_bytes(bytes(raw))
    r = subprocess.run([str(tool), "png", str(p), str(tmp_path / "o")], capture_output=True, text=True)
    assert r.returncode == 1 and "interlaced" in r.stderr


@pytest.mark.skipif(not REF_PNG.exists(), reason="/root/reference not present (GPU box)")
def test_reference_demo_map_decodes_to_the_surveyed_occupancy(tool, tmp_path):
    from PIL import Image

    occ = decode(tool, "occ", REF_PNG, tmp_path)
    want = filter_obstacles(np.asarray(Image.open(REF_PNG).convert("RGB")))
    assert occ.shape == (2603, 3976) and np.array_equal(occ, want)
    assert abs(occ.mean() - 0.346) < 0.001                      # 34.6 % obstacles
    assert occ[1300, 430] == 0 and occ[950, 3150] == 0          # demo/png_2d_planning.cpp:84-86 start and goal are free


def test_cfg_and_obj_readers(tool, tmp_path):
    cfg = tmp_path / "alpha.cfg"
    cfg.write_text("""[problem]
name = alpha_demo
robot = alpha_robot.dae
world = alpha_env.dae
start.x = 1.5
start.y = -2
start.z = 3e1
start.theta = 0.5
start.axis.x = 0
start.axis.y = 0
start.axis.z = 1
goal.x = 4
goal.y = 5
goal.z = 6
goal.theta = 3.141592653589793
goal.axis.x = 1
goal.axis.y = 0
goal.axis.z = 0
volume.min.x = -10
volume.min.y = -11
volume.min.z = -12
volume.max.x = 10
volume.max.y = 11
volume.max.z = 12

  [ planner ]
rrt.range = 12.25
this line is not a property
""")
    out = subprocess.run([str(tool), "cfg", str(cfg)], check=True, capture_output=True, text=True).stdout.splitlines()
    assert out[0] == "world=alpha_env.dae robot=alpha_robot.dae"
    start = [float(v) for v in out[1].split()[1:]]
    goal = [float(v) for v in out[2].split()[1:]]
    assert start == [0.0, 0.0, math.sin(0.25), math.cos(0.25), 1.5, -2.0, 30.0]  # Quaternion(AngleAxis): (axis sin(t/2), cos(t/2))
    assert goal[:4] == [math.sin(math.pi / 2), 0.0, 0.0, math.cos(math.pi / 2)] and goal[4:] == [4.0, 5.0, 6.0]
    assert [float(v) for v in out[3].split()[1:]] == [-10, -11, -12, 10, 11, 12]
    assert out[4] == "range 12.25" and out[5] == "missing-key-throws 1"
    obj = tmp_path / "quad.obj"
    obj.write_text("# a quad, a triangle with texture/normal indices and a relative face\nv 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nv 0 0 2\n"
                   "vn 0 0 1\nf 1 2 3 4\nf 1/1/1 2/2/1 5//1\nf -1 -2 -3\n")
    r = subprocess.run([str(tool), "obj", str(obj), str(tmp_path / "tris.raw")], check=True, capture_output=True, text=True)
    assert r.stdout.strip() == "4 triangles"
    tris = np.frombuffer((tmp_path / "tris.raw").read_bytes(), dtype=np.float32).reshape(4, 3, 3)
    v = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 2]], dtype=np.float32)
    assert np.array_equal(tris, v[[[0, 1, 2], [0, 2, 3], [0, 1, 4], [4, 3, 2]]])  # fan triangulation, relative indices


DAE = """<?xml version="1.0" encoding="utf-8"?>
<!-- a cube face (polylist quad), a triangle list with interleaved normal indices, two instances -->
<COLLADA xmlns="http://www.collada.org/2005/11/COLLADASchema" version="1.4.1">
  <asset><unit name="centimeter" meter="0.01"/><up_axis>{up}</up_axis></asset>
  <library_geometries>
    <geometry id="quad-mesh" name="quad">
      <mesh>
        <source id="quad-pos">
          <float_array id="quad-pos-array" count="15">0 0 0  1 0 0  1 1 0  0 1 0  0.5 0.5 2</float_array>
          <technique_common><accessor source="#quad-pos-array" count="5" stride="3"/></technique_common>
        </source>
        <source id="quad-nrm"><float_array id="quad-nrm-array" count="3">0 0 1</float_array></source>
        <vertices id="quad-vtx"><input semantic="POSITION" source="#quad-pos"/></vertices>
        <polylist count="2">
          <input semantic="VERTEX" source="#quad-vtx" offset="0"/>
          <input semantic="NORMAL" source="#quad-nrm" offset="1"/>
          <vcount>4 3</vcount>
          <p>0 0 1 0 2 0 3 0   0 0 1 0 4 0</p>
        </polylist>
        <triangles count="1">
          <input semantic="NORMAL" source="#quad-nrm" offset="0"/>
          <input semantic="VERTEX" source="#quad-vtx" offset="1"/>
          <p>0 2 0 3 0 4</p>
        </triangles>
      </mesh>
    </geometry>
  </library_geometries>
  <library_nodes>
    <node id="shared"><translate>0 0 10</translate><instance_geometry url="#quad-mesh"/></node>
  </library_nodes>
  <library_visual_scenes>
    <visual_scene id="Scene">
      <node id="a">
        <matrix>2 0 0 1  0 2 0 2  0 0 2 3  0 0 0 1</matrix>
        <instance_geometry url="#quad-mesh"/>
        <node id="b">
          <rotate>0 0 1 90</rotate>
          <scale>1 1 0.5</scale>
          <instance_node url="#shared"/>
        </node>
      </node>
    </visual_scene>
  </library_visual_scenes>
  <scene><instance_visual_scene url="#Scene"/></scene>
</COLLADA>
"""


def test_collada_reader(tool, tmp_path):
    """COLLADA subset of include/mptg/formats.hpp (the reference reads .dae through assimp,
    demo/se3_rigid_body_scenario.hpp:164-204): polylist / triangles with interleaved inputs, fan triangulation, node
    transforms composed from the root down, <instance_node>, up-axis conversion, recentring on the vertex mean."""
    v = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0.5, 0.5, 2]], dtype=np.float64)
    faces = [[0, 1, 2], [0, 2, 3], [0, 1, 4], [2, 3, 4]]  # quad fanned round corner 0, the polylist triangle, the <triangles> one
    A = np.array([[2, 0, 0, 1], [0, 2, 0, 2], [0, 0, 2, 3], [0, 0, 0, 1]], dtype=np.float64)
    Rz = np.array([[0, -1, 0, 0], [1, 0, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=np.float64)
    S = np.diag([1, 1, 0.5, 1.0])
    T = np.eye(4)
    T[2, 3] = 10
    ups = {"Y_UP": np.eye(4), "Z_UP": np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, -1, 0, 0], [0, 0, 0, 1.0]])}
    for up, U in ups.items():
        dae = tmp_path / f"scene_{up}.dae"
        dae.write_text(DAE.format(up=up))
        want = []
        for M in (U @ A, U @ A @ Rz @ S @ T):
            p = (M[:3, :3] @ v.T).T + M[:3, 3]
            want.append(p[faces])
        want = np.concatenate(want)
        r = subprocess.run([str(tool), "dae", str(dae), str(tmp_path / "t.raw")], check=True, capture_output=True, text=True)
        assert r.stdout.strip() == "8 triangles"
        got = np.frombuffer((tmp_path / "t.raw").read_bytes(), dtype=np.float32).reshape(-1, 3, 3)
        assert np.allclose(got, want, rtol=0, atol=1e-5), up
        # recentred (the robot mesh of the reference, :181-193): mean of the 5 + 5 instanced vertices
        subprocess.run([str(tool), "mesh", str(dae), str(tmp_path / "c.raw"), "centre"], check=True, capture_output=True)
        gotc = np.frombuffer((tmp_path / "c.raw").read_bytes(), dtype=np.float32).reshape(-1, 3, 3)
        pts = np.concatenate([(M[:3, :3] @ v.T).T + M[:3, 3] for M in (U @ A, U @ A @ Rz @ S @ T)])
        assert np.allclose(gotc, want - pts.mean(axis=0), rtol=0, atol=1e-5)
    # errors: not COLLADA, no geometry, dangling reference
    bad = tmp_path / "bad.dae"
    bad.write_text("<html><body/></html>")
    assert subprocess.run([str(tool), "dae", str(bad), str(tmp_path / "x.raw")], capture_output=True).returncode == 1
    bad.write_text('<COLLADA><library_geometries/></COLLADA>')
    r = subprocess.run([str(tool), "dae", str(bad), str(tmp_path / "x.raw")], capture_output=True, text=True)
    assert r.returncode == 1 and "does not contain meshes" in r.stderr
    bad.write_text(DAE.format(up="Y_UP").replace('url="#quad-mesh"/>\n        <node', 'url="#nothing"/>\n        <node'))
    assert subprocess.run([str(tool), "dae", str(bad), str(tmp_path / "x.raw")], capture_output=True).returncode == 1
