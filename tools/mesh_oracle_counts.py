"""Oracle traversal counts of the C5 edge wave (the numerator of the mesh roofline comes from the ORACLE's
traversal of the same inputs, not from the kernel's own counters).  CPU only; writes profiles/r2_mesh_oracle_counts.json."""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench as B
from mpt_b200 import workloads as W
from tests import oracle_binding as ob

orc = ob.load()
robot, env, vmin, vmax = W.alpha_puzzle_like(env_tris_target=4000, robot_tris_target=1000)
step = W.se3_step_size(vmin, vmax, B.SO3_W)
ea, eb = W.se3_edges(B.E_WAVE, W.EDGE_SEED, B.MESH_LO, B.MESH_HI, B.EDGE_TRANS, B.EDGE_ANGLE)
sp = ob.se3_space(B.SO3_W, B.L2_W)
og = orc.mesh_pair(robot, env, sp, step)
t0 = time.time()
ok = og.link(ea, eb)
c = og.counters()
out = {"edges": int(B.E_WAVE), "states": int(c["states"]), "bv_tests": int(c["bv_tests"]), "tri_tests": int(c["tri_tests"]),
       "valid_fraction": float(ok.mean()), "seconds": time.time() - t0, "threads": orc.threads,
       "what": "oracle (oracle.hpp MeshPair::valid under discreteMotionValid, reference visiting order, first hit ends a state and an edge) on "
               "bench.py's rank-0 edge wave: E_WAVE edges, seed EDGE_SEED"}
(ROOT / "profiles" / "r2_mesh_oracle_counts.json").write_text(json.dumps(out, indent=1) + "\n")
print(out)
