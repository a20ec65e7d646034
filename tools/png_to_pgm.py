#!/usr/bin/env python3
"""Convert an RGB image to the binary PGM occupancy map planning_demos --map reads, applying the
reference's colour filters (demo/png_2d_planning.cpp:69-72, demo/png_2d_scenario.hpp:50-69).
usage: tools/png_to_pgm.py /root/reference/demo/png_planning_input.png out.pgm"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from mpt_b200 import workloads as W  # noqa: E402

occ = W.load_reference_png(Path(sys.argv[1]))
h, w = occ.shape
with open(sys.argv[2], "wb") as f:
    f.write(f"P5\n{w} {h}\n255\n".encode())
    f.write((occ * 255).astype("uint8").tobytes())
print(f"{w}x{h}, {occ.mean():.3f} occupied")
