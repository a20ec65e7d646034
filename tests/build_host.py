"""TEST INFRASTRUCTURE: build the C++ host-layer test programs and demo mains (include/mptg/*.hpp) against
libmptg.so, in-tree.  (Lives under tests/ because the mock and the reference-parity program include oracle/ headers;
nothing under mpt_b200/ or include/ may.)

    python -m tests.build_host

Produces tests/cpp/_build/planner_test and the demo mains under demos/_build/.  They run on a GPU
box only (libmptg.so has no CPU fallback); the binaries travel with the repo snapshot.
"""
from __future__ import annotations

import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CXX = "/usr/bin/g++"
FLAGS = ["-std=c++17", "-O2", "-g", "-rdynamic", "-ffp-contract=off", "-Wall", "-Wno-unused-function", f"-I{ROOT / 'include'}"]


def _stale(out: Path, deps) -> bool:
    return not out.exists() or any(Path(d).stat().st_mtime > out.stat().st_mtime for d in deps)


def build_program(src: Path, out: Path, lib_dir: Path, lib: str, defines=()) -> Path:
    deps = [src, *sorted((ROOT / "include" / "mptg").glob("*"))]
    if _stale(out, deps):
        out.parent.mkdir(parents=True, exist_ok=True)
        rpath = "$ORIGIN/" + str(Path(*[".."] * len(out.parent.relative_to(ROOT).parts)) / lib_dir.relative_to(ROOT))
        cmd = [CXX, *FLAGS, *[f"-D{d}" for d in defines], str(src), "-o", str(out), f"-L{lib_dir}", f"-l{lib}", f"-Wl,-rpath,{rpath}", "-lpthread", "-lz"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"g++ failed for {src}:\n{r.stdout}")
    return out


def build() -> list[Path]:
    from mpt_b200 import build as b

    b.build()
    outs = [build_program(ROOT / "tests" / "cpp" / "planner_test.cpp", ROOT / "tests" / "cpp" / "_build" / "planner_test", b.LIBDIR, "mptg")]
    for src in sorted((ROOT / "demos").glob("*.cpp")):
        outs.append(build_program(src, ROOT / "demos" / "_build" / src.stem, b.LIBDIR, "mptg"))
    return outs


def build_mock() -> Path:
    """TEST ONLY: the planner test linked against a CPU mock of the C ABI (tests/cpp/mock_mptg.cpp)."""
    bdir = ROOT / "tests" / "cpp" / "_build"
    bdir.mkdir(parents=True, exist_ok=True)
    mock = bdir / "libmptg_mock.so"
    msrc = ROOT / "tests" / "cpp" / "mock_mptg.cpp"
    if _stale(mock, [msrc, *sorted((ROOT / "oracle").glob("*.hpp")), *sorted((ROOT / "include" / "mptg").glob("*"))]):
        cmd = [CXX, *FLAGS, "-fPIC", "-shared", "-fopenmp", "-mfma", "-mavx2", str(msrc), "-o", str(mock)]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"g++ failed for the mock:\n{r.stdout}")
    return build_program(ROOT / "tests" / "cpp" / "planner_test.cpp", bdir / "planner_test_mock", bdir, "mptg_mock", defines=("MPTG_TEST_MOCK_BACKEND",))


REFERENCE = Path("/root/reference")


def build_reference_parity(mock: bool = True) -> Path | None:
    """TEST ONLY: tests/cpp/reference_planner_parity.cpp -- the reference's own planner classes (compiled from
    /root/reference against the stand-in headers under oracle/shim) next to the wave planners, in one program.
    Linked against the CPU mock of the C ABI (mock=True) or against libmptg.so.  Needs /root/reference; where it is
    absent (GPU box) the binary built here is used if it travelled with the snapshot, else None."""
    bdir = ROOT / "tests" / "cpp" / "_build"
    out = bdir / ("reference_planner_parity_mock" if mock else "reference_planner_parity")
    if not REFERENCE.exists():
        return out if out.exists() else None
    src = ROOT / "tests" / "cpp" / "reference_planner_parity.cpp"
    if mock:
        build_mock()
        lib_dir, lib, defines = bdir, "mptg_mock", ["-DMPTG_TEST_MOCK_BACKEND"]
    else:
        from mpt_b200 import build as b

        b.build()
        lib_dir, lib, defines = b.LIBDIR, "mptg", []
    deps = [src, *sorted((ROOT / "include" / "mptg").glob("*")), *sorted((ROOT / "oracle" / "shim").rglob("*.hpp"))]
    if _stale(out, deps):
        rpath = "$ORIGIN/" + str(Path(*[".."] * len(out.parent.relative_to(ROOT).parts)) / lib_dir.relative_to(ROOT))
        cmd = [CXX, "-std=c++17", "-O2", "-g", "-DNDEBUG", "-ffp-contract=off", "-Wno-unused-function", "-Wno-deprecated-declarations", *defines,
               f"-I{ROOT / 'oracle' / 'shim'}", f"-I{REFERENCE / 'src'}", f"-I{REFERENCE / 'demo'}", f"-I{ROOT / 'include'}", str(src), "-o", str(out),
               f"-L{lib_dir}", f"-l{lib}", f"-Wl,-rpath,{rpath}", "-lpthread"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"g++ failed for {src}:\n{r.stdout}")
    return out


if __name__ == "__main__":
    for p in build():
        print(p)
    print(build_reference_parity(mock=False))
