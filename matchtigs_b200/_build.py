"""Builds every native artefact of the repo in-tree with explicit compiler invocations.

* ``matchtigs_b200/csrc/*.cu|*.cpp`` -> ``matchtigs_b200/libmatchtigs_b200.so``  (the product: CUDA sm_100a + C ABI)
* ``oracle/mtg_oracle.cpp``           -> ``oracle/libmtg_oracle.so``            (test infrastructure only)
* ``tools/mtg_synth.cpp``             -> ``tools/libmtg_synth.so``              (synthetic-data tooling)

The image exports CXX=/opt/gcc/bin/g++ whose OpenMP spec file is missing, so the host compiler
is resolved from PATH (``/usr/bin/g++``) instead of from $CXX.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "matchtigs_b200" / "csrc"
PRODUCT_SO = ROOT / "matchtigs_b200" / "libmatchtigs_b200.so"
ORACLE_SO = ROOT / "oracle" / "libmtg_oracle.so"
SYNTH_SO = ROOT / "tools" / "libmtg_synth.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-O3,-Wall,-pthread,-fopenmp",
    "--expt-relaxed-constexpr",
]


def _gxx() -> str:
    for cand in ("/usr/bin/g++", shutil.which("g++") or ""):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("no g++ found")


def _nvcc() -> str:
    for cand in (shutil.which("nvcc") or "", "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("no nvcc found")


def _stale(target: Path, sources: list[Path]) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(s.stat().st_mtime > t for s in sources)


def _run(cmd: list[str], verbose: bool) -> None:
    if verbose:
        print("+", " ".join(cmd), file=sys.stderr)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError(f"build failed: {' '.join(cmd[:3])} ...")
    if verbose and r.stderr:
        sys.stderr.write(r.stderr)


def build_oracle(force: bool = False, verbose: bool = False) -> Path:
    src = ROOT / "oracle" / "mtg_oracle.cpp"
    if force or _stale(ORACLE_SO, [src]):
        _run([_gxx(), "-O2", "-std=c++17", "-fPIC", "-Wall", "-pthread", "-shared", "-o", str(ORACLE_SO), str(src)], verbose)
    return ORACLE_SO


def build_synth(force: bool = False, verbose: bool = False) -> Path:
    src = ROOT / "tools" / "mtg_synth.cpp"
    if force or _stale(SYNTH_SO, [src]):
        _run([_gxx(), "-O3", "-std=c++17", "-fPIC", "-Wall", "-fopenmp", "-shared", "-o", str(SYNTH_SO), str(src)], verbose)
    return SYNTH_SO


def product_sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cpp"))


def build_product(force: bool = False, verbose: bool = False, ptxas_info: bool = False) -> Path:
    srcs = product_sources()
    deps = srcs + sorted(CSRC.glob("*.h")) + sorted(CSRC.glob("*.cuh")) + sorted((ROOT / "include").glob("*.h"))
    if force or _stale(PRODUCT_SO, deps):
        cmd = [_nvcc(), *NVCC_FLAGS, "-ccbin", _gxx(), "-I", str(ROOT / "include"), "-I", str(CSRC),
               "-shared", "-o", str(PRODUCT_SO), *[str(s) for s in srcs], "-lcudart", "-lgomp"]
        if ptxas_info:
            cmd[1:1] = ["-Xptxas", "-v"]
        _run(cmd, verbose or ptxas_info)
    return PRODUCT_SO


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_product(force, verbose)
    build_oracle(force, verbose)
    build_synth(force, verbose)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose=True)
