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
PRODUCT_A = ROOT / "matchtigs_b200" / "libmatchtigs_b200.a"
NCCL_LINK: list[str] = []  # filled in below if the library is built with its NCCL entry points
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
    """Compiles every source to an object under build/ (only the stale ones), then links the shared library the
    Python mirror and the tests load, and archives the same objects into the static library north_star names
    (``libmatchtigs_b200.a``: what a Rust host links with ``cargo:rustc-link-lib=static=matchtigs_b200``)."""
    from concurrent.futures import ThreadPoolExecutor
    srcs = product_sources()
    hdrs = sorted(CSRC.glob("*.h")) + sorted(CSRC.glob("*.cuh")) + sorted((ROOT / "include").glob("*.h"))
    objdir = ROOT / "build" / "obj"
    objdir.mkdir(parents=True, exist_ok=True)
    jobs, objs = [], []
    for src in srcs:
        obj = objdir / (src.name + ".o")
        objs.append(obj)
        if force or ptxas_info or _stale(obj, [src] + hdrs):
            cmd = [_nvcc(), *NVCC_FLAGS, "-ccbin", _gxx(), "-I", str(ROOT / "include"), "-I", str(CSRC), "-c", "-o", str(obj), str(src)]
            if ptxas_info:
                cmd[1:1] = ["-Xptxas", "-v"]
            jobs.append(cmd)
    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            list(ex.map(lambda c: _run(c, verbose or ptxas_info), jobs))
    if jobs or not PRODUCT_SO.exists() or not PRODUCT_A.exists():
        cuda_lib = str(Path(_nvcc()).resolve().parent.parent / "lib64")
        _run([_gxx(), "-shared", "-o", str(PRODUCT_SO), *[str(o) for o in objs], "-L", cuda_lib, f"-Wl,-rpath,{cuda_lib}",
              "-lcudart", "-lgomp", "-lpthread", "-ldl", *NCCL_LINK], verbose)
        if PRODUCT_A.exists():
            PRODUCT_A.unlink()
        _run(["ar", "rcs", str(PRODUCT_A), *[str(o) for o in objs]], verbose)
    return PRODUCT_SO


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_product(force, verbose)
    build_oracle(force, verbose)
    build_synth(force, verbose)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose=True)
