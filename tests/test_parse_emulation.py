"""The device-side FASTA / bcalm2 parser (csrc/parse.cu) emulated on the CPU: its kernels are compiled as host code behind
shims (tests/emu/parse_host.cpp) and checked against a line-by-line reader on random texts.  Pins the mask arithmetic of
the parser without a GPU; the GPU suite checks the real kernels against the host reader and the oracle."""
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_parser_kernels_on_random_texts(tmp_path):
    src = (ROOT / "matchtigs_b200" / "csrc" / "parse.cu").read_text()
    a = src.index("namespace {") + len("namespace {")
    b = src.index("}  // namespace\n\nvoid build_graph_from_text(")
    (tmp_path / "parse_kernels.inc").write_text(src[a:b])
    exe = tmp_path / "parse_host"
    r = subprocess.run(["/usr/bin/g++", "-O1", "-std=c++17", "-w", "-I", str(tmp_path), "-o", str(exe), str(ROOT / "tests" / "emu" / "parse_host.cpp")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([str(exe), "30000"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
    assert "30000 cases, 0 failures" in r.stdout
