"""Builds the product library with extra nvcc flags into build_variants/<name>.so (A/B experiments; load with MTG_LIB_PATH)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from matchtigs_b200 import _build as b

name, flags = sys.argv[1], sys.argv[2:]
out = b.ROOT / "build_variants" / f"{name}.so"
out.parent.mkdir(exist_ok=True)
cmd = [b._nvcc(), *b.NVCC_FLAGS, *flags, "-ccbin", b._gxx(), "-I", str(b.ROOT / "include"), "-I", str(b.CSRC), "-shared", "-o", str(out),
       *[str(s) for s in b.product_sources()], "-lcudart", "-lgomp"]
b._run(cmd, False)
print(out)
