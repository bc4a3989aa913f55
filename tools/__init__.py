"""Synthetic-data tooling (not product): ctypes front-end of ``tools/mtg_synth.cpp``."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_SO = Path(__file__).resolve().parent / "libmtg_synth.so"
_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not _SO.exists():
            import sys
            sys.path.insert(0, str(_SO.parent.parent))
            from matchtigs_b200 import _build
            _build.build_synth()
        l = C.CDLL(str(_SO))
        l.mts_free.argtypes = [C.c_void_p]
        l.mts_genome.restype = C.c_void_p
        l.mts_genome.argtypes = [C.c_size_t, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_double,
                                 C.c_uint32, C.POINTER(C.c_size_t)]
        l.mts_pangenome.restype = C.c_void_p
        l.mts_pangenome.argtypes = [C.c_char_p, C.c_size_t, C.c_uint32, C.c_uint64, C.c_double, C.c_double, C.c_double,
                                    C.POINTER(C.c_size_t)]
        l.mts_unitigs.restype = C.c_void_p
        l.mts_unitigs.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_size_t),
                                  C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        _lib = l
    return _lib


def _take(p, n) -> bytes:
    if not p:
        raise RuntimeError("synth tool failed (bad k or non-ACGT input)")
    try:
        return bytes((C.c_char * n.value).from_address(p))  # (string_at takes a C int: texts of 2 GiB and more overflow it)
    finally:
        lib().mts_free(p)


def genome(length: int, seed: int, families: int = 0, copies: int = 0, min_len: int = 300, max_len: int = 5000,
           divergence: float = 0.0, tandem_arrays: int = 0) -> bytes:
    """Uniform ACGT background with `families` repeat families of `copies` copies each."""
    n = C.c_size_t()
    p = lib().mts_genome(length, seed, families, copies, min_len, max_len, divergence, tandem_arrays, C.byref(n))
    return _take(p, n)


def pangenome(ancestor: bytes, strains: int, seed: int, snp_site_rate: float = 0.02, indel_site_rate: float = 0.0005,
              private_snp_rate: float = 0.0) -> list[bytes]:
    n = C.c_size_t()
    p = lib().mts_pangenome(ancestor, len(ancestor), strains, seed, snp_site_rate, indel_site_rate, private_snp_rate,
                            C.byref(n))
    return _take(p, n).split(b"\n")[:-1]


def unitigs(seqs, k: int, threads: int = 0) -> tuple[bytes, int, int]:
    """Compacted de Bruijn graph of `seqs` -> (bcalm2-style FASTA text, #distinct k-mers, #unitigs)."""
    if isinstance(seqs, (bytes, bytearray)):
        seqs = [bytes(seqs)]
    text = b"\n".join(seqs) + b"\n"
    n, nk, nu = C.c_size_t(), C.c_uint64(), C.c_uint64()
    p = lib().mts_unitigs(text, len(text), k, threads, C.byref(n), C.byref(nk), C.byref(nu))
    return _take(p, n), nk.value, nu.value


# ---- the BASELINE.json configs as reproducible recipes ----
def config_unitigs(name: str, scale: float = 1.0, threads: int = 0) -> tuple[bytes, int, dict]:
    """Returns (bcalm-style unitig FASTA, k, info) for a named config.

    name: "ecoli" (configs 1/2), "chr1" (3), "pangenome" (4), "human" (5).  `scale` shrinks the genome
    length (1.0 = the size BASELINE.json names) so tests can run the same recipe in seconds.
    """
    if name == "ecoli":
        k = 31
        g = genome(int(4_600_000 * scale), 1, families=20, copies=10, min_len=300, max_len=5000, divergence=0.02,
                   tandem_arrays=int(40 * scale) + 1)
        seqs = [g]
    elif name == "chr1":
        k = 31
        n = int(250_000_000 * scale)
        g = genome(n, 3, families=max(4, int(400 * scale)), copies=250, min_len=300, max_len=6000, divergence=0.10,
                   tandem_arrays=int(2000 * scale) + 1)
        seqs = [g]
    elif name == "pangenome":
        k = 31
        anc = genome(int(4_600_000 * scale), 1, families=20, copies=10, min_len=300, max_len=5000, divergence=0.02,
                     tandem_arrays=int(40 * scale) + 1)
        seqs = pangenome(anc, 100, 100, snp_site_rate=0.03, indel_site_rate=0.001, private_snp_rate=0.0002)
    elif name == "human":
        k = 51
        n = int(3_100_000_000 * scale)
        chrom = max(1, n // 23)
        seqs = [genome(chrom, 500 + c, families=max(2, int(40 * scale)), copies=250, min_len=300, max_len=6000,
                       divergence=0.10, tandem_arrays=int(200 * scale) + 1) for c in range(23)]
    else:
        raise KeyError(name)
    text, nk, nu = unitigs(seqs, k, threads)
    return text, k, {"config": name, "scale": scale, "k": k, "distinct_kmers": nk, "unitigs": nu,
                     "input_bp": sum(len(s) for s in seqs)}


# ---- on-disk cache of generated unitigs ("write the unitig FASTA once, reuse") ----
def cache_dir() -> Path:
    import os
    d = os.environ.get("MTG_CACHE_DIR")
    # outside the repository: gpurun ships the whole tree to the GPU box, half a gigabyte of unitigs must not ride along
    cands = [Path(d)] if d else [Path("/tmp/mtg_cache"), Path(__file__).resolve().parent.parent / ".cache"]
    for c in cands:
        try:
            c.mkdir(parents=True, exist_ok=True)
            probe = c / f".probe{os.getpid()}"
            probe.write_bytes(b"x")
            probe.unlink()
            return c
        except OSError:
            continue
    raise RuntimeError("no writable cache directory")


def cached_config_unitigs(name: str, scale: float = 1.0, wait_for_other: bool = False, timeout_s: float = 1800.0):
    """config_unitigs() behind a disk cache, so that arms / ranks / repeated runs on one box pay for the
    synthetic compacted-dBG construction once.  With `wait_for_other` the caller does not generate but polls
    for the file another process (rank 0) is writing."""
    import json
    import os
    import time
    d = cache_dir()
    stem = f"{name}_s{scale:g}"
    fa, meta = d / f"{stem}.unitigs.fa", d / f"{stem}.json"
    t0 = time.time()
    while wait_for_other and not meta.exists():
        if time.time() - t0 > timeout_s:
            raise TimeoutError(f"waited {timeout_s}s for {meta}")
        time.sleep(0.5)
    if meta.exists() and fa.exists():
        info = json.loads(meta.read_text())
        if fa.stat().st_size == info.get("text_bytes"):
            info["cache"] = "hit"
            return fa.read_bytes(), int(info["k"]), info
    text, k, info = config_unitigs(name, scale)
    info["text_bytes"] = len(text)
    tmp = d / f".{stem}.{os.getpid()}.tmp"
    tmp.write_bytes(text)
    os.replace(tmp, fa)
    tmpm = d / f".{stem}.{os.getpid()}.json.tmp"
    tmpm.write_text(json.dumps(info))
    os.replace(tmpm, meta)  # the meta file appears last: its presence means the text is complete
    info = dict(info, cache="miss")
    return text, k, info
