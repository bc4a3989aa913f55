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
        return C.string_at(p, n.value)
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
