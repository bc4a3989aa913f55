"""ctypes front-end of the CPU oracle (``oracle/mtg_oracle.cpp``).

TEST INFRASTRUCTURE ONLY: import this from ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` -- never from ``matchtigs_b200``.
Parity of the oracle against the real Rust binary is UNPINNED (see the .cpp header).
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

_SO = Path(__file__).resolve().parent / "libmtg_oracle.so"
_lib = None

_DTYPES = {
    "mirror": np.uint32, "edge_from": np.uint32, "edge_to": np.uint32, "edge_weight": np.uint64,
    "edge_dummy_id": np.uint32, "edge_forward": np.uint8, "edge_unitig": np.uint32,
    "out_nodes": np.uint32, "in_node_map0": np.uint8, "mult0": np.int64, "triples": np.uint32,
    "gfa": np.uint8, "fasta": np.uint8, "bitvector": np.uint8,
    "c_edge_out": np.int64, "c_insert_out": np.uint64, "c_limits": np.uint64,
}


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not _SO.exists():
            import sys
            sys.path.insert(0, str(_SO.parent.parent))
            from matchtigs_b200 import _build
            _build.build_oracle()
        l = C.CDLL(str(_SO))
        l.mto_create.restype = C.c_void_p
        l.mto_destroy.argtypes = [C.c_void_p]
        l.mto_error.restype = C.c_char_p
        l.mto_error.argtypes = [C.c_void_p]
        l.mto_load_text.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_int, C.c_int]
        l.mto_load_links.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_int]
        l.mto_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        l.mto_run_greedy.argtypes = [C.c_void_p, C.c_int]
        l.mto_get.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        l.mto_num.restype = C.c_size_t
        l.mto_num.argtypes = [C.c_void_p, C.c_char_p]
        l.mto_time.restype = C.c_double
        l.mto_time.argtypes = [C.c_void_p, C.c_char_p]
        l.mto_walk.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        l.mto_cycle.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        l.mto_candidates.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = l
    return _lib


class OracleError(RuntimeError):
    pass


class Oracle:
    """One greedy-matchtig computation on the CPU, following the reference's ``--threads 1`` semantics."""

    OUT_GFA, OUT_FASTA, OUT_BITVECTOR, OUT_CAPI = 1, 2, 4, 8

    def __init__(self, euler_fast: bool = False, outputs: int = 15, **options: int):
        """`outputs`: which outputs ``run`` produces (bit mask of OUT_*).  `options`: named switches of
        ``mto_set_option`` (the parity assumptions P1..P7, see the .cpp header)."""
        self._l = lib()
        self._h = C.c_void_p(self._l.mto_create())
        self.set_option("euler_fast", int(euler_fast))
        self.set_option("outputs", int(outputs))
        for name, value in options.items():
            self.set_option(name, int(value))

    def set_option(self, name: str, value: int) -> None:
        if self._l.mto_set_option(self._h, name.encode(), int(value)) != 0:
            raise KeyError(f"unknown oracle option {name!r}")

    def __del__(self):
        if getattr(self, "_h", None):
            self._l.mto_destroy(self._h)
            self._h = None

    def _check(self, rc: int) -> None:
        if rc != 0:
            raise OracleError(self._l.mto_error(self._h).decode())

    def load_fasta(self, text: bytes, k: int) -> "Oracle":
        """``--fa-in`` semantics (src/bin.rs:891-901)."""
        self._check(self._l.mto_load_text(self._h, text, len(text), k, 0))
        return self

    def load_bcalm(self, text: bytes, k: int) -> "Oracle":
        """``--bcalm-in`` semantics (src/bin.rs:902-912), links -> union-find numbering."""
        self._check(self._l.mto_load_text(self._h, text, len(text), k, 1))
        return self

    def load_links(self, weights, links, k: int) -> "Oracle":
        """C-API semantics (src/clib.rs:97-259): links = iterable of (a, strand_a, b, strand_b)."""
        w = np.ascontiguousarray(weights, dtype=np.uint64)
        links = list(links)
        a = np.array([l[0] for l in links], dtype=np.uint64)
        sa = np.array([l[1] for l in links], dtype=np.uint8)
        b = np.array([l[2] for l in links], dtype=np.uint64)
        sb = np.array([l[3] for l in links], dtype=np.uint8)
        self._check(self._l.mto_load_links(self._h, len(w), w.ctypes.data, len(links), a.ctypes.data, sa.ctypes.data,
                                           b.ctypes.data, sb.ctypes.data, k))
        return self

    def run(self, threads: int = 1) -> "Oracle":
        self._check(self._l.mto_run_greedy(self._h, threads))
        return self

    def array(self, name: str) -> np.ndarray:
        p, n = C.c_void_p(), C.c_size_t()
        if self._l.mto_get(self._h, name.encode(), C.byref(p), C.byref(n)) != 0:
            raise KeyError(name)
        dt = np.dtype(_DTYPES[name])
        if n.value == 0:
            return np.zeros(0, dtype=dt)
        buf = (C.c_char * (n.value * dt.itemsize)).from_address(p.value)
        return np.frombuffer(buf, dtype=dt).copy()

    def text(self, name: str) -> bytes:
        return self.array(name).tobytes()

    def num(self, name: str) -> int:
        v = self._l.mto_num(self._h, name.encode())
        if v == 2**64 - 1:
            raise KeyError(name)
        return v

    def time(self, name: str) -> float:
        return self._l.mto_time(self._h, name.encode())

    def _seq(self, fn, i: int) -> np.ndarray:
        p, n = C.c_void_p(), C.c_size_t()
        if fn(self._h, i, C.byref(p), C.byref(n)) != 0:
            raise IndexError(i)
        if n.value == 0:
            return np.zeros(0, dtype=np.uint32)
        buf = (C.c_char * (n.value * 4)).from_address(p.value)
        return np.frombuffer(buf, dtype=np.uint32).copy()

    def walks(self) -> list[np.ndarray]:
        return [self._seq(self._l.mto_walk, i) for i in range(self.num("walks"))]

    def cycles(self) -> list[np.ndarray]:
        return [self._seq(self._l.mto_cycle, i) for i in range(self.num("cycles"))]

    def candidates(self, cap: int, lo: int = 0, hi: int | None = None):
        """L(src) for sources [lo, hi): (nodes[S,cap], dists[S,cap], full_len[S])."""
        hi = self.num("sources") if hi is None else hi
        s = hi - lo
        nodes = np.zeros((s, cap), dtype=np.uint32)
        dists = np.zeros((s, cap), dtype=np.uint32)
        lens = np.zeros(s, dtype=np.uint32)
        self._check(self._l.mto_candidates(self._h, lo, hi, cap, nodes.ctypes.data, dists.ctypes.data, lens.ctypes.data))
        return nodes, dists, lens
