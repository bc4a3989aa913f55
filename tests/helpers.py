"""Shared test helpers: k-mer spectra, random inputs, output parsing."""
from __future__ import annotations

import random

_COMP = bytes.maketrans(b"ACGT", b"TGCA")


def revcomp(s: bytes) -> bytes:
    return s.translate(_COMP)[::-1]


def canon(s: bytes) -> bytes:
    r = revcomp(s)
    return r if r < s else s


def kmer_list(seq: bytes, k: int) -> list[bytes]:
    return [canon(seq[i:i + k]) for i in range(len(seq) - k + 1)]


def kmer_set(seqs, k: int) -> set[bytes]:
    out = set()
    for s in seqs:
        out.update(kmer_list(s, k))
    return out


def parse_fasta_seqs(text: bytes) -> list[bytes]:
    seqs, cur = [], None
    for line in text.split(b"\n"):
        if line.startswith(b">"):
            if cur is not None:
                seqs.append(b"".join(cur))
            cur = []
        elif line and cur is not None:
            cur.append(line.strip())
    if cur is not None:
        seqs.append(b"".join(cur))
    return seqs


def parse_gfa_seqs(text: bytes) -> list[bytes]:
    lines = text.split(b"\n")
    assert lines[0].startswith(b"H\tKL:Z:")
    out = []
    for i, l in enumerate(lines[1:]):
        if not l:
            continue
        f = l.split(b"\t")
        assert f[0] == b"S" and int(f[1]) == i + 1
        out.append(f[2])
    return out


def random_fasta(rng: random.Random, n: int, k: int, max_extra: int = 40, alphabet: bytes = b"ACGT",
                 pool: int | None = None) -> bytes:
    """Arbitrary (not dBG-valid) FASTA: n records of length k..k+max_extra.  With `pool`, record ends are
    drawn from a small pool of (k-1)-mers so the graph is densely connected (parallel edges, palindromes,
    self loops all occur)."""
    ends = None
    if pool:
        ends = [bytes(rng.choice(alphabet) for _ in range(k - 1)) for _ in range(pool)]
        # make some of them palindromic when k-1 is even
        if (k - 1) % 2 == 0:
            for i in range(0, pool, 3):
                h = ends[i][: (k - 1) // 2]
                ends[i] = h + revcomp(h)
    recs = []
    for i in range(n):
        if ends:
            a, b = rng.choice(ends), rng.choice(ends)
            if rng.random() < 0.3:
                a = revcomp(a)
            if rng.random() < 0.3:
                b = revcomp(b)
            mid = bytes(rng.choice(alphabet) for _ in range(rng.randint(0, max_extra)))
            if rng.random() < 0.35:
                # overlap the two ends so that short unitigs (weight <= k-1) are common
                ov = rng.randint(1, k - 2)
                s = a + b[ov:] if a[len(a) - ov:] == b[:ov] else a + mid + b
            else:
                s = a + mid + b
            if len(s) < k:
                s = a + bytes(rng.choice(alphabet) for _ in range(2)) + b
        else:
            s = bytes(rng.choice(alphabet) for _ in range(rng.randint(k, k + max_extra)))
        recs.append(b">" + str(i).encode() + b"\n" + s + b"\n")
    return b"".join(recs)


def check_tig_invariants(unitig_text: bytes, k: int, gfa: bytes, fasta: bytes, bitvector: bytes, dbg_valid: bool):
    """Properties that hold for any correct greedy-matchtig output."""
    unitigs = parse_fasta_seqs(unitig_text)
    tigs = parse_gfa_seqs(gfa)
    assert tigs == parse_fasta_seqs(fasta)
    # 1. spectrum preservation
    assert kmer_set(tigs, k) == kmer_set(unitigs, k)
    # 2. bitvector: one line per tig, one char per k-mer
    bv = bitvector.split(b"\n")
    assert bv[-1] == b"" and len(bv) - 1 == len(tigs)
    ones = []
    for t, line in zip(tigs, bv):
        assert len(line) == len(t) - k + 1
        assert set(line) <= set(b"01")
        km = kmer_list(t, k)
        ones.extend(km[i] for i, c in enumerate(line) if c == ord("1"))
    total_unitig_kmers = sum(len(u) - k + 1 for u in unitigs)
    assert len(ones) == total_unitig_kmers
    if dbg_valid:
        # input unitigs hold every k-mer exactly once => the '1' k-mers are pairwise distinct
        assert len(set(ones)) == len(ones) == len(kmer_set(unitigs, k))
    else:
        assert sorted(ones) == sorted(x for u in unitigs for x in kmer_list(u, k))
