"""matchtigs_b200 -- B200-native greedy-matchtig hot path behind the reference's interface.

The package holds the CUDA kernels + C ABI (``csrc/``, built into ``libmatchtigs_b200.so``) and a
thin host-side mirror of the reference's operator interface (``api``).  Importing the package does
not need a GPU; creating a ``Context`` does, and fails loudly without one.
"""
from . import api  # noqa: F401
from .api import (Context, Graph, GreedytigAlgorithm, GreedytigAlgorithmConfiguration, MatchtigsError, Unitigs,  # noqa: F401
                  read_bigraph_from_bcalm2_as_edge_centric, read_bigraph_from_fasta_as_edge_centric,
                  write_duplication_bitvector, write_walks_fasta, write_walks_gfa)

__all__ = ["Context", "Graph", "GreedytigAlgorithm", "GreedytigAlgorithmConfiguration", "MatchtigsError", "Unitigs",
           "read_bigraph_from_bcalm2_as_edge_centric", "read_bigraph_from_fasta_as_edge_centric",
           "write_duplication_bitvector", "write_walks_fasta", "write_walks_gfa"]
