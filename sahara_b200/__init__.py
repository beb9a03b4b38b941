"""sahara_b200 — B200-native search path of seqan/sahara (approximate matching over a bidirectional
FM-index with search schemes).  The product is the CUDA C ABI in libsahara_b200.so (include/sahara_b200.h)
plus the C++ host layer in sahara_b200/host/; this package is the thin Python mirror used by the tests
and bench.py.  There is no CPU fallback: importing fails when the CUDA library has not been built."""
from ._native import SaharaError  # noqa: F401
from .api import (Context, SearchScheme, decode_batch, default_policy, device_count, load_fasta_ranks, load_fasta_reads, pack_reads2, pack_reads4,  # noqa: F401
                  revcomp_ranks, set_expand_rule)
