"""retrieval_augmented_mds_b200 — B200-native exact MIPS for the non-parametric memory of
florianbaud/retrieval-augmented-mds (the `Mips.search` hot path of sotasum/mips.py feeding
sotasum/retriever_generator.py). CUDA (sm_100a) behind a C ABI; no CPU compute path."""
from .index import (METRIC_INNER_PRODUCT, METRIC_L2, B200FlatIndex, GraphedSearch, MemoryTokenStore, IndexFlat, IndexFlatIP,
                    IndexFlatL2, index_factory, merge_candidates, normalize_L2, retriever_metrics)
from .generator_ops import biased_softmax, copy_attention, copy_mixture
from .faiss_io import read_index, write_index
from .mips import Mips, MipsConfig, MipsModelOutput, RGEncoderModelOutput
from .sharded import ShardedFlatIndex, balanced_range, shard_range, weighted_ranges



def install_faiss_shim() -> str:
    """Make `import faiss` resolve to the flat-index stand-in in compat/faiss (SURVEY §8b route 1). Call it
    BEFORE `import datasets`: datasets decides at import time whether faiss exists. No-op when a real faiss
    is importable. Returns the path that was put on sys.path ('' if none)."""
    import importlib.util
    import sys
    from pathlib import Path

    if importlib.util.find_spec("faiss") is not None:
        return ""
    compat = str(Path(__file__).resolve().parent / "compat")
    if compat not in sys.path:
        sys.path.insert(0, compat)
    if "datasets.search" in sys.modules:      # imported too early: flip the flag it cached at import time
        sys.modules["datasets.search"]._has_faiss = True
    return compat


__all__ = ["copy_mixture", "copy_attention", "biased_softmax", "GraphedSearch", "MipsModelOutput", "RGEncoderModelOutput", "MemoryTokenStore", "retriever_metrics", "read_index", "write_index", "install_faiss_shim", "B200FlatIndex", "IndexFlat", "IndexFlatIP", "IndexFlatL2", "index_factory", "normalize_L2",
           "merge_candidates", "METRIC_INNER_PRODUCT", "METRIC_L2", "Mips", "MipsConfig",
           "ShardedFlatIndex", "shard_range", "balanced_range", "weighted_ranges"]
