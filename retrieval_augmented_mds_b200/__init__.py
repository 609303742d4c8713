"""retrieval_augmented_mds_b200 — B200-native exact MIPS for the non-parametric memory of
florianbaud/retrieval-augmented-mds (the `Mips.search` hot path of sotasum/mips.py feeding
sotasum/retriever_generator.py). CUDA (sm_100a) behind a C ABI; no CPU compute path."""
from .index import (METRIC_INNER_PRODUCT, METRIC_L2, B200FlatIndex, IndexFlat, IndexFlatIP, IndexFlatL2,
                    index_factory, merge_candidates, normalize_L2)
from .mips import Mips, MipsConfig
from .sharded import ShardedFlatIndex, balanced_range, shard_range

__all__ = ["B200FlatIndex", "IndexFlat", "IndexFlatIP", "IndexFlatL2", "index_factory", "normalize_L2",
           "merge_candidates", "METRIC_INNER_PRODUCT", "METRIC_L2", "Mips", "MipsConfig",
           "ShardedFlatIndex", "shard_range", "balanced_range"]
