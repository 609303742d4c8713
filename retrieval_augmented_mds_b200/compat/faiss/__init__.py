"""Opt-in `faiss` stand-in (route 1): put `retrieval_augmented_mds_b200/compat` on
`sys.path` BEFORE `import datasets` (or call `retrieval_augmented_mds_b200.install_faiss_shim()`)
and the reference's `sotasum/mips.py` and HF `datasets` run unmodified on the B200 index:
`Dataset.add_faiss_index(column, index_name, string_factory="Flat", metric_type=...)` calls
`faiss.index_factory` here and gets a GPU-resident exact index; `save_faiss_index` /
`load_faiss_index` go through `write_index` / `read_index` (flat layout, faiss_io.py).

Only what the reference touches is provided (mips.py:1,306,316,333-345,369-371,383-386,524,
536,547,665-675; retriever_lightning.py:395-404; pretrain.py:470-479): exact flat indexes. Anything
approximate (IVF / HNSW / PQ factories, GPU cloners) raises — there is no silent fallback."""
import os as _os

from retrieval_augmented_mds_b200 import faiss_io as _io
from retrieval_augmented_mds_b200 import index as _index
from retrieval_augmented_mds_b200.faiss_io import (BufferedIOReader, BufferedIOWriter,  # noqa: F401
                                                   PyCallbackIOReader, PyCallbackIOWriter, write_index)
from retrieval_augmented_mds_b200.index import (METRIC_INNER_PRODUCT, METRIC_L2, B200FlatIndex,  # noqa: F401
                                                normalize_L2)

# faiss flat indexes are exact fp32: so is the stand-in by default (fp32 rows + bf16 shadow, exact search
# at tensor-core speed). MIPS_B200_DTYPE=bf16 stores bf16 rows instead (half the memory, recall@k = 1.0
# against fp32 flat IP on the rounded inputs).
_DTYPE = _os.environ.get("MIPS_B200_DTYPE", "fp32")


def IndexFlatIP(d, **kw):
    return _index.IndexFlatIP(d, **{"dtype": _DTYPE, **kw})


def IndexFlatL2(d, **kw):
    return _index.IndexFlatL2(d, **{"dtype": _DTYPE, **kw})


def IndexFlat(d, metric=METRIC_L2, **kw):
    return _index.IndexFlat(d, metric, **{"dtype": _DTYPE, **kw})


def index_factory(d, description="Flat", metric=METRIC_L2, **kw):
    return _index.index_factory(d, description, metric, **{"dtype": _DTYPE, **kw})


def read_index(f, *flags, **kw):
    return _io.read_index(f, **{"dtype": _DTYPE, **kw})


__version__ = "1.7.4+b200"
Index = B200FlatIndex


def _unsupported(name):
    def fn(*a, **kw):
        raise NotImplementedError(f"faiss.{name} is outside the exact flat-index path this package replaces")
    fn.__name__ = name
    return fn


StandardGpuResources = _unsupported("StandardGpuResources")
index_cpu_to_gpu = _unsupported("index_cpu_to_gpu")
index_cpu_to_all_gpus = _unsupported("index_cpu_to_all_gpus")
index_cpu_to_gpus_list = _unsupported("index_cpu_to_gpus_list")


def index_gpu_to_cpu(index):
    return index
