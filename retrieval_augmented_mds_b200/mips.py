"""`Mips`-shaped host façade for the retrieval hot path — same method names, argument meaning and
return types as the reference class for the part we replace (sotasum/mips.py):

    build_index (:290-345)   search (:382-400)   np_search (:527-529)   _prepare_query (:368-375)
    l2_normalization (:521-525)   save / load (:531-549)   encode_text2's shard rule (:226-230)

Everything numerical runs on the GPU through libmips_b200.so. What is NOT here (out of scope,
SURVEY §8): text encoders, tokenisers, Arrow text gather, forcing modes, Lightning plumbing.
"""
from __future__ import annotations

import pickle
import shutil
from dataclasses import dataclass
from pathlib import Path
from typing import Optional

import numpy as np
import torch

from . import index as _index
from . import sharded as _sharded
from ._lib import OUT_AUGL2, OUT_IP
from .index import METRIC_INNER_PRODUCT, METRIC_L2, B200FlatIndex


@dataclass
class MipsConfig:
    """The MIPS knobs of the reference's ModelConfig (sotasum/model_config.py:44-72) that this
    path honours. Field names are the reference's."""
    mips_topk: int = 2
    mips_string_factory: str = "Flat"
    mips_nprobe: Optional[int] = None          # accepted, ignored (exact search)
    mips_train_size: int = -1                  # accepted, ignored (nothing to train)
    mips_metric_type: int = 0                  # 0 -> INNER_PRODUCT ; 1 -> L2 (augmented)
    mips_normalize: bool = True
    mips_db_max_size: Optional[int] = None
    mips_tmp_folder: str = "./tmp"
    # ours
    bank_dtype: str = "bf16"                   # "bf16" (tcgen05 path) or "fp32" (exact fp32 path)


class Mips:
    def __init__(self, args: Optional[MipsConfig] = None, device=None, group=None):
        self.args = args if args is not None else MipsConfig()
        if self.args.mips_string_factory.strip() != "Flat":
            raise ValueError("only the exact 'Flat' index is supported (SURVEY §2.2), got "
                             f"{self.args.mips_string_factory!r}")
        self.tmp_folder = Path(self.args.mips_tmp_folder)
        self.mips_folder = self.tmp_folder / "mips"
        self.index_file = self.mips_folder / "index.b200"
        self.max_norm_file = self.mips_folder / "max_norm.pkl"
        self.string_factory = self.args.mips_string_factory
        self.train_size = self.args.mips_train_size
        self.metric_type = self.args.mips_metric_type
        self.normalize = self.args.mips_normalize
        self.max_norm = None
        self.phi = None
        self.rebuilt_steps = [0]
        self.index_name = "mips_embeddings"
        self.embeddings_column = "embeddings"
        self.device = device
        self.group = group                      # torch.distributed group => row-sharded bank
        self.index: Optional[B200FlatIndex] = None
        self._sharded: Optional[_sharded.ShardedFlatIndex] = None

    # ------------------------------------------------------------------ build (mips.py:290-345)
    def build_index(self, embeddings, capacity: int = 0) -> None:
        """embeddings: float32 [N, d] (numpy / torch, host or device) — with a process group, THIS
        RANK'S rows (shard_range(N, rank, world)). Same order of operations as the reference:
        max_norm over raw rows, then normalise (IP ∧ normalize) or phi (L2); rows are never
        physically augmented — `|q|^2 + phi - 2<q,x>` is applied in the merge kernel."""
        if self.args.mips_db_max_size is not None:
            embeddings = embeddings[: self.args.mips_db_max_size]
        d = embeddings.shape[1]
        fuse_norm = bool(self.normalize and self.metric_type == METRIC_INNER_PRODUCT)
        # Both reference metrics rank by inner product (L2 runs on augmented vectors, which is
        # MIPS by construction, mips.py:52-70) so the bank is always searched as IP.
        self.index = B200FlatIndex(d, METRIC_INNER_PRODUCT, dtype=self.args.bank_dtype, device=self.device,
                                   capacity=max(capacity, embeddings.shape[0]))
        if self.group is not None:
            self._sharded = _sharded.ShardedFlatIndex(self.index, self.group)
            self._sharded.add_local(embeddings, normalize=fuse_norm)
            mn2 = _sharded.allreduce_max(self.index.max_norm2(), self.group, self.index.device)
        else:
            self.index.add(embeddings, normalize=fuse_norm)
            mn2 = self.index.max_norm2()
        self.max_norm = float(np.sqrt(mn2))            # mips.py:298-304
        if self.metric_type == METRIC_L2:
            self.phi = mn2                              # mips.py:316-324 (rows are not normalised here)
            self.index.phi = mn2
        if isinstance(self.args.mips_nprobe, int):
            self.index.nprobe = self.args.mips_nprobe   # mips.py:342-345

    # ------------------------------------------------------------------ query prep (mips.py:368-375)
    def l2_normalization(self, x: np.ndarray) -> np.ndarray:
        if not x.flags.c_contiguous:
            x = np.asarray(x, order="C")
        _index.normalize_L2(x)
        return x

    def _prepare_query(self, query: np.ndarray) -> np.ndarray:
        query = np.array(query, dtype=np.float32, order="C")
        if self.normalize and self.metric_type == METRIC_INNER_PRODUCT:
            query = self.l2_normalization(query)
        if self.metric_type == METRIC_L2:
            query = np.hstack((query, np.zeros((len(query), 1), dtype=np.float32)))  # augment_xq
        if not query.flags.c_contiguous:
            query = np.asarray(query, order="C")
        return query.astype(np.float32)

    # ------------------------------------------------------------------ search (mips.py:382-400)
    def _strip(self, queries):
        d = self.index.d
        if self.metric_type == METRIC_L2 and queries.shape[1] == d + 1:
            return queries[:, :d]                       # the augment_xq zero column
        return queries

    def search(self, queries: np.ndarray, ignore_indexes: list = None, k: int = 10):
        """Same contract as the reference: `queries` are already prepared (_prepare_query);
        returns (scores, indices) as numpy arrays, or as Python lists of lists when
        ignore_indexes is given. L2 scores are squared distances on the augmented vectors."""
        if self.index is None:
            raise RuntimeError("build_index() or load() first")
        queries = np.ascontiguousarray(self._strip(np.asarray(queries)), dtype=np.float32)
        out_mode = OUT_AUGL2 if self.metric_type == METRIC_L2 else OUT_IP
        if self._sharded is not None and self._sharded.world > 1:
            r = self._sharded.search(torch.from_numpy(queries), k, ignore_ids=None if ignore_indexes is None
                                     else torch.as_tensor(ignore_indexes, dtype=torch.int64), out_mode=out_mode)
            scores, indices = r["scores"].cpu().numpy(), r["ids"].cpu().numpy()
        else:
            ign = None if ignore_indexes is None else np.asarray(ignore_indexes, dtype=np.int64)
            scores, indices = self.index.search_host(queries, k, ignore_ids=ign, out_mode=out_mode)
        if ignore_indexes is not None:
            keep = indices >= 0
            scores = [s[m].tolist() for s, m in zip(scores, keep)]
            indices = [i[m].tolist() for i, m in zip(indices, keep)]
        return scores, indices

    def np_search(self, x, k: int = 2) -> tuple:
        """inner_product(x, y, k, normalize=self.normalize) of mips.py:552-560 on the stored rows:
        the reference re-normalises BOTH sides on every call when normalize is set."""
        if self.index is None:
            raise RuntimeError("build_index() or load() first")
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert len(x.shape) == 2
        if self.normalize and self.metric_type != METRIC_INNER_PRODUCT:
            # the stored rows are not unit-norm in this configuration, so the reference's
            # per-call renormalisation ranks by cosine, which this index does not store
            raise NotImplementedError("np_search with normalize=True needs the IP metric (unit-norm bank)")
        r = self.index.search_ex(torch.from_numpy(x), k, normalize_queries=bool(self.normalize),
                                 want=("scores", "ids", "cosine") if self.normalize else ("scores", "ids"))
        scores = r["cosine"] if self.normalize else r["scores"]
        return scores.cpu().numpy(), r["ids"].cpu().numpy()

    def search_device(self, queries: torch.Tensor, k: int, ignore_indexes=None, memory_seq_len: Optional[int] = None,
                      beta: float = 1.0, beta_bias: float = 0.0) -> dict:
        """Device-resident variant for the retriever-generator step (retriever_generator.py:138-193):
        query CLS stays on the GPU; returns scores, ids, cosine doc scores (`mips_scores`), the
        per-doc softmax and `memory_bias` as CUDA tensors."""
        if self.index is None:
            raise RuntimeError("build_index() or load() first")
        want = ["scores", "ids", "cosine", "doc_prob"] + (["memory_bias"] if memory_seq_len else [])
        out_mode = OUT_AUGL2 if self.metric_type == METRIC_L2 else OUT_IP
        norm_q = bool(self.normalize and self.metric_type == METRIC_INNER_PRODUCT)
        target = self._sharded if (self._sharded is not None and self._sharded.world > 1) else self.index
        fn = target.search if target is self._sharded else target.search_ex
        return fn(queries, k, ignore_ids=ignore_indexes, want=want, L=memory_seq_len,
                  normalize_queries=norm_q, out_mode=out_mode, beta=beta, beta_bias=beta_bias)

    # ------------------------------------------------------------------ save / load (mips.py:531-549)
    def save(self) -> None:
        """Rank-local checkpoint: stored rows (as float32) + metadata. (Interchange with the
        reference's index.faiss / Arrow layout is the 'next' row N4 of SURVEY §8f.)"""
        shutil.rmtree(self.mips_folder, ignore_errors=True)
        self.mips_folder.mkdir(parents=True, exist_ok=True)
        rows = self.index.reconstruct_n(0, self.index.ntotal)
        np.savez(self.index_file, rows=rows, metric_type=self.metric_type, normalize=self.normalize,
                 phi=np.float32(self.phi if self.phi is not None else 0.0), id_offset=self.index.id_offset,
                 dtype=self.index.dtype)
        with open(self.max_norm_file, "wb") as f:
            pickle.dump(self.max_norm, f)

    def load(self) -> None:
        z = np.load(str(self.index_file) + ".npz" if not str(self.index_file).endswith(".npz") else self.index_file,
                    allow_pickle=False)
        rows = z["rows"]
        self.index = B200FlatIndex(rows.shape[1], METRIC_INNER_PRODUCT, dtype=str(z["dtype"]), device=self.device,
                                   capacity=rows.shape[0], id_offset=int(z["id_offset"]))
        self.index.add(rows)                            # rows were normalised before they were stored
        if self.metric_type == METRIC_L2:
            self.phi = float(z["phi"])
            self.index.phi = self.phi
        with open(self.max_norm_file, "rb") as f:
            self.max_norm = pickle.load(f)
