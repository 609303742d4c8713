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
    # row-sharded bank only. False (default) = the reference's data-parallel semantics: every rank searches
    # with its OWN batch (each DDP rank calls self.mips(queries=...), retriever_generator.py:143-153) ->
    # ShardedFlatIndex.search_dp. True = every rank passes the SAME queries (evaluation / serving): the
    # cheaper replicated step, ShardedFlatIndex.search.
    replicated_queries: bool = False


@dataclass
class MipsModelOutput:
    """The search-result fields of the reference's `MipsModelOutput` (sotasum/mips.py:33-42), same names:
    `scores` = what `Mips.search` returned (raw index scores), `query_cls` = the queries as searched,
    `memory_input_ids` / `memory_attention_mask` = the retrieved documents' tokens (mips.py:473-505; here
    gathered on the device from a MemoryTokenStore), `metrics` = retriever_metrics (mips.py:456-463).
    `examples` carries text rows in the reference (Arrow lookup, out of scope): here it holds the retrieved
    ids [B, k], from which the caller's own table gives the rows. `mips_last_hidden_state` /
    `memory_outputs` are encoder outputs (out of scope) and stay None."""
    scores: object = None
    mips_last_hidden_state: object = None
    memory_outputs: object = None
    memory_input_ids: object = None
    memory_attention_mask: object = None
    metrics: Optional[dict] = None
    examples: object = None
    query_cls: object = None


@dataclass
class RGEncoderModelOutput:
    """The retrieval fields of the reference's `RGEncoderModelOutput` (sotasum/retriever_generator.py:29-42),
    same names: `mips_scores` = cosine doc scores [B, k] (:158-172), `faiss_scores` = raw index scores (:222),
    `memory_bias` [B, k*L] (:188-192), `memory_mask` / `copy_sequence` [B, k*L] (:187,193), `query_cls`,
    `examples` (ids, see MipsModelOutput). The encoder hidden states of the reference class are out of scope."""
    memory_mask: object = None
    memory_bias: object = None
    copy_sequence: object = None
    mips_scores: object = None
    examples: object = None
    faiss_scores: object = None
    query_cls: object = None
    doc_prob: object = None            # ours: softmax_j(beta * mips_scores_j + beta_bias), the per-doc factor


class Mips:
    def __init__(self, args: Optional[MipsConfig] = None, device=None, group=None):
        self.args = args if args is not None else MipsConfig()
        if self.args.mips_string_factory.strip() != "Flat":
            raise ValueError("only the exact 'Flat' index is supported (SURVEY §2.2), got "
                             f"{self.args.mips_string_factory!r}")
        self.tmp_folder = Path(self.args.mips_tmp_folder)
        self.mips_folder = self.tmp_folder / "mips"
        self.index_file = self.mips_folder / "index.faiss"      # same names as mips.py:162-165
        self.max_norm_file = self.mips_folder / "max_norm.pkl"
        self.embeddings_folder = self.mips_folder / "embeddings"
        self.meta_file = self.mips_folder / "b200_meta.json"
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

    # ------------------------------------------------------------------ refresh (lightning_model.py:148-180)
    # The reference rebuilds its memory every `mips_rebuild_every` steps with three barriers, a disk
    # round trip of the whole bank and a rank-0 faiss add while every other rank waits
    # (_build_mips_index2). Here the bank is double buffered in HBM: the encoder's CLS
    # rows (already on the GPU, mips.py:351-356) are ingested by K0 into the BACK shard on a side
    # stream while searches keep hitting the FRONT shard; commit is two tiny collectives and a swap.
    def needs_refresh(self, global_step: int, rebuild_every: int, frozen: bool = False) -> bool:
        """The schedule of on_train_batch_start (lightning_model.py:148-163)."""
        return (not frozen) and global_step % rebuild_every == 0 and global_step not in self.rebuilt_steps

    def begin_refresh(self, n_rows_local: int, d: Optional[int] = None) -> None:
        """Open the back buffer for `n_rows_local` rows of THIS rank (encode_text2's partition,
        mips.py:226-230). The previous back buffer is reused when it is large enough (no cudaMalloc)."""
        d = d if d is not None else self.index.d
        back = getattr(self, "_back", None)
        reuse = back is not None and back.d == d and back.capacity >= n_rows_local and back.dtype == self.args.bank_dtype
        if not reuse:
            back = B200FlatIndex(d, METRIC_INNER_PRODUCT, dtype=self.args.bank_dtype, device=self.device,
                                 capacity=max(n_rows_local, 1))
        self._back = back
        if getattr(self, "_refresh_stream", None) is None:
            self._refresh_stream = torch.cuda.Stream(device=back.device)
        # the back shard was the FRONT shard until the last commit: its reset (stream ordered) must come after
        # every search already enqueued on the caller's stream
        self._refresh_stream.wait_stream(torch.cuda.current_stream(back.device))
        if reuse:
            with torch.cuda.stream(self._refresh_stream):
                back.reset()

    def refresh_add(self, embeddings: torch.Tensor) -> None:
        """One block of freshly encoded rows (CUDA float tensor [n, d]) -> back shard, on the side stream."""
        if self.args.mips_db_max_size is not None:
            room = self.args.mips_db_max_size - self._back.ntotal
            embeddings = embeddings[: max(room, 0)]
            if embeddings.shape[0] == 0:
                return
        fuse_norm = bool(self.normalize and self.metric_type == METRIC_INNER_PRODUCT)
        # the block was produced by kernels on the caller's stream (the encoder forward): the ingest on the
        # side stream must not start before they finish
        self._refresh_stream.wait_stream(torch.cuda.current_stream(self._back.device))
        with torch.cuda.stream(self._refresh_stream):
            self._back.add(embeddings, normalize=fuse_norm)
        if isinstance(embeddings, torch.Tensor) and embeddings.is_cuda:
            embeddings.record_stream(self._refresh_stream)      # keep the block alive until K0 has read it

    def commit_refresh(self, global_step: int) -> None:
        """Make the back buffer the one searches see (collective when the bank is row-sharded)."""
        back = self._back
        torch.cuda.current_stream(back.device).wait_stream(self._refresh_stream)
        front = self.index
        self.index = back
        if self.group is not None:
            old = self._sharded
            self._sharded = _sharded.ShardedFlatIndex(back, self.group)
            if old is not None:
                self._sharded.adopt(old)            # keep the communicator / exchange buffers: no leak per refresh
            off, self._sharded.counts = _sharded.exchange_offsets(back.ntotal, self.group, back.device)
            back.id_offset = off
            mn2 = _sharded.allreduce_max(back.max_norm2(), self.group, back.device)
        else:
            mn2 = back.max_norm2()
        self.max_norm = float(np.sqrt(mn2))
        if self.metric_type == METRIC_L2:
            self.phi = mn2
            back.phi = mn2
        self._back = front                          # next refresh ingests into the old front
        self.rebuilt_steps.append(global_step)      # lightning_model.py:179

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
            fn = self._sharded.search if self.args.replicated_queries else self._sharded.search_dp
            r = fn(torch.from_numpy(queries), k, ignore_ids=None if ignore_indexes is None
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
        index = self.index
        if self.normalize and self.metric_type != METRIC_INNER_PRODUCT:
            # normalize with the L2 metric: the stored rows are NOT unit-norm, and the reference's per-call
            # renormalisation of both sides (mips.py:554-556) ranks by cosine. Do what it does, on the GPU:
            # a unit-norm copy of the bank (K0 with the fused normalisation), kept until the bank changes.
            key = (id(index), index.ntotal)
            if getattr(self, "_unit_bank_key", None) != key:
                unit = B200FlatIndex(index.d, METRIC_INNER_PRODUCT, dtype=index.dtype, device=index.device,
                                     capacity=max(index.ntotal, 1))
                for i in range(0, index.ntotal, 262144):
                    unit.add(index.reconstruct_n(i, min(262144, index.ntotal - i), as_torch=True), normalize=True)
                self._unit_bank, self._unit_bank_key = unit, key
            index = self._unit_bank
        r = index.search_ex(torch.from_numpy(x), k, normalize_queries=bool(self.normalize),
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
        if self._sharded is not None and self._sharded.world > 1:
            fn = self._sharded.search if self.args.replicated_queries else self._sharded.search_dp
        else:
            fn = self.index.search_ex
        return fn(queries, k, ignore_ids=ignore_indexes, want=want, L=memory_seq_len,
                  normalize_queries=norm_q, out_mode=out_mode, beta=beta, beta_bias=beta_bias)

    def forward(self, queries, k: int = 10, ignore_indexes=None, aid=None, aid_counts=None, row_aid=None,
                token_store=None) -> MipsModelOutput:
        """The search part of `Mips.forward` (mips.py:402-519) with the reference's output container:
        prepare the queries (:421), search (:422-426), optional retrieval metrics (:456-463: `aid` [B],
        `aid_counts` [B] and `row_aid` [N] = the `aid` column of the memory) and the retrieved documents'
        tokens (:473-505) from a MemoryTokenStore. `queries` float [B, d], numpy like the reference or a CUDA
        tensor (then nothing leaves the device). Text assembly and the encoders are out of scope."""
        if self.index is None:
            raise RuntimeError("build_index() or load() first")
        dev = self.index.device
        if isinstance(queries, torch.Tensor):
            q_dev = queries.detach().to(device=dev, dtype=torch.float32)
            query_cls = q_dev
        else:
            query_cls = np.asarray(queries, dtype=np.float32)
            q_dev = torch.from_numpy(np.ascontiguousarray(query_cls)).to(dev)
        r = self.search_device(q_dev, k, ignore_indexes=ignore_indexes)
        metrics = None
        if aid is not None and aid_counts is not None and row_aid is not None:
            m = _index.retriever_metrics(r["ids"], torch.as_tensor(row_aid).to(dev), torch.as_tensor(aid).to(dev),
                                         torch.as_tensor(aid_counts, dtype=torch.float32).to(dev))
            metrics = {"recall": m["recall"], "reciprocal_rank": m["reciprocal_rank"],
                       "average_precision": m["average_precision"]}
        mem_ids = mem_mask = None
        if token_store is not None:
            g = token_store.gather(r["ids"])
            B, L = r["ids"].shape[0], token_store.seq_len
            mem_ids = g["memory_input_ids"].view(B, k, L)                 # mips.py:499-505
            mem_mask = g["memory_attention_mask"].view(B, k, L)
        return MipsModelOutput(scores=r["scores"], memory_input_ids=mem_ids, memory_attention_mask=mem_mask,
                               metrics=metrics, examples=r["ids"], query_cls=query_cls)

    def retrieve_for_generator(self, query: torch.Tensor, k: int, ignore_indexes=None, token_store=None,
                               memory_seq_len: Optional[int] = None, beta: float = 1.0,
                               beta_bias: float = 0.0) -> RGEncoderModelOutput:
        """The retrieval block of `SotasumEncoder.forward` (retriever_generator.py:138-193) for a frozen memory
        encoder, with the reference's field names: `query` is the CLS slice [B, 1, d] or [B, d] ON THE GPU (no
        `.cpu()` round trip, :143); returns `faiss_scores` (raw search scores, :222), `mips_scores` (cosine,
        :158-172), `memory_bias` [B, k*L] (:188-192) and, with a MemoryTokenStore, `memory_mask` /
        `copy_sequence` [B, k*L] (:187,193)."""
        q = query[:, 0, :] if query.dim() == 3 else query
        L = token_store.seq_len if token_store is not None else memory_seq_len
        r = self.search_device(q, k, ignore_indexes=ignore_indexes, memory_seq_len=L, beta=beta, beta_bias=beta_bias)
        mask = copy_seq = None
        if token_store is not None:
            g = token_store.gather(r["ids"])
            B = r["ids"].shape[0]
            mask = g["memory_attention_mask"].view(B, -1)
            copy_seq = g["memory_input_ids"].view(B, -1)
        return RGEncoderModelOutput(memory_mask=mask, memory_bias=r.get("memory_bias"), copy_sequence=copy_seq,
                                    mips_scores=r["cosine"], examples=r["ids"], faiss_scores=r["scores"],
                                    query_cls=q, doc_prob=r["doc_prob"])

    # ------------------------------------------------------------------ save / load (mips.py:531-549)
    def _rank_suffix(self) -> str:
        w = self._sharded.world if self._sharded is not None else 1
        return f".rank{self._sharded.rank}of{w}" if w > 1 else ""

    def save(self) -> None:
        """Checkpoint in the reference's layout (mips.py:531-543): `mips/index.faiss` is the flat faiss
        file the reference's `load()` reads (faiss_io.py) — IndexFlatIP over the (normalised) rows, or
        IndexFlatL2 over the AUGMENTED rows [x, sqrt(phi - |x|^2)] (d+1 columns) exactly as
        `build_index` left them in the reference (mips.py:316-331) — `mips/max_norm.pkl` is a pickled
        float (cloudpickle.load reads plain pickles) and `mips/embeddings/` is the Arrow dataset with
        the `embeddings` column (written when `datasets` is importable; text columns are outside the
        hot path). A row-sharded bank writes one file set per rank (suffix `.rank{r}of{G}`)."""
        import json

        from . import faiss_io

        sfx = self._rank_suffix()
        if sfx == "" or self._sharded.rank == 0:
            shutil.rmtree(self.mips_folder, ignore_errors=True)
        self.mips_folder.mkdir(parents=True, exist_ok=True)
        if self._sharded is not None and self._sharded.world > 1:
            torch.distributed.barrier(self.group)
        n, d = self.index.ntotal, self.index.d
        l2 = self.metric_type == METRIC_L2

        def blocks():
            for i in range(0, n, 262144):
                rows = self.index.reconstruct_n(i, min(262144, n - i))
                if l2:   # augment_xb (mips.py:59-65) from the stored |x|^2: the column is never kept in HBM
                    extra = np.sqrt(np.maximum(np.float32(self.phi) - (rows.astype(np.float32) ** 2).sum(1), 0.0))
                    rows = np.hstack((rows, extra.reshape(-1, 1).astype(np.float32)))
                yield rows

        with open(str(self.index_file) + sfx, "wb") as f:
            faiss_io.write_flat(f.write, blocks(), d + 1 if l2 else d, n, self.metric_type)
        with open(str(self.max_norm_file) + sfx, "wb") as f:
            pickle.dump(float(self.max_norm), f)
        with open(str(self.meta_file) + sfx, "w") as f:
            json.dump({"bank_dtype": self.index.dtype, "id_offset": int(self.index.id_offset), "d": d,
                       "metric_type": int(self.metric_type), "normalize": bool(self.normalize),
                       "phi": None if self.phi is None else float(self.phi)}, f)
        try:
            import datasets  # optional: the reference's load() also wants the Arrow `embeddings` folder
        except ImportError:
            return
        col = self.embeddings_column
        if n:
            # streamed block by block through an Arrow file (memory-mapped back): the bank is never held twice
            from datasets.arrow_writer import ArrowWriter
            feats = datasets.Features({col: datasets.Sequence(datasets.Value("float32"))})
            tmp_arrow = str(self.mips_folder / f".embeddings_stream{sfx}.arrow")
            writer = ArrowWriter(features=feats, path=tmp_arrow)
            for blk in blocks():
                writer.write_batch({col: blk})
            writer.finalize()
            ds = datasets.Dataset.from_file(tmp_arrow)
        else:
            tmp_arrow = None
            ds = datasets.Dataset.from_dict({col: np.zeros((0, d), np.float32)})
        ds.save_to_disk(str(self.embeddings_folder) + sfx)
        if tmp_arrow is not None:
            del ds
            Path(tmp_arrow).unlink(missing_ok=True)

    def load(self) -> None:
        """Load `mips/index.faiss` (+ `max_norm.pkl`) written by `save()` OR by the reference
        (`Dataset.save_faiss_index`, mips.py:536). An L2 file holds augmented rows: the last column is
        dropped (it is a function of |x|^2 and phi) and phi is recovered as |x~|^2 of the first row."""
        import json

        from . import faiss_io

        sfx = self._rank_suffix()
        meta = {}
        if Path(str(self.meta_file) + sfx).exists():
            meta = json.loads(Path(str(self.meta_file) + sfx).read_text())
        with open(str(self.index_file) + sfx, "rb") as f:
            h = faiss_io.read_flat_header(f.read)
            l2 = h["metric_type"] == METRIC_L2
            d_file = h["d"]
            d = d_file - 1 if l2 else d_file
            self.index = B200FlatIndex(d, METRIC_INNER_PRODUCT, dtype=meta.get("bank_dtype", self.args.bank_dtype),
                                       device=self.device, capacity=max(h["ntotal"], 1),
                                       id_offset=int(meta.get("id_offset", 0)))
            phi = None
            left = h["ntotal"]
            while left > 0:
                nb = min(262144, left)
                rows = np.frombuffer(f.read(nb * d_file * 4), dtype="<f4").reshape(nb, d_file)
                if l2:
                    if phi is None:
                        phi = float((rows[0].astype(np.float64) ** 2).sum())
                    rows = np.ascontiguousarray(rows[:, :d])
                self.index.add(rows)                    # rows were normalised before they were stored
                left -= nb
        self.metric_type = h["metric_type"]
        if l2:
            self.phi = float(meta["phi"]) if meta.get("phi") is not None else phi
            self.index.phi = self.phi
        if self.group is not None:
            self._sharded = _sharded.ShardedFlatIndex(self.index, self.group)
            off, self._sharded.counts = _sharded.exchange_offsets(self.index.ntotal, self.group, self.index.device)
            self.index.id_offset = off
        with open(str(self.max_norm_file) + sfx, "rb") as f:
            self.max_norm = pickle.load(f)
