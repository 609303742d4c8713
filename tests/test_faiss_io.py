"""CPU tests of the faiss interchange layer (SURVEY §8f N4 / §8b route 1): byte layout of the flat
index file, the reader/writer wrappers datasets uses, and the opt-in `faiss` stand-in package.
No compute: the index object is a stand-in with the two methods the writer touches."""
import io
import struct

import numpy as np
import pytest

from retrieval_augmented_mds_b200 import faiss_io as fio


class _Rows:
    def __init__(self, rows, metric):
        self.rows, self.metric_type = rows, metric
        self.d, self.ntotal = rows.shape[1], rows.shape[0]

    def reconstruct_n(self, i0, n):
        return self.rows[i0:i0 + n]


@pytest.mark.parametrize("metric,fourcc", [(0, b"IxFI"), (1, b"IxF2")])
def test_flat_file_layout(metric, fourcc):
    """faiss/impl/index_write.cpp (1.7.4): fourcc, write_index_header, WRITEXBVECTOR(codes)."""
    rng = np.random.default_rng(0)
    rows = rng.standard_normal((37, 12), dtype=np.float32)
    b = fio.serialize_rows(rows, metric)
    assert b[:4] == fourcc
    d, ntotal, dummy1, dummy2 = struct.unpack_from("<iqqq", b, 4)
    assert (d, ntotal, dummy1, dummy2) == (12, 37, 1 << 20, 1 << 20)
    assert b[32] == 1                                    # is_trained
    assert struct.unpack_from("<i", b, 33)[0] == metric  # metric_type, unpadded after the bool
    assert struct.unpack_from("<Q", b, 37)[0] == 37 * 12  # number of 4-byte words
    assert len(b) == 45 + 37 * 12 * 4
    assert np.array_equal(np.frombuffer(b, dtype="<f4", offset=45).reshape(37, 12), rows)
    h = fio.read_flat_header(io.BytesIO(b).read)
    assert h == {"d": 12, "ntotal": 37, "metric_type": metric, "is_trained": True}


def test_write_index_streams_blocks_and_wrappers():
    rows = np.arange(5 * 3, dtype=np.float32).reshape(5, 3)
    buf = io.BytesIO()
    fio.write_index(_Rows(rows, 0), fio.BufferedIOWriter(fio.PyCallbackIOWriter(buf.write), bsz=7))  # datasets' form
    assert buf.getvalue() == fio.serialize_rows(rows, 0)
    rd = fio.BufferedIOReader(fio.PyCallbackIOReader(io.BytesIO(buf.getvalue()).read), bsz=5)
    assert fio.read_flat_header(rd.read)["ntotal"] == 5
    assert np.array_equal(np.frombuffer(rd.read(60), dtype="<f4").reshape(5, 3), rows)


def test_rejects_non_flat_and_corrupt_files():
    with pytest.raises(ValueError, match="not a flat faiss index"):
        fio.read_flat_header(io.BytesIO(b"IwFl" + bytes(60)).read)       # IVF fourcc: out of scope
    good = fio.serialize_rows(np.zeros((2, 4), np.float32), 0)
    bad = good[:37] + struct.pack("<Q", 9) + good[45:]
    with pytest.raises(ValueError, match="corrupt"):
        fio.read_flat_header(io.BytesIO(bad).read)
    with pytest.raises(ValueError, match="truncated"):
        fio.read_flat_header(io.BytesIO(good[:20]).read)


def test_faiss_stand_in_exports_what_the_reference_and_datasets_touch():
    import importlib
    import sys

    import retrieval_augmented_mds_b200 as m

    if importlib.util.find_spec("faiss") is not None and "b200" not in getattr(importlib.import_module("faiss"), "__version__", ""):
        pytest.skip("a real faiss is installed")
    path = m.install_faiss_shim()
    assert path in sys.path or path == ""
    faiss = importlib.import_module("faiss")
    for name in ("index_factory", "IndexFlat", "IndexFlatIP", "IndexFlatL2", "METRIC_INNER_PRODUCT", "METRIC_L2",
                 "normalize_L2", "write_index", "read_index", "BufferedIOWriter", "PyCallbackIOWriter",
                 "BufferedIOReader", "PyCallbackIOReader", "index_gpu_to_cpu"):
        assert hasattr(faiss, name), name
    assert (faiss.METRIC_INNER_PRODUCT, faiss.METRIC_L2) == (0, 1)
    with pytest.raises(NotImplementedError):
        faiss.index_cpu_to_all_gpus(None)
    with pytest.raises(ValueError, match="Flat"):
        faiss.index_factory(8, "IVF100,SQ8", 0)      # approximate factories are rejected before any GPU work


def test_flat_files_interchange_with_a_real_faiss(tmp_path):
    """N4 against the real thing where it exists: a file written by faiss.write_index is read by faiss_io and a
    file written by faiss_io is read by faiss.read_index (IndexFlatIP and IndexFlatL2). faiss-cpu is not
    installable in the build image (no network): the test runs wherever a genuine faiss is importable."""
    faiss = pytest.importorskip("faiss")
    if "b200" in getattr(faiss, "__version__", ""):
        pytest.skip("only the stand-in faiss is importable")
    from retrieval_augmented_mds_b200 import faiss_io
    rng = np.random.default_rng(0)
    xb = rng.standard_normal((257, 24), dtype=np.float32)
    for metric, cls in ((0, faiss.IndexFlatIP), (1, faiss.IndexFlatL2)):
        index = cls(24)
        index.add(xb)
        path = str(tmp_path / f"real_{metric}.faiss")
        faiss.write_index(index, path)
        data = open(path, "rb").read()
        assert data == faiss_io.serialize_rows(xb, metric)                 # byte-identical flat layout
        ours = str(tmp_path / f"ours_{metric}.faiss")
        open(ours, "wb").write(faiss_io.serialize_rows(xb, metric))
        back = faiss.read_index(ours)
        assert back.ntotal == 257 and back.d == 24 and back.metric_type == metric
        assert np.array_equal(faiss.vector_to_array(back.get_xb()).reshape(257, 24) if hasattr(back, "get_xb")
                              else back.reconstruct_n(0, 257), xb)
