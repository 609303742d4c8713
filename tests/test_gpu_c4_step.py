"""BASELINE config 4 — the retriever-generator step around the hot path: query CLS embeddings from a
BART-large-shaped encoder (random init: no checkpoints offline), MIPS k=5 with the ignore id of the
training example, cosine doc logits / per-doc softmax / memory_bias fused into the merge kernel,
batch 16 — all on the device, no host round trip (reference retriever_generator.py:138-193, where
the query goes `.detach().cpu().float().numpy()` at :143 and comes back as Python lists)."""
import numpy as np
import pytest
import torch

import retrieval_augmented_mds_b200 as pkg
from oracle import mips_oracle as o

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("bank_dtype", ["bf16", "fp32"])
def test_retriever_generator_step_d1024(cuda_device, bank_dtype):
    transformers = pytest.importorskip("transformers")
    torch.manual_seed(0)
    cfg = transformers.BartConfig(d_model=1024, encoder_layers=2, decoder_layers=1, encoder_attention_heads=16,
                                  decoder_attention_heads=16, encoder_ffn_dim=4096, decoder_ffn_dim=4096,
                                  vocab_size=5000, max_position_embeddings=128)   # BART-large widths, 2 layers: shape is what matters
    enc = transformers.BartModel(cfg).get_encoder().to(cuda_device).eval()
    B, T, k, L, N, d = 16, 64, 5, 512, 60_000, 1024
    tok = torch.randint(3, 5000, (B, T), device=cuda_device)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        query = enc(input_ids=tok).last_hidden_state[:, 0, :]           # CLS slice (retriever_generator.py:138-141)
    assert query.shape == (B, d) and query.is_cuda
    gen = torch.Generator(device=cuda_device).manual_seed(7)
    bank = torch.randn((N, d), generator=gen, device=cuda_device)
    bank[1000:1000 + B] = query.float() * 2.0 + 0.05 * torch.randn((B, d), generator=gen, device=cuda_device)  # each example's own document
    mips = pkg.Mips(pkg.MipsConfig(mips_metric_type=0, mips_normalize=True, bank_dtype=bank_dtype), device=cuda_device)
    mips.build_index(bank)
    ignore = torch.arange(1000, 1000 + B, dtype=torch.int64, device=cuda_device)   # the example's own row is excluded
    r = mips.search_device(query.float(), k, ignore_indexes=ignore, memory_seq_len=L, beta=2.0, beta_bias=0.5)
    assert mips.index.last_algo == ("tc2" if bank_dtype == "bf16" else "tcx")       # d = 1024: pair kernel
    ids, cos = r["ids"].cpu().numpy(), r["cosine"].cpu().numpy()
    assert not (ids == ignore.cpu().numpy()[:, None]).any()
    # reference arithmetic on the same stored values
    stored = mips.index.reconstruct_n(0, N)
    qn = query.float().cpu().numpy()
    qn = o.normalize_L2(qn if bank_dtype == "fp32" else qn)
    if bank_dtype == "bf16":
        qn = o.bf16_round(qn)
    D_ref, I_ref = o.exact_topk_f64(stored, qn, k, ignore=ignore.cpu().numpy())
    o.check_topk(stored, qn, r["scores"].cpu().numpy(), ids, 0, rtol=1e-4, D_ref=D_ref, I_ref=I_ref,
                 ignore=ignore.cpu().numpy(), what=f"C4 {bank_dtype}")
    want_cos = o.doc_scores(qn, stored[ids])                                         # retriever_generator.py:158-172
    np.testing.assert_allclose(cos, want_cos, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(r["doc_prob"].cpu().numpy(), o.doc_prob(want_cos, 2.0, 0.5), rtol=1e-4, atol=1e-6)
    mb = r["memory_bias"].cpu().numpy()
    assert mb.shape == (B, k * L)
    np.testing.assert_allclose(mb, o.memory_bias(want_cos, L), rtol=1e-4, atol=1e-5)  # :188-192
