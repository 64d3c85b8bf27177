"""Host-side logic that needs no GPU: mask rectangularisation, sharding, the gloo multi-process gather, and the
"no CPU fallback" contract of the product path."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, needs_reference, reference_available
from counterfactualworldmodels_b200 import dist as cdist
from counterfactualworldmodels_b200 import prediction, synthetic, vmae


def test_rectangularize_min_mode():
    torch.manual_seed(0)
    masks = torch.zeros(3, 20, dtype=torch.bool)
    masks[0, :10] = True
    masks[1, :12] = True
    masks[2, :15] = True
    before = masks.clone()
    out = prediction.RectangularizeMasks('min')(masks)
    assert out.sum(-1).tolist() == [10, 10, 10]
    assert (out & ~before).sum() == 0  # 'min' only clears bits
    assert out.data_ptr() == masks.data_ptr()  # in place, like masking.py:119-128


def test_rectangularize_noop_when_equal():
    masks = synthetic.make_mask(4, (2, 8, 8), num_clumps=2, seed=0)
    state = torch.get_rng_state()
    out = prediction.RectangularizeMasks('min')(masks.clone())
    assert torch.equal(out, masks)
    assert torch.equal(torch.get_rng_state(), state)  # no RNG draw for already-rectangular batches


def test_synthetic_mask_is_temporally_factored():
    m = synthetic.make_mask(5, (2, 28, 28), num_clumps=2, seed=3).view(5, 2, 28, 28)
    assert not m[:, 0].any()
    assert (~m[:, 1]).flatten(1).sum(-1).tolist() == [8] * 5


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 64, 1024, 1025):
        for w in (1, 2, 3, 8):
            spans = [cdist.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_forward_refuses_cpu_tensors():
    m = vmae.PretrainVisionTransformer(**synthetic.model_kwargs("tiny_4x4"))
    x = torch.zeros(1, 3, 2, 32, 32)
    mask = synthetic.make_mask(1, m.mask_size, 1)
    with pytest.raises(RuntimeError, match="no CPU"):
        m(x, mask)
    with pytest.raises(NotImplementedError):
        m.encoder.blocks[0](torch.zeros(1, 4, 128))


def test_unsupported_constructor_options_raise():
    kw = synthetic.model_kwargs("tiny_4x4")
    with pytest.raises(NotImplementedError):
        vmae.PretrainVisionTransformer(**dict(kw, decoder_depth=0))
    with pytest.raises(NotImplementedError):
        vmae.PretrainVisionTransformer(**dict(kw, embed_per_frame=True))
    with pytest.raises(NotImplementedError):
        vmae.PretrainVisionTransformer(**dict(kw, drop_path_rate=0.1))


def test_supported_constructor_options_add_the_reference_parameters():
    """layer scale and learnable positional embeddings create the same state_dict entries as the reference
    (VideoMAE/utils.py:140-144, vmae.py:68-70)."""
    kw = synthetic.model_kwargs("tiny_4x4")
    m = vmae.PretrainVisionTransformer(**dict(kw, init_values=0.1, use_learnable_pos_emb=True))
    keys = set(m.state_dict().keys())
    assert {"encoder.blocks.0.gamma_1", "encoder.blocks.1.gamma_2", "decoder.blocks.0.gamma_1",
            "encoder.pos_embed"} <= keys
    assert m.state_dict()["encoder.pos_embed"].shape == (1, 128, 128)
    assert float(m.state_dict()["decoder.blocks.0.gamma_2"][0]) == pytest.approx(0.1)
    plain = set(vmae.PretrainVisionTransformer(**kw).state_dict().keys())
    assert not any("gamma" in k or k.endswith("pos_embed") for k in plain)


def test_engine_signature_tracks_weight_changes():
    m = vmae.PretrainVisionTransformer(**synthetic.model_kwargs("tiny_4x4"))
    e = vmae._Engine()
    s0 = e._sig(m, "cpu")
    assert e._sig(m, "cpu") == s0
    with torch.no_grad():
        m.mask_token.add_(1.0)
    s1 = e._sig(m, "cpu")
    assert s1 != s0
    m.load_state_dict(m.state_dict())
    assert e._sig(m, "cpu") != s1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gather_worker(rank, world, port, n_samples, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x = torch.arange(n_samples * 6, dtype=torch.float32).reshape(n_samples, 2, 3)
        mask = torch.arange(n_samples * 4).reshape(n_samples, 4) % 2 == 0
        xs, ms, (lo, hi) = cdist.shard_samples(x, mask)
        assert xs.shape[0] == hi - lo and torch.equal(xs, x[lo:hi]) and torch.equal(ms, mask[lo:hi])
        y_local = xs * 2 + rank * 0  # stand-in for the per-rank forward
        full = cdist.gather_samples(y_local, n_samples)
        ok_all = torch.equal(full, x * 2)
        root = cdist.gather_samples(y_local, n_samples, dst=0)
        ok_root = (root is None) if rank != 0 else torch.equal(root, x * 2)
        q.put((rank, bool(ok_all and ok_root)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_samples", [7, 8, 1])
def test_shard_and_gather_gloo_world2(n_samples):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, n_samples, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    results = dict(q.get(timeout=10) for _ in range(2))
    assert results == {0: True, 1: True}
