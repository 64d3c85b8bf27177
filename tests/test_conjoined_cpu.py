"""CPU checks of the padded / conjoined (IMU-conditioned) models: the oracle
against the fixtures the REAL reference produced, known answers the reference's notebook records, state_dict /
attribute parity of the drop-in modules, and the integer padding-mask bookkeeping (bit-exact)."""
import os

import numpy as np
import pytest
import torch

import conjoined_oracle as co
import make_golden_conjoined as mgc
from conftest import load_golden_conjoined, needs_reference
from counterfactualworldmodels_b200 import conjoined_vmae as C
from counterfactualworldmodels_b200 import preprocessor, synthetic, transformer

SMALL_CASES = ["conj_padded_small_ragged", "conj_padded_small_ctxmasked", "conj_padded_small_predict",
               "conj_flow2imu_small"]


def _stream_inputs(model, x, mask, imu, mc):
    (x_m, mask_m, _), (x_c, mask_c, _) = model.get_stream_inputs(x.transpose(1, 2), mask, None, x_context=imu,
                                                                 mask_context=mc)
    return x_m, mask_m, x_c, mask_c


@pytest.mark.parametrize("case", SMALL_CASES)
def test_conjoined_oracle_matches_reference_fixture(case):
    g = load_golden_conjoined(case)
    name, B, style, wseed, x, mask, imu, mc = mgc.case_inputs(case)
    assert float(x.double().sum()) == pytest.approx(float(g["x_fingerprint"][0]), abs=1e-6)
    assert torch.equal(mask, g["mask"]) and torch.equal(mc, g["mask_ctx"])
    m = synthetic.build_conjoined(C, name)
    synthetic.init_weights_(m, seed=wseed, style=style)
    assert synthetic.weights_checksum(m) == pytest.approx(float(g["weights_checksum"][0]), abs=1e-6)
    assert sum(p.numel() for p in m.parameters()) == int(g["num_params"][0])
    x_m, mask_m, x_c, mask_c = _stream_inputs(m, x, mask, imu, mc)
    y, yc = co.conjoined_forward(m.state_dict(), x_m, mask_m, x_c, mask_c, synthetic.conjoined_oracle_cfg(name), True,
                                 True)
    assert y.shape == g["y"].shape and yc.shape == g["y_ctx"].shape
    if y.numel():
        assert (y - g["y"]).abs().max().item() < 2e-5
    assert (yc - g["y_ctx"]).abs().max().item() < 2e-5


def test_padded_oracle_matches_reference_fixture():
    g = load_golden_conjoined("padded_small_ragged")
    B, style, wseed, x, mask = mgc.padded_case_inputs("padded_small_ragged")
    from functools import partial
    m = C.PaddedVisionTransformer(norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), **mgc.PADDED_KW)
    synthetic.init_weights_(m, seed=wseed, style=style)
    assert synthetic.weights_checksum(m) == pytest.approx(float(g["weights_checksum"][0]), abs=1e-6)
    scfg = dict(enc_heads=4, dec_heads=2, max_pad=8, min_pad=0, pos="sinusoid", eps=1e-6)
    y = co.padded_forward(m.state_dict(), x.transpose(1, 2), mask, scfg)
    assert y.shape == g["y"].shape and (y - g["y"]).abs().max().item() < 2e-5
    assert int((y.abs().sum(-1) == 0).sum()) == int(g["null_rows"][0])


def test_known_answers_parameter_counts():
    """Recorded by the reference's notebook (demo/MovabilityAndMotionCovariance.ipynb:355-375)."""
    m = C.imu400_base_4x4patch_2frames_1tube()
    n = lambda mod: sum(p.numel() for p in mod.parameters())
    assert (n(m), n(m.main_stream), n(m.context_stream)) == (148265040, 92496432, 23272992)
    assert n(m) - n(m.main_stream) - n(m.context_stream) == 32495616
    assert m.main_stream.num_patches == 6272 and m.context_stream.encoder.num_tokens == 25
    f = C.imu400_8x8patch_2frames_1tube_flowbackrgb01()
    assert (n(f), n(f.main_stream), n(f.context_stream)) == (135730048, 92956480, 23272512)
    assert f.main_stream.num_patches == 784 and f.context_stream.encoder.num_tokens == 25
    assert list(m.encoder_conjoining_blocks.keys()) == ['0-0', '3-3', '6-6', '9-9']
    assert list(m.decoder_conjoining_blocks.keys()) == ['0-0', '1-1', '2-2', '3-3']
    assert list(f.encoder_conjoining_blocks.keys()) == ['0-0', '11-11']
    # attributes the wrappers read
    assert m.mask_size == (2, 56, 56) and tuple(m.patch_size) == (1, 4, 4) and m.num_frames == 2
    assert m.main_stream.max_padding_tokens == 64 and m.context_stream.max_padding_tokens == 25
    assert tuple(m.context_stream.patch_size) == (16, 1, 1)


@pytest.mark.parametrize("B,N,P,min_pad", [(3, 128, 8, 0), (4, 50, 16, 2), (2, 25, 25, 0), (1, 9, 4, 0)])
def test_padding_masks_bit_exact(B, N, P, min_pad):
    """`_set_padding_mask` of the drop-in module == the oracle restatement of conjoined_vmae.py:49-116."""
    from functools import partial
    m = C.PaddedVisionTransformer(norm_layer=partial(torch.nn.LayerNorm, eps=1e-6),
                                  **dict(mgc.PADDED_KW, encoder_depth=1, decoder_depth=1, min_padding_tokens=min_pad,
                                         max_padding_tokens=P))
    rng = np.random.RandomState(B * 100 + N)
    for trial in range(4):
        mask = torch.from_numpy(rng.rand(B, N) < 0.5)
        # keep the spread of visible counts within the padding budget, as the reference requires
        nv = (~mask).sum(-1)
        for b in range(B):
            while int(nv.max() - (~mask[b]).sum()) + min_pad > P:
                idx = torch.nonzero(mask[b]).flatten()[0]
                mask[b, idx] = False
        if trial == 3:
            mask[:] = True  # nothing visible at all: one null token per row (conjoined_vmae.py:69-82)
        m._set_padding_mask(mask, device=mask.device)
        pm, fm, nm = co.padding_masks(mask, P, min_pad)
        assert torch.equal(m.padding_mask, pm) and torch.equal(m.full_input_mask, fm) and torch.equal(m.null_mask, nm)
        assert len(set((~m.full_input_mask).sum(-1).tolist())) == 1   # every row has the same visible count
    m._reset_padding_mask()
    assert m.padding_mask is None and m.null_mask is None


def test_padding_mask_attribute_quirk():
    """`hasattr(predictor, 'padding_mask')` is False before a forward / after a reset and True
    once a mask is set, because __getattr__ forwards to main_stream and treats None as missing."""
    m = synthetic.build_conjoined(C, "conj_padded_small")
    assert not hasattr(m, "padding_mask") and hasattr(m, "max_padding_tokens")
    m.main_stream._set_padding_mask(torch.zeros(2, 128, dtype=torch.bool), device="cpu")
    assert hasattr(m, "padding_mask") and m._main_padded and m._context_padded
    m._reset_padding_mask()
    assert not hasattr(m, "padding_mask")
    with pytest.raises(AttributeError):
        m.no_such_attribute
    f = synthetic.build_conjoined(C, "conj_flow2imu_small")
    assert not hasattr(f.main_stream, "padding_mask") and not hasattr(f, "_main_padded")


def test_forward_without_gpu_raises():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    m = synthetic.build_conjoined(C, "conj_padded_small")
    with pytest.raises(RuntimeError, match="no CPU"):
        m(torch.zeros(1, 3, 2, 32, 32), torch.zeros(1, 128, dtype=torch.bool), x_context=torch.zeros(1, 6, 80),
          mask_context=torch.zeros(1, 5, dtype=torch.bool))


def test_pos_embedding_and_preprocessors():
    assert torch.equal(transformer.pos_embedding(26, 192), co.pos_embedding(26, 192))
    assert torch.equal(transformer.pos_embedding([25], 64)[0, 0], co.pos_embedding(26, 64)[0, 25])
    p = preprocessor.get_preprocessor('rgb01', unnormalize=False)
    x = torch.arange(2 * 3 * 3 * 4 * 4, dtype=torch.float32).reshape(2, 3, 3, 4, 4)
    assert torch.equal(p(x), x[:, :, :2]) and p.get_num_frames() == 2 and p.num_channels == 3
    assert torch.equal(p.get_output_frames(torch.arange(6).reshape(2, 3), temporal_dim=1), torch.tensor([[0, 1], [3, 4]]))
    imu = preprocessor.get_preprocessor('imu')
    assert imu(torch.zeros(2, 6, 400)).shape == (2, 6, 400, 1, 1) and imu.num_frames is None
    with pytest.raises(NotImplementedError, match="flow"):
        preprocessor.get_preprocessor('flowback_rgb01')(torch.zeros(1, 3, 2, 8, 8))


@needs_reference
def test_state_dict_keys_match_reference_live():
    import sys
    import ref_loader
    ref_loader.install_stubs()
    sys.path.insert(0, ref_loader.REFERENCE_ROOT)
    import cwm.models.VideoMAE.conjoined_vmae as rconj
    for name in ("conj_padded_small", "conj_flow2imu_small"):
        ref = synthetic.build_conjoined(rconj, name)
        ours = synthetic.build_conjoined(C, name)
        a = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
        b = {k: tuple(v.shape) for k, v in ours.state_dict().items()}
        assert a == b
        res = ours.load_state_dict(ref.state_dict())
        assert not res.missing_keys and not res.unexpected_keys
