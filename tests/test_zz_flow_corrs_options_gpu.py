"""compute_flow_corrs with the options the reference's only caller never sets (segmentation.py:478-547): they take the
torch-op route `_flow_corrs_general`, which is pinned bit for bit against the live reference on CPU
(tests/test_flowstats.py::test_flow_corrs_general_options_match_the_live_reference).  This file checks that route on
device tensors against the same route on CPU.  (Written after the round's GPU budget was spent: it sorts last so that a
surprise here cannot hide any other GPU test behind `pytest -x`.)"""
import pytest
import torch

import flowstats_oracle as fso

DEV = "cuda:0"


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [dict(do_spearman=True), dict(thresh=1.0, binarize=True), dict(zscore=True, use_covariance=True)])
def test_gpu_flow_corrs_general_options(kw):
    from counterfactualworldmodels_b200 import segmentation
    flows_bs, _ = fso.make_flows(2, 6, 32, 32, 4)
    want = segmentation.FlowGenerator._flow_corrs_general(fso.batch_to_samples(flows_bs, 2), downsample=4, **kw)
    got = segmentation.FlowGenerator.compute_flow_corrs(fso.batch_to_samples(flows_bs.to(DEV), 2), downsample=4, **kw)
    assert got.shape == want.shape and bool(torch.isfinite(got).all())
    assert torch.allclose(got.cpu(), want, rtol=1e-3, atol=1e-4)
