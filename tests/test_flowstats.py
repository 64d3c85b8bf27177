"""SURVEY.md section 8(f) rank 2 -- flow-derived statistics (FlowSampleFilter, compute_mean_motion_map).
CPU: the torch oracle against the fixtures the REAL reference produced (tests/golden/fs_*.npz).
GPU: csrc/flowstats.cu through the mirror classes against the fixtures and the oracle.  Tolerance: fp32 reductions in
a different order -> 1e-5 relative / 1e-6 absolute on statistics and maps; filter masks exact."""
import os

import numpy as np
import pytest
import torch

import flowstats_oracle as fso
from conftest import GOLDEN_DIR, needs_reference

FS_CASES = ["fs_b2_s6_32px", "fs_b1_s12_64px", "fs_b1_s8_224px"]
FILTER = dict(filter_methods=['patch_magnitude', 'flow_area', 'num_corners'], flow_magnitude_threshold=5.0,
              flow_area_threshold=0.75, num_corners_threshold=2)
DEV = "cuda:0"
RTOL, ATOL = 1e-5, 1e-6


def load_fs(case):
    path = os.path.join(GOLDEN_DIR, case + ".npz")
    if not os.path.exists(path):
        pytest.skip(f"golden fixture {case} missing")
    z = np.load(path)
    d = {k: z[k] for k in z.files}
    B, S, side, patch, seed = [int(v) for v in d["shape"]]
    flows_bs, centers = fso.make_flows(B, S, side, side, seed)
    d["flows_bs"] = flows_bs
    d["active"] = fso.make_active(B, S, side // patch, side // patch, centers, patch)
    d["B"] = B
    return d


@pytest.mark.parametrize("case", FS_CASES)
def test_oracle_matches_reference_fixture(case):
    d = load_fs(case)
    flows = fso.batch_to_samples(d["flows_bs"], d["B"])
    zeroed, mask, st = fso.filter_samples(flows, d["active"], FILTER["filter_methods"], 5.0, 0.75, 2)
    assert np.array_equal(mask.numpy(), d["filter_mask"])
    assert np.array_equal(st["patch_flow_mag"].numpy(), d["patch_flow_mag"])
    assert np.array_equal(st["flow_area"].numpy(), d["flow_area"])
    assert np.array_equal(fso.mean_motion_map(zeroed).numpy(), d["motion_map_filtered_plain"])
    assert np.array_equal(fso.mean_motion_map(flows, normalize_per_sample=True).numpy(), d["motion_map_raw_nps"])


@pytest.mark.gpu
@pytest.mark.parametrize("case", FS_CASES)
def test_gpu_filter_and_motion_map_match_reference_fixture(case):
    from counterfactualworldmodels_b200 import sampling
    d = load_fs(case)
    B = d["B"]
    flows_bs = d["flows_bs"].to(DEV)
    flows = fso.batch_to_samples(flows_bs, B)           # permuted VIEW, like the reference hands over
    assert not flows.is_contiguous()
    active = d["active"].to(DEV)
    filt = sampling.FlowSampleFilter(**FILTER)
    st = filt.compute_flow_statistics(flows, active)
    for key, want in (("patch_flow_mag", "patch_flow_mag"), ("flow_area", "flow_area"), ("num_corners", "num_corners"),
                      ("min", "mag_min"), ("max", "mag_max")):
        np.testing.assert_allclose(st[key].cpu().numpy(), d[want], rtol=RTOL, atol=ATOL, err_msg=key)
    raw = flows_bs.clone()
    # motion maps of the unfiltered flows
    for nps, key in ((False, "motion_map_raw_plain"), (True, "motion_map_raw_nps")):
        if key in d:
            sums = sampling.flow_magnitude_sum(flows, normalize_per_sample=nps)
            mm = sampling.motion_map_finalize(sums, flows.shape[-1])
            np.testing.assert_allclose(mm.cpu().numpy(), d[key], rtol=1e-4, atol=2e-6, err_msg=key)
    # fused: filter mask applied inside the sum, flows untouched
    mask, stats = filt.filter_mask(flows, active)
    assert np.array_equal(mask.cpu().numpy(), d["filter_mask"])
    sums = sampling.flow_magnitude_sum(flows, filter_mask=mask)
    mm = sampling.motion_map_finalize(sums, flows.shape[-1])
    np.testing.assert_allclose(mm.cpu().numpy(), d["motion_map_filtered_plain"], rtol=1e-4, atol=2e-6)
    assert torch.equal(flows_bs, raw)
    # reference API: zero in place on the view, mask returned expanded
    out, mask5 = filt(flows, active)
    assert out.data_ptr() == flows.data_ptr() and mask5.shape == flows.shape
    want, _, _ = fso.filter_samples(fso.batch_to_samples(d["flows_bs"], B), d["active"], FILTER["filter_methods"], 5.0,
                                    0.75, 2)
    assert torch.equal(out.cpu(), want)
    if "motion_map_filtered_nps" in d:
        sums = sampling.flow_magnitude_sum(out, normalize_per_sample=True)
        mm = sampling.motion_map_finalize(sums, out.shape[-1])
        np.testing.assert_allclose(mm.cpu().numpy(), d["motion_map_filtered_nps"], rtol=1e-4, atol=2e-6)


@pytest.mark.gpu
def test_gpu_flow_generator_entry_points_and_chunked_accumulation():
    from counterfactualworldmodels_b200 import sampling, segmentation, synthetic, vmae
    kw = synthetic.model_kwargs("tiny_4x4")
    kw.update(encoder_depth=1, decoder_depth=1)
    G = segmentation.FlowGenerator(predictor=vmae.PretrainVisionTransformer(**kw).to(DEV).eval())
    d = load_fs("fs_b1_s12_64px")
    flows_bs = d["flows_bs"].to(DEV)
    G.set_input(torch.zeros(1, 2, 3, 64, 64, device=DEV))
    flows = G.filter_flow_samples(flows_bs.clone(), d["active"].to(DEV))
    assert flows.shape == (1, 2, 64, 64, 12)
    mm = G.compute_mean_motion_map(flows)
    np.testing.assert_allclose(mm.cpu().numpy(), d["motion_map_filtered_plain"], rtol=1e-4, atol=2e-6)
    mm4 = G.compute_mean_motion_map(fso.batch_to_samples(flows_bs, 1).norm(dim=1, p=2).mean(-1)[:, None])
    np.testing.assert_allclose(mm4.cpu().numpy(), d["motion_map_from_distribution"], rtol=1e-4, atol=2e-6)
    # a sweep processed in chunks (or on several ranks): partial sums accumulate to the same map
    raw = fso.batch_to_samples(flows_bs, 1)
    acc = None
    for s0 in range(0, 12, 5):
        acc = sampling.flow_magnitude_sum(raw[..., s0:s0 + 5], out=acc)
    mm_chunks = sampling.motion_map_finalize(acc, 12)
    np.testing.assert_allclose(mm_chunks.cpu().numpy(), d["motion_map_raw_plain"], rtol=1e-4, atol=2e-6)
    with pytest.raises(ValueError):
        sampling.FlowSampleFilter(filter_methods=['nope']).filter_mask(raw, d["active"].to(DEV))


def test_no_cpu_fallback():
    from counterfactualworldmodels_b200 import sampling
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sampling.flow_sample_stats(torch.zeros(1, 2, 8, 8, 2))


@pytest.mark.gpu
@pytest.mark.parametrize("case,ds", [("fs_b2_s6_32px", 4), ("fs_b1_s12_64px", 8)])
def test_gpu_flow_corrs_match_reference_fixture(case, ds):
    """Motion covariance / correlation between image locations (segmentation.py:478-547)."""
    from counterfactualworldmodels_b200 import segmentation
    d = load_fs(case)
    flows = fso.batch_to_samples(d["flows_bs"].to(DEV), d["B"])
    for cov, key in ((True, f"flow_cov_ds{ds}"), (False, f"flow_corr_ds{ds}")):
        got = segmentation.FlowGenerator.compute_flow_corrs(flows, downsample=ds, use_covariance=cov)
        assert got.shape == d[key].shape
        np.testing.assert_allclose(got.cpu().numpy(), d[key], rtol=2e-4, atol=2e-5, err_msg=key)
    got = segmentation.FlowGenerator.compute_flow_corrs(flows, downsample=ds, use_covariance=True, take_top_k=3)
    np.testing.assert_allclose(got.cpu().numpy(), d[f"flow_cov_ds{ds}_top3"], rtol=2e-4, atol=2e-5)
    # properties that hold at any size: symmetric, unit diagonal of the correlation wherever the variance is non-zero
    r = segmentation.FlowGenerator.compute_flow_corrs(flows, downsample=ds, use_covariance=False)
    n = r.shape[2] * r.shape[3]
    r2 = r.reshape(-1, n, n)
    assert torch.allclose(r2, r2.transpose(1, 2), atol=1e-6)
    diag = torch.diagonal(r2, dim1=1, dim2=2)
    assert bool(((diag - 1).abs() < 1e-5).logical_or(diag == 0).all())


def test_oracle_flow_corrs_matches_reference_fixture():
    d = load_fs("fs_b2_s6_32px")
    flows = fso.batch_to_samples(d["flows_bs"], d["B"])
    assert np.array_equal(fso.flow_corrs(flows, downsample=4, use_covariance=True).numpy(), d["flow_cov_ds4"])
    assert np.array_equal(fso.flow_corrs(flows, downsample=4, use_covariance=False).numpy(), d["flow_corr_ds4"])


@needs_reference
def test_flow_corrs_general_options_match_the_live_reference():
    """The options of compute_flow_corrs the reference's only caller never sets (segmentation.py:478-547): the torch-op
    route reproduces the reference bit for bit on CPU (the public method only accepts device tensors)."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(GOLDEN_DIR), "..", "oracle"))
    import ref_loader
    ref_loader.import_reference()
    import cwm.models.segmentation as ref_seg
    import cwm.models.utils as ref_utils
    from counterfactualworldmodels_b200 import segmentation
    flows_bs, _ = fso.make_flows(2, 6, 32, 32, 4)
    flows = fso.batch_to_samples(flows_bs, 2)
    swap = fso.batch_to_samples(fso.make_flows(2, 6, 32, 32, 5)[0], 2)
    cases = [dict(do_spearman=True), dict(thresh=1.0), dict(thresh=1.0, binarize=True), dict(range_thresh=0.5),
             dict(normalize=True, use_covariance=True), dict(zscore=True), dict(flow_samples_swap=swap, take_top_k=4),
             dict(distance_func=ref_utils.ChannelL1Error(dim=1), use_covariance=True)]
    for kw in cases:
        want = ref_seg.FlowGenerator.compute_flow_corrs(flows, downsample=4, **kw)
        got = segmentation.FlowGenerator._flow_corrs_general(flows, downsample=4, **kw)
        assert torch.equal(got, want), kw
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        segmentation.FlowGenerator.compute_flow_corrs(flows, downsample=4, zscore=True)
