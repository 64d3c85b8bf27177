"""MovabilityPredictor (cwm/models/movability.py:13-360): the iteration loop over counterfactual
sweeps.  CPU: construction, keypoint distribution, the reference's error behaviour.  GPU: a 2-iteration run on the
base 8x8 predictor with RAFT-small -- consistency with the mirrored pieces it is composed of, and determinism."""
import pytest
import torch

from counterfactualworldmodels_b200 import movability, segmentation, synthetic, vmae

DEV = "cuda:0"


class _Keypoints(torch.nn.Module):
    def forward(self, x):                          # [B, T, C, H, W] -> logits [B, 1, 1, H, W]
        return (x[:, :1].mean(2, keepdim=True) - 0.5) * 8


def _raft_small():
    from counterfactualworldmodels_b200 import raft
    torch.manual_seed(0)
    args = raft.get_args("")
    args.multiframe, args.scale_inputs, args.output_dim, args.small = True, True, None, True
    return raft.RAFT(args).eval().requires_grad_(False)


def test_construction_and_keypoint_distribution_on_cpu():
    model = vmae.PretrainVisionTransformer(**synthetic.model_kwargs("tiny_8x8"))
    M = movability.MovabilityPredictor(predictor=model, imagenet_normalize_inputs=True, temporal_dim=2,
                                       keypoint_predictor=_Keypoints(), num_iters=1)
    assert isinstance(M, segmentation.ImuConditionedFlowGenerator) and M.head_motion_generator is None
    assert M.num_iters == 1 and M.movability_maps == [] and M.get_total_movability() is None
    x = synthetic.make_video(1, (64, 64), seed=0)   # one image: the reference's `value - value.amin((-2, -1))` has no
    M.set_input(x)                                  # keepdim and only broadcasts for a batch of one (prediction.py:825)
    M.set_keypoints_distribution()
    d = M.keypoints_distribution
    assert tuple(d.shape) == (1, 1, 64, 64) and float(d.amin()) == 0.0 and float(d.amax()) == pytest.approx(1.0)
    want = ((x[:, :1].mean(2, keepdim=True) - 0.5) * 8).squeeze(-3).sigmoid() ** 8
    want = want - want.amin()
    torch.testing.assert_close(d, want / want.amax().clamp(min=1e-3))
    with pytest.raises(RuntimeError):
        M.predict_keypoints_distribution(synthetic.make_video(2, (64, 64), seed=0))
    # without a keypoint predictor the default `initialize_from_keypoints=True` fails like the reference: `1 - None`
    M2 = movability.MovabilityPredictor(predictor=model, imagenet_normalize_inputs=True, temporal_dim=2)
    with pytest.raises(TypeError):
        M2(x[:1])
    with pytest.raises(NotImplementedError, match="initial patches"):
        M2(x[:1], initial_active_patches=torch.zeros(1, 128, 1, dtype=torch.bool))
    with pytest.raises(NotImplementedError):
        M2.visualize_iterations()


@pytest.mark.gpu
def test_gpu_movability_iterations():
    cfg = "base_8x8"

    def run():
        model = vmae.PretrainVisionTransformer(**synthetic.model_kwargs(cfg))
        synthetic.init_weights_(model, seed=0)
        M = movability.MovabilityPredictor(
            predictor=model.to(DEV).eval(), flow_model=_raft_small().to(DEV), imagenet_normalize_inputs=True,
            temporal_dim=2, raft_iters=3, seed=5, initialize_from_keypoints=False, num_initial_samples=4,
            num_samples_per_iteration=4, num_iters=2, sample_batch_size=4)
        x = synthetic.make_video(1, (224, 224), seed=2).to(DEV)
        return M, M(x)

    M, final = run()
    assert tuple(final.shape) == (1, 1, 224, 224) and torch.isfinite(final).all()
    assert float(final.amin()) >= 0.0 and float(final.amax()) <= 1.0 + 1e-6
    assert len(M.movability_maps) == 3 and M.it == 2
    assert [tuple(f.shape) for f in M.flow_samples_per_iter] == [(1, 2, 224, 224, 4)] * 3
    # iteration 0 moves one patch and holds none, later iterations move one and hold one (movability.py:39-45)
    n = M.predictor.mask_size[1] * M.predictor.mask_size[2]
    vis = lambda m: (~m[:, n:]).sum(1)[0].tolist()          # visible frame-1 patches per sample  # noqa: E731
    assert vis(M.active_patches_per_iter[0]) == [1] * 4 and vis(M.passive_patches_per_iter[0]) == [0] * 4
    assert vis(M.active_patches_per_iter[1]) == [1] * 4 and vis(M.passive_patches_per_iter[2]) == [1] * 4
    total = M.get_total_movability()
    want = M.compute_mean_motion_map(torch.cat(M.flow_samples_per_iter, -1))
    assert torch.equal(total, want)
    assert torch.equal(M.movability_maps[-1], M.compute_mean_motion_map(M.flow_samples_per_iter[-1]))
    assert tuple(M.get_minimum_movability().shape) == (1, 1, 224, 224)
    # same seeds, same sweep
    M2, final2 = run()
    assert all(torch.equal(a, b) for a, b in zip(M.active_patches_per_iter, M2.active_patches_per_iter))
    assert torch.equal(final, final2)
