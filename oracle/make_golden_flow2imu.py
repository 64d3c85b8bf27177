"""TEST INFRASTRUCTURE ONLY -- the FULL-SIZE flow2imu model.

Runs the REAL reference factory `conj.imu400_8x8patch_2frames_1tube_flowbackrgb01` (conjoined_vmae.py:1218-1228:
ViT-base, 8x8 patches, 224 px, 784 main tokens of the 7-channel forward-flow | backward-flow | rgb input that its
'flowback_rgb01' preprocessor builds with RAFT, 25 fully-masked IMU tokens + the dummy token) on CPU exactly as
`ImuConditionedFlowGenerator` calls it (segmentation.py:839-846: output_main=False, output_context=True) and writes
tests/golden/flow2imu_full_b2.npz: the predicted IMU tokens [B, 25, 96].

The reference loads RAFT from a checkpoint FILE (preprocessor.py:210,255-256); a seeded random-init RAFT-large state dict
is written to a temporary file for it (no checkpoint ships with the repository and there is no network).
"""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import ref_loader  # noqa: E402
import vmae_oracle as oracle  # noqa: E402
from counterfactualworldmodels_b200 import synthetic  # noqa: E402

RAFT_ITERS = 4
B = 2
WEIGHT_SEED, DATA_SEED = 31, 31
PARAMS = 135730048 + 5257536            # notebook: 135,730,048 trainable + the frozen RAFT-large inside the preprocessor


def inputs():
    """x [B, 2, 3, 224, 224] raw frames in [0, 1] (frame 1 = frame 0 with a moved square: a non-trivial flow field), the
    all-visible two-frame mask (the main stream keeps the frame-1 half), the zero IMU signal and its all-masked token
    mask (`get_fake_head_motion`, segmentation.py:814-832)."""
    x = synthetic.make_video(B, (224, 224), seed=DATA_SEED)
    mask = torch.zeros(B, 2 * 784, dtype=torch.bool)
    imu = torch.zeros(B, 6, 400)
    mask_ctx = torch.ones(B, 25, dtype=torch.bool)
    return x, mask, imu, mask_ctx


def build(conj, flow_key, flow_value):
    """Same construction for the reference (flow_model_ckpt=<path>) and the mirror (flow_model=<module>)."""
    m = conj.imu400_8x8patch_2frames_1tube_flowbackrgb01(main_input_kwargs={flow_key: flow_value, 'iters': RAFT_ITERS})
    synthetic.init_weights_(m, seed=WEIGHT_SEED, style="perturbed")
    return m.eval().requires_grad_(False)


def main():
    ref_loader.import_reference()
    import cwm.models.VideoMAE.conjoined_vmae as ref_conj
    import cwm.models.raft.raft_model as ref_raft
    torch.manual_seed(0)
    args = ref_raft.get_args("")
    args.multiframe, args.scale_inputs, args.output_dim = True, True, None
    with tempfile.TemporaryDirectory() as tmp:
        ckpt = os.path.join(tmp, "raft-large.pth")
        torch.save(ref_raft.RAFT(args).state_dict(), ckpt)
        ref = build(ref_conj, 'flow_model_ckpt', ckpt)
    n_params = sum(p.numel() for p in ref.parameters())
    assert n_params == PARAMS, n_params
    x, mask, imu, mask_ctx = inputs()
    with torch.no_grad():
        # `_preprocess` of the head-motion generator: transposed view + imagenet normalisation (prediction.py:304-312)
        y = ref(oracle.preprocess(x), mask=mask, x_context=imu, mask_context=mask_ctx, output_main=False,
                output_context=True)
    assert tuple(y.shape) == (B, 25, 96), y.shape
    path = os.path.join(ROOT, "tests", "golden", "flow2imu_full_b2.npz")
    np.savez_compressed(path, y_ctx=y.numpy().astype(np.float32), num_params=np.array([n_params]),
                        weights_checksum=np.array([synthetic.weights_checksum(ref)]))
    print(f"flow2imu_full_b2: y_ctx {tuple(y.shape)} std {y.std():.3f} absmax {y.abs().max():.3f} | "
          f"{os.path.getsize(path) / 1e3:.0f} KB")


if __name__ == "__main__":
    main()
