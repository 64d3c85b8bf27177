"""TEST INFRASTRUCTURE ONLY -- runs the REAL reference ``ImuConditionedFlowGenerator`` (cwm/models/segmentation.py:756-967)
on a seeded pair of small conjoined models (IMU-conditioned padded predictor + flow2imu with the 'flowback_rgb01'
preprocessor, both 128 px) and a seeded RAFT-large, and writes tests/golden/imu_sweep_128px.npz: the head motion the
flow2imu model predicts for the static movie, and the counterfactual videos conditioned on it.  Needs /root/reference."""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import ref_loader  # noqa: E402
from counterfactualworldmodels_b200 import synthetic  # noqa: E402
from make_golden_raft import e2e_frames  # noqa: E402

RAFT_ITERS = 3
S = 4
SHIFTS = [[1, 0], [0, -2], [-1, 1], [2, 2]]


def sweep_inputs():
    """One image (a seeded frame pair, frame 0 is what the sweep uses) and S = 4 (active, passive) 2x2 clumps on the
    32 x 32 patch grid, placed so that no shift leaves the frame and no row needs the rectangulariser."""
    x = e2e_frames(1, 128)
    T, h, w = 2, 32, 32
    active = torch.ones(1, T, h, w, S, dtype=torch.bool)
    passive = torch.zeros(1, T, h, w, S, dtype=torch.bool)
    passive[:, -1] = True
    for s in range(S):
        active[0, -1, 6 + 4 * s:8 + 4 * s, 5:7, s] = False
        passive[0, -1, 20:22, 8 + 5 * s:10 + 5 * s, s] = False
    return x, active.reshape(1, -1, S), passive.reshape(1, -1, S)


def build(conj, seg, raft_mod, flow_kwargs_key, ckpt_or_model):
    """The same construction for the reference (flow_model_ckpt=path) and the mirror (flow_model=module)."""
    pred = synthetic.build_conjoined(conj, "conj_padded_128")
    synthetic.init_weights_(pred, seed=21, style="perturbed")
    head = synthetic.build_conjoined(conj, "conj_flow2imu_128",
                                     main_input_kwargs={flow_kwargs_key: ckpt_or_model, 'iters': RAFT_ITERS})
    synthetic.init_weights_(head, seed=22, style="perturbed")
    return pred.eval().requires_grad_(False), head.eval().requires_grad_(False)


def main():
    ref_loader.import_reference()
    import cwm.models.VideoMAE.conjoined_vmae as ref_conj
    import cwm.models.raft.raft_model as ref_raft
    import cwm.models.segmentation as ref_seg
    torch.manual_seed(0)
    args = ref_raft.get_args("")
    args.multiframe, args.scale_inputs, args.output_dim = True, True, None
    with tempfile.TemporaryDirectory() as tmp:
        ckpt = os.path.join(tmp, "raft-large.pth")
        torch.save(ref_raft.RAFT(args).state_dict(), ckpt)
        pred, head = build(ref_conj, ref_seg, ref_raft, 'flow_model_ckpt', ckpt)
        flow_model = ref_raft.load_raft_model(ckpt)
    G = ref_seg.ImuConditionedFlowGenerator(predictor=pred, head_motion_predictor=head, flow_model=flow_model,
                                            imagenet_normalize_inputs=True, temporal_dim=2, raft_iters=RAFT_ITERS, seed=0)
    x, active, passive = sweep_inputs()
    with torch.no_grad():
        G.set_input(x)
        h = G.get_static_imu()
        G.reset_padding_masks()
        ys, flows = G.predict_counterfactual_videos_and_flows(x, active, passive, shifts=SHIFTS, sample_batch_size=2,
                                                              raft_iters=RAFT_ITERS)
    assert tuple(h.shape) == (1, 5, 96) and tuple(ys.shape) == (S, 2, 3, 128, 128) and tuple(flows.shape) == (S, 1, 2, 128, 128)
    path = os.path.join(ROOT, "tests", "golden", "imu_sweep_128px.npz")
    np.savez_compressed(path, head_motion=h.numpy(), ys_frame1=ys[:, 1, :, ::2, ::2].numpy(),
                        ys_frame0_equals_input=np.array(bool(torch.equal(ys[:, 0], x[:, 0].expand(S, -1, -1, -1)))),
                        flows_absmax=np.array(float(flows.abs().max())))
    print(f"imu_sweep_128px: head motion {tuple(h.shape)} |h| max {h.abs().max():.3f}, videos {tuple(ys.shape)}, "
          f"frame 0 == input: {bool(torch.equal(ys[:, 0], x[:, 0].expand(S, -1, -1, -1)))} | "
          f"{os.path.getsize(path) / 1e3:.0f} KB")


if __name__ == "__main__":
    main()
