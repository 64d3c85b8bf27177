"""TEST INFRASTRUCTURE ONLY -- pins oracle/flowstats_oracle.py against the REAL reference (FlowSampleFilter,
FlowGenerator.compute_mean_motion_map) and writes tests/golden/fs_*.npz.  Needs /root/reference."""
import os
import sys

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import flowstats_oracle as fso  # noqa: E402
import ref_loader  # noqa: E402
from counterfactualworldmodels_b200 import synthetic  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
# name: (B, S, image side, patch, seed)
CASES = {"fs_b2_s6_32px": (2, 6, 32, 4, 1), "fs_b1_s12_64px": (1, 12, 64, 8, 2), "fs_b1_s8_224px": (1, 8, 224, 8, 3)}
FILTER = dict(filter_methods=['patch_magnitude', 'flow_area', 'num_corners'], flow_magnitude_threshold=5.0,
              flow_area_threshold=0.75, num_corners_threshold=2)


def main():
    ref_vmae, _ = ref_loader.import_reference()
    import cwm.models.sampling as ref_sampling
    import cwm.models.segmentation as ref_seg
    kw = synthetic.model_kwargs("tiny_4x4")
    kw.update(encoder_depth=1, decoder_depth=1)
    G = ref_seg.FlowGenerator(predictor=ref_vmae.PretrainVisionTransformer(**kw).eval(), flow_model=nn.Identity())
    for name, (B, S, side, patch, seed) in CASES.items():
        flows_bs, centers = fso.make_flows(B, S, side, side, seed)
        active = fso.make_active(B, S, side // patch, side // patch, centers, patch)
        flows = fso.batch_to_samples(flows_bs, B)
        filt = ref_sampling.FlowSampleFilter(**FILTER)
        _, _, patch_mag, _ = filt.compute_flow_magnitude(flows, active)
        ref_flows, ref_mask = filt(flows.clone(), active)
        ref_mask_bs = ref_mask[:, 0, 0, 0, :]
        or_flows, or_mask, st = fso.filter_samples(flows, active, FILTER["filter_methods"], 5.0, 0.75, 2)
        assert torch.equal(or_mask, ref_mask_bs) and torch.equal(or_flows, ref_flows), name
        assert torch.equal(st["patch_flow_mag"], patch_mag)
        out = dict(shape=np.array([B, S, side, patch, seed]), filter_mask=ref_mask_bs.numpy(),
                   patch_flow_mag=patch_mag.numpy(), flow_area=st["flow_area"].numpy(),
                   num_corners=st["num_corners"].numpy(), mag_min=st["min"].numpy(), mag_max=st["max"].numpy())
        for tag, src in (("raw", flows), ("filtered", ref_flows)):
            for nps in (False, True):
                mm_ref = G.compute_mean_motion_map(src, normalize_per_sample=nps)
                mm_or = fso.mean_motion_map(src, normalize_per_sample=nps)
                assert torch.equal(mm_or, mm_ref), (name, tag, nps)
                if side <= 64 or (tag, nps) in (("filtered", False), ("raw", True)):
                    out[f"motion_map_{tag}_{'nps' if nps else 'plain'}"] = mm_ref.numpy()
        mm4 = G.compute_mean_motion_map(flows.norm(dim=1, p=2).mean(-1)[:, None])
        assert torch.equal(fso.mean_motion_map(flows.norm(dim=1, p=2).mean(-1)[:, None]), mm4)
        out["motion_map_from_distribution"] = mm4.numpy()
        # motion covariance / correlation (segmentation.py:478-547), stored for the small cases
        if side <= 64:
            ds = 4 if side == 32 else 8
            for cov in (True, False):
                c_ref = ref_seg.FlowGenerator.compute_flow_corrs(flows, downsample=ds, use_covariance=cov)
                c_or = fso.flow_corrs(flows, downsample=ds, use_covariance=cov)
                assert torch.equal(c_or, c_ref), (name, cov)
                out[f"flow_{'cov' if cov else 'corr'}_ds{ds}"] = c_ref.numpy().astype(np.float32)
            c_ref = ref_seg.FlowGenerator.compute_flow_corrs(flows, downsample=ds, use_covariance=True, take_top_k=3)
            assert torch.equal(fso.flow_corrs(flows, downsample=ds, use_covariance=True, take_top_k=3), c_ref)
            out[f"flow_cov_ds{ds}_top3"] = c_ref.numpy().astype(np.float32)
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: filtered {ref_mask_bs.int().tolist()} oracle == reference | {os.path.getsize(path) / 1e3:.0f} KB")


if __name__ == "__main__":
    main()
