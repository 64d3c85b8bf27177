"""Host-side mirror of ``cwm/models/movability.py`` (``MovabilityPredictor`` :13-360): the loop that turns counterfactual
sweeps into a movability map -- SURVEY.md section 3.3, the caller of everything below it.

  iteration 0:  sample patches (uniformly or from a keypoint distribution), move them, flows -> mean motion map
  iteration k:  sample active patches from the movability found so far (and passive ones from its complement),
                move them, flows -> mean motion map

Every step is a call into the mirrored ``FlowGenerator`` / ``ImuConditionedFlowGenerator`` entry points, so one iteration
is: fused counterfactual construction + VMAE prediction, RAFT, the flow-sample filter and the mean motion map on the
device.  Iterations are sequential by construction (each samples from the previous map, movability.py:336-349).
The matplotlib panel (``visualize_iterations``, :232-281) is GUI code and out of scope.
"""
from time import time

import torch

from .segmentation import ImuConditionedFlowGenerator


class MovabilityPredictor(ImuConditionedFlowGenerator):
    VERBOSE = False

    def __init__(self, *args, initialize_from_keypoints=True, iterate_from_keypoints=False, keypoints_power=8,
                 movability_power=1, num_initial_samples=16, num_initial_active_patches=1,
                 num_initial_passive_patches=0, num_samples_per_iteration=16, num_active_patches_per_sample=1,
                 num_passive_patches_per_sample=1, sample_passives_from_movable=False,
                 update_distribution_per_iteration=True, num_iters=2, sample_batch_size=4, **kwargs):
        super().__init__(*args, **kwargs)
        self.initialize_from_keypoints = initialize_from_keypoints
        self.keypoints_power = keypoints_power
        self.keypoints_distribution = None
        self.sample_batch_size = sample_batch_size
        self.movability_power = movability_power
        self.sample_passives_from_movable = sample_passives_from_movable
        self.iterate_from_keypoints = iterate_from_keypoints
        self.num_initial_samples = num_initial_samples
        self.num_initial_active_patches = num_initial_active_patches
        self.num_initial_passive_patches = num_initial_passive_patches
        self.num_samples_per_iteration = num_samples_per_iteration
        self.num_active_patches_per_sample = num_active_patches_per_sample
        self.num_passive_patches_per_sample = num_passive_patches_per_sample
        self.num_iters = num_iters
        self.update_distribution_per_iteration = update_distribution_per_iteration
        self.reset_samples()

    def set_verbosity(self, is_verbose=True):
        self.VERBOSE = is_verbose

    def set_keypoints_distribution(self, x=None):
        """movability.py:75-87."""
        x = self.x if x is None else x
        assert x is not None
        if self.keypoint_predictor is not None:
            self.keypoints_distribution = self.predict_keypoints_distribution(x, power=self.keypoints_power)
        else:
            self.keypoints_distribution = None

    def sample_and_visualize_keypoints(self, x=None, sampled_keypoints=None, sampled_passive_patches=None,
                                       num_samples=32):
        """movability.py:89-125: the sampled active (red) / passive (blue) patches blended into the image."""
        if x is None:
            assert self.x is not None
            x = self.x
        if sampled_keypoints is None:
            self.set_keypoints_distribution(x)
            sampled_keypoints = self.sample_patches_from_energy(self.keypoints_distribution, num_visible=1,
                                                                num_samples=num_samples)
        img = x.clone()
        alpha = self.get_masked_pred_patches(torch.zeros_like(x), sampled_keypoints.amin(-1), fill_value=[1, 0, 0])[:, :, 0:1]
        red = torch.cat([alpha, torch.zeros_like(alpha), torch.zeros_like(alpha)], -3)
        img = img * (1 - alpha) + 0.5 * alpha * (red + img)
        if sampled_passive_patches is not None:
            alpha = self.get_masked_pred_patches(torch.zeros_like(x), sampled_passive_patches.amin(-1),
                                                 fill_value=[0, 0, 1])[:, :, 2:3]
            blue = torch.cat(2 * [torch.zeros_like(alpha)] + [alpha], -3)
            img = img * (1 - alpha) + 0.5 * alpha * (blue + img)
        return (sampled_keypoints, img)

    def _head_motion_kwargs(self, mask_head_motion, static_head_motion):
        # an unconditioned predictor (extension, see ImuConditionedFlowGenerator.__init__) has no head-motion options
        if self.head_motion_generator is None:
            return {}
        return dict(mask_head_motion=mask_head_motion, static_head_motion=static_head_motion)

    def _sample_initial_motion_map(self, x, num_samples=None, sample_batch_size=None, do_filter=True,
                                   mask_head_motion=False, static_head_motion=True, normalize=True,
                                   patch_sampling_kwargs={}, **kwargs):
        """movability.py:127-166."""
        self.set_input(x)
        if self.initialize_from_keypoints:
            self.set_keypoints_distribution()
            sampling_dist = self.keypoints_distribution
            passive_dist = 1 - self.keypoints_distribution  # TypeError without a keypoint predictor, like the reference
        else:
            sampling_dist = passive_dist = None
        flows, motion_patches, static_patches = self.sample_counterfactual_motion_map(
            x=self.x, active_sampling_distribution=sampling_dist, passive_sampling_distribution=passive_dist,
            num_active_patches=self.num_initial_active_patches, num_passive_patches=self.num_initial_passive_patches,
            num_samples=(num_samples or self.num_initial_samples),
            sample_batch_size=(sample_batch_size or self.sample_batch_size), do_filter=do_filter,
            patch_sampling_kwargs=patch_sampling_kwargs,
            **self._head_motion_kwargs(mask_head_motion, static_head_motion), **kwargs)
        motion_map = self.compute_mean_motion_map(flows, normalize_per_sample=False, normalize=normalize)
        return (motion_map, flows, motion_patches, static_patches)

    def _iterate_motion_map(self, movability_distribution, sample_passives_from_movable=True, num_active_patches=None,
                            num_passive_patches=None, num_samples=None, sample_batch_size=None, do_filter=True,
                            mask_head_motion=False, static_head_motion=True, patch_sampling_kwargs={}, normalize=True,
                            **kwargs):
        """movability.py:168-217."""
        assert self.x is not None
        if movability_distribution is None:
            movability_distribution = torch.ones_like(self.x[:, 0:1, 0])
        movability_distribution = self.compute_mean_motion_map(movability_distribution)
        movability_distribution = movability_distribution ** self.movability_power
        if sample_passives_from_movable:
            passive_distribution = movability_distribution
        else:
            passive_distribution = (1 - movability_distribution).relu()
        if self.iterate_from_keypoints:
            self.set_keypoints_distribution(self.x)
            movability_distribution *= self.keypoints_distribution
            passive_distribution *= self.keypoints_distribution
        flows, motion_patches, static_patches = self.sample_counterfactual_motion_map(
            x=self.x, active_sampling_distribution=movability_distribution,
            passive_sampling_distribution=passive_distribution,
            num_active_patches=(num_active_patches or self.num_active_patches_per_sample),
            num_passive_patches=(num_passive_patches or self.num_passive_patches_per_sample),
            num_samples=(num_samples or self.num_samples_per_iteration),
            sample_batch_size=(sample_batch_size or self.sample_batch_size), do_filter=do_filter,
            patch_sampling_kwargs=patch_sampling_kwargs,
            **self._head_motion_kwargs(mask_head_motion, static_head_motion), **kwargs)
        motion_map = self.compute_mean_motion_map(flows, normalize_per_sample=False, normalize=normalize)
        return (motion_map, flows, motion_patches, static_patches)

    def reset_samples(self):
        self.movability_maps = []
        self.flow_samples_per_iter = []
        self.active_patches_per_iter = []
        self.passive_patches_per_iter = []

    def _update_results(self, results):
        movability, flows, active_patches, passive_patches = results
        self.movability_maps.append(movability)
        self.flow_samples_per_iter.append(flows)
        self.active_patches_per_iter.append(active_patches)
        self.passive_patches_per_iter.append(passive_patches)

    def visualize_iterations(self, *args, **kwargs):
        raise NotImplementedError("the matplotlib panel (movability.py:232-281) is GUI code, out of scope here")

    def get_total_movability(self):
        """The mean motion map over the samples of all iterations so far (movability.py:283-290)."""
        if len(self.flow_samples_per_iter) == 0:
            return None
        all_flows = torch.cat(self.flow_samples_per_iter, -1)
        return self.compute_mean_motion_map(all_flows, normalize_per_sample=False, normalize=True)

    def get_minimum_movability(self):
        if len(self.flow_samples_per_iter) == 0:
            return None
        return torch.stack([self.compute_mean_motion_map(fs) for fs in self.flow_samples_per_iter], -1).amin(-1)

    def forward(self, x, initial_active_patches=None, initial_passive_patches=None, initial_sampling_distribution=None,
                num_initial_samples=None, num_samples_per_iteration=None, sample_batch_size=None, num_iters=None,
                **kwargs):
        """movability.py:299-360."""
        self.set_input(x)
        self.reset_samples()
        self.it = 0
        t0 = time()
        if initial_active_patches is not None:
            raise NotImplementedError("pass initial patches")
        results = self._sample_initial_motion_map(x=self.x, num_samples=num_initial_samples,
                                                  sample_batch_size=sample_batch_size, **kwargs)
        self._update_results(results)
        if self.VERBOSE:
            print("Completed iter %d with %d samples in %0.3f s" % (self.it, results[1].size(-1), time() - t0))
            t0 = time()
        for self.it in range(1, (num_iters or self.num_iters) + 1):
            dist = self.get_total_movability() if self.update_distribution_per_iteration else self.movability_maps[-1]
            results = self._iterate_motion_map(dist, sample_passives_from_movable=self.sample_passives_from_movable,
                                               num_samples=num_samples_per_iteration,
                                               sample_batch_size=sample_batch_size, **kwargs)
            self._update_results(results)
            if self.VERBOSE:
                print("Completed iter %d with %d samples in %0.3f s" % (self.it, results[1].size(-1), time() - t0))
                t0 = time()
        return self.movability_maps[-1]
