"""ctypes binding of ``libcwm_b200.so`` (the C ABI declared in ``include/cwm_b200.h``).

The product path has no CPU fallback: if the shared library is missing, or the device is not sm_100, every op
raises.  The library is built in-tree by ``__graft_entry__.build()`` / ``make -C csrc``.
"""
import ctypes
import os
from ctypes import (POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcwm_b200.so")

CWM_OK = 0
EPI_F16, EPI_GELU_F16, EPI_RES_F32, EPI_F32 = 0, 1, 2, 3

# every symbol include/cwm_b200.h declares (checked by tests/test_abi.py)
ABI_VERSION = 7   # include/cwm_b200.h CWM_B200_ABI_VERSION

EXPORTED_SYMBOLS = (
    "cwm_abi_version", "cwm_last_error", "cwm_device_check", "cwm_compact_mask", "cwm_patch_gather",
    "cwm_layernorm_f16", "cwm_gemm_f16", "cwm_attention_f16", "cwm_fill_mask_tokens",
    "cwm_unpatchify_scatter", "cwm_vmae_workspace_bytes", "cwm_vmae_forward", "cwm_last_forward_launches",
    "cwm_profile_begin", "cwm_profile_end", "cwm_attention_generic_workspace_bytes", "cwm_attention_generic_f16",
    "cwm_fill_pad_rows", "cwm_block_workspace_bytes", "cwm_block_forward", "cwm_cross_block_workspace_bytes",
    "cwm_cross_block_forward", "cwm_launch_count_reset", "cwm_vmae_forward_cf", "cwm_cf_shift_masks",
    "cwm_cf_build_videos", "cwm_cf_make_static", "cwm_patch_gather_cf", "cwm_unpatchify_scatter_cf",
    "cwm_flow_sample_stats", "cwm_flow_filter_mask", "cwm_flow_zero_filtered", "cwm_flow_magnitude_sum",
    "cwm_motion_map_finalize", "cwm_flow_stats_workspace_bytes", "cwm_flow_corrs_workspace_bytes", "cwm_flow_corrs",
    "cwm_gemm_ln_parts", "cwm_rowstats_f16", "cwm_raft_corr_pyramid", "cwm_raft_corr_lookup", "cwm_raft_upsample_flow",
    "cwm_raft_corr_lookup_f16", "cwm_raft_bias_act_f16", "cwm_raft_gru_gate_f16", "cwm_raft_gru_update_f16",
    "cwm_raft_flow_update", "cwm_total_launches",
    "cwm_conv2d_weight_k", "cwm_conv2d_f16", "cwm_raft_im2col_flow", "cwm_conv2d_gru_gate_f16",
    "cwm_conv2d_gru_update_f16",
    "cwm_raft_corr_tc_workspace_bytes", "cwm_raft_corr_volume_tc", "cwm_raft_corr_pyramid_tc",
    "cwm_instnorm_workspace_bytes", "cwm_instnorm_f16", "cwm_conv2d_strided_f16", "cwm_im2col_nchw_f16", "cwm_add_act_f16",
    "cwm_conv2d_dual_f16", "cwm_raft_flow_update_taps", "cwm_act_dtype",
    "cwm_raft_corr_volume_rows_f16", "cwm_raft_corr_pyramid_rows_f16", "cwm_raft_corr_volume_rows_f16_out16",
    "cwm_raft_corr_pyramid_rows_f16_pyr16", "cwm_raft_corr_lookup_f16_pyr16",
    "cwm_philox4x32_10", "cwm_mask_uniform", "cwm_mask_energy_table", "cwm_mask_energy_sample",
    "cwm_mask_rectangularize_workspace_bytes", "cwm_mask_rectangularize",
    # tuning hooks (header section "tuning hooks")
    "cwm_debug_attention_poly", "cwm_debug_attention_war_safe", "cwm_debug_attention_persistent",
    "cwm_debug_attention_persist_map", "cwm_debug_attention_stale_max", "cwm_debug_attention_skip_idle", "cwm_debug_attn_mma_wide", "cwm_debug_attn_mma_split", "cwm_debug_gemm_cta2",
)


class CwmError(RuntimeError):
    """A libcwm_b200 entry point returned a negative status."""


class GemmEpilogue(Structure):
    _fields_ = [
        ("mode", c_int32), ("bias", c_void_p), ("scale", c_float), ("scale_cols", c_int32),
        ("res", c_void_p), ("ldr", c_int32), ("res_gather", c_void_p), ("gather_stride", c_int32),
        ("grp_rows", c_int32), ("grp_out_stride", c_int32), ("out", c_void_p), ("ldo", c_int32),
        # LayerNorm fusion (optional; zero = off)
        ("ln_x16", c_void_p), ("ln_ldx16", c_int32), ("ln_stats_out", c_void_p), ("ln_stats_in", c_void_p),
        ("ln_parts", c_int32), ("ln_colsum", c_void_p), ("ln_width", c_int32), ("ln_eps", c_float),
    ]


class BlockWeights(Structure):
    _fields_ = [(n, c_void_p) for n in (
        "ln1_g", "ln1_b", "w_qkv", "b_qkv", "w_proj", "b_proj", "ln2_g", "ln2_b", "w_fc1", "b_fc1", "w_fc2",
        "b_fc2", "w_qkv_ln", "s_qkv", "c_qkv", "w_fc1_ln", "s_fc1", "c_fc1")]


class CrossBlockWeights(Structure):
    _fields_ = [(n, c_void_p) for n in (
        "ln1_g", "ln1_b", "ln1s_g", "ln1s_b", "w_qkv", "w_qkv_s", "w_proj", "b_proj", "w_proj_s", "b_proj_s",
        "ln2_g", "ln2_b", "ln2s_g", "ln2s_b", "w_fc1", "b_fc1", "w_fc2", "b_fc2", "w_fc1_s", "b_fc1_s", "w_fc2_s",
        "b_fc2_s")]


class VmaeModel(Structure):
    _fields_ = (
        [(n, c_int32) for n in ("in_chans", "num_frames", "img_h", "img_w", "pt", "ph", "pw", "enc_dim",
                                "enc_depth", "enc_heads", "enc_hidden", "dec_dim", "dec_depth", "dec_heads",
                                "dec_hidden", "out_dim")]
        + [("ln_eps", c_float), ("enc_qk_scale", c_float), ("dec_qk_scale", c_float)]
        + [("w_patch", c_void_p), ("b_patch", c_void_p), ("pos_enc", c_void_p),
           ("enc_blocks", POINTER(BlockWeights)), ("enc_norm_g", c_void_p), ("enc_norm_b", c_void_p),
           ("w_e2d", c_void_p), ("mask_token", c_void_p), ("pos_dec", c_void_p),
           ("dec_blocks", POINTER(BlockWeights)), ("dec_norm_g", c_void_p), ("dec_norm_b", c_void_p),
           ("w_head", c_void_p), ("b_head", c_void_p)]
    )


class CfSource(Structure):
    """cwm_cf_source: the virtual counterfactual video (include/cwm_b200.h, section 8(f) rank 1)."""
    _fields_ = [("x", c_void_p), ("xs", c_int64 * 5), ("sample_image", c_void_p), ("shift_px", c_void_p),
                ("shifted_active", c_void_p), ("frame", c_int32), ("static_frame", c_int32)]


class ProfileEntry(Structure):
    _fields_ = [("name", ctypes.c_char * 32), ("launches", c_int32), ("ms", ctypes.c_double),
                ("flops", ctypes.c_double), ("bytes", ctypes.c_double)]


_lib = None


def _declare(lib):
    i64x5 = POINTER(c_int64)
    lib.cwm_abi_version.restype = c_int
    lib.cwm_last_error.restype = c_char_p
    lib.cwm_device_check.restype = c_int
    lib.cwm_compact_mask.argtypes = [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.cwm_patch_gather.argtypes = [c_void_p, i64x5, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                     c_void_p, c_int, c_int, POINTER(c_float), POINTER(c_float), c_void_p,
                                     c_void_p]
    lib.cwm_layernorm_f16.argtypes = [c_void_p, c_int, c_int, c_void_p, c_void_p, c_float, c_int, c_int, c_int,
                                      c_void_p, c_void_p]
    lib.cwm_gemm_f16.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, POINTER(GemmEpilogue), c_void_p]
    lib.cwm_gemm_ln_parts.argtypes = [c_int]
    lib.cwm_rowstats_f16.argtypes = [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]
    lib.cwm_attention_f16.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]
    lib.cwm_fill_mask_tokens.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                         c_void_p]
    lib.cwm_unpatchify_scatter.argtypes = [c_void_p, c_void_p, i64x5, c_void_p, c_int, c_int, c_int, c_int,
                                           c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]
    lib.cwm_vmae_workspace_bytes.argtypes = [POINTER(VmaeModel), c_int, c_int]
    lib.cwm_vmae_workspace_bytes.restype = c_size_t
    lib.cwm_vmae_forward.argtypes = [POINTER(VmaeModel), c_void_p, i64x5, c_int, POINTER(c_float),
                                     POINTER(c_float), c_void_p, c_int, c_void_p, c_void_p, c_size_t, c_void_p]
    lib.cwm_last_forward_launches.restype = c_int
    lib.cwm_launch_count_reset.restype = c_int
    lib.cwm_total_launches.restype = ctypes.c_longlong
    for name in ("cwm_debug_attention_poly", "cwm_debug_attention_war_safe", "cwm_debug_attention_persistent",
                 "cwm_debug_attention_persist_map", "cwm_debug_attention_stale_max", "cwm_debug_attention_skip_idle", "cwm_debug_attn_mma_wide", "cwm_debug_attn_mma_split"):
        getattr(lib, name).argtypes = [c_int]
        getattr(lib, name).restype = None
    lib.cwm_debug_gemm_cta2.argtypes = [c_int]
    lib.cwm_attention_generic_workspace_bytes.argtypes = [c_int] * 5
    lib.cwm_attention_generic_workspace_bytes.restype = c_size_t
    lib.cwm_attention_generic_f16.argtypes = [c_void_p, c_void_p, c_void_p] + [c_int] * 11 + [c_void_p, c_int, c_void_p,
                                                                                           c_size_t, c_void_p]
    lib.cwm_fill_pad_rows.argtypes = [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]
    lib.cwm_block_workspace_bytes.argtypes = [c_int] * 6
    lib.cwm_block_workspace_bytes.restype = c_size_t
    lib.cwm_block_forward.argtypes = [POINTER(BlockWeights), c_void_p] + [c_int] * 6 + [c_float, c_float, c_void_p,
                                                                                      c_size_t, c_void_p]
    lib.cwm_cross_block_workspace_bytes.argtypes = [c_int] * 9
    lib.cwm_cross_block_workspace_bytes.restype = c_size_t
    lib.cwm_cross_block_forward.argtypes = ([POINTER(CrossBlockWeights), c_void_p, c_void_p] + [c_int] * 9 +
                                            [c_float, c_float, c_void_p, c_size_t, c_void_p])
    lib.cwm_vmae_forward_cf.argtypes = [POINTER(VmaeModel), POINTER(CfSource), c_int, POINTER(c_float),
                                        POINTER(c_float), c_void_p, c_int, c_void_p, c_void_p, c_size_t, c_void_p]
    lib.cwm_cf_shift_masks.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                       c_void_p, c_void_p]
    lib.cwm_cf_build_videos.argtypes = [POINTER(CfSource)] + [c_int] * 7 + [c_void_p, c_void_p]
    lib.cwm_cf_make_static.argtypes = [c_void_p, i64x5, c_void_p] + [c_int] * 7 + [c_void_p, c_void_p]
    lib.cwm_patch_gather_cf.argtypes = [POINTER(CfSource)] + [c_int] * 8 + [c_void_p, c_int, c_int, POINTER(c_float),
                                                                          POINTER(c_float), c_void_p, c_void_p]
    lib.cwm_unpatchify_scatter_cf.argtypes = [c_void_p, POINTER(CfSource), c_void_p] + [c_int] * 9 + [c_void_p,
                                                                                                   c_void_p]
    i64x3 = POINTER(c_int64)
    lib.cwm_flow_corrs_workspace_bytes.argtypes = [c_int] * 5
    lib.cwm_flow_corrs_workspace_bytes.restype = c_size_t
    lib.cwm_flow_corrs.argtypes = [c_void_p, i64x5, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                   c_size_t, c_void_p]
    lib.cwm_flow_stats_workspace_bytes.argtypes = [c_int] * 4
    lib.cwm_flow_stats_workspace_bytes.restype = c_size_t
    lib.cwm_flow_sample_stats.argtypes = [c_void_p, i64x5, c_int, c_int, c_int, c_int, c_void_p, i64x3, c_int, c_int,
                                          c_float, c_void_p, c_void_p, c_size_t, c_void_p]
    lib.cwm_flow_filter_mask.argtypes = [c_void_p, c_int, c_int, c_int, c_float, c_float, c_float, c_void_p, c_void_p]
    lib.cwm_flow_zero_filtered.argtypes = [c_void_p, i64x5, c_int, c_int, c_int, c_int, c_void_p, c_void_p]
    lib.cwm_flow_magnitude_sum.argtypes = [c_void_p, i64x5, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int,
                                           c_float, c_int, c_void_p, c_void_p, c_size_t, c_void_p]
    lib.cwm_motion_map_finalize.argtypes = [c_void_p, c_int, c_int, c_int, c_float, c_int, c_float, c_void_p, c_void_p]
    lib.cwm_raft_corr_pyramid.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, POINTER(c_void_p),
                                          c_void_p]
    lib.cwm_raft_corr_tc_workspace_bytes.argtypes = [c_int] * 4
    lib.cwm_raft_corr_tc_workspace_bytes.restype = c_size_t
    lib.cwm_raft_corr_volume_tc.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t,
                                            c_void_p]
    lib.cwm_raft_corr_pyramid_tc.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, POINTER(c_void_p),
                                             c_void_p, c_size_t, c_void_p]
    lib.cwm_raft_corr_lookup.argtypes = [POINTER(c_void_p), c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p,
                                         c_void_p]
    lib.cwm_raft_upsample_flow.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]
    c_ll = ctypes.c_longlong
    lib.cwm_raft_corr_lookup_f16.argtypes = [POINTER(c_void_p), c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_int,
                                             c_void_p]
    lib.cwm_raft_bias_act_f16.argtypes = [c_void_p, c_int, c_void_p, c_int, c_int, c_ll, c_void_p, c_int, c_void_p, c_int,
                                          c_void_p, c_int, c_int, c_void_p]
    lib.cwm_raft_gru_gate_f16.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_ll, c_void_p, c_void_p, c_int,
                                          c_void_p]
    lib.cwm_raft_gru_update_f16.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_ll, c_void_p, c_void_p]
    lib.cwm_raft_flow_update.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]
    lib.cwm_conv2d_weight_k.argtypes = [c_int] * 3
    lib.cwm_conv2d_f16.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                   c_void_p, c_int, c_void_p, c_int, c_void_p]
    lib.cwm_conv2d_gru_gate_f16.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int,
                                            c_int, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p]
    lib.cwm_conv2d_gru_update_f16.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int,
                                              c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p]
    lib.cwm_conv2d_strided_f16.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int,
                                           c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p]
    lib.cwm_im2col_nchw_f16.argtypes = [c_void_p, c_ll, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_float,
                                        c_void_p, c_int, c_void_p]
    lib.cwm_add_act_f16.argtypes = [c_void_p, c_void_p, c_ll, c_int, c_void_p, c_void_p]
    lib.cwm_conv2d_dual_f16.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                        c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p]
    lib.cwm_raft_corr_volume_rows_f16.argtypes = [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]
    lib.cwm_raft_corr_pyramid_rows_f16.argtypes = [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                                   POINTER(c_void_p), c_void_p]
    lib.cwm_raft_corr_volume_rows_f16_out16.argtypes = [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]
    lib.cwm_raft_corr_pyramid_rows_f16_pyr16.argtypes = [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                                         POINTER(c_void_p), c_void_p]
    lib.cwm_raft_corr_lookup_f16_pyr16.argtypes = [POINTER(c_void_p), c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p,
                                                   c_int, c_void_p]
    lib.cwm_raft_flow_update_taps.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int,
                                              c_void_p, c_int, c_void_p]
    lib.cwm_raft_im2col_flow.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p]
    lib.cwm_instnorm_workspace_bytes.argtypes = [c_int, c_int]
    lib.cwm_instnorm_workspace_bytes.restype = c_size_t
    lib.cwm_instnorm_f16.argtypes = [c_void_p, c_int, c_int, c_int, c_float, c_int, c_void_p, c_int, c_void_p, c_void_p,
                                     c_size_t, c_void_p]
    c_u64, c_u32p = ctypes.c_uint64, POINTER(ctypes.c_uint32)
    lib.cwm_philox4x32_10.argtypes = [c_u32p, c_u32p, c_u32p]
    lib.cwm_mask_uniform.argtypes = [c_u64] + [c_int] * 8 + [c_void_p, c_void_p]
    lib.cwm_mask_energy_table.argtypes = [c_void_p, c_int, c_int, c_float, c_void_p, c_void_p]
    lib.cwm_mask_energy_sample.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_u64, c_int, c_int, c_int, c_int,
                                           c_void_p, c_void_p]
    lib.cwm_mask_rectangularize_workspace_bytes.argtypes = [c_int]
    lib.cwm_mask_rectangularize_workspace_bytes.restype = c_size_t
    lib.cwm_mask_rectangularize.argtypes = [c_void_p, c_int, c_int, c_int, c_u64, c_int, c_void_p, c_size_t, c_void_p]
    lib.cwm_profile_begin.restype = c_int
    lib.cwm_profile_end.argtypes = [POINTER(ProfileEntry), c_int, POINTER(c_int)]
    for name in EXPORTED_SYMBOLS:   # a missing export fails here, at load time, not at first use
        getattr(lib, name)
    return lib


def lib_path():
    """The build this process uses: ``CWM_DTYPE=bf16`` selects the bf16-operand twin (same sources, -DCWM_ACT_BF16); the
    default is the f16 build, the only one inside the parity bar."""
    mode = os.environ.get("CWM_DTYPE", "f16").lower()
    if mode not in ("f16", "fp16", "bf16"):
        raise CwmError(f"CWM_DTYPE={mode!r}: expected f16 or bf16")
    return LIB_PATH.replace("libcwm_b200.so", "libcwm_b200_bf16.so") if mode == "bf16" else LIB_PATH


def load():
    """Loads (once) and returns the ctypes handle.  Raises if the library has not been built."""
    global _lib
    if _lib is None:
        path = lib_path()
        if not os.path.exists(path):
            raise CwmError(
                f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU or PyTorch fallback for the CWM VMAE path)")
        import torch  # noqa: F401  -- makes sure libcudart.so.12 is already mapped (same SONAME is reused)
        _lib = _declare(ctypes.CDLL(path, mode=ctypes.RTLD_GLOBAL))
    return _lib


def act_dtype():
    """torch dtype of every 16-bit activation / weight tensor handed to the library (f16, or bf16 in the bf16 build)."""
    import torch
    return torch.bfloat16 if load().cwm_act_dtype() == 1 else torch.float16


def require_f16(what):
    """Paths that exist for f16 operands only (RAFT, the conjoined models) refuse the bf16 build instead of mis-reading it."""
    if load().cwm_act_dtype() != 0:
        raise CwmError(f"{what} is available in the f16 build only (unset CWM_DTYPE=bf16)")


def check(rc):
    if rc != CWM_OK:
        msg = load().cwm_last_error()
        raise CwmError(f"libcwm_b200 error {rc}: {msg.decode() if msg else '?'}")


def profile_begin():
    check(load().cwm_profile_begin())


def profile_end():
    """-> list of dict(name, launches, ms, flops, bytes), one per kernel class, in first-launch order."""
    arr = (ProfileEntry * 32)()
    n = c_int(0)
    check(load().cwm_profile_end(arr, 32, ctypes.byref(n)))
    return [dict(name=arr[i].name.decode(), launches=arr[i].launches, ms=arr[i].ms, flops=arr[i].flops,
                 bytes=arr[i].bytes) for i in range(n.value)]


def strides5(t):
    """Element strides of a 5-D tensor as the ``int64_t[5]`` the C ABI wants."""
    assert t.dim() == 5, t.shape
    return (c_int64 * 5)(*t.stride())


def float_array(vals):
    if vals is None:
        return None
    return (c_float * len(vals))(*[float(v) for v in vals])
