"""Loader for libndnet_b200.so (the C ABI declared in include/ndnet_b200.h).

There is deliberately no fallback: if the CUDA library is missing or cannot be loaded the import of any
product module fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NDNET_B200_LIB", os.path.join(os.path.dirname(HERE), "lib", "libndnet_b200.so"))

INFO_DTYPE = np.dtype([
    ("status", "<i4"), ("prune_status", "<i4"), ("evaluations", "<i4"), ("len", "<u4", 3),
    ("num_voxels", "<u4"), ("num_valid", "<u4"), ("num_kl", "<u4"), ("num_kl_after", "<u4"),
    ("num_out", "<u4"), ("num_survivors", "<u4"),
    ("voxel_size", "<f8"), ("offset", "<f8", 3), ("limits", "<f8", 6),
])
assert INFO_DTYPE.itemsize == 128

F32, F64 = 0, 1
NAN_TO_NUM = 1
LABELS_U8 = 2
TEXTBOOK_KL = 4

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python __graft_entry__.py` (or make -C ndt-net_b200/csrc). "
            "ndnet_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i, l, u = C.c_void_p, C.c_int, C.c_long, C.c_uint
    L.ndnet_b200_version.restype = C.c_char_p
    L.ndnet_b200_create.restype = i
    L.ndnet_b200_create.argtypes = [C.POINTER(vp), i]
    L.ndnet_b200_destroy.restype = None
    L.ndnet_b200_destroy.argtypes = [vp]
    L.ndnet_b200_last_error.restype = C.c_char_p
    L.ndnet_b200_last_error.argtypes = [vp]
    batch_args = [vp, vp, i, vp, i, l, i, l, u, vp, vp, vp, vp, vp, vp]
    L.ndnet_b200_downsample_batch.restype = i
    L.ndnet_b200_downsample_batch.argtypes = batch_args
    L.ndnet_b200_downsample_batch_host.restype = i
    L.ndnet_b200_downsample_batch_host.argtypes = batch_args
    L.ndnet_b200_keep_point_voxels.restype = i
    L.ndnet_b200_keep_point_voxels.argtypes = [vp, i]
    L.ndnet_b200_last_point_voxels.restype = i
    L.ndnet_b200_last_point_voxels.argtypes = [vp, vp, vp]
    L.ndnet_b200_last_kl_list.restype = l
    L.ndnet_b200_last_kl_list.argtypes = [vp, i, vp, vp, vp, l]
    L.ndnet_b200_selftest_div.restype = l
    L.ndnet_b200_selftest_div.argtypes = [l, u]
    L.ndnet_b200_launch_count.restype = l
    L.ndnet_b200_launch_count.argtypes = []
    L.ndnet_b200_stage_timing.restype = i
    L.ndnet_b200_stage_timing.argtypes = [vp, i]
    L.ndnet_b200_last_search_passes.restype = i
    L.ndnet_b200_last_search_passes.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.ndnet_b200_stage_times.restype = i
    L.ndnet_b200_stage_times.argtypes = [vp, vp, i, vp]
    L.ndnet_b200_model_create.restype = i
    L.ndnet_b200_model_create.argtypes = [vp, C.POINTER(vp), i, i, C.POINTER(C.c_char_p), C.POINTER(vp), C.POINTER(vp),
                                          C.POINTER(i)]
    L.ndnet_b200_model_input_dim.restype = i
    L.ndnet_b200_model_input_dim.argtypes = [vp]
    L.ndnet_b200_model_destroy.restype = None
    L.ndnet_b200_model_destroy.argtypes = [vp]
    L.ndnet_b200_model_forward.restype = i
    L.ndnet_b200_model_forward.argtypes = [vp, vp, vp, i, i, vp, vp]
    L.ndnet_b200_model_set_fused_head.restype = i
    L.ndnet_b200_model_set_fused_head.argtypes = [vp, i]
    L.ndnet_b200_model_tap.restype = l
    L.ndnet_b200_model_tap.argtypes = [vp, vp, C.c_char_p, vp, l, vp]
    L.ndnet_b200_test_fail_next_reserve.restype = i
    L.ndnet_b200_test_fail_next_reserve.argtypes = [vp]
    L.ndnet_b200_infer_host.restype = i
    L.ndnet_b200_infer_host.argtypes = [vp, vp, vp, i, vp, i, l, i, l, vp, l, vp]
    L.ndnet_b200_infer_host_u8.restype = i
    L.ndnet_b200_infer_host_u8.argtypes = [vp, vp, vp, i, vp, i, l, i, l, vp, l, vp]
    L.ndnet_b200_infer_host_async.restype = i
    L.ndnet_b200_infer_host_async.argtypes = [vp, vp, vp, i, vp, i, i, l, i, l, vp, l, vp]
    L.ndnet_b200_infer_wait.restype = i
    L.ndnet_b200_infer_wait.argtypes = [vp, vp]
    L.ndnet_b200_infer_device.restype = i
    L.ndnet_b200_infer_device.argtypes = [vp, vp, vp, i, vp, i, l, i, l, vp, l, vp]
    L.ndnet_b200_set_pipeline.restype = i
    L.ndnet_b200_set_pipeline.argtypes = [vp, i, i]
    L.ndnet_b200_set_device_chunk.restype = i
    L.ndnet_b200_set_device_chunk.argtypes = [vp, i]
    L.ndnet_b200_set_stagger.restype = i
    L.ndnet_b200_set_stagger.argtypes = [vp, i]
    L.ndnet_b200_set_ndt_graph.restype = i
    L.ndnet_b200_set_ndt_graph.argtypes = [vp, i]
    L.ndnet_b200_ply_load.restype = i
    L.ndnet_b200_ply_load.argtypes = [i, vp, C.c_size_t, i, i, i, vp, C.POINTER(vp), C.POINTER(C.c_ulong), C.POINTER(l),
                                      C.POINTER(l)]
    L.ndnet_b200_ply_num_points.restype = l
    L.ndnet_b200_ply_num_points.argtypes = [vp]
    L.ndnet_b200_ply_sample.restype = i
    L.ndnet_b200_ply_sample.argtypes = [vp, vp, C.c_size_t, i, vp, vp, vp, vp]
    L.ndnet_b200_ply_free.restype = None
    L.ndnet_b200_ply_free.argtypes = [vp]
    L.ndnet_b200_trainer_create.restype = i
    L.ndnet_b200_trainer_create.argtypes = [i, i, C.POINTER(C.c_char_p), C.POINTER(vp), C.POINTER(i), C.POINTER(vp)]
    L.ndnet_b200_keep_kl_list.restype = i
    L.ndnet_b200_keep_kl_list.argtypes = [vp, i]
    L.ndnet_b200_onehot_to_labels.restype = i
    L.ndnet_b200_onehot_to_labels.argtypes = [vp, l, i, vp, vp]
    L.ndnet_b200_labels_to_onehot.restype = i
    L.ndnet_b200_labels_to_onehot.argtypes = [vp, l, i, vp, vp]
    L.ndnet_b200_trainer_info.restype = i
    L.ndnet_b200_trainer_info.argtypes = [vp, C.POINTER(i), C.POINTER(i), C.POINTER(i)]
    L.ndnet_b200_trainer_forward.restype = i
    L.ndnet_b200_trainer_forward.argtypes = [vp, vp, i, i, C.POINTER(vp), vp, i, vp]
    L.ndnet_b200_trainer_backward.restype = i
    L.ndnet_b200_trainer_backward.argtypes = [vp, vp, C.POINTER(vp), C.POINTER(vp), vp]
    L.ndnet_b200_trainer_backward_flat.restype = i
    L.ndnet_b200_trainer_backward_flat.argtypes = [vp, vp, C.POINTER(vp), vp, vp]
    L.ndnet_b200_trainer_grad_layout.restype = l
    L.ndnet_b200_trainer_grad_layout.argtypes = [vp, vp, i]
    L.ndnet_b200_trainer_num_buckets.restype = i
    L.ndnet_b200_trainer_num_buckets.argtypes = [vp]
    L.ndnet_b200_trainer_bucket_range.restype = i
    L.ndnet_b200_trainer_bucket_range.argtypes = [vp, i, C.POINTER(l), C.POINTER(l)]
    L.ndnet_b200_trainer_set_deferred_copy.restype = i
    L.ndnet_b200_trainer_set_deferred_copy.argtypes = [vp, i]
    L.ndnet_b200_trainer_bucket_ready.restype = i
    L.ndnet_b200_trainer_bucket_ready.argtypes = [vp, i, vp, vp]
    L.ndnet_b200_trainer_set_graph.restype = i
    L.ndnet_b200_trainer_set_graph.argtypes = [vp, i]
    L.ndnet_b200_trainer_set_precision.restype = i
    L.ndnet_b200_trainer_set_precision.argtypes = [vp, i]
    L.ndnet_b200_debug_train_gemm.restype = i
    L.ndnet_b200_debug_train_gemm.argtypes = [i, vp, l, vp, l, vp, l, i, i, i, vp, i, vp]
    L.ndnet_b200_trainer_last_error.restype = C.c_char_p
    L.ndnet_b200_trainer_last_error.argtypes = [vp]
    L.ndnet_b200_trainer_debug_buffer.restype = l
    L.ndnet_b200_trainer_debug_buffer.argtypes = [vp, C.c_char_p, vp, vp]
    L.ndnet_b200_trainer_destroy.restype = None
    L.ndnet_b200_trainer_destroy.argtypes = [vp]
    _lib = L
    return L


EXPORTED = [
    "ndt_downsample", "prune_nds", "to_point_cloud", "free_nds", "free_kl_divergences", "print_matrix",
    "ndnet_b200_create", "ndnet_b200_destroy", "ndnet_b200_last_error", "ndnet_b200_version",
    "ndnet_b200_downsample_batch", "ndnet_b200_downsample_batch_host", "ndnet_b200_keep_point_voxels", "ndnet_b200_last_point_voxels",
    "ndnet_b200_last_kl_list", "ndnet_b200_selftest_div", "ndnet_b200_launch_count", "ndnet_b200_stage_timing", "ndnet_b200_stage_times",
    "ndnet_b200_last_search_passes",
    "ndnet_b200_model_create", "ndnet_b200_model_input_dim", "ndnet_b200_model_destroy", "ndnet_b200_model_forward",
    "ndnet_b200_model_tap", "ndnet_b200_model_set_fused_head", "ndnet_b200_test_fail_next_reserve",
    "ndnet_b200_infer_host", "ndnet_b200_infer_host_u8", "ndnet_b200_infer_host_async", "ndnet_b200_infer_wait", "ndnet_b200_infer_device", "ndnet_b200_set_pipeline", "ndnet_b200_set_device_chunk", "ndnet_b200_set_stagger", "ndnet_b200_set_ndt_graph",
    "ndnet_b200_ply_load", "ndnet_b200_ply_num_points", "ndnet_b200_ply_sample", "ndnet_b200_ply_free",
    "ndnet_b200_keep_kl_list", "ndnet_b200_onehot_to_labels", "ndnet_b200_labels_to_onehot", "ndnet_b200_trainer_create", "ndnet_b200_trainer_info", "ndnet_b200_trainer_forward", "ndnet_b200_trainer_backward", "ndnet_b200_trainer_last_error",
    "ndnet_b200_trainer_backward_flat", "ndnet_b200_trainer_grad_layout", "ndnet_b200_trainer_set_graph",
    "ndnet_b200_trainer_num_buckets", "ndnet_b200_trainer_bucket_range", "ndnet_b200_trainer_set_deferred_copy", "ndnet_b200_trainer_bucket_ready",
    "ndnet_b200_trainer_set_precision", "ndnet_b200_debug_train_gemm", "ndnet_b200_trainer_debug_buffer", "ndnet_b200_trainer_destroy",
]
