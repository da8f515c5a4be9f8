"""ctypes binding of the C-ABI library (include/ptb200.h).  There is NO fallback: if the
library is missing or a call fails the error is raised immediately."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "_C", "libptb200.so")

c_void_p, c_int, c_ll, c_float = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float
c_fp = ctypes.POINTER(ctypes.c_float)

# name -> (restype, argtypes); must list every symbol include/ptb200.h declares
SIGNATURES = {
    "pt_last_error": (ctypes.c_char_p, []),
    "pt_abi_version": (c_int, []),
    "pt_build_arch": (ctypes.c_char_p, []),
    "pt_bag_gen": (c_int, [c_void_p, c_ll, c_void_p, c_int, c_fp, c_int, c_fp, c_int, c_float, c_void_p,
                           c_void_p, c_int, c_void_p]),
    "pt_make_rois": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "pt_neg_weight": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p]),
    "pt_box_iou_rotated": (c_int, [c_void_p, c_int, c_void_p, c_int, c_ll, c_ll, c_int, c_int, c_int, c_void_p,
                                   c_void_p]),
    "pt_bbox_overlaps": (c_int, [c_void_p, c_int, c_void_p, c_int, c_ll, c_ll, c_int, c_int, c_float, c_void_p,
                                 c_void_p]),
    "pt_nchw_to_nhwc": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "pt_nchw_to_nhwc_ex": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "pt_roi_align_forward": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_ll, c_int, c_int, c_int, c_int, c_int,
                                     c_int, c_int, c_float, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "pt_map_roi_levels": (c_int, [c_void_p, c_int, c_int, c_float, c_int, c_void_p, c_void_p]),
    "pt_roi_rescale": (c_int, [c_void_p, c_int, c_int, c_float, c_float, c_void_p, c_void_p]),
    "pt_fc_gemm_workspace_bytes": (c_ll, [c_int]),
    "pt_fc_gemm_bf16": (c_int, [c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_void_p, c_ll, c_int, c_int, c_int,
                                c_int, c_int, c_void_p, c_ll, c_int, c_int, c_void_p]),
    "pt_prep_fc1_weight": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_ll, c_int, c_void_p]),
    "pt_cast_weight_bf16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_ll, c_int, c_void_p]),
    "pt_reg_decode": (c_int, [c_void_p, c_int, c_ll, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_int, c_int, c_float, c_float, c_float, c_float, c_float, c_void_p,
                              c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "pt_small_heads_bf16": (c_int, [c_void_p, c_ll, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int,
                                    c_void_p, c_void_p, c_void_p]),
    "pt_cls_ins_heads": (c_int, [c_void_p, c_int, c_ll, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                 c_void_p, c_void_p, c_void_p]),
    "pt_score_select": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_int, c_void_p]),
    "pt_neg_loss": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "pt_finalize_losses": (c_int, [c_void_p, c_int, c_int, c_float, c_float, c_float, c_float, c_void_p, c_void_p]),
    "pt_split_bf16x3": (c_int, [c_void_p, c_void_p, c_ll, c_int, c_void_p]),
    "pt_focal_cost_table": (c_int, [c_void_p, c_ll, c_float, c_float, c_float, c_float, c_void_p, c_void_p]),
    "pt_topk_pre": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_float, c_int, c_void_p, c_void_p,
                            c_void_p, c_void_p]),
    "pt_topk_second": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int,
                               c_void_p, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pt_bbox_metric": (c_int, [c_void_p, c_int, c_void_p, c_int, c_ll, c_ll, c_int, c_int, c_float, c_void_p, c_void_p]),
    "pt_max_iou_assign": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_float, c_float,
                                  c_float, c_float, c_float, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p]),
    "pt_decode_ltrb": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "pt_pseudo_aggregate": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int,
                                    c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_void_p]),
    "pt_ltrb_targets": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                c_void_p]),
    "pt_nms_rotated_workspace_bytes": (c_ll, [c_int]),
    "pt_nms_rotated": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p, c_ll,
                               c_void_p]),
    "pt_black_paper_select": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_void_p]),
    "pt_black_paper_select_ex": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_void_p, c_void_p]),
    "pt_fill_polys": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p]),
    "pt_fc_gemm_bf16_ex": (c_int, [c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_void_p, c_ll, c_int, c_int, c_int, c_int,
                                   c_int, c_void_p, c_ll, c_void_p, c_ll, c_int, c_int, c_void_p]),
    "pt_fc_gemm_bf16_mn": (c_int, [c_void_p, c_ll, c_int, c_void_p, c_ll, c_int, c_void_p, c_void_p, c_ll, c_int, c_int,
                                   c_int, c_int, c_int, c_void_p, c_ll, c_void_p, c_ll, c_int, c_int, c_void_p]),
    "pt_reg_loss_grad": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_float, c_float,
                                 c_float, c_float, c_void_p, c_void_p, c_float, c_void_p, c_void_p]),
    "pt_reg_loss_grad_ex": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_float, c_float,
                                    c_float, c_float, c_void_p, c_void_p, c_float, c_void_p, c_int, c_void_p]),
    "pt_roi_align_rotated_backward": (c_int, [c_void_p, c_ll, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_int,
                                              c_int, c_int, c_void_p, c_void_p]),
    "pt_bag_loss_grad": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int,
                                 c_void_p, c_void_p, c_float, c_float, c_void_p, c_void_p]),
    "pt_head_bwd": (c_int, [c_void_p, c_int, c_void_p, c_ll, c_int, c_void_p, c_int, c_void_p, c_ll, c_void_p, c_void_p,
                            c_void_p]),
    "pt_transpose_pad_bf16": (c_int, [c_void_p, c_ll, c_int, c_int, c_void_p, c_ll, c_void_p]),
    "pt_unpermute_dw1": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "pt_colsum_bf16": (c_int, [c_void_p, c_ll, c_int, c_int, c_void_p, c_void_p]),
    "pt_nhwc_to_nchw_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "pt_roi_align_backward": (c_int, [c_void_p, c_ll, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_int, c_int,
                                      c_void_p, c_void_p]),
    "pt_roi_align_backward_ex": (c_int, [c_void_p, c_ll, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_int,
                                         c_int, c_void_p, c_void_p, c_int, c_void_p]),
    "pt_roi_align_rotated_backward_ex": (c_int, [c_void_p, c_ll, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float,
                                                 c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    "pt_augment_param_stride": (c_int, []),
    "pt_augment_image": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "pt_augment_coords": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                                  c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p]),
    "pt_sigmoid_focal_loss": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_float, c_float, c_int, c_int, c_void_p,
                                      c_void_p, c_void_p, c_void_p]),
    "pt_rotated_iou_loss": (c_int, [c_void_p, c_void_p, c_int, c_int, c_float, c_int, c_float, c_void_p, c_void_p,
                                    c_void_p]),
    "pt_aligned_iou_mean": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
}

_LIB = None


class PTB200Error(RuntimeError):
    pass


def load():
    """Load libptb200.so; raises if it has not been built (python -m point_teacher_b200.build)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(SO_PATH):
        raise PTB200Error(
            f"{SO_PATH} is missing: build it with `python -m point_teacher_b200.build` "
            "(nvcc, sm_100a).  There is no CPU / PyTorch fallback for this path.")
    lib = ctypes.CDLL(SO_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


LAUNCHES = {"count": 0}   # kernels enqueued through the C-ABI (every launching entry point = one kernel)
_NO_KERNEL = {"pt_last_error", "pt_abi_version", "pt_build_arch", "pt_fc_gemm_workspace_bytes",
              "pt_nms_rotated_workspace_bytes", "pt_augment_param_stride"}


# optional per-call device timing (bench.py's in-step breakdown): CUDA events on the launching stream around every
# C-ABI call; only meaningful while the host runs ahead of the device (the bench parks the device behind a sleep first)
TRACE = {"on": False, "events": []}


def call(name, *args):
    lib = load()
    if TRACE["on"] and name not in _NO_KERNEL:
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib, name)(*args)
        e1.record()
        TRACE["events"].append((name, e0, e1))
    else:
        rc = getattr(lib, name)(*args)
    if name not in _NO_KERNEL:
        LAUNCHES["count"] += 1
    if rc != 0:
        raise PTB200Error(f"{name} failed (code {rc}): {lib.pt_last_error().decode()}")
    return rc
