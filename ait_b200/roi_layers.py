"""Drop-in for the reference's `model.roi_layers` (lib/model/roi_layers/{nms,roi_align}.py): same
names, argument meaning and error behaviour, backed by libaitb200 instead of `model._C`.

    from ait_b200.roi_layers import nms, ROIAlign, roi_align
"""
import torch
from torch import nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable
from torch.nn.modules.utils import _pair

from . import ops


def nms(dets, scores, threshold):
    """Greedy NMS with the legacy +1 box convention; returns the kept ORIGINAL indices in ascending
    order as int64 -- the contract of `_C.nms` on CUDA tensors (csrc/nms.h:10-28 ->
    csrc/cuda/nms.cu:70-131: suppress when IoU > threshold, final ascending sort :127-130).
    Empty input returns an empty CPU long tensor exactly like csrc/nms.h:17-18."""
    if dets.numel() == 0:
        return torch.empty((0,), dtype=torch.int64, device="cpu")
    if not dets.is_cuda or not scores.is_cuda:
        raise RuntimeError("ait_b200.nms: dets and scores must be CUDA tensors (no CPU path)")
    if dets.dtype != torch.float32 or scores.dtype != torch.float32:
        raise RuntimeError("ait_b200.nms: float32 boxes/scores required (nms.cu is float-only)")
    if dets.dim() != 2 or dets.size(1) != 4 or scores.numel() != dets.size(0):
        raise RuntimeError("ait_b200.nms: expected dets [N,4] and scores [N]")
    n = dets.size(0)
    order = ops.topk_desc(scores.reshape(1, n), n)                       # descending, stable
    keep, n_keep, _ = ops.nms_batched(dets.reshape(1, n, 4), order, threshold, n, mode=1)
    return keep[0, : int(n_keep.item())]


class _ROIAlign(Function):
    """autograd wrapper with the reference's signature (roi_layers/roi_align.py:12-43)."""

    @staticmethod
    def forward(ctx, input, roi, output_size, spatial_scale, sampling_ratio):
        ctx.save_for_backward(roi)
        ctx.output_size = _pair(output_size)
        ctx.spatial_scale = spatial_scale
        ctx.sampling_ratio = sampling_ratio
        ctx.input_shape = input.size()
        if not input.is_cuda or not roi.is_cuda:
            raise RuntimeError("ait_b200.roi_align: input and rois must be CUDA tensors (no CPU path)")
        if input.dtype not in (torch.float32, torch.bfloat16):
            raise RuntimeError("ait_b200.roi_align: float32 or bfloat16 input required, got %s" % input.dtype)
        if roi.dim() != 2 or roi.size(1) != 5:
            raise RuntimeError("ait_b200.roi_align: rois must be [K, 5] (batch_idx, x1, y1, x2, y2)")
        b, c, h, w = input.shape
        ph, pw = ctx.output_size
        if roi.size(0) == 0:
            return input.new_empty((0, c, ph, pw))
        nhwc = ops.transpose_cs(input.reshape(b, c, h * w), to_channels_last=True).view(b, h, w, c)
        return ops.roi_align_forward(nhwc, roi, spatial_scale, ph, pw, sampling_ratio, token_major=False)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        rois, = ctx.saved_tensors
        ph, pw = ctx.output_size
        bs, ch, h, w = ctx.input_shape
        grad_input = ops.roi_align_backward(grad_output, rois, ctx.spatial_scale, ph, pw, bs, ch, h, w,
                                            ctx.sampling_ratio)
        return grad_input.to(grad_output.dtype), None, None, None, None


roi_align = _ROIAlign.apply


class ROIAlign(nn.Module):
    def __init__(self, output_size, spatial_scale, sampling_ratio):
        super().__init__()
        self.output_size = output_size
        self.spatial_scale = spatial_scale
        self.sampling_ratio = sampling_ratio

    def forward(self, input, rois):
        return roi_align(input, rois, self.output_size, self.spatial_scale, self.sampling_ratio)

    def extra_repr(self):
        return "output_size=%s, spatial_scale=%s, sampling_ratio=%s" % (
            self.output_size, self.spatial_scale, self.sampling_ratio)


__all__ = ["nms", "roi_align", "ROIAlign"]
