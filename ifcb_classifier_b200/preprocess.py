"""Host wrapper of the fused ROI preprocess kernel (K1, ``ifcb_preprocess``).

Replaces the per-item transform chain of ``IfcbBinDataset.__getitem__``
(reference neuston_data.py:456-464) for all ROIs of a bin in one launch.
"""
import ctypes as C
import torch

from . import _lib

OUT_F32_NCHW, OUT_BF16_NCHW, OUT_U8_GRAY = _lib.IFCB_OUT_F32_NCHW, _lib.IFCB_OUT_BF16_NCHW, _lib.IFCB_OUT_U8_GRAY
PASS_PILLOW12, PASS_HV = _lib.IFCB_PASS_PILLOW12, _lib.IFCB_PASS_HV
FRAME_H, FRAME_W = 1034, 1380        # IFCB camera frame: upper bound of any ROI


def parse_imgnorm(img_norm_arg):
    """``--img-norm MEAN STD`` -> (mean[3], std[3]); mirrors neuston_data.py:331-339."""
    mean = [float(m) for m in img_norm_arg[0].split(',')]
    if len(mean) == 1:
        mean = 3 * mean
    std = [float(s) for s in img_norm_arg[1].split(',')]
    if len(std) == 1:
        std = 3 * std
    assert len(mean) == len(std) == 3, '--img-norm invalid: {}'.format(img_norm_arg)
    return mean, std


def preprocess_rois(packed, offsets, heights, widths, R, img_norm=None, out_mode=OUT_F32_NCHW,
                    out=None, max_h=FRAME_H, max_w=FRAME_W, pass_rule=PASS_PILLOW12, status=None):
    """Resizes/normalises ``n`` ROIs on the GPU.

    packed   uint8 cuda tensor: raw ``.roi`` bytes
    offsets  int64 cuda tensor [n] (START_BYTE), heights/widths int32 cuda tensors [n]
    img_norm None, the reference's ``[str, str]``, or ``(mean[3], std[3])``
    status   optional int32 cuda tensor [1]: flags of ROIs the kernel refused (table entry outside the packed bytes,
             or larger than max_h / max_w): their output slot is zeroed, ``check_status`` raises
    Returns a cuda tensor: float32/bfloat16 [n,3,R,R] or uint8 [n,R,R].
    """
    if not packed.is_cuda:
        raise RuntimeError('preprocess_rois: inputs must be CUDA tensors (no CPU fallback)')
    n = int(offsets.shape[0])
    dev = packed.device
    assert packed.dtype == torch.uint8 and offsets.dtype == torch.int64
    assert heights.dtype == torch.int32 and widths.dtype == torch.int32
    if out is None:
        if out_mode == OUT_U8_GRAY:
            out = torch.empty((n, R, R), dtype=torch.uint8, device=dev)
        else:
            out = torch.empty((n, 3, R, R), device=dev,
                              dtype=torch.float32 if out_mode == OUT_F32_NCHW else torch.bfloat16)
    mean_p = std_p = None
    if img_norm:
        mean, std = parse_imgnorm(img_norm) if isinstance(img_norm[0], str) else img_norm
        mean_p = (C.c_float * 3)(*[float(m) for m in mean])
        std_p = (C.c_float * 3)(*[float(s) for s in std])
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    rc = _lib.lib().ifcb_preprocess(packed.data_ptr(), packed.numel(), offsets.data_ptr(), heights.data_ptr(),
                                    widths.data_ptr(), n, int(max_h), int(max_w), int(R), mean_p, std_p,
                                    int(out_mode), out.data_ptr(), int(pass_rule),
                                    status.data_ptr() if status is not None else None, stream)
    _lib.check(rc, 'preprocess')
    return out


def check_status(status):
    """Raises if a preprocess launch refused a ROI (reads the device word: synchronises)."""
    v = int(status.item())
    if v:
        status.zero_()
        why = [t for b, t in ((_lib.IFCB_PRE_BAD_TABLE, 'a ROI table entry points outside the .roi bytes'),
                              (_lib.IFCB_PRE_TOO_LARGE, 'a ROI is larger than the declared max_h / max_w')) if v & b]
        raise RuntimeError('preprocess: ' + '; '.join(why))
