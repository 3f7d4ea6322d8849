"""The callers either side of the hot path, as thin torch glue over the B200 class:

* DenoisingModel.padding_input / crop_output (Experimental_root/models/denoising_model.py:133-168):
  reflect-pad H and W up to multiples of 4, crop back afterwards;
* temp_denoise (Experimental_root/models/validation_seq_infer.py:10-31): constant sigma map, forward,
  clamp to [0,1].

`denoise_sequence` goes through ONE C-ABI call (bsvd_denoise_clip): reflect padding and the constant
noise map are synthesised by the first kernel's loads, clamp and crop by the last kernel's stores.
`denoise_sequence_unfused` keeps the same steps as separate torch ops (the reference's structure);
the tests check the two against each other.  `psnr_per_frame` is calculate_psnr_float
(BasicSR/basicsr/metrics/psnr_ssim.py:130-168) and `ssim_per_frame` calculate_ssim (:49-128) on the device.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def pad_to_multiple_of_4(seq: torch.Tensor):
    """[F,C,H,W] -> reflect-padded tensor and the (pad_h, pad_w) that were added
    (denoising_model.py:133-159 pads bottom/right with 'reflect')."""
    h, w = seq.shape[-2:]
    ph, pw = (4 - h % 4) % 4, (4 - w % 4) % 4
    if ph or pw:
        seq = F.pad(seq, (0, pw, 0, ph), mode="reflect")
    return seq, (ph, pw)


def _sigma_scalar(sigma):
    if torch.is_tensor(sigma):
        s = float(sigma.flatten()[0])
        assert abs(float(sigma.float().mean()) - s) < 1e-5     # validation_seq_infer.py:19
        return s
    return float(sigma)


def denoise_sequence(net, noisy: torch.Tensor, sigma: float | torch.Tensor) -> torch.Tensor:
    """noisy: [F,3,H,W] in [0,1]; sigma: noise std in [0,1] (scalar or constant map).  Denoised
    [F,3,H,W] clamped to [0,1] — what DenoisingModel.test produces for one validation folder
    (pad -> denoise_seq/temp_denoise -> crop), fused into the first/last kernels."""
    return net.denoise_sequence(noisy, _sigma_scalar(sigma))


def psnr_per_frame(a: torch.Tensor, b: torch.Tensor, crop_border: int = 0) -> torch.Tensor:
    """[F,C,H,W] x2 -> [F] PSNR in dB on the device (bsvd_psnr)."""
    from . import capi
    assert a.shape == b.shape and a.is_cuda and b.is_cuda
    a = a.float().contiguous(); b = b.float().contiguous()
    Fr, C, H, W = a.shape
    out = torch.empty(Fr, dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        capi.check(capi.load_library().bsvd_psnr(
            a.data_ptr(), b.data_ptr(), Fr, C, H, W, crop_border, out.data_ptr(),
            torch.cuda.current_stream(a.device).cuda_stream))
    return out


def ssim_per_frame(a: torch.Tensor, b: torch.Tensor, crop_border: int = 0, data_range: float = 1.0) -> torch.Tensor:
    """[F,C,H,W] x2 -> [F] SSIM on the device (bsvd_ssim): calculate_ssim
    (BasicSR/basicsr/metrics/psnr_ssim.py:49-128) per frame.  data_range 1 for [0,1] floats, 255 for the
    reference's [0,255] images (e.g. uint8 frames converted with .float())."""
    from . import capi
    assert a.shape == b.shape and a.is_cuda and b.is_cuda
    a = a.float().contiguous(); b = b.float().contiguous()
    Fr, C, H, W = a.shape
    out = torch.empty(Fr, dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device):
        capi.check(capi.load_library().bsvd_ssim(
            a.data_ptr(), b.data_ptr(), Fr, C, H, W, crop_border, float(data_range), out.data_ptr(),
            torch.cuda.current_stream(a.device).cuda_stream))
    return out


def denoise_sequence_unfused(net, noisy: torch.Tensor, sigma: float | torch.Tensor) -> torch.Tensor:
    """noisy: [F,3,H,W] in [0,1]; sigma: noise std in [0,1] (scalar or [F,1,H,W] constant map).
    Returns the denoised [F,3,H,W] clamped to [0,1] — what DenoisingModel.test produces for one
    validation folder (pad -> denoise_seq/temp_denoise -> crop)."""
    Fr, C, H, W = noisy.shape
    assert C == 3
    dev = torch.device("cuda", torch.cuda.current_device()) if not noisy.is_cuda else noisy.device
    x, (ph, pw) = pad_to_multiple_of_4(noisy.to(dev).float())
    if torch.is_tensor(sigma):
        s = float(sigma.flatten()[0])
        assert abs(float(sigma.float().mean()) - s) < 1e-5     # validation_seq_infer.py:19
    else:
        s = float(sigma)
    nmap = torch.full((Fr, 1, x.shape[-2], x.shape[-1]), s, dtype=torch.float32, device=dev)
    with torch.no_grad():
        out = net(x[None], noise_map=nmap[None])[0]
    out = out.clamp(0.0, 1.0)
    if ph or pw:
        out = out[..., :H, :W]
    return out
