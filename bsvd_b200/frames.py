"""Frame I/O either side of the hot path: image folders in, PNGs and metrics out.

Mirrors, as thin host glue over the device entry points, what the reference does around BSVD.forward when
`run_test.py` validates a folder:

* `get_imagenames` / `open_sequence` (Experimental_root/data/utils_common.py:78-123): ordered file list
  (sorted by the digits in the path), cv2.imread -> uint8 BGR frames;
* `ValFolderDataset.__getitem__` (Experimental_root/data/video_dali_dataset.py:222-249): gt = frames / 255 in
  RGB, noisy = gt + N(0, valnoisestd/255) (float, not re-quantised), constant sigma map;
* `DenoisingModel.test` + `dist_validation` (Experimental_root/models/denoising_model.py:133-190, 276-310):
  pad / denoise / clamp / crop, `tensor2img` (clamp, x255, round, uint8 BGR) + `imwrite` of every frame,
  `calculate_psnr_float` on the float result, `calculate_ssim` on the uint8 images.

Decoding and encoding stay on the host (cv2 codecs: no arithmetic of this path); everything between the
decoded uint8 frames and the uint8 result runs on the device: /255 + RGB swap, noise synthesis, the fused
`bsvd_denoise_clip`, PSNR (`bsvd_psnr`), SSIM (`bsvd_ssim`), quantisation.
"""
from __future__ import annotations

import glob
import os

import numpy as np
import torch

IMAGETYPES = ('*.bmp', '*.png', '*.jpg', '*.jpeg', '*.tif')     # utils_common.py:24


def get_imagenames(seq_dir: str, pattern: str | None = None):
    """Ordered list of image files of a sequence folder (utils_common.py:78-95: sorted by the integer formed
    by ALL digits of the path)."""
    files = []
    for typ in IMAGETYPES:
        files.extend(glob.glob(os.path.join(seq_dir, typ)))
    if pattern is not None:
        files = [f for f in files if pattern in os.path.split(f)[-1]]
    files.sort(key=lambda f: int(''.join(filter(str.isdigit, f)) or 0))
    return files


def read_sequence_u8(seq_dir: str, max_num_fr: int = 100) -> np.ndarray:
    """uint8 [F,H,W,3] in cv2's BGR order (open_sequence without the float conversion)."""
    import cv2
    files = get_imagenames(seq_dir)[:max_num_fr]
    if not files:
        raise FileNotFoundError(f"no image files in {seq_dir}")
    frames = []
    for f in files:
        img = cv2.imread(f)
        if img is None:
            raise IOError(f"cv2.imread failed for {f}")
        frames.append(img)
    if any(fr.shape != frames[0].shape for fr in frames):
        raise ValueError("frames of a sequence must share one size")
    return np.stack(frames, axis=0)


def write_sequence_u8(frames_bgr: np.ndarray, out_dir: str, suffix: str = "") -> list:
    """One PNG per frame, `{idx:08d}{suffix}.png` (denoising_model.py:300-304)."""
    import cv2
    os.makedirs(out_dir, exist_ok=True)
    paths = []
    for i, fr in enumerate(frames_bgr):
        p = os.path.join(out_dir, f"{i:08d}{suffix}.png")
        if not cv2.imwrite(p, np.ascontiguousarray(fr)):
            raise IOError(f"cv2.imwrite failed for {p}")
        paths.append(p)
    return paths


def to_float_rgb(frames_bgr_u8: torch.Tensor) -> torch.Tensor:
    """device uint8 [F,H,W,3] BGR -> float [F,3,H,W] RGB in [0,1] (img2tensor / open_image)."""
    return frames_bgr_u8.flip(-1).permute(0, 3, 1, 2).float().div_(255.0).contiguous()


def to_u8_bgr(x: torch.Tensor) -> torch.Tensor:
    """float [F,3,H,W] RGB -> device uint8 [F,H,W,3] BGR: clamp, x255, round (tensor2img, img_util.py:38-95)."""
    return (x.clamp(0, 1) * 255.0).round().to(torch.uint8).permute(0, 2, 3, 1).flip(-1).contiguous()


def denoise_folder(net, seq_dir: str, out_dir: str | None = None, valnoisestd: float | None = 20.0,
                   add_noise: bool = True, seed: int = 0, max_num_fr: int = 100, crop_border: int = 0,
                   suffix: str = "", device=None) -> dict:
    """Validate one sequence folder the way `run_test.py` does.

    add_noise=True : frames are clean ground truth; noisy = gt + N(0, valnoisestd/255) on the device
                     (torch.Generator(seed): same distribution as ValFolderDataset, not its CPU stream),
                     sigma = valnoisestd/255 for a non-blind model; PSNR / SSIM against the ground truth.
    add_noise=False: frames are the noisy input themselves; uint8 in, uint8 out through the fused
                     frame entry (bsvd_denoise_clip_u8); no metrics.
    Returns {'frames', 'psnr' [F], 'ssim' [F], 'paths'} (metrics None without ground truth)."""
    from . import pipeline
    dev = device or torch.device("cuda", torch.cuda.current_device())
    frames = torch.from_numpy(read_sequence_u8(seq_dir, max_num_fr))
    fr_dev = frames.pin_memory().to(dev, non_blocking=True)
    blind = bool(getattr(net, "blind", False))
    sigma = None if blind else float(valnoisestd) / 255.0
    psnr = ssim = None
    if add_noise:
        gt = to_float_rgb(fr_dev)
        g = torch.Generator(device=dev)
        g.manual_seed(seed)
        noisy = gt + torch.randn(gt.shape, generator=g, device=dev) * (float(valnoisestd) / 255.0)
        with torch.no_grad():
            res = net.denoise_sequence(noisy, sigma)
        psnr = pipeline.psnr_per_frame(res, gt, crop_border).cpu()
        out_u8 = to_u8_bgr(res)
        # calculate_ssim runs on the uint8 images (tensor2img of result and gt), data range 255
        ssim = pipeline.ssim_per_frame(out_u8.permute(0, 3, 1, 2).float(), fr_dev.permute(0, 3, 1, 2).float(),
                                       crop_border, data_range=255.0).cpu()
    else:
        with torch.no_grad():
            out_u8 = net.denoise_frames_u8(fr_dev, sigma, bgr=True)
    out_host = out_u8.cpu().numpy()
    paths = write_sequence_u8(out_host, out_dir, suffix) if out_dir else []
    return {"frames": int(frames.shape[0]), "psnr": psnr, "ssim": ssim, "paths": paths, "result_u8": out_host}
