"""Recording and seeding helpers behind the reference's names (``save_video``, ``set_seeds`` of
torchdriveenv/helpers.py): frames recorded by ``BirdviewRecordingWrapper`` (tde_render_view) go to an mp4 file,
and one call seeds every generator the host side uses."""
from __future__ import annotations

import os
import random
import shutil
import subprocess
import tempfile
from typing import Iterable, Optional

import numpy as np
import torch

_FOURCC = "mp4v"


def _frames_bgr(imgs: Iterable, batch_index: int):
    """B x 3 x H x W uint8 frames (torch tensors on any device, or numpy arrays) -> H x W x 3 BGR arrays for OpenCV."""
    for img in imgs:
        frame = img[batch_index]
        frame = frame.detach().cpu().numpy() if torch.is_tensor(frame) else np.asarray(frame)
        yield np.ascontiguousarray(frame.astype(np.uint8, copy=False)[::-1].transpose(1, 2, 0))   # RGB planes -> BGR pixels


def save_video(imgs, filename, batch_index=0, fps=10, web_browser_friendly=False):
    """Writes the recorded frames as an mp4 (codec mp4v) at ``fps`` frames per second; ``batch_index`` picks the env
    of a batched recording.  ``web_browser_friendly`` re-encodes the file to H.264 with ffmpeg (must be on PATH)."""
    import cv2
    writer = None
    try:
        for frame in _frames_bgr(imgs, batch_index):
            if writer is None:
                height, width = frame.shape[:2]
                writer = cv2.VideoWriter(filename, cv2.VideoWriter_fourcc(*_FOURCC), fps, (width, height))
            writer.write(frame)
    finally:
        if writer is not None:
            writer.release()
    if writer is None:
        raise ValueError("save_video: no frames")
    if web_browser_friendly:
        if shutil.which("ffmpeg") is None:
            raise RuntimeError("save_video(web_browser_friendly=True) needs ffmpeg on PATH")
        fd, tmp = tempfile.mkstemp(suffix=".mp4", dir=os.path.dirname(os.path.abspath(filename)))
        os.close(fd)
        os.replace(filename, tmp)
        try:
            subprocess.run(["ffmpeg", "-y", "-i", tmp, "-hide_banner", "-loglevel", "error", "-vcodec", "libx264", "-f", "mp4", filename], check=True)
        finally:
            os.remove(tmp)


def set_seeds(seed: Optional[int], logger=None) -> int:
    """Seeds python, numpy and torch (CPU and, when present, CUDA) generators; draws a seed when none is given."""
    chosen = int(np.random.randint(0, 2**32 - 1)) if seed is None else seed
    if logger is not None:
        logger.info(f"seed: {chosen}")
    for seeder in (random.seed, np.random.seed, torch.manual_seed):
        seeder(chosen)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(chosen)
    return chosen
