"""Counterparts of torchdriveenv/helpers.py: ``save_video`` :7-36 and ``set_seeds`` :39-50."""
import os
import random

import numpy as np
import torch


def save_video(imgs, filename, batch_index=0, fps=10, web_browser_friendly=False):
    """imgs: list of B x 3 x H x W uint8 tensors (``BirdviewRecordingWrapper.get_birdviews()``); mp4v, `fps` frames
    per second, as the reference writes it.  ``web_browser_friendly`` re-encodes with ffmpeg when it is installed."""
    import cv2
    stack = [cv2.cvtColor(np.ascontiguousarray(img[batch_index].cpu().numpy().astype(np.uint8).transpose(1, 2, 0)), cv2.COLOR_RGB2BGR)
             for img in imgs]
    h, w = stack[0].shape[0], stack[0].shape[1]
    out = cv2.VideoWriter(filename=filename, fourcc=cv2.VideoWriter_fourcc(*'mp4v'), fps=fps, frameSize=(w, h))
    for frame in stack:
        out.write(frame)
    out.release()
    if web_browser_friendly:
        import shutil
        import uuid
        if shutil.which("ffmpeg") is None:
            raise RuntimeError("save_video(web_browser_friendly=True) needs ffmpeg on PATH")
        tmp = os.path.join(os.path.dirname(filename), str(uuid.uuid4()) + '.mp4')
        os.rename(filename, tmp)
        os.system(f"ffmpeg -y -i {tmp} -hide_banner -loglevel error -vcodec libx264 -f mp4 {filename}")
        os.remove(tmp)


def set_seeds(seed, logger=None):
    if seed is None:
        seed = np.random.randint(low=0, high=2**32 - 1)
    if logger is not None:
        logger.info(f"seed: {seed}")
    torch.manual_seed(seed)
    np.random.seed(seed)
    random.seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)
    return seed
