"""Inference-side shim for the reference's `Trainer` (trainer.py:19-47, :224-247): exposes `gen`,
`gen_ema` and `load_checkpoint` the way test_fullframework.py:47-49 uses them, with the CUDA
Generator. Unlike the reference it does NOT wrap the generators in nn.DataParallel (that wrapper
breaks `model.mot_embedding` on a CUDA machine). Training (losses, optimiser, EMA
updates) is out of scope for this hot-path library."""
from __future__ import annotations

import copy

import numpy as np
import torch
from torch import nn

from .model import Generator


class Trainer(nn.Module):
    def __init__(self, config, precision: str = "fp32"):
        super().__init__()
        self.gen = Generator(config["model"], precision=precision)
        self.gen_ema = copy.deepcopy(self.gen)
        self.model_dir = config.get("model_dir")
        self.config = config
        parents = np.array(config["dataset"]["mocha"]["parents"])
        self.parents = np.concatenate([[-1], parents + 1])
        self.device = "cpu"
        if torch.cuda.is_available():
            self.device = torch.cuda.current_device()
            self.gen = self.gen.to(self.device)
            self.gen_ema = self.gen_ema.to(self.device)

    def load_checkpoint(self, model_path=None, resume=False):
        if resume:
            raise NotImplementedError("optimizer state is training-only")
        state_dict = torch.load(model_path, map_location="cpu")
        self.gen.load_state_dict(state_dict["gen"])
        self.gen_ema.load_state_dict(state_dict["gen_ema"])
        epochs = int(model_path[-6:-3])
        print("Load from epoch %d" % epochs)
        return epochs
