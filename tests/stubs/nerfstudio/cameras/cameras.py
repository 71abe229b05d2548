"""Stand-in for nerfstudio.cameras.cameras.Cameras (test infrastructure, see tests/stubs/README.md): the attributes
and methods QEDSplatterModel.get_outputs touches (model.py:207-250)."""
from __future__ import annotations

import torch


class Cameras:
    def __init__(self, camera_to_worlds, fx, fy, cx, cy, width, height, metadata=None):
        C = camera_to_worlds.shape[0]
        self.camera_to_worlds = camera_to_worlds  # [C,3,4]

        def col(v, dtype):
            return torch.as_tensor(v, dtype=dtype, device=camera_to_worlds.device).reshape(-1, 1).expand(C, 1).clone()

        self.fx, self.fy, self.cx, self.cy = (col(v, torch.float32) for v in (fx, fy, cx, cy))
        self.width, self.height = col(width, torch.int64), col(height, torch.int64)
        self.metadata = metadata

    @property
    def shape(self):
        return self.camera_to_worlds.shape[:-2]

    @property
    def device(self):
        return self.camera_to_worlds.device

    def rescale_output_resolution(self, scaling_factor: float) -> None:
        """nerfstudio: intrinsics scale linearly, image size is rounded to nearest."""
        self.fx = self.fx * scaling_factor
        self.fy = self.fy * scaling_factor
        self.cx = self.cx * scaling_factor
        self.cy = self.cy * scaling_factor
        self.height = torch.floor(self.height * scaling_factor + 0.5).to(torch.int64)
        self.width = torch.floor(self.width * scaling_factor + 0.5).to(torch.int64)

    def get_intrinsics_matrices(self) -> torch.Tensor:
        K = torch.zeros(*self.shape, 3, 3, dtype=torch.float32, device=self.device)
        K[..., 0, 0] = self.fx.squeeze(-1)
        K[..., 1, 1] = self.fy.squeeze(-1)
        K[..., 0, 2] = self.cx.squeeze(-1)
        K[..., 1, 2] = self.cy.squeeze(-1)
        K[..., 2, 2] = 1.0
        return K
