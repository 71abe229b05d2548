"""Stand-in for nerfstudio.models.splatfacto (test infrastructure, see tests/stubs/README.md).

Only what `QEDSplatterModel` (/root/reference/qed_splatter/model.py) inherits and touches: the Gaussian parameter
properties, `step`, `crop_box`, `camera_optimizer`, `_get_downscale_factor`, `_get_background_color`, `get_gt_img`,
`composite_with_background` and the parent `get_loss_dict` (splatfacto's RGB loss: (1 - ssim_lambda) * L1 +
ssim_lambda * (1 - SSIM), both images multiplied by `batch["mask"]` when present).  Hand-written from the public
nerfstudio 1.1.5 interface; hyper-parameter defaults as SURVEY.md A.8."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Type

import torch
import torch.nn.functional as F
from torch import nn


def _ssim(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """pytorch_msssim.SSIM(data_range=1.0, size_average=True, channel=3) on [B,3,H,W]: 11x11 Gaussian window
    (sigma 1.5), separable valid convolution, mean of the SSIM map."""
    ch = x.shape[1]
    coords = torch.arange(11, dtype=torch.float64) - 5
    g = torch.exp(-(coords ** 2) / (2 * 1.5 ** 2))
    g = (g / g.sum()).to(x.dtype).to(x.device)

    def filt(t):
        t = F.conv2d(t, g.view(1, 1, 11, 1).expand(ch, 1, 11, 1), groups=ch)
        return F.conv2d(t, g.view(1, 1, 1, 11).expand(ch, 1, 1, 11), groups=ch)

    c1, c2 = 0.01 ** 2, 0.03 ** 2
    mu1, mu2 = filt(x), filt(y)
    s1 = filt(x * x) - mu1 * mu1
    s2 = filt(y * y) - mu2 * mu2
    s12 = filt(x * y) - mu1 * mu2
    cs = (2 * s12 + c2) / (s1 + s2 + c2)
    return (((2 * mu1 * mu2 + c1) / (mu1 * mu1 + mu2 * mu2 + c1)) * cs).flatten(2).mean(-1).mean()


class _NoCameraOptimizer:
    """CameraOptimizer with mode "off" (splatfacto's default): poses pass through, no gradient."""

    def apply_to_camera(self, camera):
        return camera.camera_to_worlds


@dataclass
class SplatfactoModelConfig:
    _target: Type = field(default_factory=lambda: SplatfactoModel)
    warmup_length: int = 500
    refine_every: int = 100
    resolution_schedule: int = 3000
    background_color: str = "random"
    num_downscales: int = 2
    cull_alpha_thresh: float = 0.1
    densify_grad_thresh: float = 0.0008
    sh_degree_interval: int = 1000
    sh_degree: int = 3
    ssim_lambda: float = 0.2
    rasterize_mode: str = "classic"
    use_bilateral_grid: bool = False
    use_scale_regularization: bool = False
    max_gauss_ratio: float = 10.0
    output_depth_during_training: bool = False

    def setup(self, **kwargs):
        return self._target(self, **kwargs)


class SplatfactoModel(nn.Module):
    config: SplatfactoModelConfig

    def __init__(self, config, gauss_params: Dict[str, torch.Tensor], **kwargs):
        """`gauss_params`: means [N,3], scales [N,3] (log), quats [N,4], features_dc [N,3], features_rest [N,15,3],
        opacities [N,1] (logit) -- the real model builds them from seed points in populate_modules()."""
        super().__init__()
        self.config = config
        self.gauss_params = nn.ParameterDict({k: nn.Parameter(v.clone()) for k, v in gauss_params.items()})
        self.populate_modules()

    def populate_modules(self):
        self.step = 0
        self.crop_box = None
        self.camera_optimizer = _NoCameraOptimizer()
        if self.config.background_color == "random":
            self.background_color = torch.tensor([0.1490, 0.1647, 0.2157])
        elif self.config.background_color == "white":
            self.background_color = torch.ones(3)
        else:
            self.background_color = torch.zeros(3)

    # -- parameters ---------------------------------------------------------------------------
    means = property(lambda self: self.gauss_params["means"])
    scales = property(lambda self: self.gauss_params["scales"])
    quats = property(lambda self: self.gauss_params["quats"])
    features_dc = property(lambda self: self.gauss_params["features_dc"])
    features_rest = property(lambda self: self.gauss_params["features_rest"])
    opacities = property(lambda self: self.gauss_params["opacities"])

    @property
    def num_points(self):
        return self.means.shape[0]

    @property
    def device(self):
        return self.means.device

    # -- helpers the subclass calls --------------------------------------------------------------
    def _get_downscale_factor(self):
        if self.training:
            return 2 ** max(self.config.num_downscales - self.step // self.config.resolution_schedule, 0)
        return 1

    def _downscale_if_required(self, image):
        d = self._get_downscale_factor()
        if d > 1:  # nerfstudio resize_image: d x d box filter
            w = (1.0 / (d * d)) * torch.ones((1, 1, d, d), dtype=torch.float32, device=image.device)
            return F.conv2d(image.float().permute(2, 0, 1)[:, None, ...], w, stride=d).squeeze(1).permute(1, 2, 0)
        return image

    def get_gt_img(self, image: torch.Tensor):
        if image.dtype == torch.uint8:
            image = image.float() / 255.0
        return self._downscale_if_required(image).to(self.device)

    def composite_with_background(self, image, background):
        if image.shape[2] == 4:
            alpha = image[..., -1].unsqueeze(-1).repeat((1, 1, 3))
            return alpha * image[..., :3] + (1 - alpha) * background
        return image

    def _get_background_color(self):
        if self.config.background_color == "random":
            if self.training:
                return torch.rand(3, device=self.device)
            return self.background_color.to(self.device)
        return self.background_color.to(self.device)

    def get_empty_outputs(self, width, height, background):
        rgb = background.repeat(height, width, 1)
        depth = background.new_ones(*rgb.shape[:2], 1) * 10
        return {"rgb": rgb, "depth": depth, "accumulation": background.new_zeros(*rgb.shape[:2], 1), "background": background}

    # -- splatfacto's RGB loss (what QEDSplatterModel.get_loss_dict calls through super(), model.py:83-85) ----
    def get_loss_dict(self, outputs, batch, metrics_dict=None) -> Dict[str, torch.Tensor]:
        gt_img = self.composite_with_background(self.get_gt_img(batch["image"]), outputs["background"])
        pred_img = outputs["rgb"]
        if "mask" in batch:
            mask = self._downscale_if_required(batch["mask"]).to(self.device)
            assert mask.shape[:2] == gt_img.shape[:2] == pred_img.shape[:2]
            gt_img = gt_img * mask
            pred_img = pred_img * mask
        Ll1 = torch.abs(gt_img - pred_img).mean()
        simloss = 1 - _ssim(gt_img.permute(2, 0, 1)[None, ...], pred_img.permute(2, 0, 1)[None, ...])
        scale_reg = torch.tensor(0.0).to(self.device)
        return {"main_loss": (1 - self.config.ssim_lambda) * Ll1 + self.config.ssim_lambda * simloss, "scale_reg": scale_reg}
