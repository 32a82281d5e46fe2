"""Identity similarity (drop-in for spi/criteria/id_loss/id_loss.py:7-75).  Metrics only in SPI (never in a gradient path,
SURVEY.md §2 row 6b).  Kept behaviour: the crop [35:223, 32:220] is applied to whatever resolution is passed (512^2 images
from `Metric.run`, base_coach.py:145), then adaptive average pooling to 112^2."""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from .model_irse import Backbone


class IDLoss(nn.Module):
    def __init__(self, path_ir_se50=None, num_scales=1):
        super().__init__()
        self.facenet = Backbone(input_size=112, num_layers=50, drop_ratio=0.6, mode='ir_se')
        if path_ir_se50 and os.path.isfile(path_ir_se50):
            self.facenet.load_state_dict(torch.load(path_ir_se50, map_location='cpu'))
        self.face_pool = torch.nn.AdaptiveAvgPool2d((112, 112))
        self.facenet.eval()
        self.num_scales = num_scales

    def extract_feats(self, x):
        x = x[:, :, 35:223, 32:220]
        x = self.face_pool(x)
        return self.facenet(x)

    def calculate_similarity(self, x, y):
        assert x.shape[0] == 1
        return self.extract_feats(x)[0].dot(self.extract_feats(y)[0])

    def calculate_batch_similarity(self, x, y):
        return torch.mean(torch.sum(self.extract_feats(x) * self.extract_feats(y), dim=-1))

    def forward(self, x, y):
        n_samples = x.shape[0]
        loss = 0.0
        for _scale in range(self.num_scales):
            x_feats, y_feats = self.extract_feats(x), self.extract_feats(y)
            for i in range(n_samples):
                loss = loss + 1 - y_feats[i].dot(x_feats[i])
            if _scale != self.num_scales - 1:
                x = F.interpolate(x, mode='bilinear', scale_factor=0.5, align_corners=False, recompute_scale_factor=True)
                y = F.interpolate(y, mode='bilinear', scale_factor=0.5, align_corners=False, recompute_scale_factor=True)
        return loss / n_samples
