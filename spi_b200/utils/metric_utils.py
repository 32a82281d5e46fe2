"""L2 / LPIPS / ID metrics of a finished inversion (spi/utils/metric_utils.py:6-17)."""
import torch

from ..configs import global_config, paths_config
from ..criteria.id_loss.id_loss import IDLoss
from ..criteria.l2_loss import l2_loss
from ..criteria.lpips.lpips import LPIPS


class Metric:
    def __init__(self, lpips_loss=None, id_loss=None):
        self.lpips_loss = lpips_loss if lpips_loss is not None else LPIPS(net_type='vgg').to(global_config.device).eval()
        self.id_loss = id_loss if id_loss is not None else IDLoss(paths_config.IDLOSS_PATH).to(global_config.device).eval()

    @torch.no_grad()
    def run(self, gt, fake):
        l2 = l2_loss(gt, fake)
        lpips = self.lpips_loss(gt, fake)
        id_sim = self.id_loss.calculate_similarity(gt, fake)
        return l2.item(), lpips.item(), id_sim.item()
