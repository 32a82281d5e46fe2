"""L2 / LPIPS / ID metrics of a finished inversion (spi/utils/metric_utils.py:6-17)."""
import torch

from ..configs import global_config, paths_config
from ..criteria.id_loss.id_loss import IDLoss
from ..criteria.l2_loss import l2_loss
from ..criteria.lpips.lpips import LPIPS


class Metric:
    def __init__(self, lpips_loss=None, id_loss=None):
        self.lpips_loss = lpips_loss if lpips_loss is not None else LPIPS(net_type='vgg').to(global_config.device).eval()
        if id_loss is None:
            import os
            if not os.path.isfile(paths_config.IDLOSS_PATH):
                print(f'[WARNING]: {paths_config.IDLOSS_PATH} not found: the ID similarity column of metric_log.txt comes from a randomly '
                      'initialised IR-SE50 and is only meaningful for plumbing runs')
            id_loss = IDLoss(paths_config.IDLOSS_PATH).to(global_config.device).eval()
        self.id_loss = id_loss

    @torch.no_grad()
    def run(self, gt, fake):
        l2 = l2_loss(gt, fake)
        lpips = self.lpips_loss(gt, fake)
        id_sim = self.id_loss.calculate_similarity(gt, fake)
        return l2.item(), lpips.item(), id_sim.item()


_KEYS = ('l2', 'lpips', 'id', 'l2_m', 'lpips_m', 'id_m')


def format_metric_log(coach_name, hp, metric_dic):
    """The text `BaseCoach.log_metric` appends to metric_log.txt (base_coach.py:157-197), byte for byte: header, then per
    mode one line per image and the plain means."""
    out = (f'Coach name: {coach_name}\n'
           f'hyperparameters.use_encoder: {hp.use_encoder}\n'
           f'hyperparameters.first_inv_type: {hp.first_inv_type}\n'
           f'hyperparameters.first_inv_steps: {hp.first_inv_steps}\n'
           f'hyperparameters.G_1_step: {hp.G_1_step}\n'
           f'hyperparameters.G_2_step: {hp.G_2_step}\n'
           '\n')
    for key, cur in metric_dic.items():
        msg = f'Mode: {key}\n'
        cnt = len(cur['l2'])
        tot = [0, 0, 0, 0, 0, 0]
        for i in range(cnt):
            row = [cur[k][i] for k in _KEYS]
            msg += (f'ID: {i} L2: {row[0]:.6f}; Lpips: {row[1]:.6f}; ID Sim: {row[2]:.6f}; L2 M: {row[3]:.6f}; '
                    f'Lpips M: {row[4]:.6f}; ID Sim M: {row[5]:.6f};\n')
            tot = [t + v for t, v in zip(tot, row)]         # same left-to-right float accumulation as the reference
        avg = [t / cnt for t in tot]
        msg += f'Mode: {key} AVG\n'
        msg += (f'L2: {avg[0]:.6f}; Lpips: {avg[1]:.6f}; ID Sim: {avg[2]:.6f}; L2 M: {avg[3]:.6f}; Lpips M: {avg[4]:.6f}; '
                f'ID Sim M: {avg[5]:.6f};\n')
        out += msg + '\n'
    return out


def merge_metric_dics(parts):
    """Concatenate per-rank metric tables in rank order.  Rank r holds dataset block r+1 of W (contiguous,
    images_dataset.py:149-158), so rank order is dataset order and the merged table equals a single-process run's."""
    merged = {}
    for part in parts:
        for mode, cur in part.items():
            dst = merged.setdefault(mode, {k: [] for k in _KEYS})
            for k in _KEYS:
                dst[k].extend(cur.get(k, []))
    return merged


def gather_metric_dic(metric_dic):
    """-> (merged table, this rank writes the file).  The only collective of a multi-GPU inversion run besides the barrier:
    an object all-gather of ~50 bytes per image (NCCL when the ranks own GPUs, gloo otherwise)."""
    import os
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world == 1:
        return metric_dic, True
    import torch.distributed as dist
    if not dist.is_initialized():
        dist.init_process_group('nccl' if torch.cuda.is_available() else 'gloo')
    parts = [None] * world
    dist.all_gather_object(parts, metric_dic)
    return merge_metric_dics(parts), dist.get_rank() == 0
