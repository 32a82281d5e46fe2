"""Oracle (CPU, test infrastructure): ArcFace IR-SE50 identity similarity (metrics only in the reference).

Follows spi/criteria/id_loss/id_loss.py:17-28 (crop [35:223, 32:220] of whatever resolution is passed, adaptive average
pool to 112^2, backbone, dot product of the L2-normalised 512-d features) and the backbone of
spi/criteria/id_loss/model_irse.py:10-49 + helpers.py:23-121 (IR-SE bottlenecks, eval-mode BatchNorm, PReLU, SE gate),
functional over a state dict with the reference's module names.
"""
import torch
import torch.nn.functional as F

IRSE50_BLOCKS = [(64, 64, 3), (64, 128, 4), (128, 256, 14), (256, 512, 3)]     # (in_channel, depth, units), first unit stride 2


def unit_list():
    units = []
    for cin, depth, n in IRSE50_BLOCKS:
        units.append((cin, depth, 2))
        units += [(depth, depth, 1)] * (n - 1)
    return units


def _bn(x, sd, p):
    return F.batch_norm(x, sd[p + 'running_mean'], sd[p + 'running_var'], sd[p + 'weight'], sd[p + 'bias'], False, 0.0, 1e-5)


def _prelu(x, w):
    return F.prelu(x, w)


def bottleneck_ir_se(x, sd, p, cin, depth, stride):
    """helpers.py:99-121."""
    if cin == depth:
        shortcut = F.max_pool2d(x, 1, stride)
    else:
        shortcut = _bn(F.conv2d(x, sd[p + 'shortcut_layer.0.weight'], stride=stride), sd, p + 'shortcut_layer.1.')
    r = _bn(x, sd, p + 'res_layer.0.')
    r = F.conv2d(r, sd[p + 'res_layer.1.weight'], padding=1)
    r = _prelu(r, sd[p + 'res_layer.2.weight'])
    r = F.conv2d(r, sd[p + 'res_layer.3.weight'], stride=stride, padding=1)
    r = _bn(r, sd, p + 'res_layer.4.')
    s = F.adaptive_avg_pool2d(r, 1)
    s = F.relu(F.conv2d(s, sd[p + 'res_layer.5.fc1.weight']))
    s = torch.sigmoid(F.conv2d(s, sd[p + 'res_layer.5.fc2.weight']))
    return r * s + shortcut


def backbone(x, sd, prefix=''):
    """Backbone.forward (model_irse.py:44-48), input 112^2, eval mode (Dropout = identity)."""
    p = prefix
    x = _prelu(_bn(F.conv2d(x, sd[p + 'input_layer.0.weight'], padding=1), sd, p + 'input_layer.1.'), sd[p + 'input_layer.2.weight'])
    for i, (cin, depth, stride) in enumerate(unit_list()):
        x = bottleneck_ir_se(x, sd, f'{p}body.{i}.', cin, depth, stride)
    x = _bn(x, sd, p + 'output_layer.0.')
    x = x.reshape(x.shape[0], -1)
    x = F.linear(x, sd[p + 'output_layer.3.weight'], sd[p + 'output_layer.3.bias'])
    x = F.batch_norm(x, sd[p + 'output_layer.4.running_mean'], sd[p + 'output_layer.4.running_var'], sd[p + 'output_layer.4.weight'],
                     sd[p + 'output_layer.4.bias'], False, 0.0, 1e-5)
    return x / torch.norm(x, 2, 1, True)


def extract_feats(x, sd):
    """id_loss.py:17-21."""
    x = x[:, :, 35:223, 32:220]
    x = F.adaptive_avg_pool2d(x, (112, 112))
    return backbone(x, sd, 'facenet.')


def similarity(x, y, sd):
    """IDLoss.calculate_similarity (id_loss.py:23-28)."""
    return extract_feats(x, sd)[0].dot(extract_feats(y, sd)[0])
