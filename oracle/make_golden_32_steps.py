"""Golden vectors of the full loop steps at the BENCH depth resolution (32 + 32 samples per ray): the same reference runs as
`make_golden.golden_steps` (two `mir` projector steps, one PTI step, one RotBbox `i % 4 == 0` step with all four branches; ~8 min of
CPU) with `rendering_kwargs['depth_resolution'(_importance)] = 32`.  Writes tests/golden/steps_32_full.npz.

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden_32_steps
"""
import json
import os
import shutil

import torch

from . import make_golden as MG
from . import ref_shim, weights


def main():
    assert ref_shim.available(), 'reference tree not found'
    ref_shim.install()
    torch.manual_seed(0)
    sd = weights.generator_state_dict(0)
    G = ref_shim.build_reference_generator()
    G.load_state_dict(sd, strict=True)
    G.rendering_kwargs = dict(G.rendering_kwargs, depth_resolution=32, depth_resolution_importance=32)
    nets = MG.make_nets()
    keep = os.path.join(MG.OUT, 'steps.npz')
    aside = keep + '.keep'
    shutil.copyfile(keep, aside)
    try:
        before = dict(MG.REPORT)
        MG.golden_steps(G, sd, nets, fast=False)
        os.replace(keep, os.path.join(MG.OUT, 'steps_32_full.npz'))
    finally:
        os.replace(aside, keep)
    path = os.path.join(MG.OUT, 'REPORT.json')
    old = json.load(open(path)) if os.path.exists(path) else {}
    old.update({k.replace('/', '32/', 1) if not k.startswith('_') else k: v for k, v in MG.REPORT.items() if k not in before})
    json.dump(old, open(path, 'w'), indent=1, sort_keys=True)
    print('wrote', path)


if __name__ == '__main__':
    main()
