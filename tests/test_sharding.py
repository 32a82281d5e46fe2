"""CPU, world_size 2, gloo: the multi-GPU host logic.  Images are independent, so ranks only partition the image list
(`--dataset_block (rank+1)/W`, spi/data/images_dataset.py:149-158) and gather a report; no data-path collective."""
import os
import tempfile

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, root, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from spi_b200 import run_inversion
    args = run_inversion.shard_for_rank(run_inversion.parse_args(['--data_root', root, '--first_inv_type', 'mir', '--G_1_type', 'RotBbox']))
    dataset, _ = run_inversion.build_dataset(args)
    names = [os.path.dirname(p).split('/')[-1] for p in dataset.source_paths]
    gathered = [None] * world
    dist.all_gather_object(gathered, names)
    t = torch.tensor([float(len(names))])
    dist.all_reduce(t)                      # the only collective of a real run: the end-of-run report
    if rank == 0:
        torch.save({'gathered': gathered, 'total': t.item(), 'block': args.dataset_block}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_dataset_block_sharding_two_ranks():
    with tempfile.TemporaryDirectory() as root:
        for i in range(7):
            os.makedirs(os.path.join(root, 'crop', f'{i:03d}'))
        out = os.path.join(root, 'out.pt')
        mp.spawn(_worker, args=(2, root, 29533, out), nprocs=2, join=True)
        res = torch.load(out)
        a, b = res['gathered']
        assert res['block'] == '1/2'
        assert sorted(a + b) == [f'{i:03d}' for i in range(7)] and not set(a) & set(b)
        assert res['total'] == 7.0
        assert a == [f'{i:03d}' for i in range(4)]          # contiguous blocks of len//W + 1, as the reference slices them
