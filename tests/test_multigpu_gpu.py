"""-m gpu, needs >= 2 GPUs (skipped otherwise): the multi-GPU paths over NCCL with the real kernels.
  * batch sharding: every rank attacks its slice, the gathered perturbation equals the per-shard single-GPU runs bit for bit;
  * one surrogate per GPU (ShardedEnsembleModel): two NCCL all-reduces per iteration; with K = 2 the perturbation equals the
    single-device EnsembleModel's bit for bit, and all ranks stay in lockstep."""
import os
import sys
import tempfile

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

pytestmark = pytest.mark.gpu


def _need2():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")


def _net(arch, seed, dev):
    import torchvision
    torch.manual_seed(seed)
    return getattr(torchvision.models, arch)(weights=None).eval().to(dev)


def _data(B=4):
    g = torch.Generator().manual_seed(1)
    return torch.rand(B, 3, 224, 224, generator=g), torch.randint(0, 1000, (B,), generator=g)


def _worker(rank, world, init_file, out_dir, case):
    import torch.distributed as dist
    import transferattack_b200 as tab
    from transferattack_b200 import multigpu
    from helpers import make_attack
    torch.cuda.set_device(rank)
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.deterministic = True
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", init_method="file://" + init_file, rank=rank, world_size=world, device_id=dev)
    try:
        x, y = _data()
        if case == "shard":
            atk = make_attack(tab, "mifgsm", _net("resnet18", 0, dev), epoch=4)
            out = multigpu.run_sharded(atk, x, y, gather=True)
        elif case == "shard_dim":
            atk = make_attack(tab, "ditimi", _net("resnet18", 0, dev), epoch=4)
            out = multigpu.run_sharded(atk, x, y, seed=5 + rank, gather=True)
        elif case in ("ens", "ens_rs"):
            member = tab.utils.wrap_model(_net("resnet18" if rank == 0 else "mobilenet_v2", 0 if rank == 0 else 3, dev))
            torch.cuda.manual_seed(100 + rank)         # per-rank device generators differ: a random start must still agree
            atk = multigpu.make_ens_attack(tab.load_attack_class("ens"), member, epoch=4, random_start=(case == "ens_rs"))
            out = atk(x, y)
        elif case == "p2p":
            member = tab.utils.wrap_model(_net("resnet18" if rank == 0 else "mobilenet_v2", 0 if rank == 0 else 3, dev))
            run = multigpu.make_fused_p2p_ens(tab.load_attack_class("ens"), member, epoch=4)
            out = run(x, y)
            out2 = run(*_data(4))                       # second batch through the cached symmetric buffers
            assert torch.equal(out, out2)
        np.save(os.path.join(out_dir, "%s_rank%d.npy" % (case, rank)), out.detach().cpu().numpy())
    finally:
        dist.destroy_process_group()


def _run(case, world=2):
    import torch.multiprocessing as mp
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, os.path.join(d, "init"), d, case), nprocs=world, join=True)
        return [np.load(os.path.join(d, "%s_rank%d.npy" % (case, r))) for r in range(world)]


def test_batch_sharded_equals_single_gpu_shards():
    _need2()
    import transferattack_b200 as tab
    from helpers import make_attack
    outs = _run("shard")
    assert np.array_equal(outs[0], outs[1])
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.deterministic = True
    x, y = _data()
    atk = make_attack(tab, "mifgsm", _net("resnet18", 0, "cuda:0"), epoch=4)
    for lo, hi in ((0, 2), (2, 4)):
        assert np.array_equal(outs[0][lo:hi], atk(x[lo:hi], y[lo:hi]).cpu().numpy())


def test_batch_sharded_di_ti_mi_shares_host_rng():
    _need2()
    import transferattack_b200 as tab
    from helpers import make_attack, seed_all
    outs = _run("shard_dim")
    assert np.array_equal(outs[0], outs[1])
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.deterministic = True
    x, y = _data()
    atk = make_attack(tab, "ditimi", _net("resnet18", 0, "cuda:0"), epoch=4)
    for lo, hi in ((0, 2), (2, 4)):
        seed_all(5)                                       # rank 0's seed
        assert np.array_equal(outs[0][lo:hi], atk(x[lo:hi], y[lo:hi]).cpu().numpy())


def test_sharded_ensemble_equals_single_device_ensemble():
    _need2()
    import transferattack_b200 as tab
    from helpers import make_attack
    outs = _run("ens")
    assert np.array_equal(outs[0], outs[1])
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.deterministic = True
    x, y = _data()
    nets = [_net("resnet18", 0, "cuda:0"), _net("mobilenet_v2", 3, "cuda:0")]
    ref = make_attack(tab, "ens", nets, epoch=4)(x, y).cpu().numpy()
    assert np.array_equal(outs[0], ref)


def test_fused_p2p_ensemble_equals_single_device_ensemble():
    """ta_fused_allreduce_update_linf: reduce-scatter + update + all-gather in one kernel over NVLink peer memory."""
    _need2()
    import transferattack_b200 as tab
    from helpers import make_attack
    outs = _run("p2p")
    assert np.array_equal(outs[0], outs[1])
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.deterministic = True
    x, y = _data()
    nets = [_net("resnet18", 0, "cuda:0"), _net("mobilenet_v2", 3, "cuda:0")]
    atk = make_attack(tab, "ens", nets, epoch=4)
    atk.mean_mode = "exact"
    ref = atk(x, y).cpu().numpy()
    assert np.array_equal(outs[0], ref), int((outs[0] != ref).sum())


def test_sharded_ensemble_random_start_is_broadcast():
    """the sharded-ensemble attack broadcasts rank 0's random start, so the replicated updates stay in lockstep"""
    _need2()
    outs = _run("ens_rs")
    assert np.array_equal(outs[0], outs[1]) and np.abs(outs[0]).max() <= 16 / 255 + 1e-7
