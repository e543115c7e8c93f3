"""world_size-2 tests of transferattack_b200.multigpu on CPU (gloo): batch sharding without a data-path collective, the
host-RNG synchronisation DIM needs, and the one-surrogate-per-rank ensemble (logits all-reduce forward, input-gradient
all-reduce backward). The kernels are replaced by the oracle stand-in in every process (tests/oracle_backend.py)."""
import os
import sys
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def _worker(rank, world, init_file, out_dir, case):
    import transferattack_b200 as tab
    from transferattack_b200 import multigpu, ops
    from oracle_backend import OracleBackend
    from helpers import make_attack, tiny_net
    torch.set_num_threads(1)
    ops._install_backend_for_tests(OracleBackend())
    dist.init_process_group("gloo", init_method="file://" + init_file, rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(1)
        x = torch.rand(5 if case == "uneven" else 4, 3, 32, 32, generator=g)
        y = torch.randint(0, 10, (x.shape[0],), generator=g)
        if case in ("mifgsm", "uneven"):
            atk = make_attack(tab, "mifgsm", tiny_net(0), epoch=4)
            out = multigpu.run_sharded(atk, x, y, gather=True)
        elif case == "targeted":
            atk = make_attack(tab, "mifgsm", tiny_net(0), epoch=3, targeted=True)
            out = multigpu.run_sharded(atk, x, torch.stack([y, (y + 1) % 10]), gather=True)
        elif case == "dim":
            atk = make_attack(tab, "dim", tiny_net(0), epoch=4)
            out = multigpu.run_sharded(atk, x, y, seed=7 + 100 * rank, gather=True)      # rank 0's seed must win
        elif case == "admix":
            atk = make_attack(tab, "admix", tiny_net(0), epoch=1)
            try:
                multigpu.run_sharded(atk, x, y)
                out = torch.zeros(1)
            except RuntimeError as e:
                out = torch.ones(1) if "replicas" in str(e) else torch.zeros(1)
        elif case in ("ens", "ens_rs"):
            member = tab.utils.wrap_model(tiny_net(0 if rank == 0 else 3))
            torch.manual_seed(100 + rank)          # different generator states per rank: a random start must still agree
            atk = multigpu.make_ens_attack(tab.load_attack_class("ens"), member, epoch=4, random_start=(case == "ens_rs"))
            out = atk(x, y)
        elif case == "asr":
            model = tab.utils.wrap_model(tiny_net(0))
            batches = [(x[i:i + 1], y[i:i + 1], ["f%d" % i]) for i in range(x.shape[0])] + [(x[:3], y[:3], ["a", "b", "c"])]
            out = torch.tensor([multigpu.sharded_asr(model, batches, False, "cpu"),
                                multigpu.sharded_asr(model, [(b[0], torch.stack([b[1], b[1]]), b[2]) for b in batches], True, "cpu")])
        else:
            raise ValueError(case)
        np.save(os.path.join(out_dir, "%s_rank%d.npy" % (case, rank)), out.detach().cpu().numpy())
    finally:
        dist.destroy_process_group()


def _run(case, world=2):
    with tempfile.TemporaryDirectory() as d:
        init_file = os.path.join(d, "init")
        mp.spawn(_worker, args=(world, init_file, d, case), nprocs=world, join=True)
        return [np.load(os.path.join(d, "%s_rank%d.npy" % (case, r))) for r in range(world)]


def _inputs(n=4):
    g = torch.Generator().manual_seed(1)
    x = torch.rand(n, 3, 32, 32, generator=g)
    return x, torch.randint(0, 10, (n,), generator=g)


def test_shard_bounds_cover_the_batch():
    from transferattack_b200.multigpu import shard_bounds
    for n in (0, 1, 5, 64, 255, 256):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


def test_batch_sharded_mifgsm_equals_per_shard_reference():
    from oracle import torch_ref
    from helpers import tiny_net
    outs = _run("mifgsm")
    assert np.array_equal(outs[0], outs[1]) and outs[0].shape == (4, 3, 32, 32)
    x, y = _inputs()
    for lo, hi in ((0, 2), (2, 4)):
        ref = torch_ref.ref_mifgsm(torch_ref.ref_wrap_model(tiny_net(0)), epoch=4)(x[lo:hi], y[lo:hi]).numpy()
        assert np.array_equal(outs[0][lo:hi], ref)
    # vs the unsharded batch: the CE-mean factor B/B_shard = 2 cancels exactly in g / mean|g|
    full = torch_ref.ref_mifgsm(torch_ref.ref_wrap_model(tiny_net(0)), epoch=4)(x, y).numpy()
    assert (np.abs(outs[0] - full) > 1e-6).mean() <= 1e-3


def test_uneven_batch_and_targeted_labels():
    outs = _run("uneven")
    assert outs[0].shape == (5, 3, 32, 32) and np.array_equal(outs[0], outs[1])
    outs = _run("targeted")
    assert outs[0].shape == (4, 3, 32, 32) and np.array_equal(outs[0], outs[1])


def test_dim_shards_share_the_host_rng():
    from oracle import torch_ref
    from helpers import tiny_net, seed_all
    outs = _run("dim")
    assert np.array_equal(outs[0], outs[1])
    x, y = _inputs()
    for lo, hi in ((0, 2), (2, 4)):
        ref = torch_ref.RefDIM(torch_ref.ref_wrap_model(tiny_net(0)), epoch=4)
        seed_all(7)
        r = ref(x[lo:hi], y[lo:hi]).numpy()
        # same coin / size / pad sequence on both shards; DIM's blend is tolerance-level vs ATen's CPU kernel
        assert (np.abs(outs[0][lo:hi] - r) > 1e-6).mean() <= 5e-3


def test_admix_refuses_to_shard():
    outs = _run("admix")
    assert outs[0][0] == 1 and outs[1][0] == 1


def test_sharded_ensemble_matches_single_device_ensemble():
    from oracle import torch_ref
    from helpers import tiny_net
    outs = _run("ens")
    assert np.array_equal(outs[0], outs[1])          # replicated update stays in lockstep
    x, y = _inputs()
    ens = torch_ref.RefEnsemble([torch_ref.ref_wrap_model(tiny_net(0)), torch_ref.ref_wrap_model(tiny_net(3))])
    ref = torch_ref.ref_mifgsm(ens, epoch=4)(x, y).numpy()
    assert np.array_equal(outs[0], ref)              # K = 2: two-term sums commute → bit-identical


def test_sharded_ensemble_random_start_is_broadcast():
    """with random_start every rank drew its own delta; the sharded-ensemble attack now broadcasts rank 0's draw,
    so the replicated updates stay in lockstep (different per-rank generator states on purpose)."""
    outs = _run("ens_rs")
    assert np.array_equal(outs[0], outs[1])
    assert np.abs(outs[0]).max() <= 16 / 255 + 1e-7


def test_sharded_asr_equals_single_process_pass():
    """--eval with the batches dealt round-robin over the ranks returns the ASR a single pass returns"""
    import transferattack_b200 as tab
    from transferattack_b200 import multigpu, ops
    from oracle_backend import OracleBackend
    from helpers import tiny_net
    outs = _run("asr")
    assert np.array_equal(outs[0], outs[1])
    ops._install_backend_for_tests(OracleBackend())
    try:
        x, y = _inputs()
        model = tab.utils.wrap_model(tiny_net(0))
        with torch.no_grad():
            pred = torch.cat([model(x[i:i + 1]).argmax(1) for i in range(4)] + [model(x[:3]).argmax(1)])
        lab = torch.cat([y, y[:3]])
        acc = float((pred == lab).sum()) / 7
        assert abs(outs[0][0] - (1 - acc) * 100) < 1e-9 and abs(outs[0][1] - acc * 100) < 1e-9
    finally:
        ops._install_backend_for_tests(None)
