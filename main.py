#!/usr/bin/env python
"""Command-line driver with the reference's flags (reference main.py:10-98): attack mode writes adversarial PNGs, `--eval`
reports the attack success rate of saved PNGs on the eight victim models.

Differences from the reference driver, none visible in the outputs:
  * `--eps --alpha --epoch --momentum --random_start` ARE forwarded to the attack constructor when given (the reference parses
    them and then builds the attacker with only model_name / targeted, main.py:41); unset flags keep each attack's defaults;
  * batches are uploaded one batch AHEAD from pinned host memory on a side stream (`PrefetchLoader`); the uint8 quantisation +
    NHWC transpose of save_images runs on the device (`ta_quantize_u8`), only bytes come back, on a side stream, and the PNGs
    are encoded by a thread pool while the next batch is attacked (`AsyncImageWriter`; `--sync_io` restores the serial path);
  * `--eval` under torchrun deals the batches round-robin over the ranks (`multigpu.sharded_asr`), same ASR row;
  * `--gpus N` (under torchrun) shards every batch across ranks with `multigpu.run_sharded` (no data-path collective);
  * `--random_weights` builds surrogates / victims with `weights=None` (offline smoke tests; there is no network here).
"""
import argparse
import os

import torch
import tqdm

import transferattack_b200 as transferattack
from transferattack_b200 import multigpu
from transferattack_b200.utils import *  # noqa: F401,F403


def get_parser():
    p = argparse.ArgumentParser(description='Generating transferable adversarial examples (B200 engine)')
    p.add_argument('-e', '--eval', action='store_true', help='attack/evaluation')
    p.add_argument('--attack', default='mifgsm', type=str, choices=transferattack.attack_zoo.keys())
    p.add_argument('--epoch', default=None, type=int)
    p.add_argument('--batchsize', default=32, type=int)
    p.add_argument('--eps', default=None, type=float)
    p.add_argument('--alpha', default=None, type=float)
    p.add_argument('--momentum', default=None, type=float, help='decay factor of momentum based attacks')
    p.add_argument('--model', default='resnet50', type=str)
    p.add_argument('--ensemble', action='store_true')
    p.add_argument('--random_start', default=None, type=bool)
    p.add_argument('--input_dir', default='./data', type=str)
    p.add_argument('--output_dir', default='./results', type=str)
    p.add_argument('--targeted', action='store_true')
    p.add_argument('--GPU_ID', default='0', type=str)
    p.add_argument('--random_weights', action='store_true', help='weights=None surrogates/victims (offline smoke test)')
    p.add_argument('--num_workers', default=4, type=int)
    p.add_argument('--sync_io', action='store_true', help='serial upload / save_images as in the reference (no prefetch, no async writer)')
    p.add_argument('--eval_bf16', action='store_true', help='run the victim models under bf16 autocast in --eval (inference only)')
    return p.parse_args()


def _attack_kwargs(args):
    kw = {}
    for flag, name in (('epoch', 'epoch'), ('eps', 'epsilon'), ('alpha', 'alpha'), ('momentum', 'decay'), ('random_start', 'random_start')):
        v = getattr(args, flag)
        if v is not None:
            kw[name] = v
    return kw


def _build_attacker(args):
    cls = transferattack.load_attack_class(args.attack)
    model_name = args.model.split(',') if (args.ensemble or len(args.model.split(',')) > 1) else args.model
    if args.random_weights:
        def load_model(self, names):
            one = lambda n: wrap_model(models.__dict__[n](weights=None).eval().cuda())
            return EnsembleModel([one(n) for n in names]) if isinstance(names, list) else one(names)
        # plain torchvision nets: the class supplying the surrogate declares it capturable (attack.py: _GRAPH_HOOKS)
        cls = type(cls.__name__, (cls,), {"load_model": load_model, "graph_safe": True})
    return cls(model_name=model_name, targeted=args.targeted, **_attack_kwargs(args))


def main():
    args = get_parser()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        import torch.distributed as dist
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        os.environ["CUDA_VISIBLE_DEVICES"] = args.GPU_ID
    rank = int(os.environ.get("RANK", "0"))
    os.makedirs(args.output_dir, exist_ok=True)

    dataset = AdvDataset(input_dir=args.input_dir, output_dir=args.output_dir, targeted=args.targeted, eval=args.eval)
    loader = torch.utils.data.DataLoader(dataset, batch_size=args.batchsize, shuffle=False, num_workers=args.num_workers,
                                         pin_memory=True)
    device = torch.device("cuda", torch.cuda.current_device())
    if not args.eval:
        attacker = _build_attacker(args)
        batches = loader if (args.sync_io or world > 1) else PrefetchLoader(loader, device)
        writer = None if args.sync_io else AsyncImageWriter(workers=max(2, args.num_workers))
        try:
            for batch_idx, (images, labels, filenames) in tqdm.tqdm(enumerate(batches), disable=rank != 0):
                if args.targeted and isinstance(labels, (list, tuple)):
                    labels = torch.stack(list(labels))
                if world > 1:
                    delta = multigpu.run_sharded(attacker, images, labels, seed=batch_idx, gather=True)
                else:
                    delta = attacker(images, labels)
                if rank == 0:
                    x_dev = images.to(delta.device, non_blocking=True)
                    if writer is None:
                        save_images(args.output_dir, x_dev, filenames, delta=delta)
                    else:
                        writer.submit(args.output_dir, x_dev, filenames, delta=delta)
        finally:
            if writer is not None:
                writer.close()
    else:
        res = '|'
        victims = [(n, models.__dict__[n](weights=None if args.random_weights else "DEFAULT")) for n in cnn_model_paper] \
            if args.random_weights else load_pretrained_model(cnn_model_paper, vit_model_paper)
        for model_name, model in victims:
            model = wrap_model(model.eval().cuda())
            for p_ in model.parameters():
                p_.requires_grad = False
            asr = multigpu.sharded_asr(model, loader, args.targeted, device, dtype=torch.bfloat16 if args.eval_bf16 else None)
            if rank == 0:
                print(f'{model_name}: {asr:.1f}')
            res += f' {asr:.1f} |'
        if rank == 0:
            print(res)
            with open('results_eval.txt', 'a') as f:
                f.write(args.output_dir + res + '\n')
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
