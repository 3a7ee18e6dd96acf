# -*- coding: utf-8 -*-
"""Command line of the reference (main.py:10-76), driving the B200 hot path.

Same flags and defaults, same loop over the three (Step, Ne, Nc, Ner, Ncr) presets; every flag
given on the command line overrides ALL presets, as in the reference.  Extra flags: --variant
(1-4 = model_1..model_4, default 2 = the model main.py imports), --root (directory holding
Adjset/ and dataset/), --seed, --no_q2 (feed each batch its own entity->hunk maps instead of the
reference's first-Mini_batch slice, model_2.py:376-381).
Multi-GPU: launch with `python -m torch.distributed.run --nproc-per-node N main.py ...`.
"""
import argparse
import os

import torch


def main(argv=None):
    steps = [2, 3, 5]
    entity_nodes = [200, 250, 250]
    hunk_nodes = [74, 114, 150]
    entity_edges = [39800, 62250, 62250]
    hunk_edges = [5402, 12882, 22350]
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and not torch.distributed.is_initialized():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        torch.distributed.init_process_group("nccl")
    from hdgnn_b200.model import graph2graph

    results = []
    for step, entity_node, hunk_node, entity_edge, hunk_edge in zip(steps, entity_nodes, hunk_nodes, entity_edges, hunk_edges):
        print(step, entity_node, hunk_node, entity_edge, hunk_edge)
        parser = argparse.ArgumentParser(description='')
        parser.add_argument('--epoch', type=int, default=50, help='number of training epochs')
        parser.add_argument('--Ds', type=int, default=1, help='The State Dimention')
        parser.add_argument('--Ds_inter', type=int, default=1, help='The State Dimention of inter state')
        parser.add_argument('--Dr', type=int, default=2, help='The Relationship Dimension')
        parser.add_argument('--Dr_inter', type=int, default=2, help='The Relationship Dimension of inter state')
        parser.add_argument('--De_e', type=int, default=20, help='The Effect Dimension on entity')
        parser.add_argument('--De_er', type=int, default=20, help='The Effect Dimension on entity Relations')
        parser.add_argument('--Mini_batch', type=int, default=50, help='The training mini_batch')
        parser.add_argument('--checkpoint_dir', dest='checkpoint_dir', default='./checkpoint40/', help='models are saved here')
        parser.add_argument('--Ne', type=int, default=entity_node, help='The Number of entities')
        parser.add_argument('--Nc', type=int, default=hunk_node, help='The Number of code changes')
        parser.add_argument('--Ner', type=int, default=entity_edge, help='The Number of entity Relations')
        parser.add_argument('--Ncr', type=int, default=hunk_edge, help='The Number of code change Relations')
        parser.add_argument('--Step', type=int, default=step, help='the number of commits/groups')
        parser.add_argument('--Repo', type=str, default='glide', help='the name of repository')
        parser.add_argument('--Type', dest='Type', default='train', help='train or test')
        # additions (no reference flag renamed or removed)
        parser.add_argument('--variant', type=int, default=2, help='1..4 = model_1..model_4 (reference imports model_2)')
        parser.add_argument('--root', type=str, default='.', help='directory holding Adjset/ and dataset/')
        parser.add_argument('--seed', type=int, default=None, help='weight-init seed (the reference is unseeded)')
        parser.add_argument('--no_q2', action='store_true', help='per-batch entity->hunk maps (fixes model_2.py:376-381)')
        args = parser.parse_args(argv)

        ck = args.checkpoint_dir if os.path.isabs(args.checkpoint_dir) else os.path.join(args.root, args.checkpoint_dir)
        if not os.path.exists(ck):
            os.makedirs(ck, exist_ok=True)
        model = graph2graph(None,
                            Ds=args.Ds,
                            Ne=args.Ne, Nc=args.Nc,
                            Ner=args.Ner, Ncr=args.Ncr,
                            Dr=args.Dr,
                            De_e=args.De_e, De_er=args.De_er,
                            Mini_batch=args.Mini_batch,
                            checkpoint_dir=args.checkpoint_dir,
                            epoch=args.epoch,
                            Ds_inter=args.Ds_inter, Dr_inter=args.Dr_inter,
                            Step=args.Step,
                            Repo=args.Repo,
                            variant=args.variant, seed=args.seed)
        if args.Type == 'train':
            results.append(model.train(args, root=args.root, quirk_q2=not args.no_q2))
        if args.Type == 'test':
            results.append(model.test(args, root=args.root, quirk_q2=not args.no_q2))
        model.engine.close()
    return results


if __name__ == '__main__':
    main()
