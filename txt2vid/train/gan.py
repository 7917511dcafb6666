"""txt2vid/train/gan.py entry point (main + CLI flags, train/gan.py:28-221) on the B200-native stack.

Same flags, same construction / init order (the RNG stream and therefore the initial weights match the
reference for a given --seed), same reflection-based class selection.  Differences: no apex import, the
optimisers are the fused multi-tensor Adam kernel, `--ngpu` is honoured through torch.distributed (launch with
torchrun: one process per GPU), and `--cuda_graphs` replays the iteration as CUDA graphs.
"""
import argparse

import torch
import torch.optim as optim

from txt2vid.train.setup import setup
from txt2vid_b200.gan import CondGan, MixedGanLoss
from txt2vid_b200.optim import FusedAdam
from txt2vid_b200.parallel import DistContext
from txt2vid_b200.trainer import add_params_to_parser, test, train
from txt2vid_b200.util import create_object, init, load, status


def main(args):
    dist = DistContext()
    if dist.enabled and args.seed is None:
        # every rank must build the same initial weights: agree on rank 0's random seed
        import random
        pick = torch.tensor([random.randint(1, 100000)], dtype=torch.int64)
        if torch.distributed.get_backend() == "nccl":
            pick = pick.cuda()
        torch.distributed.broadcast(pick, src=0)
        args.seed = int(pick[0])
    seed, device = setup(args)
    status("%d cuda devices available" % torch.cuda.device_count())
    vocab = load(args.vocab) if args.vocab else None

    txt_encoder = None
    if not args.dont_use_sent:
        if args.sent_weights:
            status("Loading pre-trained sentence model from %s" % args.sent_weights)
            txt_encoder = torch.load(args.sent_weights, weights_only=False)
            if 'txt' in txt_encoder:
                txt_encoder = txt_encoder['txt'].to(device)
        else:
            status("Using random init sentence encoder")
            txt_encoder = create_object(args.sent, vocab_size=len(vocab)).to(device)
            if args.sent_init_method is None:
                args.sent_init_method = args.init_method
            init(txt_encoder, init_method=args.sent_init_method)
    cond_dim = txt_encoder.encoder.encoding_size if txt_encoder is not None else 0

    gen = create_object(args.G, cond_dim=cond_dim).to(device)
    discrims = [create_object(d, cond_dim=cond_dim).to(device) for d in args.D]
    init(gen, init_method=args.init_method)
    for discrim in discrims:
        init(discrim, init_method=args.init_method)
    sample_mapping = None
    if args.M:
        sample_mapping = create_object(args.M).to(device)
        init(sample_mapping, init_method=args.init_method)

    D_params = [{"params": d.parameters()} for d in discrims]
    G_params = [{"params": gen.parameters()}]
    if args.end2end and txt_encoder is not None:
        D_params.append({"params": txt_encoder.parameters()})
        G_params.append({"params": txt_encoder.parameters()})
    if args.sgd:
        optD = optim.SGD(D_params, lr=args.D_lr, momentum=args.D_beta1)
        optG = optim.SGD(G_params, lr=args.G_lr, momentum=args.G_beta1)
    else:
        optD = FusedAdam(D_params, lr=args.D_lr, betas=(args.D_beta1, args.D_beta2))
        optG = FusedAdam(G_params, lr=args.G_lr, betas=(args.G_beta1, args.G_beta2))

    gan = CondGan(gen=gen, discrims=discrims, cond_encoder=txt_encoder, sample_mapping=sample_mapping,
                  discrim_names=args.D_names, discrim_lambdas=args.D_lambdas)
    if args.weights is not None:
        status("Loading weights from %s" % args.weights)
        to_load = torch.load(args.weights, weights_only=False)
        gan.load_from_dict(to_load)
        if 'optD' in to_load:
            optD.load_state_dict(to_load['optD'])
        if 'optG' in to_load:
            optG.load_state_dict(to_load['optG'])

    dataset = create_object(args.data, vocab=vocab, anno=args.anno) if args.data else None
    if args.G_loss is None:
        args.G_loss = args.D_loss
    losses = MixedGanLoss(g_loss=create_object(args.G_loss), d_loss=create_object(args.D_loss))
    if dist.enabled:
        # one process per GPU: identical replicas (same seed, then rank 0's weights broadcast to be safe), per-rank
        # z / caption-permutation streams with SHARED frame offsets, a rank-strided view of the batches, and the
        # summed gradient penalty rescaled for gradient averaging
        for m in [gen, txt_encoder, sample_mapping] + list(discrims):
            dist.broadcast_module(m)
        dist.seed_ranks(seed)
        if dataset is not None:
            dataset = dist.shard(dataset)
        args.gp_lambda = dist.gp_lambda_for(args.gp_lambda, discrims)
    print("GAN has %d parameters" % gan.count_params())
    if args.test:
        test(gan=gan, num_samples=args.num_samples, dataset=dataset, device=device, params=args,
             channel_first=not args.sequence_first, vocab=vocab)
    else:
        train(gan=gan, num_epoch=args.epochs, dataset=dataset, device=device, optD=optD, optG=optG, params=args,
              losses=losses, vocab=vocab, channel_first=not args.sequence_first, end2end=args.end2end,
              dist=dist if dist.enabled else None)


def build_parser():
    parser = argparse.ArgumentParser()
    add_params_to_parser(parser)
    a = parser.add_argument
    a('--test', action='store_true')
    a('--num_samples', type=int, default=1)
    a('--seed', type=int, default=None)
    a('--cuda', action='store_true')
    a('--workers', type=int, default=2)
    a('--ngpu', type=int, default=1)
    a('--frame_sizes', type=int, nargs='+', default=[64])
    a('--num_channels', type=int, default=1)
    a('--random_frames', type=int, default=0)
    a('--opt_level', type=str, default='O2')
    a('--epochs', type=int, default=5)
    a('--batch_size', type=int, default=64)
    a('--init_method', type=str, default='xavier')
    a('--G_loss', type=str, default=None)
    a('--G_lr', type=float, default=0.0001)
    a('--G_beta1', type=float, default=0.5)
    a('--G_beta2', type=float, default=0.9)
    a('--D_loss', type=str, default='txt2vid.gan.losses.VanillaGanLoss')
    a('--D_lr', type=float, default=0.0001)
    a('--D_beta1', type=float, default=0.5)
    a('--D_beta2', type=float, default=0.9)
    a('--weights', type=str, default=None)
    a('--sent_weights', type=str, default=None)
    a('--data', type=str, required=True)
    a('--anno', type=str, default=None)
    a('--vocab', type=str, default=None)
    a('--M', type=str, default=None)
    a('--G', type=str, default=None, required=True)
    a('--D', type=str, default=None, nargs='+', required=True)
    a('--D_names', type=str, default=None, nargs='+')
    a('--D_lambdas', type=float, default=None, nargs='+')
    a('--sent', type=str, default=None)
    a('--sent_init_method', type=str, default=None)
    a('--dont_use_sent', action='store_true', default=False)
    a('--end2end', action='store_true', default=False)
    a('--sgd', action='store_true', default=False)
    a('--sequence_first', action='store_true', default=False)
    a('--debug', action='store_true', default=False)
    a('--cuda_graphs', action='store_true', default=False, help='replay the iteration as CUDA graphs')
    return parser


if __name__ == '__main__':
    main(build_parser().parse_args())
