"""GPU debug aid: the discriminator trunk stage by stage in one precision mode against the oracle's functions on the
same weights (forward values, first-order gradients, gradient-penalty style double backward), per pyramid level.
usage: python scripts/debug_d_stages.py fp32|bf16 [perturb]"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import oracle.txt2vid_oracle as O  # noqa: E402
from helpers import build_product_models, l2rel, state_to_cpu  # noqa: E402
from test_product_vs_oracle_cpu import perturb_models  # noqa: E402
from txt2vid_b200 import ops  # noqa: E402

DEV = "cuda" if torch.cuda.is_available() else "cpu"
if DEV == "cpu":                       # build container: exercise the script on the executable spec
    import cpu_kernels
    ops.K = cpu_kernels


def oracle_stages(sd, p, x):
    out = {}
    m = p + "res_block.inner_module"
    h = O._conv(x, sd, m + ".0", 1)
    out["stem"] = F.relu(h)
    h = O._conv(F.relu(h), sd, m + ".2", 1)
    h = F.avg_pool3d(h, (1, 2, 2), 2)
    s = O._conv(F.avg_pool3d(x, (1, 2, 2), 2), sd, p + "res_block.identity_map.1", 0)
    out["skip"] = s
    x = s + h
    out["res"] = x
    i = 0
    while (p + "down.%d.main.inner_module.1.weight" % i) in sd or (p + "down.%d.gamma" % i) in sd:
        if (p + "down.%d.gamma" % i) in sd:
            x = O.attention3d(x, sd, p + "down.%d" % i)
        else:
            x = O.down_block(x, sd, p + "down.%d" % i)
        out["down%d" % i] = x
        i += 1
    out["feat"] = torch.sum(x, [2, 3, 4])
    return out


def product_stages(d, x):
    out = {}
    m = d.res_block.inner_module
    xc = ops.to_cl(x)
    h = ops.conv(ops.im2col3(x), ops.stem_weight_2d(m[0].weight), m[0].bias, relu=True, relu_later=True)
    out["stem"] = ops.from_cl(h, 64)
    c1 = d.res_block.identity_map[1]
    pk, ps = (1, 2, 2), (2, 2, 2)
    skip = ops.conv(ops.avg_pool(xc, pk, ps), c1.weight, c1.bias)
    out["skip"] = ops.from_cl(skip, 64)
    h = ops.conv(h, m[2].weight, m[2].bias, x_relu=True)
    h = ops.avg_pool(h, pk, ps, residual=skip)
    out["res"] = ops.from_cl(h, 64)
    for i, blk in enumerate(d.down):
        h = blk.forward_cl(h)
        out["down%d" % i] = ops.from_cl(h, h.shape[-1] if not hasattr(blk, "out_channels") else blk.out_channels)
    out["feat"] = ops.sum_spatial(h)
    return out


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "fp32"
    perturb = len(sys.argv) > 2
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    txt, gen, dis = build_product_models(True, V=100, seed=100)
    if perturb:
        perturb_models(gen, dis)
    sd = O.as_leaves(state_to_cpu(dis))
    p = O.discrim_prefix(sd)
    dis = dis.to(DEV)
    d = dis.single_discrim.module
    names = [n for n, _ in dis.named_parameters()]
    ops.set_precision(mode)
    g = torch.Generator().manual_seed(5)
    for shape in [(8, 3, 16, 8, 8), (4, 3, 8, 16, 16), (2, 3, 4, 32, 32), (1, 3, 2, 64, 64), (16, 3, 16, 8, 8)]:
        print("=== level", shape, mode)
        x = torch.rand(*shape, generator=g) * 2 - 1
        cond = torch.randn(shape[0], 256, generator=g)
        with torch.no_grad():
            so = oracle_stages(sd, p, x)
            sp = product_stages(d, x.to(DEV))
        for k in so:
            print("  stage %-6s rel %.3e" % (k, l2rel(sp[k].float().cpu(), so[k])))
        # first order: heads + weights
        xo = x.clone().requires_grad_(True)
        co = cond.clone().requires_grad_(True)
        u, c, feat = O.resnet3d(sd, p, xo, co)
        r = torch.randn(feat.shape, generator=g)
        lo = (feat * r).sum() * 1e-3 + u.sum() + 0.5 * c.sum()
        go = torch.autograd.grad(lo, [xo, co] + [sd[n] for n in names], allow_unused=True)
        xp = x.to(DEV).requires_grad_(True)
        cp = cond.to(DEV).requires_grad_(True)
        up, cpred, featp = d(xp, cond=cp)
        lp = (featp * r.to(DEV)).sum() * 1e-3 + up.sum() + 0.5 * cpred.sum()
        gp = torch.autograd.grad(lp, [xp, cp] + list(dis.parameters()), allow_unused=True)
        print("  u %.3e c %.3e feat %.3e" % (l2rel(up.cpu().view(-1), u.view(-1)), l2rel(cpred.cpu().view(-1), c.view(-1)),
                                              l2rel(featp.cpu(), feat)))
        for n, a, b in zip(["x", "cond"] + names, gp, go):
            if b is None:
                continue
            e = l2rel(a.float().cpu(), b)
            if e > (2e-4 if mode == "fp32" else 2e-2):
                print("  grad1 %-50s rel %.3e" % (n, e))
        # gradient penalty style double backward
        u, c, feat = O.resnet3d(sd, p, xo, co)
        gx, = torch.autograd.grad([u, c], [xo], [torch.ones_like(u), torch.ones_like(c)], create_graph=True)
        pen_o = (gx ** 2).sum()
        go2 = torch.autograd.grad(pen_o, [sd[n] for n in names], allow_unused=True)
        up, cpred, featp = d(xp, cond=cp)
        gxp, = torch.autograd.grad([up, cpred], [xp], [torch.ones_like(up), torch.ones_like(cpred)], create_graph=True)
        pen_p = (gxp ** 2).sum()
        gp2 = torch.autograd.grad(pen_p, list(dis.parameters()), allow_unused=True)
        print("  gp: gx %.3e penalty %.6e vs %.6e" % (l2rel(gxp.float().cpu(), gx), float(pen_p), float(pen_o)))
        for n, a, b in zip(names, gp2, go2):
            if b is None or a is None:
                if (a is None) != (b is None) and float((a if b is None else b).abs().max()) > 0:
                    print("  grad2 %-50s None mismatch (product %s, oracle %s)" % (n, a is None, b is None))
                continue
            e = l2rel(a.float().cpu(), b)
            if e > (5e-4 if mode == "fp32" else 5e-2):
                print("  grad2 %-50s rel %.3e" % (n, e))
    # ---- generator: forward levels and first-order gradients under random cotangents
    print("=== generator", mode)
    sdg = O.as_leaves(state_to_cpu(gen))
    gen = gen.to(DEV)
    B = 8
    z, cond = torch.randn(B, 256, generator=g), torch.randn(B, 256, generator=g)
    zo, co = z.clone().requires_grad_(True), cond.clone().requires_grad_(True)
    abs_o = []
    fo = O.gen_forward(sdg, zo, co, [1, 0, 1], True, 16, {}, abstract=abs_o)
    rs = [torch.randn(f.shape, generator=g) for f in fo]
    gnames = [n for n, _ in gen.named_parameters()]
    go = torch.autograd.grad(sum((f * r).sum() for f, r in zip(fo, rs)), [zo, co] + [sdg[n] for n in gnames],
                             allow_unused=True)
    draws = iter([1, 0, 1])
    gen.subsample.draw = lambda: next(draws)
    gen.train()
    zp, cp = z.to(DEV).requires_grad_(True), cond.to(DEV).requires_grad_(True)
    fp, abs_p = gen(zp, cond=cp, return_abstract_maps=True)
    for i, (a, b) in enumerate(zip(abs_p, abs_o)):
        print("  level %d abstract map rel %.3e" % (i, l2rel(a.float().cpu(), b)))
    for i, (a, b) in enumerate(zip(fp, fo)):
        print("  level %d fake rel %.3e" % (i, l2rel(a.float().cpu(), b)))
    # the temporal generator alone
    with torch.no_grad():
        x0 = torch.randn(8, 1024, 1, 1, generator=g)
        ho = torch.stack(O.conv_lstm(x0, sdg, "clstm.cell0", 16), 1)              # (B,T,C,1,1)
        hp, _ = gen.clstm(x0.to(DEV))
        hp = torch.stack(hp, 1)
        for t in (0, 1, 7, 15):
            print("  conv_lstm step %d rel %.3e" % (t, l2rel(hp[:, t].float().cpu(), ho[:, t])))
    gp = torch.autograd.grad(sum((f * r.to(DEV)).sum() for f, r in zip(fp, rs)), [zp, cp] + list(gen.parameters()),
                             allow_unused=True)
    for n, a, b in zip(["z", "cond"] + gnames, gp, go):
        if b is None or a is None:
            continue
        e = l2rel(a.float().cpu(), b)
        if e > (2e-4 if mode == "fp32" else 2e-2):
            print("  gradG %-50s rel %.3e" % (n, e))
    ops.set_precision("bf16")


if __name__ == "__main__":
    main()
