"""GPU: the one-launch Adam (csrc/lgs_adam.cu through lgs_b200.optim.Adam) against goldens of torch.optim.Adam on CUDA, against
torch.optim.Adam live on the same device, and through the optimizer-state surgery the reference's densification does."""
import numpy as np
import pytest
import torch

from test_oracle_adam_golden import GOLD, ids

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def same_bits(a, b):
    return torch.equal(a.detach().view(torch.int32), b.detach().view(torch.int32))


@pytest.mark.parametrize("path", GOLD, ids=ids)
def test_matches_torch_adam_goldens_bit_for_bit(path):
    from lgs_b200 import optim
    g = np.load(path)
    n = len(g["in_lrs"])
    ps = [torch.nn.Parameter(torch.from_numpy(g[f"in_param{i}"]).to(DEV)) for i in range(n)]
    opt = optim.Adam([dict(params=[p], lr=float(lr), name=f"g{i}") for i, (p, lr) in enumerate(zip(ps, g["in_lrs"]))],
                     lr=0.0, eps=float(g["in_eps"]), betas=tuple(float(b) for b in g["in_betas"]))
    for s in range(int(g["in_steps"])):
        for i, p in enumerate(ps):
            p.grad = torch.from_numpy(g[f"in_grad{i}_s{s}"]).to(DEV)
        opt.step()
        for i, p in enumerate(ps):
            st = opt.state[p]
            for got, key in ((p, "param"), (st["exp_avg"], "exp_avg"), (st["exp_avg_sq"], "exp_avg_sq")):
                want = g[f"{key}{i}_s{s}"]
                assert np.array_equal(got.detach().cpu().numpy().view(np.uint32), want.view(np.uint32)), (key, i, s)
            assert float(st["step"]) == s + 1


def _groups(seed, scale=1):
    """the reference's parameter groups at a small scale (scene/gaussian_model.py:351-388)"""
    g = torch.Generator().manual_seed(seed)
    A = 5000 * scale
    shapes = dict(anchor=(A, 3), offset=(A, 10, 3), anchor_feat=(A, 32), opacity=(A, 1), scaling=(A, 6), rotation=(A, 4),
                  mlp_w1=(32, 35), mlp_b1=(32,), mlp_w2=(7, 32), mlp_b2=(7,))
    lrs = dict(anchor=0.0, offset=0.01, anchor_feat=0.0075, opacity=0.02, scaling=0.007, rotation=0.002, mlp_w1=0.002,
               mlp_b1=0.002, mlp_w2=0.004, mlp_b2=0.004)
    return {k: torch.randn(s, generator=g) for k, s in shapes.items()}, lrs, g


def test_live_against_torch_adam_with_state_surgery_and_state_dict():
    from lgs_b200 import optim
    init, lrs, gen = _groups(3)
    mk = lambda cls: (lambda ps: (ps, cls([dict(params=[ps[k]], lr=lrs[k], name=k) for k in ps], lr=0.0, eps=1e-15)))(
        {k: torch.nn.Parameter(v.clone().to(DEV)) for k, v in init.items()})
    (pa, oa), (pb, ob) = mk(optim.Adam), mk(torch.optim.Adam)

    def step_both(skip=()):
        for k in pa:
            if k in skip:
                pa[k].grad = pb[k].grad = None      # a parameter without a gradient is left alone, its step does not advance
                continue
            gr = torch.randn(pa[k].shape, generator=gen).to(DEV)
            pa[k].grad, pb[k].grad = gr.clone(), gr.clone()
        oa.step()
        ob.step()
        for k in pa:
            assert same_bits(pa[k], pb[k]), k
            if pa[k] in oa.state:
                assert same_bits(oa.state[pa[k]]["exp_avg"], ob.state[pb[k]]["exp_avg"]), k
                assert same_bits(oa.state[pa[k]]["exp_avg_sq"], ob.state[pb[k]]["exp_avg_sq"]), k
                assert float(oa.state[pa[k]]["step"]) == float(ob.state[pb[k]]["step"]), k

    step_both()
    step_both(skip=("rotation",))
    for grp in oa.param_groups:                       # the scheduler's per-step lr update (gaussian_model.py:437-470)
        grp["lr"] *= 0.5
    for grp in ob.param_groups:
        grp["lr"] *= 0.5
    step_both()

    # densification: prune half of the anchors and append new ones, state rebuilt as _prune_anchor_optimizer /
    # cat_tensors_to_optimizer do it (scene/gaussian_model.py:567-650)
    def surgery(ps, opt):
        for grp in opt.param_groups:
            if grp["name"].startswith("mlp"):
                continue
            old = grp["params"][0]
            st = opt.state.pop(old)
            keep = torch.arange(old.shape[0], device=DEV) % 2 == 0
            ext = torch.zeros((100,) + tuple(old.shape[1:]), device=DEV)
            new = torch.nn.Parameter(torch.cat([old.detach()[keep], ext + 0.25]).contiguous())
            st["exp_avg"] = torch.cat([st["exp_avg"][keep], torch.zeros_like(ext)])
            st["exp_avg_sq"] = torch.cat([st["exp_avg_sq"][keep], torch.zeros_like(ext)])
            grp["params"][0] = new
            opt.state[new] = st
            ps[grp["name"]] = new
    surgery(pa, oa)
    surgery(pb, ob)
    step_both()

    # a torch.optim.Adam loads our state and continues exactly like we do (checkpoints stay interchangeable)
    import copy
    sd = copy.deepcopy(oa.state_dict())      # load_state_dict() does not copy: without this the two would share moment tensors
    oc = torch.optim.Adam([dict(params=[torch.nn.Parameter(pa[k].detach().clone())], lr=0.0, name=k) for k in pa], lr=0.0, eps=1e-15)
    oc.load_state_dict(sd)
    pc = {grp["name"]: grp["params"][0] for grp in oc.param_groups}
    for k in pa:
        gr = torch.randn(pa[k].shape, generator=gen).to(DEV)
        pa[k].grad, pc[k].grad = gr.clone(), gr.clone()
    oa.step()
    oc.step()
    for k in pa:
        assert same_bits(pa[k], pc[k]), k


def test_many_tensors_unaligned_views_and_errors():
    from lgs_b200 import optim
    g = torch.Generator().manual_seed(11)
    # more tensors than one launch takes, odd sizes, and parameters that are 4-byte-aligned views into one flat buffer
    sizes = [1, 3, 4097, 5, 8191, 2, 7] * 9
    flat_a = torch.randn(sum(sizes) + 1, generator=g).to(DEV)
    flat_b = flat_a.clone()
    pa, pb, o = [], [], 1
    for n in sizes:
        pa.append(torch.nn.Parameter(flat_a[o:o + n]))
        pb.append(torch.nn.Parameter(flat_b[o:o + n]))
        o += n
    oa, ob = optim.Adam(pa, lr=0.01, eps=1e-15), torch.optim.Adam(pb, lr=0.01, eps=1e-15)
    for _ in range(3):
        for x, y in zip(pa, pb):
            gr = torch.randn(x.shape, generator=g).to(DEV)
            x.grad, y.grad = gr.clone(), gr.clone()
        oa.step()
        ob.step()
    assert same_bits(flat_a, flat_b)
    with pytest.raises(RuntimeError):
        p = torch.nn.Parameter(torch.zeros(4))      # CPU parameter: no fallback
        p.grad = torch.zeros(4)
        optim.Adam([p], lr=0.1).step()
    with pytest.raises(NotImplementedError):
        p = torch.nn.Parameter(torch.zeros(4, device=DEV))
        p.grad = torch.zeros(4, device=DEV)
        optim.Adam([p], lr=0.1, weight_decay=0.1).step()
    p = torch.nn.Parameter(torch.ones(4, device=DEV))
    optim.Adam([p], lr=0.1).step()                   # nothing has a gradient: a no-op
    assert float(p.detach().sum()) == 4.0


def test_more_than_one_launch_worth_of_tensors_with_empty_ones():
    """Host batching of lgs_adam_step: empty tensors are skipped inside a batch of 48, so a batch can consume more than 48
    table entries; the next batch must start where the previous one stopped, not 48 entries further (which would apply the
    update twice to the tensors in between)."""
    from lgs_b200 import optim
    g = torch.Generator().manual_seed(12)
    sizes = ([0, 5, 0, 9] * 13 + [17] * 30 + [0, 3] * 10)  # 102 entries, 26 + ... empties inside the first batches
    pa = [torch.nn.Parameter(torch.randn(n, generator=g).to(DEV)) for n in sizes]
    pb = [torch.nn.Parameter(x.detach().clone()) for x in pa]
    oa, ob = optim.Adam(pa, lr=0.01, eps=1e-15), torch.optim.Adam(pb, lr=0.01, eps=1e-15)
    for _ in range(2):
        for x, y in zip(pa, pb):
            gr = torch.randn(x.shape, generator=g).to(DEV)
            x.grad, y.grad = gr.clone(), gr.clone()
        oa.step()
        ob.step()
    for i, (x, y) in enumerate(zip(pa, pb)):
        assert same_bits(x.detach(), y.detach()), i
