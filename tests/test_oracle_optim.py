"""CPU: the optimizer oracle against torch.optim.AdamW (pin), internal consistency of the pytorch-transformers 1.x
restatement, and the schedules against transformers' current equivalents."""
import torch

from oracle import optim_oracle as OO


def test_torch_semantics_oracle_matches_torch_adamw():
    torch.manual_seed(0)
    p0 = torch.randn(37, 11)
    grads = [torch.randn(37, 11) * (0.1 + i) for i in range(5)]
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([ref], lr=3e-3, betas=(0.9, 0.98), eps=1e-8, weight_decay=0.05)
    p, state = p0.clone(), {}
    for g in grads:
        ref.grad = g.clone()
        opt.step()
        OO.adamw_torch(p, g, state, 3e-3, (0.9, 0.98), 1e-8, 0.05)
        assert (p - ref.detach()).abs().max().item() < 1e-6


def test_hf1_restatement_consistency():
    # with weight_decay = 0 and eps -> 0 the two algorithms coincide (the eps placement is the only other difference)
    torch.manual_seed(1)
    p0 = torch.randn(64)
    a, b, sa, sb = p0.clone(), p0.clone(), {}, {}
    for i in range(4):
        g = torch.randn(64)
        OO.adamw_hf1(a, g, sa, 1e-2, (0.9, 0.999), 0.0, 0.0, True)
        OO.adamw_torch(b, g, sb, 1e-2, (0.9, 0.999), 0.0, 0.0)
        assert (a - b).abs().max().item() < 1e-6
    # decay is applied AFTER the update in the 1.x optimizer: p1 = (p0 - lr*sign-ish) * (1 - lr*wd)
    p, st = torch.ones(3), {}
    g = torch.tensor([1.0, -2.0, 0.5])
    OO.adamw_hf1(p, g, st, 0.1, (0.9, 0.999), 1e-6, 0.5, True)
    # first step, bias-corrected: p - lr * g / (|g| + eps / sqrt(1 - b2)), then the decay
    want = 1.0 - 0.1 * g / (g.abs() + 1e-6 / 0.001 ** 0.5)
    assert torch.allclose(p, want * (1 - 0.1 * 0.5), atol=1e-5)
    # correct_bias=False: the first step is lr * (1-b1) g / (sqrt((1-b2) g^2) + eps)
    p, st = torch.zeros(1), {}
    OO.adamw_hf1(p, torch.tensor([2.0]), st, 1.0, (0.9, 0.999), 0.0, 0.0, False)
    assert abs(p.item() + 0.1 * 2.0 / (0.001 ** 0.5 * 2.0)) < 1e-5


def test_schedules_match_transformers_successors():
    from transformers.optimization import get_constant_schedule_with_warmup, get_linear_schedule_with_warmup
    from cpt_b200.optimization import WarmupConstantSchedule, WarmupLinearSchedule, get_lr_sched, warmup_linear

    def lrs(make):
        p = torch.nn.Parameter(torch.zeros(1))
        opt = torch.optim.SGD([p], lr=1.0)
        sch = make(opt)
        out = []
        for _ in range(30):
            out.append(opt.param_groups[0]["lr"])
            opt.step()
            sch.step()
        return out

    a = lrs(lambda o: WarmupLinearSchedule(o, warmup_steps=5, t_total=25))
    b = lrs(lambda o: get_linear_schedule_with_warmup(o, 5, 25))
    assert max(abs(x - y) for x, y in zip(a, b)) < 1e-12
    a = lrs(lambda o: WarmupConstantSchedule(o, warmup_steps=7))
    b = lrs(lambda o: get_constant_schedule_with_warmup(o, 7))
    assert max(abs(x - y) for x, y in zip(a, b)) < 1e-12

    class Opts:
        learning_rate, warmup_steps, num_train_steps = 2e-5, 10, 100
    assert get_lr_sched(5, Opts) == 2e-5 * 0.5 and get_lr_sched(100, Opts) == 1e-8
    assert warmup_linear(55, 10, 100) == 0.5
