"""The call sequences of the reference's trainers, replayed on the GPU against this package (the GPU box has no
reference checkout; tests/test_reference_trainers.py drives the real trainer classes where it exists).

QATTrainer._train_epoch (trainer/quantization_aware_training_trainer.py:31-111): model.train(); every
`update_qparams_period` steps `model.apply(reset_qparams)`; zero_grad, forward, criterion, backward, optional
clip_grad_norm_, optimizer.step, lr_scheduler.step, loss.item(); validation under eval() + no_grad().
FSPTQTrainer.train (trainer/fsptq_trainer.py:28-112): per block, cache the quantised model's block input and the
full-precision block output over the calibration set, fit the block with Adam on 64-sample mini-batches."""
import copy

import pytest
import torch
from torch import nn

pytestmark = pytest.mark.gpu

CFG = {"weight": {"enable": True, "type": "minmax_channel", "args": {"n_bits": 4, "signed": True, "ch_axis": 0}},
       "input": {"enable": True, "type": "minmax_tensor", "args": {"n_bits": 4, "signed": False}},
       "exclude_layers": [], "override_options": [], "momentum": 0.1}


class Net(nn.Module):
    def __init__(self, classes=10):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 16, 3, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(16)
        self.block = nn.Sequential(nn.Conv2d(16, 16, 3, padding=1, bias=False), nn.BatchNorm2d(16), nn.ReLU(),
                                   nn.Conv2d(16, 32, 3, padding=1, stride=2, bias=False), nn.BatchNorm2d(32), nn.ReLU())
        self.linear = nn.Linear(32, classes)

    def forward(self, x):
        x = torch.relu(self.bn1(self.conv1(x)))
        return self.linear(self.block(x).mean((2, 3)))


def _data(n_batches, batch, seed):
    proto = torch.randn(10, 3, 16, 16, generator=torch.Generator().manual_seed(1234))     # the classes: same in every split
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(n_batches):
        t = torch.randint(0, 10, (batch,), generator=g)
        out.append(((proto[t] + 0.3 * torch.randn(batch, 3, 16, 16, generator=g)).cuda(), t.cuda()))
    return out


@pytest.mark.parametrize("fused", [False, True])
def test_qat_trainer_call_sequence(fused):
    from dlmc_quant_b200 import quantize_model
    from dlmc_quant_b200.fuse import fuse_bn_act_quant
    from dlmc_quant_b200.quantize import group_weight_quantizers
    from dlmc_quant_b200.scalar.modules.base import QBase
    torch.manual_seed(2333)
    model = Net().cuda()
    if fused:
        model = model.to(memory_format=torch.channels_last)
    quantize_model(model, copy.deepcopy(CFG), None)
    optimizer = torch.optim.SGD(model.parameters(), lr=0.05, momentum=0.9, nesterov=True)      # built BEFORE any forward
    scheduler = torch.optim.lr_scheduler.StepLR(optimizer, 50, 0.5)
    criterion = nn.CrossEntropyLoss()
    if fused:
        group_weight_quantizers(model)
        fuse_bn_act_quant(model)
    loader, valid = _data(12, 32, 1), _data(2, 32, 2)
    period, grad_clip = 5, 5.0
    losses, scales_seen = [], []
    for epoch in range(1, 4):
        model.train()
        for batch_idx, (data, target) in enumerate(loader):
            if (epoch * len(loader) + batch_idx) % period == 1:                 # qat_trainer.py:44-48
                model.apply(lambda m: m.reset_qparams() if hasattr(m, "reset_qparams") else None)
            if fused:
                data = data.contiguous(memory_format=torch.channels_last)
            optimizer.zero_grad()
            output = model(data)
            loss = criterion(output, target)
            loss.backward()
            torch.nn.utils.clip_grad.clip_grad_norm_(model.parameters(), grad_clip)
            optimizer.step()
            scheduler.step()
            losses.append(loss.item())
            scales_seen.append(float(model.block[0].in_scale.detach().reshape(-1)[0]))
    assert all(torch.isfinite(torch.tensor(losses)))
    assert sum(losses[-6:]) / 6 < 0.7 * sum(losses[:6]) / 6, (losses[:6], losses[-6:])       # it trains
    qs = [m for m in model.modules() if isinstance(m, QBase)]
    assert all(m._host_init == {"in": True, "wt": True} and float(m.in_init_state) == 1 for m in qs)
    assert all(p.grad is not None for n, p in model.named_parameters() if n.endswith("scale"))   # scales are trained
    assert len(set(round(s, 9) for s in scales_seen)) > 10                                        # ... and move
    # validation (qat_trainer.py:113-133) and a checkpoint round trip in the middle of training
    model.eval()
    with torch.no_grad():
        acc = []
        for data, target in valid:
            if fused:
                data = data.contiguous(memory_format=torch.channels_last)
            acc.append(float((model(data).argmax(1) == target).float().mean()))
    assert sum(acc) / len(acc) > 0.5
    clone = Net().cuda()
    if fused:
        clone = clone.to(memory_format=torch.channels_last)
    quantize_model(clone, copy.deepcopy(CFG), None)
    clone.load_state_dict(model.state_dict(), strict=True)
    clone.eval()
    with torch.no_grad():
        d = valid[0][0].contiguous(memory_format=torch.channels_last) if fused else valid[0][0]
        assert torch.allclose(clone(d), model(d), rtol=1e-4, atol=1e-5)


def test_fsptq_trainer_call_sequence_matches_the_naive_procedure():
    """recon.FSPTQReconstructor (fp outputs of all blocks cached in one pass, quantised pass stopped at the block,
    caches in HBM) against the reference's procedure restated literally (both models run to the end for every block,
    caches moved through the CPU): identical caches, same fitted parameters given the same mini-batch order."""
    from dlmc_quant_b200 import quantize_model
    from dlmc_quant_b200.recon import FSPTQReconstructor, l2_loss
    from dlmc_quant_b200.scalar.FSPTQuant.base import FSPTQBase
    cfg = copy.deepcopy(CFG)
    cfg["weight"]["args"]["n_bits"] = cfg["input"]["args"]["n_bits"] = 8
    torch.manual_seed(7)
    fp = Net().cuda().eval()
    batches = [b for b, _ in _data(4, 32, 3)]

    def build():
        m = copy.deepcopy(fp)
        quantize_model(m, copy.deepcopy(cfg), None, quantization_type="FSPTQ")
        return m
    ours, naive = build(), build()
    epochs = 30
    rec = FSPTQReconstructor(ours, fp, block_types=(nn.Sequential,), epochs=epochs)
    g1 = torch.Generator().manual_seed(11)
    hist = rec.run(batches, generator=g1)
    assert list(hist) == ["conv1", "block", "linear"]

    # the reference procedure, literally (fsptq_trainer.py:36-100), on the second model
    g2 = torch.Generator().manual_seed(11)
    for (name, module), fp_module in zip(naive.named_modules(), fp.modules()):
        if not ((isinstance(module, FSPTQBase) and name in ["conv1", "linear"]) or type(module) is nn.Sequential):
            continue
        cin, cout = [], []
        h1 = fp_module.register_forward_hook(lambda m, i, o: cout.append(o.cpu()))
        h2 = module.register_forward_hook(lambda m, i, o: cin.append(i[0].cpu()))
        naive.eval()
        for data in batches:
            with torch.no_grad():
                fp(data); naive(data)
        block_input, block_output = torch.cat(cin).cuda(), torch.cat(cout).cuda()
        h1.remove(); h2.remove()
        opt, sched = rec.generate_optimizer(module)
        naive.train()
        for i in range(epochs):
            idx = torch.randperm(block_input.size(0), generator=g2)[:64].cuda()
            opt.zero_grad()
            loss = l2_loss(block_output[idx], module(block_input[idx]))
            loss.backward()
            opt.step(); sched.step()
    # same caches, same mini-batch order, same optimizer: the fitted parameters agree up to what Adam makes of cuDNN's
    # run-to-run gradient noise (every Adam step moves a weight by at most its learning rate, 1e-5, whatever the gradient)
    for (n, p), (_, q) in zip(ours.named_parameters(), naive.named_parameters()):
        assert torch.allclose(p, q, rtol=1e-3, atol=2 * epochs * 1e-5), n
    for name in hist:
        assert hist[name][-1] <= hist[name][0] * 1.05 + 1e-9, (name, hist[name])
