"""debug: graph/_smallnet.py operators one by one against torch.nn.functional on the GPU"""
import importlib, os, sys
import torch, torch.nn as nn, torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
S = importlib.import_module("musicgeneration_vae-torch_b200.graph._smallnet")
torch.manual_seed(0)
dev = "cuda"

def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))

def check_conv(cin, cout, k, s, p, H, W, act=False, transposed=False, bias=False):
    m = (nn.ConvTranspose2d if transposed else nn.Conv2d)(cin, cout, k, s, p, bias=bias).to(dev)
    x = torch.randn(3, cin, H, W, device=dev).to(torch.bfloat16).float().requires_grad_(True)
    with torch.no_grad():
        m.weight.copy_(m.weight.to(torch.bfloat16).float())
    ref = m(x)
    if act: ref = F.relu(ref)
    g = torch.randn_like(ref).to(torch.bfloat16).float()
    ref.backward(g)
    rw, rx = m.weight.grad.clone(), x.grad.clone()
    m.weight.grad = None
    x2 = x.detach().permute(0, 2, 3, 1).contiguous().requires_grad_(True)
    out = S.conv(x2, m, act=act, out_f32=not act)
    out.backward(g.permute(0, 2, 3, 1).contiguous().to(out.dtype))
    print("conv %d->%d k%s s%s T=%s act=%s: out %.2e dw %.2e dx %.2e" % (cin, cout, k, s, transposed, act,
          rel(out.permute(0, 3, 1, 2).float(), ref), rel(m.weight.grad, rw), rel(x2.grad.permute(0, 3, 1, 2), rx)))

def check_bn(C, H, W, act, training, xf32):
    bn = nn.BatchNorm2d(C, momentum=0.01).to(dev).train(training)
    with torch.no_grad():
        bn.weight.normal_(1, 0.3); bn.bias.normal_(0, 0.3); bn.running_mean.normal_(0, 0.2); bn.running_var.uniform_(0.5, 1.5)
    import copy
    bn2 = copy.deepcopy(bn)
    x = (torch.randn(4, C, H, W, device=dev) * 2 + 3)
    if not xf32: x = x.to(torch.bfloat16).float()
    x.requires_grad_(True)
    ref = bn(x)
    if act: ref = F.relu(ref)
    g = torch.randn_like(ref).to(torch.bfloat16).float()
    ref.backward(g)
    x2 = x.detach().permute(0, 2, 3, 1).contiguous()
    if not xf32: x2 = x2.to(torch.bfloat16)
    x2.requires_grad_(True)
    out = S.batch_norm(x2, bn2, act=act)
    out.backward(g.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16))
    print("bn C%d act=%s train=%s f32=%s: out %.2e dx %.2e dgamma %.2e dbeta %.2e rm %.2e rv %.2e" % (C, act, training, xf32,
          rel(out.permute(0, 3, 1, 2).float(), ref), rel(x2.grad.permute(0, 3, 1, 2).float(), x.grad), rel(bn2.weight.grad, bn.weight.grad),
          rel(bn2.bias.grad, bn.bias.grad), rel(bn2.running_mean, bn.running_mean), rel(bn2.running_var, bn.running_var)))

for args in [(1, 8, (3, 1), (2, 1), (1, 0), 192, 12), (8, 16, (3, 1), (2, 1), (1, 0), 96, 12), (16, 16, 1, 1, 0, 48, 12),
             (16, 32, 3, 2, 1, 48, 12), (32, 64, 3, 2, 1, 24, 6), (1, 8, 3, (2, 1), 1, 192, 1), (8, 8, 3, (2, 1), 1, 96, 1),
             (1, 8, (1, 4), (1, 2), (0, 1), 192, 60), (8, 8, (4, 1), (2, 1), (1, 0), 192, 30), (1, 8, (4, 1), (2, 1), (1, 0), 192, 60),
             (16, 8, 1, 1, 0, 96, 30), (8, 8, 3, 1, 1, 96, 30), (8, 16, 3, 2, 1, 96, 30), (1, 2, 4, 1, 2, 96, 60), (2, 8, 4, 1, 2, 48, 30)]:
    check_conv(*args)
    check_conv(*args, act=True)
check_conv(8, 2, 4, 2, 1, 24, 15, transposed=True)
check_conv(2, 1, 4, 2, 1, 48, 30, transposed=True)
for C in (1, 2, 8, 16, 64):
    for act in (False, True):
        for training in (True, False):
            check_bn(C, 24, 15, act, training, True)
    check_bn(C, 24, 15, False, True, False)
