"""Developer probe: InstanceNorm2d(affine=False) on channels_last fp32 discriminator features -- stock instance_norm
(which round-trips through NCHW copies) vs group_norm(C groups) vs a var_mean formulation; forward + backward."""
import torch, torch.nn.functional as F

def t(fn, x, n=20):
    for _ in range(3):
        y = fn(x); y.backward(torch.ones_like(y)); x.grad = None
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        y = fn(x); y.backward(torch.ones_like(y)); x.grad = None
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

def vm(x):
    var, mean = torch.var_mean(x, dim=(2, 3), keepdim=True, correction=0)
    return (x - mean) * torch.rsqrt(var + 1e-5)

for shape in [(2, 128, 129, 257), (2, 256, 65, 129), (2, 512, 66, 130)]:
    for cl in (True, False):
        x = torch.randn(shape, device="cuda")
        if cl:
            x = x.contiguous(memory_format=torch.channels_last)
        x.requires_grad_(True)
        ref = F.instance_norm(x)
        r = {"instance_norm": t(F.instance_norm, x), "group_norm": t(lambda v: F.group_norm(v, v.shape[1]), x), "var_mean": t(vm, x)}
        err = {"group_norm": float((F.group_norm(x, shape[1]) - ref).abs().max()), "var_mean": float((vm(x) - ref).abs().max())}
        print(shape, "channels_last" if cl else "nchw", {k: round(v, 3) for k, v in r.items()}, err,
              "gn out cl:", F.group_norm(x, shape[1]).is_contiguous(memory_format=torch.channels_last))
