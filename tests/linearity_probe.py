"""Developer probe: gradient of a batch of two vs the two single-image gradients (mean loss and sum loss)."""
import importlib, sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jpdse_b200
nw = importlib.import_module("jpd-se_b200.ctu.models.pix2pixHD_networks.networks")
cuda = torch.device('cuda')
def cos(a,b):
    a,b=a.double().flatten(),b.double().flatten(); return float((a*b).sum()/(a.norm()*b.norm()+1e-30))
def run(n_down, n_blocks, H, W, scale1):
    torch.manual_seed(1234)
    net = nw.define_G(39, 3, 64, "global", n_down, n_blocks, 1, 3, "instance", gpu_ids=[0]).train()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 39, H, W, generator=g).to(cuda); tgt = (torch.rand(2, 3, H, W, generator=g) - 0.5).to(cuda)
    def grads(xs, ts, s):
        for p in net.parameters(): p.grad = None
        y = net(xs)
        (s * (y - ts).abs().sum()).backward()
        return y.detach().clone(), [p.grad.clone() for p in net.parameters()]
    yb, both = grads(x, tgt, 1e-6)
    y0, g0 = grads(x[:1].contiguous(), tgt[:1].contiguous(), scale1)
    y1, g1 = grads(x[1:].contiguous(), tgt[1:].contiguous(), scale1)
    print("config", n_down, n_blocks, H, W, "single-image loss scale", scale1, "forward equal:", torch.equal(yb[:1], y0), torch.equal(yb[1:], y1))
    for (name,_), a, b0, b1 in zip(net.named_parameters(), both, g0, g1):
        m = (b0+b1) * (1e-6 / scale1)
        if float(m.abs().max())==0: continue
        print("  %-32s lin %.7f" % (name, cos(a,m)))
run(4, 1, 256, 512, 1e-6)
run(4, 1, 256, 512, 2e-6)
