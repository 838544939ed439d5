"""Debug: run the h=300 GCN stack and compare every tcgen05 GEMM call against an fp64 product of ITS OWN inputs."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graph_hscn_b200 import gemm, pyg, synthetic

def rel(a, r):
    return float((a.double() - r).abs().max() / r.abs().max().clamp_min(1e-30))

orig_tn, orig_nn, orig_prep = gemm.gemm3x_tn, gemm.gemm3x, gemm.gemm3x_prep
images = {}
def prep(w, transpose=False):
    img = orig_prep(w, transpose)
    images[img.data_ptr()] = (w.detach().clone(), transpose)
    return img
def nn(a, image, n_out, bias=None, relu=False):
    c = orig_nn(a, image, n_out, bias, relu)
    w, tr = images[image.data_ptr()]
    ref = a.double() @ (w.double() if tr else w.double().t())
    if bias is not None: ref = ref + bias.double()
    print(f"gemm3x    a={tuple(a.shape)} stride={a.stride()} tr={tr}: err={rel(c, ref):.3e}  |a|max={float(a.abs().max()):.3g} fp32err={rel(a @ (w if tr else w.t()) + (0 if bias is None else bias), ref):.3e}")
    return c
def tn(p, q):
    out = orig_tn(p, q)
    ref = p.double().t() @ q.double()
    print(f"gemm3x_tn p={tuple(p.shape)} {p.stride()} q={tuple(q.shape)} {q.stride()}: err={rel(out, ref):.3e} fp32err={rel(p.t() @ q, ref):.3e} "
          f"|p|max={float(p.abs().max()):.3g} |q|max={float(q.abs().max()):.3g} |ref|max={float(ref.abs().max()):.3g} sum|p||q|max={float((p.abs().double().t() @ q.abs().double()).max()):.3g}")
    return out
gemm.gemm3x_tn, gemm.gemm3x, gemm.gemm3x_prep = tn, nn, prep

dev = torch.device("cuda")
p_ns = pyg.namespace()
b = synthetic.peptides_batch(40, seed=90)
torch.manual_seed(3)
layers = [p_ns.GCNConv(9, 300, add_self_loops=False).to(dev), p_ns.GCNConv(300, 300, add_self_loops=False).to(dev),
          p_ns.GCNConv(300, 300, add_self_loops=False).to(dev)]
x = b.x.float().to(dev)
ei = b.edge_index.to(dev)
for l in layers:
    x = torch.relu(l(x, ei))
x.sum().backward()
