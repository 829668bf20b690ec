import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import rel_l2, load_circuit_cases
from quanonet_b200.ops import hea_expval, hea_expval_backward
from oracle import hea_oracle as orc
dev = torch.device("cuda:0")
z, meta = load_circuit_cases()
tag = "c2_q5_net40"
blocks = [tuple(b) for b in meta[tag]["blocks"]]; depths = [d for _, d in blocks]
x = torch.tensor(z[tag + "/x"], device=dev); w = torch.tensor(z[tag + "/w"], device=dev); g = torch.tensor(z[tag + "/g"], device=dev)
o, gx, gw = hea_expval_backward(g, x, w, 5, depths, None, 0, 0.0, 1.0, 0, True)
print("c2 synthetic: out", rel_l2(o[:, 0].cpu().numpy(), z[tag + "/e"]), "gx", rel_l2(gx.cpu().numpy(), z[tag + "/gx"]),
      "gw", rel_l2(gw.cpu().numpy(), z[tag + "/gw"]))
o2 = hea_expval(x, w, 5, depths, None, 0, 0.0, 1.0, 0)
print("fwd-only out", rel_l2(o2[:, 0].cpu().numpy(), z[tag + "/e"]))
# bigger random batch vs oracle + TQ-faithful
rng = np.random.default_rng(5)
B = 256
xb = rng.uniform(-np.pi, np.pi, (B, 300)).astype(np.float32); gb = rng.standard_normal(B).astype(np.float32)
e, egx, egw = orc.hea_forward_backward(xb, z[tag + "/w"], 5, blocks, orc.ham_from_bound(5), grad_out=gb)
o, gx, gw = hea_expval_backward(torch.tensor(gb, device=dev), torch.tensor(xb, device=dev), w, 5, depths, None, 0, 0.0, 1.0, 0, True)
print("B=256 ours : out", rel_l2(o[:, 0].cpu().numpy(), e), "gx", rel_l2(gx.cpu().numpy(), egx), "gw", rel_l2(gw.cpu().numpy(), egw))
from oracle.tq_faithful import tq_forward_backward
to, tgx, tgw = tq_forward_backward(torch.tensor(xb), torch.tensor(z[tag + "/w"]), 5, blocks, torch.tensor(gb), ham_offset=0.0, ham_coeff=1.0)
print("B=256 TQf32: out", rel_l2(to[:, 0].numpy(), e), "gx", rel_l2(tgx.numpy(), egx), "gw", rel_l2(tgw.numpy(), egw))
print("ours vs TQf32: out", rel_l2(o[:, 0].cpu().numpy(), to[:, 0].numpy()), "gx", rel_l2(gx.cpu().numpy(), tgx.numpy()), "gw", rel_l2(gw.cpu().numpy(), tgw.numpy()))
# sincos accuracy
t = torch.tensor(xb[0], device=dev)
