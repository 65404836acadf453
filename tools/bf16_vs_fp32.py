"""How far the bf16 tensor-core mode is from the fp32 parity mode on the bench workload (mvM, 64 x 1024):
same weights, same batch, same host-generator subsets and dropout masks; prints loss and gradient deviations."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pcoe
dev = torch.device("cuda:0")
B, N = 64, 1024
torch.manual_seed(1000)
m32 = pcoe.PointNetPPMvM(precision="fp32").to(dev).train()
m16 = pcoe.PointNetPPMvM(precision="bf16").to(dev).train()
m16.load_state_dict(m32.state_dict())
xyz = pcoe.synthetic.clouds(1, B, N, 0).to(dev)
gt, K = pcoe.synthetic.mvm_targets(B, 0)
gt, K = gt.to(dev), K.to(device=dev, dtype=torch.int32)
res = []
for m in (m32, m16):
    torch.manual_seed(42)
    mu, kap, w = m(xyz)
    loss = pcoe.match_loss(mu, kap, w, gt, None, K).mean()
    loss.backward()
    res.append((float(loss), {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}))
(l32, g32), (l16, g16) = res
print(f"loss fp32 {l32:.6f} bf16 {l16:.6f} rel {abs(l16 - l32) / abs(l32):.2e}")
tot32 = torch.cat([g.flatten() for g in g32.values()]); tot16 = torch.cat([g16[n].flatten() for n in g32])
print(f"all gradients: rel L2 {float((tot16 - tot32).norm() / tot32.norm()):.2e}, norm ratio {float(tot16.norm() / tot32.norm()):.4f}, "
      f"cosine {float(torch.dot(tot16, tot32) / (tot16.norm() * tot32.norm())):.5f}")
for pre in ("sa1", "sa2", "sa3", "fc", "head"):
    a = torch.cat([g32[n].flatten() for n in g32 if n.startswith(pre)]); b = torch.cat([g16[n].flatten() for n in g32 if n.startswith(pre)])
    print(f"  {pre:5s} rel L2 {float((b - a).norm() / a.norm()):.2e}  cosine {float(torch.dot(a, b) / (a.norm() * b.norm())):.5f}")
