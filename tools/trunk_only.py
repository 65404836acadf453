import sys, os, torch
sys.path.insert(0, "/root/repo")
import pcoe
dev = torch.device("cuda:0")
torch.manual_seed(0)
B = 64
model = pcoe.PointNetPPMvM(sampler="randperm_device").to(dev).train()
feat = torch.randn(B, 1024, device=dev)
model._sa_features = lambda xyz: (feat + 0.0 * xyz.sum())       # constant SA3 output: trunk + head + loss only
engine = pcoe.dp.DataParallel(model)
opt = pcoe.optim.FusedAdam(engine, lr=1e-3, max_grad_norm=1.0, zero_grad_in_step=True)
gt, K = pcoe.synthetic.mvm_targets(B, 0)
gt, K = gt.to(dev), K.to(torch.int32).to(dev)
xyz = pcoe.synthetic.clouds(1, B, 1024, 0).to(dev)
g = pcoe.GraphedTrainStep(model, lambda res, gt, K: pcoe.match_loss(res[0], res[1], res[2], gt, None, K).mean(), opt, xyz, (gt, K), clip_norm=1.0, engine=engine, warmup=2)
for _ in range(5): g(xyz, gt, K)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200): g.graph.replay()
e1.record(); torch.cuda.synchronize()
print("trunk+head+loss+adam graph: %.1f us/step" % (e0.elapsed_time(e1) / 200 * 1e3))
