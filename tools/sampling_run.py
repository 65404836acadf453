"""One launch of every sampling / grouping / gather kernel at the c2 (64 x 1024) and c4 (32 x 8192) shapes, for an
`ncu --set full` capture (profiles/r02_sampling_ncu.txt):
    ncu --set full --clock-control none -k regex:"fps_|knn_|ball_|gather_|random_subset|square_|resample_" -o out python tools/sampling_run.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pcoe
dev = torch.device("cuda:0")
for B, N in ((64, 1024), (32, 8192)):
    xyz = pcoe.synthetic.clouds(7, B, N).to(dev)
    start = torch.zeros(B, dtype=torch.int32, device=dev)
    for rep in range(2):                      # first pass warms up (attribute setting, L2); ncu takes the second (--launch-skip)
        idx, new_xyz = pcoe.ops.farthest_point_sample(xyz, 128, start, return_xyz=True, int32=True)
        nbr = pcoe.ops.knn_int32(new_xyz, xyz, 32)
        pcoe.ops.ball_query_int32(0.2, 32, xyz, new_xyz)
        pcoe.ops.ball_query_multi_int32([0.1, 0.2, 0.4], [16, 32, 128], xyz, new_xyz)
        pcoe.ops.gather_points(xyz, idx)
        pcoe.ops.gather_points(xyz, nbr.reshape(B, -1))
        pcoe.ops.random_subset(B, N, 128, 1, 0, dev, xyz=xyz)
        pcoe.square_distance(new_xyz, xyz)
        torch.cuda.synchronize()
print("done")
