"""Bisect the chained training forward (DenseNetworkTrainer) against the float64 oracle, stage by stage."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn.functional as F
from lisec_b200.train import DenseNetworkTrainer
from lisec_b200.weights import synthetic_network_pack, conv3d_blocks, rpn_blocks
from oracle import train_oracle as TO

nx, ny, B = 24, 40, 2
pack = {k: np.asarray(v, dtype=np.float32) for k, v in synthetic_network_pack(3).items()}
for k in pack:
    if k.endswith("/kernel"):
        pack[k] = torch.from_numpy(pack[k]).to(torch.bfloat16).float().numpy()
g = torch.Generator(device="cpu").manual_seed(43)
grid = torch.rand((B, 8, nx, ny, 64), generator=g).to(torch.bfloat16)
net = DenseNetworkTrainer(pack, B, nx, ny)
net.forward(grid.cuda())
torch.cuda.synchronize()
p = TO.to_params(pack)


def rel(got, want):
    got, want = got.double().cpu(), want.double()
    return float(((got - want) ** 2).sum().sqrt() / (want ** 2).sum().sqrt())


with torch.no_grad():
    x = grid.double().permute(0, 4, 1, 2, 3)
    stats = {}
    for (st, conv, bn, dense), (_, _, _, stride, pad) in zip(net.c3, conv3d_blocks()):
        w = p[conv + "/kernel"].permute(4, 3, 0, 1, 2)
        z = F.conv3d(x, w, p[conv + "/bias"], stride=stride, padding=pad)
        print("%-12s conv out z   rel-L2 %.4f" % (conv, rel(st.z.permute(0, 4, 1, 2, 3), z)))
        u = TO._bn_train(z, p, bn, stats, channels_last=False)
        print("%-12s BN out       rel-L2 %.4f   mean err %.2e invstd rel err %.2e" % (
            conv, rel(st.bn.y.permute(0, 4, 1, 2, 3), u),
            float((st.bn.mean.cpu().double() - stats[bn][0]).abs().max()),
            float((st.bn.invstd.cpu().double() * torch.sqrt(stats[bn][1] + 1e-3) - 1).abs().max())))
        x = torch.relu(torch.einsum("ncdhw,ck->nkdhw", u, p[dense + "/kernel"]))
        print("%-12s block out    rel-L2 %.4f" % (conv, rel(st.y.permute(0, 4, 1, 2, 3), x)))
    x = x[:, :, 0]
    for (stages, tail, tname, s, dy_t, x_out), (convs, _) in zip(net.blocks, rpn_blocks()):
        for (st, conv, bn), (_, _, _, _, stride) in zip(stages, convs):
            w = p[conv + "/kernel"].permute(3, 2, 0, 1)
            z = F.conv2d(x, w, p[conv + "/bias"], stride=stride, padding=1)
            x = torch.relu(TO._bn_train(z, p, bn, stats, channels_last=False))
            print("%-12s z rel-L2 %.4f   out rel-L2 %.4f" % (conv, rel(st.z[:, 0].permute(0, 3, 1, 2), z), rel(st.bn.y[:, 0].permute(0, 3, 1, 2), x)))
