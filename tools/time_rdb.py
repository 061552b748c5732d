"""Timeline of the persistent RDB kernel at config-2 size (G forward only)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from trainner_b200.architectures import RRDBNet_arch
torch.manual_seed(0)
net = RRDBNet_arch.RRDBNet(3, 3, 64, 1).cuda()
x = torch.rand(16, 3, 64, 64, device="cuda")
with torch.no_grad():
    for _ in range(3): net(x)
torch.cuda.synchronize()
dbg = torch.zeros(148 * 32, dtype=torch.int64, device="cuda")
os.environ["B200_RDB_DBG_PTR"] = str(dbg.data_ptr())
with torch.no_grad(): net(x)
torch.cuda.synchronize()
del os.environ["B200_RDB_DBG_PTR"]
d = dbg.view(148, 32).cpu()
for cta in (0, 1, 70, 136):
    t0 = int(d[cta, 0])
    row = []
    for j in range(5):
        row.append("s%d[flags %d A %d mma %d st %d pub %d]" % tuple([j] + [int(d[cta, 1 + j * 6 + k]) - t0 for k in range(5)]))
    print("cta %3d: " % cta + " ".join(row))
