import os, sys
sys.path.insert(0, "/root/repo")
import torch
import strainer_b200 as sb
L = sb._lib
lib = L.init(0)
st = L.P(torch.cuda.current_stream().cuda_stream)
n = 1 << 26
v = torch.empty(n, dtype=torch.float32, device="cuda").normal_().exp_()
ws = torch.empty(lib.sg_sort_workspace_bytes(n), dtype=torch.uint8, device="cuda")
so = torch.empty(n, dtype=torch.float32, device="cuda")
order = torch.empty(n, dtype=torch.int32, device="cuda")
for _ in range(2):
    L.check(lib.sg_sort_f32(L.P(v.data_ptr()), n, L.P(so.data_ptr()), L.P(order.data_ptr()), L.P(ws.data_ptr()), st))
torch.cuda.synchronize()
