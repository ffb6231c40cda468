"""ncu target: the LayerNorm kernels at the bench shape [23968, 768], each variant launched a few times.
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        -k regex:ln_ --csv python tools/ncu_ln.py
(gpu__time_duration is the only trustworthy clock for ~20 us kernels: an event pair also sees the host's launch gap)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from speechmix_b200 import kernels as K  # noqa: E402

R, C = 23968, 768
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(R, C, device="cuda", generator=g).to(torch.bfloat16)
r = torch.randn(R, C, device="cuda", generator=g).to(torch.bfloat16)
dy = torch.randn(R, C, device="cuda", generator=g).to(torch.bfloat16)
gamma = 1 + 0.1 * torch.randn(C, device="cuda", generator=g)
beta = 0.1 * torch.randn(C, device="cuda", generator=g)
for _ in range(3):
    y, s, mean, rstd = K.layernorm_fwd(x, gamma, beta)                          # x -> y
    y, s, mean, rstd = K.layernorm_fwd(x, gamma, beta, res=r, want_sum=True)    # x + res -> sum, y
    K.layernorm_bwd(dy, s, gamma, mean, rstd, want_colsum=True)                 # dy, x -> dx (+ column sums)
    K.layernorm_bwd(dy, s, gamma, mean, rstd, dres=r)                           # dy, x, dres -> dx
torch.cuda.synchronize()
