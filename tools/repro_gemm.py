import sys, torch
sys.path.insert(0, ".")
from mmt_b200 import _lib
dev = "cuda"
Bt, Hh, S, dh = int(sys.argv[1]), 4, 218, 128
d = Hh * dh; Sp = 220
g = torch.Generator().manual_seed(1)
qkv = torch.randn(Bt * S, 3 * d, generator=g).to(dev)
Pn = torch.softmax(torch.randn(Bt, Hh, S, Sp, generator=g), -1).to(dev)
ctx = torch.zeros(Bt * S, d, device=dev)
bsP, bsQ = (Hh * S * Sp, S * Sp), (S * 3 * d, dh)
_lib.gemm(S, dh, S, Pn, Sp, 1, qkv, 1, 3 * d, ctx, d, b_off=2 * d, a_bs=bsP, b_bs=bsQ, c_bs=(S * d, dh),
          batch=Bt * Hh, batch_inner=Hh, precision=_lib.PREC_TF32)
torch.cuda.synchronize()
print("ok", float(ctx.abs().sum()))
