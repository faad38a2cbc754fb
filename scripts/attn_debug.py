import sys, torch
sys.path.insert(0, ".")
from facet_b200 import ops
torch.manual_seed(0)
bsz = 2
qkv = (torch.randn(bsz * 257, 3072, device="cuda") * 1.5).to(torch.bfloat16)
got = ops.vit_attention(qkv, bsz).float().reshape(bsz, 257, 16, 64)
torch.cuda.synchronize()
q, k, v = qkv.float().reshape(bsz, 257, 3, 16, 64).unbind(2)
att = torch.softmax(torch.einsum("bqhd,bkhd->bhqk", q, k) * 0.125, dim=-1)
ref = torch.einsum("bhqk,bkhd->bqhd", att, v)
err = (got - ref).abs()
print("max err", float(err.max()), "mean", float(err.mean()))
print("err by token block:", [float(err[:, a:b].max()) for a, b in [(0, 128), (128, 256), (256, 257)]])
print("err by d half:", float(err[..., :32].max()), float(err[..., 32:].max()))
# variant without key 256 / structure hints
print("sample got", got[0, 5, 3, :6].tolist()); print("sample ref", ref[0, 5, 3, :6].tolist())
