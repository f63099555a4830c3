"""Runs the dominant GEMM (ViT mlp.c_fc shape, bias + QuickGELU, bf16 out) a few times -- target of `ncu --set full`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from generativeimage2text_b200 import _lib
lib = _lib.load()
M, N, K = 64 * 197, 3072, 768
a = torch.randn(M, K, device='cuda').to(torch.bfloat16)
w = (torch.randn(N, K, device='cuda') * 0.03).to(torch.bfloat16)
b = torch.randn(N, device='cuda')
o = torch.empty(M, N, dtype=torch.bfloat16, device='cuda')
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(5):
        rc = lib.gitb200_op_gemm(a.data_ptr(), w.data_ptr(), b.data_ptr(), None, o.data_ptr(), M, N, K, 1, 1, 0, 1, 0, s.cuda_stream)
        assert rc == 0
torch.cuda.synchronize()
print('done')
