"""Run one GEMM configuration a few times (for ncu): python scripts/gemm_one.py M N K epi bn cg"""
import sys, torch
sys.path.insert(0, '/root/repo')
from clip_based_cross_modal_hash_b200 import _lib
lib = _lib.lib()
M, N, K, epi, bn, cg = (int(v) for v in sys.argv[1:7])
a = torch.randn(M, K, device='cuda').to(torch.bfloat16)
w = (torch.randn(N, K, device='cuda') * K ** -0.5).to(torch.bfloat16)
bias = torch.randn(N, device='cuda')
out = torch.zeros(M, N, dtype=torch.float32 if epi >= 2 else torch.bfloat16, device='cuda')
st = torch.cuda.current_stream().cuda_stream
lib.cmh_gemm_force_tile(bn, cg)
for _ in range(4):
    rc = lib.cmh_gemm_bf16(a.data_ptr(), M, K, K, w.data_ptr(), N, K, bias.data_ptr(), epi, out.data_ptr(), N,
                           out.data_ptr() if epi == 2 else None, N if epi == 2 else 0, st)
    assert rc == 0
torch.cuda.synchronize()
