"""Per-CTA timeline of one GEMM (SM clock stamps): python scripts/gemm_trace.py M N K epi bn cg"""
import sys, torch
sys.path.insert(0, '/root/repo')
from clip_based_cross_modal_hash_b200 import _lib
lib = _lib.lib()
M, N, K, epi, bn, cg = (int(v) for v in sys.argv[1:7])
epi_code = epi
a = torch.randn(M, K, device='cuda').to(torch.bfloat16)
w = (torch.randn(N, K, device='cuda') * K ** -0.5).to(torch.bfloat16)
bias = torch.randn(N, device='cuda')
out = torch.zeros(M, N, dtype=torch.float32 if epi >= 2 else torch.bfloat16, device='cuda')
st = torch.cuda.current_stream().cuda_stream
lib.cmh_gemm_force_tile(bn, cg)
if len(sys.argv) > 7:
    lib.cmh_gemm_force_units(int(sys.argv[7]))
trace = torch.zeros(148, 64, dtype=torch.int64, device='cuda')
call = lambda: lib.cmh_gemm_bf16(a.data_ptr(), M, K, K, w.data_ptr(), N, K, bias.data_ptr(), epi, out.data_ptr(), N,
                                 out.data_ptr() if epi == 2 else None, N if epi == 2 else 0, st)
for _ in range(3):
    call()
lib.cmh_gemm_set_trace(trace.data_ptr())
call()
torch.cuda.synchronize()
lib.cmh_gemm_set_trace(None)
t = trace.cpu()
kb = (K + 63) // 64
r = t[0]
per = [(int(r[5 + it * 4]) - int(r[4 + it * 4])) / kb for it in range(8) if int(r[5 + it * 4])]
epi = [int(r[35 + it * 2]) - int(r[34 + it * 2]) for it in range(8) if int(r[35 + it * 2])]
print('M=%d N=%d K=%d epi=%d bn=%d cg=%d units=%s | pdl_wait %d | first load %d | cycles per k-block %s | epilogue cycles %s | exit-last_mma %d' % (
    M, N, K, epi_code, bn, cg, sys.argv[7] if len(sys.argv) > 7 else 'all', int(r[2]) - int(r[1]), int(r[4]) - int(r[3]),
    ' '.join('%.0f' % v for v in per), ' '.join(str(v) for v in epi), int(r[63]) - max(int(r[5 + it * 4]) for it in range(8))))
