"""Tile sweep of the tcgen05 GEMM on the encoder's shapes: python scripts/gemm_bench.py"""
import sys, torch
sys.path.insert(0, '/root/repo')
from clip_based_cross_modal_hash_b200 import _lib
lib = _lib.lib()


def run(M, N, K, epi, bn, cg, iters=20):
    a = torch.randn(M, K, device='cuda').to(torch.bfloat16)
    w = (torch.randn(N, K, device='cuda') * K ** -0.5).to(torch.bfloat16)
    bias = torch.randn(N, device='cuda')
    out = torch.zeros(M, N, dtype=torch.float32 if epi == 2 else torch.bfloat16, device='cuda')
    st = torch.cuda.current_stream().cuda_stream
    lib.cmh_gemm_force_tile(bn, cg)
    call = lambda: lib.cmh_gemm_bf16(a.data_ptr(), M, K, K, w.data_ptr(), N, K, bias.data_ptr(), epi, out.data_ptr(), N,
                                     out.data_ptr() if epi == 2 else None, N if epi == 2 else 0, st)
    for _ in range(3):
        assert call() == 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        call()
    e1.record()
    torch.cuda.synchronize()
    lib.cmh_gemm_force_tile(0, 0)
    return e0.elapsed_time(e1) / iters


shapes = [(12800, 2304, 768, 0, 'qkv'), (12800, 768, 768, 2, 'out+res'), (12800, 3072, 768, 1, 'fc+gelu'), (12800, 768, 3072, 2, 'proj+res'),
          (12544, 768, 3072, 0, 'patch'), (8192, 1536, 512, 0, 't.qkv'), (8192, 512, 512, 2, 't.out'), (8192, 2048, 512, 1, 't.fc'), (8192, 512, 2048, 2, 't.proj')]
for M, N, K, epi, name in shapes:
    line = '%-9s M=%5d N=%4d K=%4d |' % (name, M, N, K)
    for cg in (1, 2):
        for bn in (128, 192, 256):
            ms = run(M, N, K, epi, bn, cg)
            line += ' cg%d/bn%d %5.1fus %4.0fTF |' % (cg, bn, ms * 1e3, 2 * M * N * K / ms / 1e9)
    ms = run(M, N, K, epi, 0, 0)
    line += ' auto %5.1fus %4.0fTF' % (ms * 1e3, 2 * M * N * K / ms / 1e9)
    print(line, flush=True)
