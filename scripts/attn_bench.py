import sys, torch
sys.path.insert(0, '/root/repo')
from clip_based_cross_modal_hash_b200 import _lib
lib = _lib.lib()
st = torch.cuda.current_stream().cuda_stream
for B, L, H, causal in ((256, 50, 12, 0), (256, 32, 8, 1)):
    qkv = torch.randn(B * L, 3 * H * 64, device='cuda').to(torch.bfloat16)
    out = torch.empty(B * L, H * 64, device='cuda', dtype=torch.bfloat16)
    x = torch.randn(B * L, H * 64, device='cuda'); g = torch.randn(H * 64, device='cuda'); h = torch.empty_like(out)
    for name, fn in (('attention', lambda: lib.cmh_attention_bf16(qkv.data_ptr(), B, L, H, None, causal, out.data_ptr(), st)),
                     ('layernorm', lambda: lib.cmh_layernorm(x.data_ptr(), B * L, H * 64, g.data_ptr(), g.data_ptr(), 1e-5, h.data_ptr(), 0, st))):
        for _ in range(3): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 50
        nbytes = (qkv.numel() + out.numel()) * 2 if name == 'attention' else x.numel() * 4 + h.numel() * 2
        print('%s B=%d L=%d H=%d: %.1f us, %.2f TB/s' % (name, B, L, H, ms * 1e3, nbytes / ms / 1e9))
