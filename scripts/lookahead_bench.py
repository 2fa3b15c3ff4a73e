import sys, torch
sys.path.insert(0, '/root/repo')
from clip_based_cross_modal_hash_b200 import _lib, encoder, synth
lib = _lib.lib()
sd = synth.clip_state_dict(synth.VIT_B32, seed=0)
bb = encoder.ClipBackbone(sd)
img = synth.random_images(256, 1).cuda()
def timeit(fn, iters=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for la in (1, 2, 3, 4, 6):
    lib.cmh_gemm_mma_lookahead(la)
    print('lookahead %d: encode_image %.3f ms' % (la, timeit(lambda: bb.encode_image(img))), flush=True)
