"""Quick timing of the encoder towers (CUDA events, warm, inputs resident): python scripts/encoder_bench.py [B]"""
import sys, torch
sys.path.insert(0, '/root/repo')
from clip_based_cross_modal_hash_b200 import encoder, synth
from oracle import clip_port as port

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
sd = synth.clip_state_dict(synth.VIT_B32, seed=0)
bb = encoder.ClipBackbone(sd)
img = synth.random_images(B, 1).cuda()
txt, pad = synth.random_captions(B, 2)
txt = txt.cuda()


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


ms = timeit(lambda: bb.encode_image(img))
fl = port.flops_image() * B
print('encode_image B=%d: %.3f ms  %.0f img/s  %.0f TFLOP/s' % (B, ms, B / ms * 1e3, fl / ms / 1e9))
ms = timeit(lambda: bb.encode_text(txt))
fl = port.flops_text() * B
print('encode_text  B=%d: %.3f ms  %.0f cap/s  %.0f TFLOP/s' % (B, ms, B / ms * 1e3, fl / ms / 1e9))
g = torch.cuda.CUDAGraph()
out = bb.encode_image(img)
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    bb.encode_image(img)
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s):
        out = bb.encode_image(img)
ms = timeit(lambda: g.replay())
print('encode_image (CUDA graph) B=%d: %.3f ms  %.0f img/s  %.0f TFLOP/s' % (B, ms, B / ms * 1e3, port.flops_image() * B / ms / 1e9))
