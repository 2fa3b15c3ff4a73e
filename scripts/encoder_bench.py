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

# image and text towers on two streams (do they fill each other's tails?)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s1):
        bb.encode_image(img)
    with torch.cuda.stream(s2):
        bb.encode_text(txt)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
ms = timeit(both)
print('image + text on two streams B=%d: %.3f ms' % (B, ms))
def seq():
    bb.encode_image(img); bb.encode_text(txt)
ms = timeit(seq)
print('image then text on one stream B=%d: %.3f ms' % (B, ms))
