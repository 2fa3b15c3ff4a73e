"""Does splitting the batch over two streams fill the GEMM tails?  python scripts/split_bench.py"""
import sys, torch
sys.path.insert(0, '/root/repo')
from clip_based_cross_modal_hash_b200 import encoder, synth
sd = synth.clip_state_dict(synth.VIT_B32, seed=0)
B = 256
img = synth.random_images(B, 1).cuda()
def timeit(fn, iters=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
bb = encoder.ClipBackbone(sd)
print('one stream, B=256: %.3f ms' % timeit(lambda: bb.encode_image(img)))
for parts in (2, 3, 4):
    bbs = [encoder.ClipBackbone(sd) for _ in range(parts)]   # separate workspaces
    streams = [torch.cuda.Stream() for _ in range(parts)]
    chunks = list(torch.chunk(img, parts))
    def split():
        cur = torch.cuda.current_stream()
        for s in streams: s.wait_stream(cur)
        for b, s, c in zip(bbs, streams, chunks):
            with torch.cuda.stream(s):
                b.encode_image(c)
        for s in streams: cur.wait_stream(s)
    print('%d streams x B=%d: %.3f ms' % (parts, B // parts, timeit(split)))
