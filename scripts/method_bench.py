"""Per-method encode time at B = 256 (images resident): tower + head + packed codes.  python scripts/method_bench.py"""
import sys, torch
sys.path.insert(0, '/root/repo')
from clip_based_cross_modal_hash_b200 import models, synth
sd = synth.clip_state_dict(synth.VIT_B32, seed=0)
B = 256
img = synth.random_images(B, 1).cuda()
txt, pad = synth.random_captions(B, 2)
txt, pad = txt.cuda(), pad.cuda()
def timeit(fn, iters=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for name, cls, hs, nbits in (('DSPH', models.DSPH, synth.dsph_head_state_dict, 64), ('DCMHT', models.DCMHT, synth.dcmht_head_state_dict, 64),
                             ('MITH', models.MITH, synth.mith_head_state_dict, 128)):
    m = cls(sd, hs(512, nbits, seed=1))
    ti = timeit(lambda: m.encode_image_packed(img))
    tt = timeit(lambda: (m.encode_text_packed(txt, pad) if name == 'MITH' else m.encode_text_packed(txt)))
    print('%-5s %3d bit: image %.3f ms (%.0f img/s)  text %.3f ms (%.0f cap/s)' % (name, nbits, ti, B / ti * 1e3, tt, B / tt * 1e3), flush=True)
m = models.MITH(sd, synth.mith_head_state_dict(512, 128, seed=1))
m.hash.split_precision = False; m.hash.refresh()
print('MITH 128 bit, plain bf16 token MLPs: image %.3f ms' % timeit(lambda: m.encode_image_packed(img)))
