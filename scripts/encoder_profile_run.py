"""One warm encode_image + encode_text at B=256 (for ncu launch lists): python scripts/encoder_profile_run.py [B] [iters]"""
import sys, torch
sys.path.insert(0, '/root/repo')
from clip_based_cross_modal_hash_b200 import encoder, synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
sd = synth.clip_state_dict(synth.VIT_B32, seed=0)
bb = encoder.ClipBackbone(sd)
img = synth.random_images(B, 1).cuda()
txt, pad = synth.random_captions(B, 2)
txt = txt.cuda()
for _ in range(iters):
    bb.encode_image(img)
    bb.encode_text(txt)
torch.cuda.synchronize()
