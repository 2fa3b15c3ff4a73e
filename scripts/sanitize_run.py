"""Small end-to-end pass for compute-sanitizer (memcheck / racecheck): tiny CLIP towers, heads, MITH, retrieval."""
import sys, torch
sys.path.insert(0, '/root/repo')
from clip_based_cross_modal_hash_b200 import calc_utils, models, synth
sd = synth.clip_state_dict(synth.TINY, seed=3)
image = synth.random_images(5, seed=5)
text, pad = synth.random_captions(5, seed=6, vocab=synth.TINY["vocab_size"])
for cls, hs in ((models.DSPH, synth.dsph_head_state_dict), (models.DCMHT, synth.dcmht_head_state_dict), (models.MITH, synth.mith_head_state_dict)):
    m = cls(sd, hs(synth.TINY["embed_dim"], 32, seed=4))
    ci = m.encode_image_packed(image)
    ct = m.encode_text_packed(text, pad) if cls is models.MITH else m.encode_text_packed(text)
    torch.cuda.synchronize()
    print(cls.__name__, ci.shape, ct.shape)
qB, rB = synth.random_codes(40, 64, 1), synth.random_codes(3000, 64, 2)
qL, rL = synth.random_labels(40, 80, 3), synth.random_labels(3000, 80, 4)
print('mAP', float(calc_utils.calc_map_k(qB, rB, qL, rL, 100)), calc_utils.hamming_topk(qB, rB, 20)[0].shape)
torch.cuda.synchronize()
