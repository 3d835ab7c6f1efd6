import sys, numpy as np, torch
sys.path.insert(0, '.')
from gyre_b200.images import to_png_bytes, resize
from oracle import png as opng
from oracle.safety import synthetic_image
rng = np.random.default_rng(3)
for img in (synthetic_image(64, 96), rng.integers(0, 256, (33, 17, 3), dtype=np.uint8), np.zeros((40, 40, 3), np.uint8),
            rng.integers(0, 256, (20, 30, 4), dtype=np.uint8), synthetic_image(70, 300), synthetic_image(3, 10000), synthetic_image(300, 1)):
    got = to_png_bytes(torch.from_numpy(np.ascontiguousarray(img[None])).cuda())[0]
    assert got == opng.encode_png(img), img.shape
x = torch.rand(1, 1, 128, 128).cuda()
resize(x, (1 / 8, 1 / 8)); resize(x, (2.0, 0.5), sharpness=2)
torch.cuda.synchronize()
print("ok")

# safety-check front end and tail (resample_u8, clip_normalize, patchify, vision_embed, cosine_scores)
import os
from gyre_b200.safety_checker import B200FeatureExtractor, B200SafetyChecker
gold = torch.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "safety.pt"))
m = gold["models"]["tiny"]
sc = B200SafetyChecker({"vision_config": m["vision_config"], "projection_dim": m["projection_dim"]}).load_state_dict(m["state_dict"])
fx = B200FeatureExtractor(size=m["vision_config"]["image_size"])
pv = fx(torch.from_numpy(synthetic_image(100, 37)[None]).cuda()).pixel_values
assert sc(clip_input=m["clip_input"].cuda(), images=None)[1] == m["flags"]
sc.scores(pv)
torch.cuda.synchronize()
print("safety ok")
