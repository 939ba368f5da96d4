import sys, numpy as np
sys.path.insert(0, '.')
import zosimos_b200 as Z
from zosimos_b200 import _ffi, ops
from zosimos_b200.buffer import *
use_tma = int(sys.argv[1])
W,H,w,h = 512,512,157,151
d = lambda w,h: Descriptor(ByteLayout(w,h,w*4,4), Color.SRGB, Texel.new_u8(SampleParts.RgbA))
rng = np.random.default_rng(0)
bg = rng.integers(0,256,(H,W*4),dtype=np.uint8); fg = rng.integers(0,256,(h,w*4),dtype=np.uint8)
ctx = Z.Context(0)
below, above, dst = ctx.upload(d(W,H), bg), ctx.upload(d(w,h), fg), ctx.image(d(W,H))
ops.compose(ctx, below, above, dst, ops.compose_params(sel=(0,0,w,h), tgt=(13,7,w,h), use_tma=bool(use_tma)))
ctx.sync()
got = dst.download().reshape(H,W,4)
exp = bg.reshape(H,W,4).copy(); exp[7:7+h,13:13+w] = fg.reshape(h,w,4)
print("use_tma", use_tma, "equal", np.array_equal(got, exp))
