import io, sys
import numpy as np, torch
from PIL import Image
sys.path.insert(0, ".")
from facet_b200 import ops
from facet_b200.synth import synth_image_bgr
from facet_b200.utils import jpeg as fj
H, W = 4000, 6000
ri = int(sys.argv[1]) if len(sys.argv) > 1 else 25
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16
datas = []
for i in range(2):
    buf = io.BytesIO()
    Image.fromarray(synth_image_bgr(2000 + i, H, W)[:, :, ::-1].copy()).save(buf, "JPEG", quality=90, restart_marker_blocks=ri)
    datas.append(np.frombuffer(buf.getvalue(), np.uint8).copy())
streams = [datas[i % 2] for i in range(n)]
infos = [fj.parse(s) for s in streams]
slot = (max(len(s) for s in streams) + 255) & ~255
buf = torch.empty(n * slot + 256, dtype=torch.uint8, device="cuda")
for k, s in enumerate(streams):
    buf[k * slot:k * slot + s.size].copy_(torch.from_numpy(s))
torch.cuda.synchronize()
for _ in range(2):
    ops.jpeg_decode_device(buf, slot, infos)
torch.cuda.synchronize()
