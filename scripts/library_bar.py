"""The "library kernels" bar of SURVEY 8(d): the same DanSpeechPrimary-shaped network (3 conv + 9 x 1200 bi-GRU + fc +
softmax, batch 64 x 15 s) written with stock torch.nn modules, i.e. cuDNN convolutions / cuDNN GRU / cuBLAS on the same
B200, timed with CUDA events next to this repo's hand-written path.  Not part of the product or of bench.py; it uses
neither oracle/ nor the reference.  Prints one JSON line.

    python scripts/library_bar.py            # fp32 (TF32 allowed, torch defaults for cuDNN) and bf16 autocast
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402

B, SECONDS, H, LAYERS, C = 64, 15, 1200, 9, 33
T = 1 + SECONDS * 16000 // 160
dev = torch.device("cuda")


class Net(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv = nn.Sequential(
            nn.Conv2d(1, 32, (41, 11), stride=(2, 2), padding=(20, 5)), nn.BatchNorm2d(32), nn.Hardtanh(0, 20),
            nn.Conv2d(32, 32, (21, 11), stride=(2, 1), padding=(10, 5)), nn.BatchNorm2d(32), nn.Hardtanh(0, 20),
            nn.Conv2d(32, 96, (21, 11), stride=(2, 1), padding=(10, 5)), nn.BatchNorm2d(96), nn.Hardtanh(0, 20))
        self.norms = nn.ModuleList([nn.BatchNorm1d(H) for _ in range(LAYERS - 1)])
        self.rnns = nn.ModuleList([nn.GRU(21 * 96 if i == 0 else H, H, bidirectional=True) for i in range(LAYERS)])
        self.fc = nn.Sequential(nn.BatchNorm1d(H), nn.Linear(H, C, bias=False))

    def forward(self, x):
        x = self.conv(x)                                   # equal lengths: the time mask is the identity
        x = x.view(x.size(0), -1, x.size(3)).permute(2, 0, 1).contiguous()
        for i, rnn in enumerate(self.rnns):
            if i:
                t, b, h = x.shape
                x = self.norms[i - 1](x.view(t * b, h)).view(t, b, h)
            x, _ = rnn(x)
            x = x[..., :H] + x[..., H:]
        t, b, h = x.shape
        x = self.fc(x.view(t * b, h)).view(t, b, C).transpose(0, 1)
        return torch.softmax(x, dim=-1).argmax(dim=-1)


def timed(fn, warm=3, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


torch.manual_seed(0)
net = Net().to(dev).eval()
x = torch.randn(B, 1, 161, T, device=dev)
out = {"workload": "stock torch.nn DeepSpeech2 of the DanSpeechPrimary shape, batch %d x %d s, model forward + argmax" % (B, SECONDS),
       "torch": torch.__version__, "cudnn": torch.backends.cudnn.version(),
       "cudnn_allow_tf32": torch.backends.cudnn.allow_tf32, "matmul_allow_tf32": torch.backends.cuda.matmul.allow_tf32}
with torch.no_grad():
    ms = timed(lambda: net(x))
    out["fp32"] = {"ms_per_batch": ms, "rtfx": B * SECONDS / (ms / 1e3)}
    with torch.autocast("cuda", dtype=torch.bfloat16):
        ms = timed(lambda: net(x))
    out["bf16_autocast"] = {"ms_per_batch": ms, "rtfx": B * SECONDS / (ms / 1e3)}
    half = Net().to(dev).eval().half()
    xh = x.half()
    ms = timed(lambda: half(xh))
    out["fp16_weights"] = {"ms_per_batch": ms, "rtfx": B * SECONDS / (ms / 1e3)}
print(json.dumps(out))
