"""Developer probe (GPU box): how does this torch divide a CUDA tensor by a Python scalar?"""
import numpy as np
import torch
x = torch.rand(1 << 20) - 0.5
for b in (0.28209479177387814, 448, 256.0, 3.0):
    g = (x.cuda() / b).cpu()
    true_div = x / b
    recip = x * (np.float32(1.0) / np.float32(b)).item()
    print(b, "cuda==cpu true division:", bool(torch.equal(g, true_div)), " cuda==x*(1/b):", bool(torch.equal(g, recip)),
          " mismatches vs true:", int((g != true_div).sum()), " vs recip:", int((g != recip).sum()))
print(torch.__version__)
