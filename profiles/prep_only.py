import sys, os
sys.path.insert(0, os.getcwd())
import torch
from gptq_gguf_toolkit_b200 import ops
n = 4096
x = torch.randn(2 * n, n, device="cuda")
H = (x.T @ x) / n
W = torch.randn(256, n, device="cuda")
ops.prepare(H.clone(), W, 0.01)
torch.cuda.synchronize()
