# SPDX-License-Identifier: Apache-2.0
"""How does the SM clock behave around short kernels? (bring-up only)"""
import time
import torch

def probe(cycles=2_000_000):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); torch.cuda._sleep(cycles); b.record(); torch.cuda.synchronize()
    return cycles / (a.elapsed_time(b) * 1e-3) / 1e9

torch.cuda.init()
x = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
print("cold:", [round(probe(), 2) for _ in range(5)])
for n in (5, 20, 100, 400):
    for _ in range(n):
        x @ x
    torch.cuda.synchronize()
    print(f"after {n} matmuls:", [round(probe(), 2) for _ in range(3)])
time.sleep(0.05)
print("after 50 ms idle:", [round(probe(), 2) for _ in range(3)])
time.sleep(0.5)
print("after 500 ms idle:", [round(probe(), 2) for _ in range(3)])
# flush-dominated loop (like bench): clock while mostly memory-bound fills run
for _ in range(200):
    flush.fill_(1)
torch.cuda.synchronize()
print("after 200 fills:", [round(probe(), 2) for _ in range(3)])
for i in range(5):
    for _ in range(20):
        flush.fill_(1); torch.cuda._sleep(200000)
    print("  fill+sleep loop:", round(probe(200000), 2))
