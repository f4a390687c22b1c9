import time, torch, sys
sys.path.insert(0, '/root/repo')
from prior_flow_b200 import corr as pcorr, ops
g = torch.Generator(device="cuda").manual_seed(0)
f1 = torch.randn(1, 256, 64, 128, device="cuda", generator=g); f2 = torch.randn(1, 256, 64, 128, device="cuda", generator=g)
with torch.no_grad():
    for fresh in (True, False):
        d = pcorr.DCCL(4, 4)
        ts = []
        for i in range(12):
            if fresh:
                d = pcorr.DCCL(4, 4)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            p = d.build_pyramid(pcorr.corr(f1, f2))
            t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
            ts.append(((t1 - t0) * 1e3, (t2 - t0) * 1e3))
        print("fresh DCCL per call" if fresh else "one DCCL", "cpu ms / total ms:", [f"{a:.2f}/{b:.2f}" for a, b in ts[2:8]])
    t0 = time.perf_counter(); 
    for _ in range(10): pcorr.available_device_memory(f1.device)
    print("available_device_memory ms", (time.perf_counter() - t0) * 100)
